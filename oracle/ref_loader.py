"""TEST INFRASTRUCTURE (container only) -- import the unmodified reference from /root/reference.

Used by `oracle/mint_goldens.py` and by the container-only tests that pin `oracle/vmg_oracle.py`
against the real reference.  The GPU box has no /root/reference; everything that must run there
uses the committed fixtures under tests/golden/ instead.
"""
import os
import sys

REF_ROOT = os.environ.get('VMG_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REF_ROOT, 'models', 'vmg.py'))


def load_reference_vmg_class():
    if not available():
        raise RuntimeError(f'reference not found at {REF_ROOT}')
    from . import ref_shims
    ref_shims.install()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        from models.vmg import VMG  # noqa: the reference's own class
    return VMG
