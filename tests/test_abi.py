"""C-ABI checks that need no GPU: the library loads and exports every symbol include/vmgb.h declares."""
import ctypes
import os

import pytest

from vmg_b200 import ops


def test_header_parses():
    protos = ops.parse_header()
    assert len(protos) >= 35
    for name in ('vmgb_conv2d_nhwc', 'vmgb_conv2d_nhwc_tc', 'vmgb_ltam_wins', 'vmgb_flow_warp_bilinear',
                 'vmgb_morph_permute', 'vmgb_layernorm', 'vmgb_final_output', 'vmgb_version'):
        assert name in protos


def test_library_exports_every_declared_symbol():
    assert os.path.isfile(ops.LIB_PATH), 'run __graft_entry__.build() first'
    L = ctypes.CDLL(ops.LIB_PATH)
    missing = [n for n in ops.parse_header() if not hasattr(L, n)]
    assert not missing, missing


def test_version_and_arch_without_gpu():
    L = ops.lib()
    assert L.vmgb_version() >= 100
    assert L.vmgb_arch() == b'sm_100a'


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(RuntimeError):
        ops.layernorm(torch.zeros(4, 8), torch.ones(8), torch.zeros(8))
    from tests import common
    m = common.build('few', T=7)
    with pytest.raises(RuntimeError):
        m(torch.rand(1, 7, 3, 64, 64))


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dp, _, fs in os.walk(os.path.join(root, 'vmg_b200')):
        for f in fs:
            if f.endswith('.py'):
                src = open(os.path.join(dp, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src and 'emu_ops' not in src, f
