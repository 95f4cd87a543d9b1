"""TEST INFRASTRUCTURE (build container only) -- mint golden vectors from the UNMODIFIED reference.

    python -m oracle.mint_goldens

For each model family: build vmg_b200.VMG with seed 0 (its random init is the shared state_dict),
load that state_dict into the reference `models.vmg.VMG` (strict=True), run the reference forward twice
on a seeded input (the second call exercises the in-place Morph-FC weight decay) and store sub-sampled
outputs + statistics under tests/golden/.  The fixtures travel to the GPU box; /root/reference does not.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader                      # noqa: E402
from tests import common                           # noqa: E402
import tests.emu_ops as emu                        # noqa: E402
import vmg_b200.model as vm                        # noqa: E402

CASES = [('few', (1, 7, 3, 64, 64)), ('full', (1, 7, 3, 64, 64)), ('few', (2, 4, 3, 72, 68)), ('full', (1, 6, 3, 68, 80))]


def main():
    Ref = ref_loader.load_reference_vmg_class()
    for family, shape in CASES:
        kw = common.kwargs_of(family)
        mine = common.build(family, seed=0)
        sd0 = {k: v.clone() for k, v in mine.state_dict().items()}
        ref = Ref(**kw).eval()
        ref.load_state_dict(sd0, strict=True)
        ref.num_out_frames = shape[1]
        x = common.seeded_input(shape, seed=1)
        with torch.no_grad():
            y1 = ref(x)
            y2 = ref(x)
        sd2 = ref.state_dict()
        name = f"{family}_{'x'.join(map(str, shape))}.pt"
        torch.save({
            'family': family, 'shape': shape, 'seed_w': 0, 'seed_x': 1,
            'sd_checksum': common.sd_checksum(sd0), 'sd_checksum_after2': common.sd_checksum(sd2),
            'x_checksum': float(x.double().sum()),
            'y1_sub': y1[..., ::4, ::4].clone(), 'y2_sub': y2[..., ::4, ::4].clone(),
            'y1_stats': [y1.mean().item(), y1.std().item(), y1.min().item(), y1.max().item()],
            'y2_stats': [y2.mean().item(), y2.std().item(), y2.min().item(), y2.max().item()],
            'torch': str(torch.__version__),
        }, os.path.join(common.GOLDEN, name))
        print('minted', name, 'mean %.8f' % y1.mean().item(), 'frames_mirror', ref.frames_mirror)


if __name__ == '__main__':
    main()
