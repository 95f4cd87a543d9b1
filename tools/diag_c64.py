"""Diagnostic: the C = 64 residual-block sequence on the forced CTA-pair kernel vs cuDNN, with / without syncs."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import emu_ops as emu
from vmg_b200 import ops
DEV = 'cuda'
torch.backends.cudnn.allow_tf32 = False
def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(*shape, generator=g) * 2 - 1) * scale
print('env', {k: v for k, v in os.environ.items() if k.startswith('VMGB')})
for (n, H, W, C, nb) in ((2, 128, 128, 64, 4), (2, 128, 128, 112, 4), (2, 128, 128, 128, 2)):
    ws = [(rnd(C, C, 3, 3, seed=10 + i, scale=1.5 * (C * 9) ** -0.5).bfloat16().float(), rnd(C, seed=50 + i, scale=0.2)) for i in range(2 * nb)]
    pcs = [ops.pack_conv(w.to(DEV), b.to(DEV), 1, tc=True) for w, b in ws]
    pes = [emu.pack_conv(w.to(DEV), b.to(DEV)) for w, b in ws]
    x = rnd(n, H, W, C, seed=3).bfloat16().to(DEV)
    for variant in (13, 0):
        for sync in (False, True):
            y, yr = x, x
            errs = []
            for k in range(nb):
                t = ops.conv2d(y, pcs[2 * k], act=ops.ACT_RELU, tc_variant=variant)
                if sync: torch.cuda.synchronize()
                y2 = ops.conv2d(t, pcs[2 * k + 1], epi=ops.EPI_RESADD, res=y, res_scale=0.1, tc_variant=variant)
                if sync: torch.cuda.synchronize()
                tr = emu.conv2d(yr, pes[2 * k], act=ops.ACT_RELU)
                yr2 = emu.conv2d(tr, pes[2 * k + 1], epi=ops.EPI_RESADD, res=yr, res_scale=0.1)
                torch.cuda.synchronize()
                errs.append(((t.float() - tr.float()).abs().max().item(), (y2.float() - yr2.float()).abs().max().item()))
                y, yr = y2, yr2
            print(f'{n}x{H}x{W}x{C} variant {variant} sync {sync}:', ' '.join(f'({a:.3f},{b:.3f})' for a, b in errs), flush=True)
