"""One launch of every hot kernel family at its headline shape (fewer frames), for ONE `ncu --set full` capture:
    ncu --set full --clock-control none --import-source on -k regex:'vmgb::(conv|layernorm|rownorm|morph|channel_sum|mix3|ltam|flow_warp|resample|tab_)' \
        -o gpurun_out/suite python tools/ncu_suite.py
Each case is launched exactly once after the weights are packed (ncu replays it itself).  `--frames N` sets the image batch."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vmg_b200 import ops  # noqa: E402

DEV = 'cuda'


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=6)
    ap.add_argument('--only', default=None)
    a = ap.parse_args()
    N, H, W, C = a.frames, 184, 320, 112
    g = torch.Generator(device=DEV).manual_seed(0)
    rnd = lambda *s: torch.randn(*s, device=DEV, generator=g)
    x = rnd(N, H, W, C).bfloat16()
    gam, bet = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)

    def pk(cout, cin, r, groups=1):
        return ops.pack_conv(rnd(cout, cin // groups, r, r) * 0.05, rnd(cout) * 0.1, groups, tc=True)
    fc1, fc2, lin, proj = pk(672, 112, 3, 4), pk(112, 672, 1), pk(112, 112, 1), pk(112, 112, 1)
    traj_in = pk(112, 224, 3)
    up2 = pk(256, 112, 3)
    hidden = rnd(N, H, W, 672).bfloat16()
    A = ops.morph_gather(x, 8, 0)
    a3 = torch.rand(1, 3, C, device=DEV)
    T = 18
    keys = rnd(T, 2, H, W, C).bfloat16()
    loc = torch.zeros(2, H, W, 2 * T, device=DEV)
    for k in range(T):
        ops.loc_identity_(loc, 2 * k)
    loc2 = torch.empty_like(loc)
    flow = rnd(2, H, W, 2) * 3
    rel, dec = torch.zeros(4, 4, 4, device=DEV), torch.full((4,), 0.97, device=DEV)
    tab_pcs = tuple(pk(112, 112, 1) for _ in range(4))
    tab_a = torch.full((1, 3, C), 1.0 / 3, device=DEV)
    rgb = rnd(N, H, W, 3)
    rgb_pc = ops.pack_conv(rnd(112, 3, 3, 3) * 0.1, rnd(112) * 0.1, 1, tc=False)
    x1 = rnd(4 * N, 92, 160, 224).bfloat16()                 # level-1 activations (chunk 16, C = 224): the unfused TAB glue
    lin1, proj1 = pk(224, 224, 1), pk(224, 224, 1)
    A1 = ops.morph_gather(x1, 16, 0)
    a31 = torch.rand(1, 3, 224, device=DEV)
    torch.cuda.synchronize()
    cases = {
        'tab_mix_phase0': lambda: ops.tab_mix(x, gam, bet, 1e-5, tab_pcs, 0, 1),
        'tab_mix_phase1': lambda: ops.tab_mix(x, gam, bet, 1e-5, tab_pcs, 1, 1, a=tab_a),
        'rgb_input_conv': lambda: ops.conv2d(rgb, rgb_pc, act=ops.ACT_LRELU, alpha=0.01, out_dtype=torch.bfloat16, use_tc=False),
        'l1_linear_relu_224': lambda: ops.linear(x1, lin1, act=ops.ACT_RELU, out_scale=1.0 / 224),
        'l1_proj_gate_224': lambda: ops.linear(x1, proj1, epi=ops.EPI_GATE, res=x1, aux=x1),
        'l1_morph_gather_16': lambda: ops.morph_gather(x1, 16, 1),
        'l1_morph_scatter_16': lambda: ops.morph_scatter(A1, 4 * N, 92, 160, 224, 16, 0),
        'l1_channel_sum': lambda: ops.channel_sum(x1, x1, x1, 4 * N, 1, 92 * 160, 1.0),
        'l1_mix3': lambda: ops.mix3(x1, x1, x1, a31, 4 * N, 1, 92 * 160),
        'l1_layernorm_224': lambda: ops.layernorm(x1, torch.ones(224, device=DEV), torch.zeros(224, device=DEV)),
        'ffn_fc1': lambda: ops.conv2d(x, fc1, act=ops.ACT_GELU),
        'ffn_fc2': lambda: ops.linear(hidden, fc2, epi=ops.EPI_RESADD, res=x),
        'linear_relu': lambda: ops.linear(x, lin, act=ops.ACT_RELU, out_scale=1.0 / C),
        'proj_gate': lambda: ops.linear(x, proj, epi=ops.EPI_GATE, res=x, aux=x),
        'traj_input_conv': lambda: ops.conv2d((x[:2], x[2:4]), traj_in, act=ops.ACT_LRELU, alpha=0.1),
        'upconv2_ps': lambda: ops.conv2d(x[:2], up2, act=ops.ACT_LRELU, alpha=0.1, pixel_shuffle=2),
        'layernorm': lambda: ops.layernorm(x, gam, bet),
        'morph_gather': lambda: ops.morph_gather(x, 8, 1),
        'morph_scatter': lambda: ops.morph_scatter(A, N, H, W, C, 8, 0),
        'channel_sum': lambda: ops.channel_sum(x, x, x, N, 1, H * W, 1.0),
        'mix3': lambda: ops.mix3(x, x, x, a3, N, 1, H * W),
        'l2norm': lambda: ops.l2norm(x),
        'resample_ln_down': lambda: ops.resample_ln(x, None, torch.ones(4 * C, device=DEV), torch.zeros(4 * C, device=DEV), up=False),
        'ltam_K17': lambda: ops.ltam_wins(keys[0], keys[0], keys[0], 2 * H * W * C, 17, loc, rel, dec, 4),
        'flow_warp_fused': lambda: ops.flow_warp_fused(keys[1], flow, loc, 2 * 17, loc2),
    }
    for name, fn in cases.items():
        if a.only and a.only not in name:
            continue
        fn()
        torch.cuda.synchronize()
        print('ran', name, flush=True)


if __name__ == '__main__':
    main()
