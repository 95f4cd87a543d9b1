"""Per-kernel microbenchmarks (CUDA events, L2 flushed between iterations) for the shapes that dominate the
VMG forward.  Usage: python tools/microbench.py [conv|rows|all] [--iters N]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vmg_b200 import ops  # noqa: E402

DEV = 'cuda'


def timeit(fn, iters, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.add_(1.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def conv_cases():
    # name, N,H,W,Cin,Cout,R,groups,n_src,epi,act,ps
    return [
        ('traj_res_112', 1, 184, 320, 112, 112, 3, 1, 1, ops.EPI_RESADD, ops.ACT_NONE, 0),
        ('traj_res_112_store', 1, 184, 320, 112, 112, 3, 1, 1, ops.EPI_STORE, ops.ACT_RELU, 0),
        ('traj_res_144', 1, 184, 320, 144, 144, 3, 1, 1, ops.EPI_RESADD, ops.ACT_NONE, 0),
        ('traj_in_224to112', 1, 184, 320, 224, 112, 3, 1, 2, ops.EPI_STORE, ops.ACT_LRELU, 0),
        ('traj_res_112_b2', 2, 184, 320, 112, 112, 3, 1, 1, ops.EPI_RESADD, ops.ACT_NONE, 0),
        ('traj_res_112_b4', 4, 184, 320, 112, 112, 3, 1, 1, ops.EPI_RESADD, ops.ACT_NONE, 0),
        ('traj_res_112_b4_store', 4, 184, 320, 112, 112, 3, 1, 1, ops.EPI_STORE, ops.ACT_RELU, 0),
        ('tiny_1round_112', 1, 64, 296, 112, 112, 3, 1, 1, ops.EPI_STORE, ops.ACT_RELU, 0),
        ('ffn_fc1_g4_112', 25, 184, 320, 112, 672, 3, 4, 1, ops.EPI_STORE, ops.ACT_GELU, 0),
        ('ffn_fc1_g4_112_relu', 25, 184, 320, 112, 672, 3, 4, 1, ops.EPI_STORE, ops.ACT_RELU, 0),
        ('ffn_fc1_g4_112_none', 25, 184, 320, 112, 672, 3, 4, 1, ops.EPI_STORE, ops.ACT_NONE, 0),
        ('ffn_fc1_L1_g4_224', 25, 92, 160, 224, 1344, 3, 4, 1, ops.EPI_STORE, ops.ACT_GELU, 0),
        ('ffn_fc2_672to112', 1, 1, 25 * 184 * 320, 672, 112, 1, 1, 1, ops.EPI_RESADD, ops.ACT_NONE, 0),
        ('linear_112', 1, 1, 25 * 184 * 320, 112, 112, 1, 1, 1, ops.EPI_STORE, ops.ACT_RELU, 0),
        ('gate_proj_112', 1, 1, 25 * 184 * 320, 112, 112, 1, 1, 1, ops.EPI_GATE, ops.ACT_NONE, 0),
        ('gate_proj_224', 1, 1, 25 * 92 * 160, 224, 224, 1, 1, 1, ops.EPI_GATE, ops.ACT_NONE, 0),
        ('linear_448', 1, 1, 100 * 23 * 40, 448, 448, 1, 1, 1, ops.EPI_STORE, ops.ACT_RELU, 0),
        ('linear_224', 1, 1, 100 * 92 * 160, 224, 224, 1, 1, 1, ops.EPI_STORE, ops.ACT_RELU, 0),
        ('linear_224_L2', 1, 1, 100 * 46 * 80, 224, 224, 1, 1, 1, ops.EPI_STORE, ops.ACT_RELU, 0),
        ('ffn_fc2_672to112_100f', 1, 1, 100 * 184 * 320, 672, 112, 1, 1, 1, ops.EPI_RESADD, ops.ACT_NONE, 0),
        ('ffn_fc2_L1_1344to224', 1, 1, 25 * 92 * 160, 1344, 224, 1, 1, 1, ops.EPI_RESADD, ops.ACT_NONE, 0),
        ('upconv2_112to256_ps', 4, 360, 640, 112, 256, 3, 1, 1, ops.EPI_STORE, ops.ACT_LRELU, 2),
        ('hrconv_64', 2, 720, 1280, 64, 64, 3, 1, 1, ops.EPI_STORE, ops.ACT_LRELU, 0),
        ('conv_last_64to3', 2, 720, 1280, 64, 3, 3, 1, 1, ops.EPI_STORE, ops.ACT_NONE, 0),
        ('spynet_8to32_7x7', 99, 192, 320, 8, 32, 7, 1, 1, ops.EPI_STORE, ops.ACT_RELU, 0),
        ('spynet_32to64_7x7', 99, 192, 320, 32, 64, 7, 1, 1, ops.EPI_STORE, ops.ACT_RELU, 0),
        ('spynet_64to32_7x7', 99, 192, 320, 64, 32, 7, 1, 1, ops.EPI_STORE, ops.ACT_RELU, 0),
        ('spynet_32to16_7x7', 99, 192, 320, 32, 16, 7, 1, 1, ops.EPI_STORE, ops.ACT_RELU, 0),
        ('spynet_16to2_7x7', 99, 192, 320, 16, 2, 7, 1, 1, ops.EPI_RESADD, ops.ACT_NONE, 0),
        ('spynet_8to32_7x7_l3', 99, 24, 40, 8, 32, 7, 1, 1, ops.EPI_STORE, ops.ACT_RELU, 0),
    ]


def run_conv(iters, only=None, variants=(0,)):
    flush = torch.zeros(64 * 1024 * 1024, device=DEV)      # 256 MB > L2
    out = []
    for name, N, H, W, Cin, Cout, R, groups, n_src, epi, act, ps in conv_cases():
        if only and only not in name:
            continue
        w = torch.randn(Cout, Cin // groups, R, R, device=DEV) * 0.05
        b = torch.randn(Cout, device=DEV)
        pc = ops.pack_conv(w, b, groups, tc=True)
        xs = tuple(torch.randn(N, H, W, Cin // n_src, device=DEV).bfloat16() for _ in range(n_src))
        oshape = (N, 2 * H, 2 * W, Cout // 4) if ps else (N, H, W, Cout)
        o = torch.empty(oshape, device=DEV, dtype=torch.float32 if Cout == 3 else torch.bfloat16)
        res = torch.randn(oshape, device=DEV).bfloat16() if epi != ops.EPI_STORE else None
        aux = torch.randn(oshape, device=DEV).bfloat16() if epi == ops.EPI_GATE else None
        ref = None
        if len(variants) > 1:            # cross-check every variant against the per-tap kernel
            ops.conv2d(xs if n_src > 1 else xs[0], pc, act=act, alpha=0.1, epi=epi, res=res, aux=aux, res_scale=0.1, pixel_shuffle=ps, out=o, tc_variant=1)
            ref = o.float().clone()
        for variant in variants:
            if (1 < variant < 8 and R != 3) or (variant == 8 and R not in (3, 7)) or (variant > 8 and variant != 13 and variant < 32 and R != 3):
                continue
            fn = lambda: ops.conv2d(xs if n_src > 1 else xs[0], pc, act=act, alpha=0.1, epi=epi, res=res, aux=aux, res_scale=0.1, pixel_shuffle=ps, out=o, tc_variant=variant)
            try:
                ms = timeit(fn, iters, flush)
            except RuntimeError as e:
                print(json.dumps({'case': name, 'variant': variant, 'error': str(e)[:80]}), flush=True)
                continue
            fl = 2.0 * N * H * W * Cout * R * R * (Cin // groups)
            rec = {'case': name, 'variant': variant, 'ms': round(ms, 4), 'tflops': round(fl / ms / 1e9, 1), 'gflop': round(fl / 1e9, 2)}
            if ref is not None:
                rec['maxdiff_vs_v1'] = float((o.float() - ref).abs().max())
            print(json.dumps(rec), flush=True)
            out.append(rec)
    return out


def run_rows(iters):
    flush = torch.zeros(64 * 1024 * 1024, device=DEV)
    N, H, W, C = 25, 184, 320, 112
    x = torch.randn(N, H, W, C, device=DEV).bfloat16()
    g, b = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    nbytes = x.numel() * 2
    cases = {
        'layernorm': (lambda: ops.layernorm(x, g, b), 2 * nbytes),
        'morph_gather_h8': (lambda: ops.morph_gather(x, 8, 0), 2 * nbytes),
        'morph_gather_w16': (lambda: ops.morph_gather(x, 16, 1), 2 * nbytes),
        'add': (lambda: ops.add(x, x), 3 * nbytes),
        'channel_sum3': (lambda: ops.channel_sum(x, x, x, N, 1, H * W, 1.0), 3 * nbytes),
        'l2norm': (lambda: ops.l2norm(x), 2 * nbytes),
    }
    A = ops.morph_gather(x, 8, 0)
    cases['morph_scatter_h8'] = (lambda: ops.morph_scatter(A, N, H, W, C, 8, 0), 2 * nbytes)
    a3 = torch.rand(1, 3, C, device=DEV)
    cases['mix3'] = (lambda: ops.mix3(x, x, x, a3, N, 1, H * W), 4 * nbytes)
    flow = torch.randn(1, H, W, 2, device=DEV) * 4
    x1 = x[:1].contiguous()
    cases['flow_warp_bilinear'] = (lambda: ops.flow_warp_bilinear(x1, flow), 2 * x1.numel() * 2)
    T = 34
    keys = torch.randn(T, 1, H, W, C, device=DEV).bfloat16()
    loc = torch.zeros(1, H, W, 2 * T, device=DEV)
    for k in range(T):
        ops.loc_identity_(loc, 2 * k)
    rel, dec = torch.zeros(4, 4, 4, device=DEV), torch.full((4,), 0.97, device=DEV)
    for K in (4, 17, 33):
        cases[f'ltam_K{K}'] = ((lambda K=K: ops.ltam_wins(keys[0], keys[0], keys[0], H * W * C, K, loc, rel, dec, 4)),
                               (2 * K * C + 2 * C) * H * W * 2)
    for name, (fn, nb) in cases.items():
        ms = timeit(fn, iters, flush)
        print(json.dumps({'case': name, 'ms': round(ms, 4), 'GBps': round(nb / ms / 1e6, 1)}), flush=True)


def run_tab(iters, frames=25):
    """The fused Morph-FC mixing half (csrc/tab_mix.cu) at the level-0 shape: phase 0 (stats) and phase 1 (apply)."""
    flush = torch.zeros(64 * 1024 * 1024, device=DEV)
    N, H, W, C = frames, 184, 320, 112
    x = torch.randn(N, H, W, C, device=DEV).bfloat16()
    g, b = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    pcs = tuple(ops.pack_conv(torch.randn(C, C, device=DEV) * 0.1, torch.randn(C, device=DEV) * 0.1, 1, tc=True) for _ in range(4))
    a = torch.full((1, 3, C), 1.0 / 3, device=DEV)
    nbytes = x.numel() * 2
    for name, fn, nb in (('tab_mix_phase0', lambda: ops.tab_mix(x, g, b, 1e-5, pcs, 0, 1), nbytes),
                         ('tab_mix_phase1', lambda: ops.tab_mix(x, g, b, 1e-5, pcs, 1, 1, a=a), 2 * nbytes)):
        ms = timeit(fn, iters, flush)
        print(json.dumps({'case': name, 'frames': N, 'ms': round(ms, 4), 'GBps': round(nb / ms / 1e6, 1)}), flush=True)


def run_chain(iters):
    """The 15-block residual chain of one trajectory step: ONE persistent kernel vs 30 per-layer launches."""
    flush = torch.zeros(64 * 1024 * 1024, device=DEV)
    for n, H, W, C, nb in ((2, 184, 320, 112, 15), (2, 128, 128, 112, 15), (2, 184, 320, 144, 15), (2, 128, 128, 144, 15)):
        layers = [ops.pack_conv(torch.randn(C, C, 3, 3, device=DEV) * 0.03, torch.randn(C, device=DEV) * 0.1, 1, tc=True) for _ in range(2 * nb)]
        pc = ops.pack_resblock_chain(layers)
        x = torch.randn(n, H, W, C, device=DEV).bfloat16()
        out = torch.empty_like(x)
        flags = ops.chain_flags(x)
        scratch = (torch.empty_like(x), torch.empty_like(x))
        state = {'i': 0}

        def chain():
            i = state['i']
            state['i'] += 1
            ops.resblock_chain(x, pc, 0.1, out, flags[i & 1], flags[(i + 1) & 1], scratch)

        def per_layer():
            y = x
            for k in range(nb):
                t = ops.conv2d(y, layers[2 * k], act=ops.ACT_RELU)
                y = ops.conv2d(t, layers[2 * k + 1], epi=ops.EPI_RESADD, res=y, res_scale=0.1)
        fl = 2.0 * n * H * W * C * C * 9 * 2 * nb
        for name, fn in (('chain', chain), ('per_layer', per_layer)):
            if name == 'chain' and not ops.resblock_chain_supported(pc, x):
                continue
            ms = timeit(fn, iters, flush)
            print(json.dumps({'case': f'{name}_{n}x{H}x{W}x{C}x{nb}', 'ms': round(ms, 4), 'us_per_layer': round(1e3 * ms / (2 * nb), 2),
                              'tflops': round(fl / ms / 1e9, 1)}), flush=True)


def run_ltam(iters):
    """LTAM at the headline trajectory shape (2 images of 184 x 320, C = 112) with flow-like sampling locations
    (identity + a few pixels of jitter, some outside the frame), projection bias and anchor as in the model."""
    flush = torch.zeros(64 * 1024 * 1024, device=DEV)
    n, H, W, C, T = 2, 184, 320, 112, 34
    g = torch.Generator(device='cpu').manual_seed(0)
    keys = torch.nn.functional.normalize(torch.randn(T, n, H, W, C, generator=g), dim=-1).bfloat16().to(DEV)
    vals = torch.randn(T, n, H, W, C, generator=g).bfloat16().to(DEV)
    loc = torch.zeros(n, H, W, 2 * T, device=DEV)
    for k in range(T):
        ops.loc_identity_(loc, 2 * k)
    loc += (torch.randn(n, H, W, 2 * T, generator=g) * 3).round().to(DEV)
    rel = (torch.randn(4, 4, 4, generator=g) * 0.05).to(DEV)
    dec = torch.full((4,), 0.97, device=DEV)
    anchor = vals[1]
    bias = torch.zeros(C, device=DEV)
    outs = {}
    for K in (1, 4, 9, 17, 33):
        fn = lambda K=K: ops.ltam_wins(keys[0], keys[0], vals[0], n * H * W * C, K, loc, rel, dec, 4, out_bias=bias, anchor=anchor)
        ms = timeit(fn, iters, flush)
        outs[K] = fn().float()
        nb = (2 * K * C + 3 * C) * n * H * W * 2
        print(json.dumps({'case': f'ltam_K{K}', 'ms': round(ms, 4), 'GBps': round(nb / ms / 1e6, 1),
                          'checksum': float(outs[K].double().abs().mean())}), flush=True)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('what', nargs='?', default='all')
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--only', default=None)
    ap.add_argument('--variants', default='0')
    a = ap.parse_args()
    if a.what in ('conv', 'all'):
        run_conv(a.iters, a.only, tuple(int(v) for v in a.variants.split(',')))
    if a.what in ('rows', 'all'):
        run_rows(a.iters)
    if a.what in ('chain', 'all'):
        run_chain(a.iters)
    if a.what in ('tab', 'all'):
        run_tab(a.iters)
    if a.what == 'ltam':
        run_ltam(a.iters)
    if a.what == 'tab_small':
        run_tab(1, frames=4)
