"""Time every conv2d / linear launch of one eager forward with CUDA events and aggregate by problem shape.
Usage: python tools/conv_breakdown.py [--workload NAME]"""
import argparse
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vmg_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default=bench.DEFAULT_WORKLOAD)
    a = ap.parse_args()
    from vmg_b200.model import VMG
    import vmg_b200.model as vm
    _kind, yml, B, T, H, W = bench.WORKLOADS[a.workload]
    torch.manual_seed(0)
    m = VMG(**bench.model_kwargs(yml), precision='bf16').eval().cuda()
    m.num_out_frames = T
    m.use_cuda_graph = False
    x = bench.synthetic_clip(B, T, H, W, 1).cuda()
    m(x)
    recs = []
    real = ops.conv2d

    def timed(xx, pc, **kw):
        xs = xx if isinstance(xx, (tuple, list)) else (xx,)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        y = real(xx, pc, **kw)
        e.record()
        N, H_, W_, _ = xs[0].shape
        key = (N * H_ * W_, pc.Cin, pc.Cout, pc.R, pc.groups, len(xs), kw.get('act', 0), kw.get('epi', 0), kw.get('pixel_shuffle', 0), str(xs[0].dtype)[6:])
        recs.append((key, s, e, 2.0 * N * H_ * W_ * pc.Cout * pc.R * pc.S * (pc.Cin // pc.groups)))
        return y
    ops.conv2d = timed
    m(x)
    torch.cuda.synchronize()
    ops.conv2d = real
    agg = collections.OrderedDict()
    for key, s, e, fl in recs:
        v = agg.setdefault(key, [0, 0.0, 0.0])
        v[0] += 1
        v[1] += s.elapsed_time(e)
        v[2] += fl
    tot = sum(v[1] for v in agg.values())
    print('| M (pixels) | Cin | Cout | R | groups | srcs | act | epi | ps | dtype | launches | total ms | TFLOP/s | share |')
    print('|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|---:|---:|---:|---:|')
    for key, (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print('| ' + ' | '.join(str(k) for k in key) + f' | {n} | {ms:.2f} | {fl / ms / 1e9:.0f} | {100 * ms / tot:.1f} % |')
    print(f'total conv/linear ms: {tot:.1f} over {len(recs)} launches')


if __name__ == '__main__':
    main()
