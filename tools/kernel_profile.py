"""Per-kernel device-time breakdown of one eager VMG forward via torch.profiler (CUPTI) -- cheap enough for
the full headline workload (ncu launch lists are taken on smaller workloads).
Usage: python tools/kernel_profile.py [--workload NAME] [--out profiles/x.md]"""
import argparse
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default=bench.DEFAULT_WORKLOAD)
    ap.add_argument('--precision', default='bf16')
    ap.add_argument('--out', default=None)
    a = ap.parse_args()
    from vmg_b200.model import VMG
    _kind, yml, B, T, H, W = bench.WORKLOADS[a.workload]
    torch.manual_seed(0)
    m = VMG(**bench.model_kwargs(yml), precision=a.precision).eval().cuda()
    m.num_out_frames = T
    m.use_cuda_graph = False
    x = bench.synthetic_clip(B, T, H, W, 1).cuda()
    m(x)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        m(x)
        torch.cuda.synchronize()
    agg = collections.OrderedDict()
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            k = ev.name.split('(')[0][:90]
            e = agg.setdefault(k, [0, 0.0])
            e[0] += 1
            e[1] += ev.device_time_total if hasattr(ev, 'device_time_total') else ev.cuda_time_total
    tot = sum(v[1] for v in agg.values()) or 1.0
    lines = [f'workload {a.workload} ({a.precision}), one eager forward, CUPTI kernel times', '',
             '| kernel | launches | total ms | mean us | share |', '|---|---:|---:|---:|---:|']
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f'| `{k}` | {n} | {us / 1e3:.2f} | {us / n:.1f} | {100 * us / tot:.1f} % |')
    lines.append(f'| **total** | {sum(v[0] for v in agg.values())} | {tot / 1e3:.2f} | | |')
    text = '\n'.join(lines)
    print(text)
    from vmg_b200 import ops
    print('bf16 launches on the generic kernel:', sorted(set(ops.FALLBACK_LOG)))
    if a.out:
        open(a.out, 'w').write(text + '\n')


if __name__ == '__main__':
    main()
