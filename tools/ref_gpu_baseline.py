"""The reference's own PyTorch-CUDA arithmetic on this B200, for context (the north-star's "x3" denominator):
runs the oracle restatement (same ATen / cuDNN calls as the reference forward, PyTorch defaults: cuDNN TF32 on,
matmul TF32 off) with all tensors on cuda:0 and reports 720p frames/s.  Test/measurement infrastructure only.
Usage: python tools/ref_gpu_baseline.py [--workload NAME] [--bf16]"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import vmg_oracle  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default=bench.DEFAULT_WORKLOAD)
    ap.add_argument('--bf16', action='store_true', help='torch.autocast(bfloat16)')
    ap.add_argument('--steps', type=int, default=2)
    a = ap.parse_args()
    from vmg_b200.model import VMG
    _kind, yml, B, T, H, W = bench.WORKLOADS[a.workload]
    kw = bench.model_kwargs(yml)
    torch.manual_seed(0)
    sd = {k: v.clone().cuda() for k, v in VMG(**kw).state_dict().items()}
    hp = vmg_oracle.Hyper(**kw)
    hp.num_out_frames = T
    x = bench.synthetic_clip(B, T, H, W, 1).cuda()
    # the oracle builds its meshgrids on the CPU default device: run it under a cuda default device
    torch.set_default_device('cuda')
    times = []
    with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16, enabled=a.bf16):
        for i in range(1 + a.steps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            y = vmg_oracle.forward(sd, hp, x, skip_dead_flow_levels=False)
            torch.cuda.synchronize()
            if i:
                times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    print(json.dumps({'what': 'reference arithmetic (oracle on cuda, PyTorch eager)', 'workload': a.workload, 'bf16_autocast': a.bf16,
                      'seconds_per_forward': sec, 'fps_720p': B * T * H * W / bench.REF_PIX / sec,
                      'peak_mem_GB': torch.cuda.max_memory_allocated() / 1e9}))


if __name__ == '__main__':
    main()
