#!/usr/bin/env python
"""bench.py -- benchmarks of the VMG model-forward hot path (BASELINE.json metric: 720p output frames/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

Default workload = BASELINE.json configs[1]: VMG-REDS full model, one synthetic REDS4-shaped clip of 100 frames
180x320 -> 720x1280 per GPU, bf16 tcgen05 path.  A "step" is one VMG forward over that clip.  Multi-GPU = one process
per GPU (torchrun), one clip per rank, no data-path collective (clips are independent): weak scaling.

ONE JSON line on rank 0:
  value         720p frames/s, inputs resident in HBM (CUDA events, max over ranks)
  e2e           same through the public nn.Module call with pinned-host input and output copies in the timed region
                (the read-back of step i overlaps the forward of step i+1 on a copy stream; all copies end inside the region)
  roofline      tensor-pipe fraction of the dominant kernel from per-launch CUDA events of one instrumented forward
  cpu_baseline  the CPU oracle (oracle/vmg_oracle.py, a port of the reference forward) on a bounded sample (N=1 only)
  gpu_reference the reference's own arithmetic on this GPU: the oracle restatement (the same ATen / cuDNN calls as
                models/vmg.py) with all tensors on cuda:0 and PyTorch defaults -- the north-star's "x3" denominator --
                timed OUTSIDE the product's timed region (N=1 only); `speedup` = value / its fps
  parity        max-abs / dPSNR of the benched model's output against that oracle in strict fp32 on the same weights
  peak_mem_GB   torch.cuda.max_memory_allocated of the product arm
  tester_route  BASELINE.json configs[2] next to it: 4 clips x 100 frames through the reference's eval route
                (3 temporal windows x 6 tiles = 72 units of [1,50,3,128,128], few_levels model), units sharded over the
                N ranks, tiles of a clip that spans two ranks gathered with NCCL send/recv, blended + quantised on device:
                STRONG scaling (total work fixed)

Other workloads (--workload): reds4_few_100x180x320, reds4_few_tester_4x100 (the tester route as the main line),
vimeo_{full,few}_7x64x112_B{8,16,32,64} (BASELINE configs[3]: many short clips batched along B),
uhd_full_16x540x960_tiles (BASELINE configs[4]: one 540x960 clip as 45 overlapped tiles over the ranks, outputs gathered
with NCCL send/recv), tile_full_50x128x128, small_full_7x64x64.

`--impl reference` times the CPU oracle as the reference arm (the reference is pure PyTorch; its own CPU path is what
tools/cpu_test.sh runs).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = '720p output frames/sec (4x VSR, REDS shape)'
UNIT = 'frames/s'
REF_PIX = 180 * 320           # LQ pixels behind one 720p output frame

WORKLOADS = {
    # name: (kind, config yml, B, T, H, W)
    'reds4_full_100x180x320': ('forward', 'VMG-REDS.yml', 1, 100, 180, 320),
    'reds4_few_100x180x320': ('forward', 'VMG-REDS-few_levels.yml', 1, 100, 180, 320),
    'tile_full_50x128x128': ('forward', 'VMG-REDS.yml', 1, 50, 128, 128),
    'small_full_7x64x64': ('forward', 'VMG-REDS.yml', 1, 7, 64, 64),
    # Tester route (tools/Tester.py semantics): n_clips clips of T x H x W, units sharded over the ranks (strong scaling)
    'reds4_few_tester_4x100': ('tester', 'VMG-REDS-few_levels.yml', 4, 100, 180, 320),
    'uhd_full_16x540x960_tiles': ('tester', 'VMG-REDS.yml', 1, 16, 540, 960),
}
for _fam, _yml in (('full', 'VMG-REDS.yml'), ('few', 'VMG-REDS-few_levels.yml')):
    for _b in (8, 16, 32, 64):
        WORKLOADS[f'vimeo_{_fam}_7x64x112_B{_b}'] = ('forward', _yml, _b, 7, 64, 112)
DEFAULT_WORKLOAD = 'reds4_full_100x180x320'
TESTER_ROUTE = 'reds4_few_tester_4x100'
# dataset section of the eval YAMLs (configs/VMG-few_levels-eval-REDS4.yml:18-41 of the reference)
TESTER_DATASET = {
    'reds4_few_tester_4x100': {'name': 'REDS4', 'num_frames': 50, 'overlapped_mode': 'mid', 'wins': [128, 128], 'overlapped_spatial_length': 20},
    'uhd_full_16x540x960_tiles': {'name': 'REDS4', 'num_frames': 16, 'overlapped_mode': 'mid', 'wins': [128, 128], 'overlapped_spatial_length': 20},
}
CPU_SAMPLE = (1, 7, 64, 64)   # bounded CPU sample of the same model (FLOPs are linear in T*H*W)


def synthetic_clip(B, T, H, W, seed):
    """Deterministic smooth moving field in [0,1] (sum of seeded low-frequency sinusoids translating
    <= 2 px/frame) -- realistic scatter for the warp / trajectory kernels, unlike white noise."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing='ij')
    clip = torch.zeros(B, T, 3, H, W)
    for b in range(B):
        vel = (torch.rand(2, generator=g) * 4 - 2)
        for c in range(3):
            acc = torch.zeros(T, H, W)
            for _ in range(8):
                fx, fy = (torch.rand(2, generator=g) * 0.25 + 0.01).tolist()
                ph = (torch.rand(1, generator=g) * 6.28).item()
                amp = (torch.rand(1, generator=g) * 0.5 + 0.1).item()
                t = torch.arange(T, dtype=torch.float32)[:, None, None]
                acc += amp * torch.sin(fx * (xx[None] + vel[0] * t) + fy * (yy[None] + vel[1] * t) + ph)
            clip[b, :, c] = acc
    clip = (clip - clip.amin()) / (clip.amax() - clip.amin())
    return clip.contiguous()


def model_kwargs(yml):
    from vmg_b200 import config as cfg
    c = cfg.parse_config(os.path.join(ROOT, 'configs', yml), is_train=False)
    return cfg.model_kwargs(c)


class ClockSampler:
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.FIELDS}',
                                          '--format=csv,noheader,nounits', '-lms', '200'], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([s.strip() for s in line.split(',')])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace('.', '').isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith('active') for r in self.rows)]
        busy = [s for s in sm if s > 0]
        return {'sm_mhz': busy[len(busy) // 2] if busy else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(self.rows)}


def run_cpu_oracle(kw, steps, warmup, threads):
    """Times the CPU oracle (port of the reference forward) on the bounded sample.  Returns (fps, s/step)."""
    from oracle import vmg_oracle
    from vmg_b200.model import VMG
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    sd = {k: v.clone() for k, v in VMG(**kw).state_dict().items()}
    B, T, H, W = CPU_SAMPLE
    x = synthetic_clip(B, T, H, W, seed=1)
    hp = vmg_oracle.Hyper(**kw)
    hp.num_out_frames = T
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            vmg_oracle.forward(sd, hp, x)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return B * T * H * W / REF_PIX / sec, sec


def psnr(a, b):
    mse = torch.mean((a.double().clamp(0, 1) - b.double().clamp(0, 1)) ** 2).item()
    return 99.0 if mse == 0 else 10.0 * torch.log10(torch.tensor(1.0 / mse)).item()


def run_gpu_reference(kw, sd0, x_host, y_product, timed=2):
    """The reference's own arithmetic on this GPU: the oracle restatement on cuda:0.
    O2 = PyTorch defaults (cuDNN TF32 on): 1 warm + `timed` forwards -> fps, peak memory.
    O3 = strict fp32 (allow_tf32 off): the parity reference for the benched output `y_product` (host tensor)."""
    from oracle import vmg_oracle
    hp = vmg_oracle.Hyper(**kw)
    hp.num_out_frames = x_host.shape[1]
    B, T, _, H, W = x_host.shape
    old_default = torch.get_default_device()
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    out = {}
    try:
        torch.set_default_device('cuda')          # the oracle builds its meshgrids on the default device
        x = x_host.cuda()
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
        times = []
        with torch.no_grad():
            for i in range(1 + timed):
                sd = {k: v.detach().clone().cuda() for k, v in sd0.items()}
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                y = vmg_oracle.forward(sd, hp, x)
                torch.cuda.synchronize()
                if i:
                    times.append(time.perf_counter() - t0)
                del y
            sec = sum(times) / len(times)
            out['gpu_reference'] = {'fps': B * T * H * W / REF_PIX / sec, 'seconds': sec, 'peak_mem_GB': torch.cuda.max_memory_allocated() / 1e9,
                                    'what': 'oracle restatement of models/vmg.py (same ATen / cuDNN calls) on cuda:0, PyTorch defaults '
                                            '(cuDNN TF32 convolutions, fp32 matmuls), eager, 1 warm + %d timed forwards' % timed}
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
            sd = {k: v.detach().clone().cuda() for k, v in sd0.items()}
            y_ref = vmg_oracle.forward(sd, hp, x).cpu()
        tgt = torch.rand(y_ref.shape, generator=torch.Generator(device='cpu').manual_seed(8), device='cpu')
        per_t = (y_product - y_ref).abs().amax(dim=(0, 2, 3, 4))
        out['parity'] = {'max_abs': per_t.max().item(), 'mean_abs': (y_product - y_ref).abs().mean().item(),
                         'dpsnr': abs(psnr(y_product, tgt) - psnr(y_ref, tgt)), 'finite': bool(torch.isfinite(y_product).all()),
                         'max_abs_first_last_10_frames': [per_t[:10].max().item(), per_t[-10:].max().item()],
                         'against': 'oracle on cuda:0, strict fp32 (cudnn.allow_tf32=False), same random-init state_dict and input',
                         'tolerance': {'max_abs': 5e-3, 'dpsnr': 0.02}}
    finally:
        torch.set_default_device(old_default)
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------------------------
# Tester route (BASELINE configs[2] and [4])
# ----------------------------------------------------------------------------------------------------------------
def run_tester_route(name, steps, warmup, rank, world, precision='bf16'):
    """Returns the result dict of the Tester-route workload `name` (all ranks call; timing = max over ranks)."""
    import torch.distributed as dist
    from vmg_b200 import frame_io, shard
    from vmg_b200.model import VMG
    from vmg_b200.tester import Tester
    _, yml, n_clips, T, H, W = WORKLOADS[name]
    kw = model_kwargs(yml)
    torch.manual_seed(0)
    model = VMG(**kw, precision=precision).eval().cuda()
    tester = Tester({'scale': 4, 'dataset': TESTER_DATASET[name]}, model=model)
    units = shard.enumerate_units(tester, n_clips, T, H, W)
    n_units = len(units)
    frames_host = [(synthetic_clip(1, T, H, W, seed=1 + c)[0].permute(0, 2, 3, 1) * 255.0).round().to(torch.uint8).contiguous().pin_memory()
                   for c in range(n_clips)]
    up = frame_io.FrameUploader()
    clips = [up.upload(f) for f in frames_host]
    torch.cuda.synchronize()
    owners = sorted({units[u][0]: u * world // n_units for u in reversed(range(n_units))}.items())
    out_host = {c: torch.empty((T, 4 * H, 4 * W, 3), dtype=torch.uint8).pin_memory() for c, o in owners if o == rank}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # the stateful weight decay continues from pass to pass like one long sequential reference run (the unit index keeps growing)
    state = {'base': 0, 'done': 0}

    def one_pass(e2e):
        cl = [up.upload(f) for f in frames_host] if e2e else clips
        res = shard.run_tester_route_sharded(tester, cl, rank, world, exact_decay=True, state=state)
        if e2e:
            for c, v in res.items():
                out_host[c].copy_(v, non_blocking=True)
        return res

    for _ in range(max(1, warmup)):
        one_pass(False)
    from vmg_b200 import ops
    barrier()
    n0 = ops.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        one_pass(False)
    ev1.record()
    barrier()
    n_blk = len(shard.assign_blocks(n_units, rank, world))
    launches = ops.launch_count() - n0 + steps * n_blk * getattr(model, 'last_graph_launches', 0) * (1 if model.use_cuda_graph else 0)
    ms = max_over_ranks(ev0.elapsed_time(ev1)) / steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        one_pass(True)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    frames = n_clips * T * (H * W / REF_PIX)
    moved = sum(1 for u in range(n_units) if (u * world // n_units) != dict(owners)[units[u][0]])
    return {'workload': name, 'model_config': yml, 'clips': n_clips, 'clip': [T, 3, H, W], 'units': n_units, 'unit_shape': [1, min(T, TESTER_DATASET[name]['num_frames']), 3, min(H, 128), min(W, 128)],
            'dataset': TESTER_DATASET[name], 'scaling': 'strong', 'value': frames / (ms * 1e-3), 'unit': UNIT, 'ms_per_pass': ms,
            'e2e': {'value': frames / (e2e_ms * 1e-3), 'unit': UNIT, 'ms_per_pass': e2e_ms,
                    'h2d_bytes_per_step': sum(f.numel() for f in frames_host), 'd2h_bytes_per_step': sum(v.numel() for v in out_host.values())},
            'units_per_rank': n_blk, 'tiles_moved_over_nvlink': moved, 'gather': 'NCCL isend/irecv of finished tiles to the clip owner' if world > 1 else 'none (1 rank)',
            'decay': 'exact: unit n runs on the weights of the n-th forward of a sequential reference run',
            'gpu_launches': launches, 'peak_mem_GB': torch.cuda.max_memory_allocated() / 1e9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='vmg_b200', choices=['vmg_b200', 'reference'])
    ap.add_argument('--workload', default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-gpu-reference', action='store_true')
    ap.add_argument('--no-tester-route', action='store_true')
    ap.add_argument('--profile-mode', action='store_true', help='warm-up + timed steps only (for ncu runs)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    kind, yml, B, T, H, W = WORKLOADS[args.workload]
    kw = model_kwargs(yml)
    config = {'workload': args.workload, 'model_config': yml, 'clip': [B, T, 3, H, W], 'clips_per_gpu': 1,
              'parallelism': f'clip-sharded x{world}' if world > 1 else 'single GPU',
              'l2_policy': 'activations per step are GBs (>> 126 MB L2); no explicit flush needed'}

    # ------------------------------------------------------------------ reference arm (CPU oracle)
    if args.impl == 'reference':
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        fps, sec = run_cpu_oracle(kw, max(1, args.steps), max(1, min(args.warmup, 1)), threads)
        sample = f'oracle forward on a {CPU_SAMPLE[1]}x{CPU_SAMPLE[2]}x{CPU_SAMPLE[3]} crop of the workload, same model, fp32, warm'
        print(json.dumps({'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
                          'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak' if kind == 'forward' else 'strong',
                          'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config,
                          'cpu_baseline': {'value': fps, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample},
                          'e2e': {'value': fps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return

    # ------------------------------------------------------------------ product arm
    import torch.distributed as dist
    from vmg_b200 import ops
    from vmg_b200.model import VMG
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    if kind == 'tester':
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        r = run_tester_route(args.workload, max(1, args.steps), max(1, min(args.warmup, 2)), rank, world, args.precision)
        clocks = sampler.stop() if rank == 0 else None
        if rank == 0:
            config.update({'clips': r['clips'], 'units': r['units'], 'unit_shape': r['unit_shape'], 'dataset': r['dataset'],
                           'parallelism': f'{r["units"]} (clip, window, tile) units in contiguous blocks over {world} rank(s); ' + r['gather']})
            config.pop('clips_per_gpu')
            print(json.dumps({'metric': METRIC, 'value': r['value'], 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                              'ms_per_step': r['ms_per_pass'], 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
                              'dtype': args.precision, 'data': 'synthetic', 'config': config, 'clocks': clocks, 'e2e': r['e2e'],
                              'gpu_launches': r['gpu_launches'], 'peak_mem_GB': r['peak_mem_GB'], 'tester_route': r}))
        if world > 1:
            dist.destroy_process_group()
        return

    torch.manual_seed(0)
    model = VMG(**kw, precision=args.precision).eval().cuda()
    model.num_out_frames = T
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}      # random-init weights before any in-place decay
    x_host = synthetic_clip(B, T, H, W, seed=1 + rank).pin_memory()
    x_dev = x_host.cuda(non_blocking=True)
    out_host = torch.empty((B, T, 3, 4 * H, 4 * W), dtype=torch.float32).pin_memory()
    frames_per_step = B * T * H * W / REF_PIX

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # parity input: the very first forward (random-init weights, decay index 0) is kept for the comparison with the oracle
    y_first = model(x_dev)
    finite = bool(torch.isfinite(y_first).all())
    want_ref = rank == 0 and world == 1 and not args.no_gpu_reference and not args.profile_mode and args.precision == 'bf16'
    y_first_host = y_first.cpu() if want_ref else None
    del y_first
    for _ in range(max(0, args.warmup - 1)):
        model(x_dev)
    # ---- device-resident timing
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    n0 = ops.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        y = model(x_dev)
    ev1.record()
    barrier()
    launches = ops.launch_count() - n0 + args.steps * getattr(model, 'last_graph_launches', 0) * (1 if model.use_cuda_graph else 0)
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = world * frames_per_step / (ms_step * 1e-3)
    if args.profile_mode:
        if rank == 0:
            print(json.dumps({'profile_mode': True, 'ms_per_step': ms_step, 'value': value, 'gpu_launches': launches}))
        return
    # ---- end to end: pinned host input -> H2D -> forward -> D2H of the output frames
    # The read-back of step i (1.1 GB over PCIe) runs on a copy stream while step i+1 computes, into alternating pinned
    # buffers -- what a serving loop around the public call does; every copy starts and finishes inside the timed region
    # (the end event waits for the copy stream).
    model(x_host.cuda(non_blocking=True))
    out_hosts = [out_host, torch.empty_like(out_host).pin_memory()]
    copy_stream = torch.cuda.Stream()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        xd = x_host.cuda(non_blocking=True)
        y = model(xd)
        done = torch.cuda.Event()
        done.record()
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done)
            out_hosts[i & 1].copy_(y, non_blocking=True)
            y.record_stream(copy_stream)
    torch.cuda.current_stream().wait_stream(copy_stream)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    e2e = {'value': world * frames_per_step / (e2e_ms * 1e-3), 'unit': UNIT,
           'h2d_bytes_per_step': x_host.numel() * 4, 'd2h_bytes_per_step': out_host.numel() * 4, 'ms_per_step': e2e_ms}
    peak_mem = torch.cuda.max_memory_allocated() / 1e9
    # ---- roofline of the dominant kernel: per-launch CUDA events on the launching stream (instrumented forward)
    roofline = None
    if rank == 0:
        recs = []
        real = ops.lib().vmgb_conv2d_nhwc_tc
        real_chain = ops.lib().vmgb_resblock_chain

        def timed(*a):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            rc = real(*a)
            e.record()
            # C prototype positions 17..24 = n_img,H,W,Cin,Cout,R,S,groups ; 25/27 = act / epilogue
            n_img, h, w, cin, cout, r, s_, g = a[17:25]
            nbytes = 2.0 * n_img * h * w * (cin + cout + (cout if a[28] else 0)) + 2.0 * r * r * (cin // g) * cout
            recs.append((s, e, 2.0 * n_img * h * w * cout * r * s_ * (cin // g), ('conv', n_img, h, w, cin, cout, r, g), nbytes, 1))
            return rc

        def timed_chain(*a):
            # C prototype: y, y_alt, t_tmp, out, w, Kp, bias, flags, flags_next, n_img, H, W, C, n_blocks, res_scale, stream
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            rc = real_chain(*a)
            e.record()
            n_img, h, w, c, nb = a[9:14]
            if rc == 0:       # one launch = 2*nb dependent 3x3 convolutions; compulsory bytes: chain input + output + all weights
                recs.append((s, e, 2.0 * n_img * h * w * c * 9 * c * 2 * nb, ('chain', n_img, h, w, c, c, 3, 1),
                             2.0 * n_img * h * w * 2 * c + 2.0 * 2 * nb * 9 * c * c, 2 * nb))
            return rc
        ops.lib().vmgb_conv2d_nhwc_tc = timed
        ops.lib().vmgb_resblock_chain = timed_chain
        use_graph = model.use_cuda_graph
        model.use_cuda_graph = False          # eager launches so that each one can be bracketed by events
        try:
            model(x_dev)
            torch.cuda.synchronize()
        finally:
            ops.lib().vmgb_conv2d_nhwc_tc = real
            ops.lib().vmgb_resblock_chain = real_chain
            model.use_cuda_graph = use_graph
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        # the dominant kernel is timed inside a long step (thousands of back-to-back launches): sustained figure
        peak = peaks.get('bf16_tflops_sustained', 1400.0)
        src = 'MEASURED_PEAKS.json bf16_tflops_sustained (of measured)' if 'bf16_tflops_sustained' in peaks else 'fallback 1.4 PFLOP/s sustained (of fallback)'
        if recs:
            # group the tcgen05 launches by (kernel family, shape); the dominant kernel = the group with the most device time
            # (the two epilogue flavours of a convolution shape -- ReLU store / residual add -- are one kernel)
            by_shape = {}
            for s, e, fl, key, nbytes, layers in recs:
                b = by_shape.setdefault(key, [0, 0.0, 0.0, 0.0, layers])
                b[0] += 1; b[1] += s.elapsed_time(e); b[2] += fl; b[3] += nbytes
            all_ms = sum(v[1] for v in by_shape.values())
            all_fl = sum(v[2] for v in by_shape.values())
            dom_key, dom = max(by_shape.items(), key=lambda kv: kv[1][1])
            ach = dom[2] / (dom[1] * 1e-3) / 1e12
            fam, n_img, h, w, cin, cout, r, g = dom_key
            if fam == 'chain':
                kname = 'resblock_chain_kernel'
            else:
                kname = 'conv3x3_ws2_kernel' if (r == 3 and cin // g <= 192) else 'conv_tc_kernel'
            traffic, traffic_src = None, None
            for capname in ('r2_dominant_kernel_ncu.json', 'r1_dominant_kernel_ncu.json'):
                try:                              # dram bytes per launch of the same kernel and shape from the committed ncu capture
                    cap = json.load(open(os.path.join(ROOT, 'profiles', capname)))
                    if cap.get('shape') == [n_img, h, w, cin, cout, r, g] and cap.get('kernel', kname) == kname:
                        traffic = cap['dram_bytes_read'] + cap['dram_bytes_write']
                        traffic_src = 'profiles/' + capname
                        break
                except Exception:
                    pass
            roofline = {'kernel': kname, 'shape': {'n_img': n_img, 'H': h, 'W': w, 'Cin': cin, 'Cout': cout, 'R': r, 'groups': g,
                                                   'dependent_conv_layers_per_launch': dom[4]},
                        'bound': 'tensor', 'achieved': ach, 'peak': peak, 'unit': 'TFLOP/s', 'frac': ach / peak,
                        'traffic': traffic, 'traffic_source': traffic_src,
                        'algorithmic_bytes_per_launch': dom[3] / dom[0],
                        'launches': dom[0], 'us_per_launch': 1e3 * dom[1] / dom[0], 'kernel_ms_per_step': dom[1],
                        'share_of_step': dom[1] / ms_step, 'algorithmic_tflop_per_launch': dom[2] / dom[0] / 1e12,
                        'all_tcgen05_launches': {'launches': len(recs), 'ms_per_step': all_ms, 'achieved': all_fl / (all_ms * 1e-3) / 1e12,
                                                 'share_of_step': all_ms / ms_step},
                        'timing': 'CUDA events around every tcgen05 launch of one eager forward (graphs off) on the launching stream',
                        'peak_source': src}
    # ---- the reference's own arithmetic on this GPU + parity of the benched output (rank 0, N = 1 only)
    gpu_reference, parity = None, {'finite': finite}
    if want_ref:
        model._graphs.clear()                     # release the graph's private pool before the 88 GB reference forward
        del y
        torch.cuda.empty_cache()
        try:
            r = run_gpu_reference(kw, sd0, x_host, y_first_host)
            gpu_reference, parity = r['gpu_reference'], r['parity']
            gpu_reference['speedup'] = value / gpu_reference['fps']
            gpu_reference['speedup_e2e'] = e2e['value'] / gpu_reference['fps']
        except Exception as ex:                   # e.g. out of memory on a smaller part: report, do not fail the bench
            gpu_reference = {'unavailable': repr(ex)[:300]}
    # ---- CPU baseline beside it (rank 0, N = 1 only)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        fps, sec = run_cpu_oracle(kw, 2, 1, threads)
        cpu_baseline = {'value': fps, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'seconds_per_sample': sec,
                        'sample': f'oracle (CPU port of the reference forward) on a {CPU_SAMPLE[1]}x{CPU_SAMPLE[2]}x{CPU_SAMPLE[3]} crop, same model, fp32, 2 warm calls'}
    # ---- BASELINE configs[2] next to the headline: the 72-unit Tester route, units sharded over the ranks (strong scaling)
    tester_route = None
    if args.workload == DEFAULT_WORKLOAD and not args.no_tester_route:
        del model
        torch.cuda.empty_cache()
        tester_route = run_tester_route(TESTER_ROUTE, 1, 1, rank, world, args.precision)
    if rank == 0:
        print(json.dumps({'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                          'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                          'dtype': args.precision, 'data': 'synthetic', 'config': config, 'clocks': clocks, 'e2e': e2e,
                          'gpu_launches': launches, 'peak_mem_GB': peak_mem, 'roofline': roofline, 'cpu_baseline': cpu_baseline,
                          'gpu_reference': gpu_reference, 'parity': parity, 'tester_route': tester_route}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
