"""Shared helpers for the test-suite: model construction from the repo's configs, seeded inputs."""
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
CONFIGS = {'few': 'VMG-REDS-few_levels', 'full': 'VMG-REDS'}


def kwargs_of(family):
    from vmg_b200 import config as cfg
    c = cfg.parse_config(os.path.join(ROOT, 'configs', CONFIGS[family] + '.yml'), is_train=False)
    return cfg.model_kwargs(c)


def build(family, seed=0, precision='fp32', T=None):
    """vmg_b200.VMG with deterministic CPU-RNG random init (identical on every box of this image)."""
    from vmg_b200.model import VMG
    torch.manual_seed(seed)
    m = VMG(**kwargs_of(family), precision=precision).eval()
    if T is not None:
        m.num_out_frames = T
    return m


def seeded_input(shape, seed=1):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(*shape, generator=g)


def sd_checksum(sd):
    return float(sum(v.double().sum().item() for v in sd.values()))


def psnr(a, b):
    mse = torch.mean((a.double().clamp(0, 1) - b.double().clamp(0, 1)) ** 2).item()
    return 99.0 if mse == 0 else 10.0 * torch.log10(torch.tensor(1.0 / mse)).item()
