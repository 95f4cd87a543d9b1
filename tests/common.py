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


def build(family, seed=0, precision='fp32', T=None, coord='cpu'):
    """vmg_b200.VMG with deterministic CPU-RNG random init (identical on every box of this image).
    coord: rounding of the flow_warp coordinate normalisation -- 'cpu' for comparisons with the CPU oracle / the
    CPU-minted goldens, 'cuda' (the model default) for comparisons with the reference arithmetic run on a GPU."""
    from vmg_b200.model import VMG
    torch.manual_seed(seed)
    m = VMG(**kwargs_of(family), precision=precision).eval()
    m.coord_semantics = coord
    if T is not None:
        m.num_out_frames = T
    return m


def oracle_on_cuda(sd, kw, x, allow_tf32=False):
    """The oracle restatement (same ATen / cuDNN calls as the reference forward) with every tensor on cuda:0:
    SURVEY.md 8c oracle O3 (allow_tf32=False: strict fp32, the large-shape checker) or O2 (PyTorch defaults: cuDNN
    TF32 on, the reference's own GPU arithmetic).  Returns (y, new_state_dict) on the device."""
    from oracle import vmg_oracle
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    old_default = torch.get_default_device()
    try:
        torch.set_default_device('cuda')          # the oracle builds its meshgrids on the default device
        sd_dev = {k: v.detach().clone().cuda() for k, v in sd.items()}
        hp = vmg_oracle.Hyper(**kw)
        hp.num_out_frames = x.shape[1]
        with torch.no_grad():
            y = vmg_oracle.forward(sd_dev, hp, x.cuda())
        return y, sd_dev
    finally:
        torch.set_default_device(old_default)
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


def seeded_input(shape, seed=1):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(*shape, generator=g)


def sd_checksum(sd):
    return float(sum(v.double().sum().item() for v in sd.values()))


def psnr(a, b):
    mse = torch.mean((a.double().clamp(0, 1) - b.double().clamp(0, 1)) ** 2).item()
    return 99.0 if mse == 0 else 10.0 * torch.log10(torch.tensor(1.0 / mse)).item()
