"""Pins the oracle (oracle/vmg_oracle.py): against golden vectors minted from the UNMODIFIED reference
(always) and against the reference itself when /root/reference is present (build container)."""
import glob
import os

import pytest
import torch

from oracle import ref_loader, vmg_oracle
from tests import common

GOLDENS = sorted(glob.glob(os.path.join(common.GOLDEN, '*.pt')))


@pytest.mark.parametrize('path', GOLDENS, ids=[os.path.basename(p) for p in GOLDENS])
def test_oracle_matches_reference_goldens(path):
    g = torch.load(path)
    m = common.build(g['family'], seed=g['seed_w'])
    sd = m.state_dict()
    assert common.sd_checksum(sd) == pytest.approx(g['sd_checksum'], rel=0, abs=1e-6), 'random init is not reproducible'
    x = common.seeded_input(g['shape'], seed=g['seed_x'])
    assert float(x.double().sum()) == pytest.approx(g['x_checksum'], abs=1e-9)
    y1, sd1 = vmg_oracle.run(sd, common.kwargs_of(g['family']), x)
    assert (y1[..., ::4, ::4] - g['y1_sub']).abs().max().item() <= 2e-6
    assert y1.mean().item() == pytest.approx(g['y1_stats'][0], abs=1e-6)
    # second forward on the mutated weights (in-place Morph-FC decay, function.py:765-781)
    y2, sd2 = vmg_oracle.run(sd1, common.kwargs_of(g['family']), x)
    assert (y2[..., ::4, ::4] - g['y2_sub']).abs().max().item() <= 2e-6
    assert common.sd_checksum(sd2) == pytest.approx(g['sd_checksum_after2'], abs=1e-4)
    assert (y2 - y1).abs().max().item() > 0, 'forward must be stateful'


@pytest.mark.skipif(not ref_loader.available(), reason='/root/reference not present (GPU box)')
@pytest.mark.parametrize('family,shape', [('few', (1, 5, 3, 64, 72)), ('full', (1, 4, 3, 64, 64))])
def test_oracle_matches_live_reference(family, shape):
    Ref = ref_loader.load_reference_vmg_class()
    kw = common.kwargs_of(family)
    torch.manual_seed(3)
    ref = Ref(**kw).eval()
    ref.num_out_frames = shape[1]
    sd0 = {k: v.clone() for k, v in ref.state_dict().items()}
    x = common.seeded_input(shape, seed=5)
    if shape[1] % 2 == 0:       # palindromic clip -> frames_mirror path (vmg.py:426-432, 461-462)
        x = torch.cat([x[:, :shape[1] // 2], x[:, :shape[1] // 2].flip(1)], 1)
    with torch.no_grad():
        y_ref = ref(x)
    y, sd1 = vmg_oracle.run(sd0, kw, x)
    assert (y - y_ref).abs().max().item() <= 2e-6
    after = ref.state_dict()
    assert max((sd1[k] - after[k]).abs().max().item() for k in after) == 0.0


def test_oracle_rejects_wrong_T_and_small_frames():
    m = common.build('few')
    hp = vmg_oracle.Hyper(**common.kwargs_of('few'))
    with pytest.raises(RuntimeError):
        vmg_oracle.forward(m.state_dict(), hp, torch.rand(1, hp.num_out_frames + 1, 3, 64, 64))
    hp.num_out_frames = 2
    with pytest.raises(AssertionError):
        vmg_oracle.forward(m.state_dict(), hp, torch.rand(1, 2, 3, 32, 64))
