"""Host-side logic of the drop-in (vmg_b200/model.py) on CPU: the kernels are replaced by the plain
PyTorch op references of tests/emu_ops.py (test double, injected here only), so layouts, key-frame
bookkeeping, weight handling and the state_dict contract are checked against the oracle / goldens
without a GPU.  The CUDA kernels themselves are checked by the `-m gpu` tests."""
import os

import pytest
import torch

import vmg_b200.model as vm
from oracle import vmg_oracle
from tests import common, emu_ops


@pytest.fixture()
def emu(monkeypatch):
    monkeypatch.setattr(vm, '_ops', emu_ops)


@pytest.mark.parametrize('name', ['full_1x6x3x68x80.pt'])   # the few family is covered by the batched mirror test below
def test_forward_matches_reference_golden(emu, name):
    g = torch.load(os.path.join(common.GOLDEN, name))
    m = common.build(g['family'], seed=g['seed_w'], T=g['shape'][1])
    x = common.seeded_input(g['shape'], seed=g['seed_x'])
    y1 = m(x)
    assert y1.shape == (g['shape'][0], g['shape'][1], 3, 4 * g['shape'][3], 4 * g['shape'][4]) and y1.dtype == torch.float32
    assert (y1[..., ::4, ::4] - g['y1_sub']).abs().max().item() <= 5e-6
    y2 = m(x)       # second call: decayed weights persisted in the parameters
    assert (y2[..., ::4, ::4] - g['y2_sub']).abs().max().item() <= 5e-6
    assert common.sd_checksum(m.state_dict()) == pytest.approx(g['sd_checksum_after2'], abs=1e-4)


def test_batched_mirror_clip_matches_oracle(emu):
    m = common.build('few', seed=2, T=4)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    half = common.seeded_input((2, 2, 3, 66, 70), seed=9)
    x = torch.cat([half, half.flip(1)], 1)                 # palindrome -> frames_mirror
    y = m(x)
    assert m.frames_mirror is True
    y_ref, _ = vmg_oracle.run(sd0, common.kwargs_of('few'), x)
    assert (y - y_ref).abs().max().item() <= 5e-6


@pytest.mark.parametrize('family,n_keys,n_params', [('few', 560, 26059711), ('full', 724, 37546975)])
def test_state_dict_schema(family, n_keys, n_params):
    m = common.build(family)
    sd = m.state_dict()
    assert len(sd) == n_keys
    assert sum(p.numel() for p in m.parameters()) == n_params
    for k in ('spynet.mean', 'spynet.basic_module.5.basic_module.4.conv.weight', 'input_proj.proj.0.weight',
              'encoder_layers.0.mlp_blocks.0.spatial_mixing.gamma_h', 'encoder_layers.0.mlp_blocks.0.spatial_mixing.mlp_h.0.weight',
              'encoder_layers.0.traj_mixing.LTAM.relative_pos_encoding', 'encoder_layers.0.traj_mixing.LTAM.decay_v',
              'encoder_layers.0.traj_mixing.resblocks.main.2.14.conv2.bias', 'encoder_layers.0.traj_mixing.fusion.weight',
              'downsample.0.norm.weight', 'upsample.0.linear.bias', 'upconv1.weight', 'conv_last.bias'):
        assert k in sd, k
    if family == 'few':
        assert 'local_cnn.weight' in sd and 'encoder_layers.1.local_cnn.bias' in sd
        assert 'encoder_layers.0.mlp_blocks.0.spatial_mixing.mlp_c.body.3.conv_du.2.weight' in sd
    else:
        assert 'sc_64_16.1.weight' in sd and 'sc_32_8.0.bias' in sd
        assert sd['encoder_layers.2.mlp_blocks.0.spatial_mixing.gamma_h'].shape == (228, 228)
        assert sd['encoder_layers.0.mlp_blocks.0.channel_mixing.fc1.weight'].shape == (672, 28, 3, 3)
    assert len(m.mlp_wd_param) == sum(1 for n, _ in m.named_parameters() if '.mlp_blocks.' in n)
    m2 = common.build(family, seed=5)
    m2.load_state_dict({'module.' + k if False else k: v for k, v in sd.items()}, strict=True)


def test_constructor_contract():
    from vmg_b200.model import VMG
    kw = common.kwargs_of('few')
    assert VMG(**kw).num_out_frames == kw['num_frames']
    for bad in ({'temporal_empty': False}, {'ffn_type': 'vanilla'}, {'traj_mode': 'max'}, {'if_concat': True},
                {'back_RBs': 2}, {'spynet_pretrained': None}, {'mixer_type': ['mbconv', 'mlps']},
                {'temporal_type': [True, None]}, {'relu_scale_norm': True}):
        with pytest.raises(NotImplementedError):
            VMG(**{**kw, **bad})


def test_forward_contract_errors(emu):
    m = common.build('few', T=3)
    with pytest.raises(AssertionError):
        m(torch.rand(1, 3, 3, 32, 64))
    with pytest.raises(RuntimeError):
        m(torch.rand(1, 4, 3, 64, 64))
    m.train()
    with pytest.raises(NotImplementedError):
        m(torch.rand(1, 3, 3, 64, 64))


def test_config_factory_missing_keys_use_defaults(repo_root):
    from vmg_b200 import config as cfg
    c = cfg.parse_config(os.path.join(repo_root, 'configs', 'VMG-REDS.yml'), is_train=False)
    kw = cfg.model_kwargs(c)
    assert 'channel_mixer' not in kw and 'if_local_fuse' not in kw and kw['image_size'] == [64, 64]
    m = cfg.create_model(c)
    assert m.channel_mixer == 'vanilla' and m.local_fuse is False and m.mdsc is True
    assert m.chunk_h == [8, 16, 12, 8]
    with pytest.raises(NotImplementedError):
        cfg.model_kwargs({**c, 'model': 'other'})
