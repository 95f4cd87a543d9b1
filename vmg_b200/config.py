"""YAML config -> VMG constructor kwargs.

Mirrors the reference factory `models/__init__.py:6-49` (create_model) and the two fields of
`utils/config.py:25-79` (parse_config) that the model path reads (`is_train`, `scale`).

One deliberate difference: the shipped YAMLs omit keys the reference
factory indexes (`if_print` everywhere; `ltam`, `deform_groups`, `max_res_scale`, `if_local_fuse`,
`channel_mixer` in VMG-REDS.yml), so the reference raises KeyError.  Here a missing YAML key means
"use the constructor default of models/vmg.py:177-210".
"""
from collections import OrderedDict

import yaml

# YAML `network:` key -> VMG kwarg (models/__init__.py:11-43)
_KEYMAP = OrderedDict([
    ('embed_dim', 'embed_dim'), ('depths', 'depths'), ('mlp_ratio', 'mlp_ratio'),
    ('n_groups', 'n_groups'), ('num_heads', 'num_heads'), ('window_sizes', 'window_sizes'),
    ('num_frames', 'num_frames'), ('back_RBs', 'back_RBs'), ('spynet', 'spynet_pretrained'),
    ('if_print', 'if_print'), ('ltam', 'ltam'), ('traj_win', 'traj_win'),
    ('traj_keyframes_n', 'traj_keyframes_n'), ('traj_heads', 'traj_heads'),
    ('temporal_type', 'temporal_type'), ('temporal_empty', 'temporal_empty'),
    ('traj_res_n', 'traj_res_n'), ('deform_groups', 'deform_groups'),
    ('max_res_scale', 'max_residual_scale'), ('spatial_type', 'spatial_type'),
    ('use_mdsc', 'mdsc'), ('if_concat', 'if_concat'), ('flow_smooth', 'flow_smooth'),
    ('smooth_region_range', 'smooth_region_range'), ('ret_decay', 'retention_decay'),
    ('non_linear', 'non_linear'), ('gating', 'gating'), ('if_symm', 'symm'),
    ('symm_act', 'symm_act'), ('relu_scale', 'relu_scale'), ('relu_scale_norm', 'relu_scale_norm'),
    ('ffn_type', 'ffn_type'), ('mixer_type', 'mixer_type'), ('mixer_n', 'mixer_n'),
    ('r_scaling', 'r_scaling'), ('chunk_ratios', 'chunk_ratios'), ('traj_mode', 'traj_mode'),
    ('twins', 'twins'), ('traj_scale', 'traj_scale'), ('traj_refine', 'traj_refine'),
    ('m_scaling', 'm_scaling'), ('if_local_fuse', 'if_local_fuse'),
    ('channel_mixer', 'channel_mixer'),
])


def parse_config(config_path, is_train=True):
    """Subset of utils/config.py:25-79 that the model path needs (no path expansion)."""
    with open(config_path, 'r') as f:
        config = yaml.safe_load(f)
    config['is_train'] = is_train
    config['network']['scale'] = config['scale']
    return config


def model_kwargs(config):
    """The kwargs create_model passes to VMG (models/__init__.py:9-43); missing key -> default."""
    if config.get('model') != 'VMG':
        raise NotImplementedError('Model [{}] not recognized.'.format(config.get('model')))
    net = config['network']
    shape_r = config['dataset']['image_shape_r']
    scale = config['scale']
    kw = {'image_size': [int(shape_r[1] / scale), int(shape_r[2] / scale)],
          'is_train': config.get('is_train', False)}
    for ykey, kwarg in _KEYMAP.items():
        if ykey in net:
            kw[kwarg] = net[ykey]
    return kw


def create_model(config, **engine_kw):
    """Drop-in for models/__init__.py:create_model; returns vmg_b200.VMG."""
    from .model import VMG
    return VMG(**model_kwargs(config), **engine_kw)
