"""TEST INFRASTRUCTURE (container only) -- glue shims so the UNMODIFIED reference imports.

The reference (/root/reference, read-only, not present on the GPU box) imports `timm`, `mmcv`
and `thop`, none of which is installed.  On the model-forward hot path they are used only as glue
(SURVEY.md section 8c): DropPath (identity in eval), trunc_normal_, ConvModule (= conv + optional
ReLU, sub-module name `conv`), load_checkpoint (network download -> no-op offline), loggers.
This file fabricates those modules in `sys.modules`.  Nothing here is product code and nothing
here restates reference arithmetic.
"""
import logging
import sys
import types

import torch
import torch.nn as nn


def _module(name):
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


class _DropPath(nn.Module):
    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        raise RuntimeError("shim DropPath: training-mode stochastic depth is out of scope")


class _ConvModule(nn.Module):
    """mmcv.cnn.ConvModule restricted to what vmg.py:131-170 uses (norm_cfg=None)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0,
                 norm_cfg=None, act_cfg=dict(type='ReLU'), **kw):
        super().__init__()
        assert norm_cfg is None
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding)
        nn.init.kaiming_normal_(self.conv.weight, a=0, mode='fan_out', nonlinearity='relu')
        nn.init.constant_(self.conv.bias, 0)
        self.with_activation = act_cfg is not None
        if self.with_activation:
            assert act_cfg['type'] == 'ReLU'
            self.activate = nn.ReLU(inplace=True)

    def forward(self, x):
        x = self.conv(x)
        if self.with_activation:
            x = self.activate(x)
        return x


def install():
    if 'timm' in sys.modules and getattr(sys.modules['timm'], '_vmgb_shim', False):
        return
    timm = _module('timm'); timm._vmgb_shim = True
    tm = _module('timm.models'); timm.models = tm
    tl = _module('timm.models.layers'); tm.layers = tl
    tl.DropPath = _DropPath
    tl.trunc_normal_ = lambda t, mean=0., std=1., a=-2., b=2.: nn.init.trunc_normal_(t, mean, std, a, b)
    tl.to_2tuple = lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    mmcv = _module('mmcv'); mmcv.__version__ = '1.3.9'
    cnn = _module('mmcv.cnn'); mmcv.cnn = cnn
    cnn.ConvModule = _ConvModule
    cnn.constant_init = lambda m, val, bias=0: None
    cnn.kaiming_init = lambda m, **kw: None
    runner = _module('mmcv.runner'); mmcv.runner = runner
    runner.load_checkpoint = lambda model, filename, **kw: None   # offline: random-init SPyNet
    utils = _module('mmcv.utils'); mmcv.utils = utils
    utils.get_logger = lambda name, log_file=None, log_level=logging.INFO: logging.getLogger(name)

    thop = _module('thop')
    thop.profile = lambda *a, **k: (0, 0)
