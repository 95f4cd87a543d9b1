"""VMG -- drop-in for the reference `models/vmg.py:VMG` on its forward hot path, running on libvmgb.

Same constructor (models/vmg.py:177-210), same `state_dict` keys/shapes, same
`forward(x[B,T,3,H,W]) -> [B,T,3,4H,4W]` including the reference's in-place Morph-FC weight decay
(models/function.py:765-781).  The nn.Modules below only HOLD parameters under the reference's names;
all arithmetic is issued from `VMG.forward` as vmgb_* kernels (vmg_b200/ops.py).  Unsupported
constructor values raise NotImplementedError at construction -- there is no fallback path.

Internal layout: channels-last, frames time-major: activations are [T*B, H, W, C].
"""
from fractions import Fraction
import math
import os

import torch
import torch.nn as nn

from . import ops as _ops


# ---------------------------------------------------------------------------------------------
# parameter holders (names mirror the reference module tree)
# ---------------------------------------------------------------------------------------------
class _ConvModule(nn.Module):          # mmcv.cnn.ConvModule as used at vmg.py:131-170 (key `.conv`)
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 7, 1, 3)
        nn.init.kaiming_normal_(self.conv.weight, a=0, mode='fan_out', nonlinearity='relu')
        nn.init.constant_(self.conv.bias, 0)


class _SPyNetBasicModule(nn.Module):   # vmg.py:126-173
    def __init__(self):
        super().__init__()
        ch = [8, 32, 64, 32, 16, 2]
        self.basic_module = nn.Sequential(*[_ConvModule(ch[i], ch[i + 1]) for i in range(5)])


class _SPyNet(nn.Module):              # vmg.py:18-37
    def __init__(self):
        super().__init__()
        self.basic_module = nn.ModuleList([_SPyNetBasicModule() for _ in range(6)])
        self.register_buffer('mean', torch.Tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1))
        self.register_buffer('std', torch.Tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1))


class _InputProj(nn.Module):           # layers.py:674-717
    def __init__(self, cin, dim):
        super().__init__()
        self.proj = nn.Sequential(nn.Conv2d(cin, dim, 3, 1, 1), nn.LeakyReLU(inplace=True))


class _CALayer(nn.Module):             # function.py:542-558
    def __init__(self, c, reduction):
        super().__init__()
        self.conv_du = nn.Sequential(nn.Conv2d(c, c // reduction, 1), nn.ReLU(inplace=True),
                                     nn.Conv2d(c // reduction, c, 1), nn.Sigmoid())


class _RCAB(nn.Module):                # function.py:561-583
    def __init__(self, c, reduction=8):
        super().__init__()
        self.body = nn.Sequential(nn.Conv2d(c, c, 3, 1, 1), nn.ReLU(True), nn.Conv2d(c, c, 3, 1, 1),
                                  _CALayer(c, reduction))


class _Mlp(nn.Module):                 # function.py:30-47
    def __init__(self, cin, hidden, cout):
        super().__init__()
        self.fc1 = nn.Linear(cin, hidden)
        self.fc2 = nn.Linear(hidden, cout)


def _form_decay(chunk, S):
    """gamma buffer of Enhanced_MorphFCs_decay (function.py:651-732): entry ((m,s),(i,s')) =
    mean_k d_k^(|m-i|+1), d_k = 1 - 2^(-5-(chunk-1-k)), powers built by repeated multiplication."""
    d = torch.log(1 - 2 ** (-5 - torch.arange(chunk - 1, -1, -1, dtype=torch.float))).exp()
    p = d.unsqueeze(-1).repeat(1, chunk)
    for i in range(1, chunk):
        p[:, i] = p[:, i] * p[:, i - 1]
    idx = (torch.arange(chunk)[:, None] - torch.arange(chunk)[None, :]).abs()
    G = p[:, idx]                                            # [k, m, i]
    G = G[:, :, None, :, None].expand(chunk, chunk, S, chunk, S).reshape(chunk, chunk * S, chunk * S)
    return torch.mean(G, 0)


class _MorphFC(nn.Module):             # Enhanced_MorphFCs_decay, function.py:596-741
    def __init__(self, dim, chunk_h, chunk_w, qkv_bias, channel_mixer):
        super().__init__()
        self.chunk_h, self.chunk_w = chunk_h, chunk_w
        self.Ch = int(math.ceil(dim / chunk_h)) * chunk_h
        self.Cw = int(math.ceil(dim / chunk_w)) * chunk_w
        self.mlp_h = nn.Sequential(nn.Linear(self.Ch, self.Ch, bias=qkv_bias), nn.ReLU())
        self.mlp_w = nn.Sequential(nn.Linear(self.Cw, self.Cw, bias=qkv_bias), nn.ReLU())
        if channel_mixer == 'vanilla':
            self.mlp_c = nn.Sequential(nn.Linear(dim, dim, bias=qkv_bias), nn.ReLU())
        else:
            self.mlp_c = _RCAB(dim)
        self.reweight = _Mlp(dim, dim // 4, dim * 3)
        self.proj = nn.Linear(dim, dim)
        self.register_buffer('gamma_h', _form_decay(chunk_h, self.Ch // chunk_h))
        self.register_buffer('gamma_w', _form_decay(chunk_w, self.Cw // chunk_w))


class _MlpCnn(nn.Module):              # function.py:50-79
    def __init__(self, dim, ratio, groups):
        super().__init__()
        self.hidden = int(dim * ratio)
        self.fc1 = nn.Conv2d(dim, self.hidden, 3, 1, 1, groups=groups)
        self.fc2 = nn.Linear(self.hidden, dim)


class _TAB(nn.Module):                 # function.py:1139-1217
    def __init__(self, dim, chunk_h, chunk_w, ratio, groups, qkv_bias, channel_mixer):
        super().__init__()
        self.norm2 = nn.LayerNorm(dim)
        self.spatial_mixing = _MorphFC(dim, chunk_h, chunk_w, qkv_bias, channel_mixer)
        self.norm3 = nn.LayerNorm(dim)
        self.channel_mixing = _MlpCnn(dim, ratio, groups)


class _LTAM(nn.Module):                # trajectory.py:493-531
    def __init__(self, dim, heads, twins):
        super().__init__()
        self.proj = nn.Linear(dim, dim)
        self.relative_pos_encoding = nn.Parameter(torch.zeros(heads, twins[0] * twins[1], twins[0] * twins[1]))
        nn.init.trunc_normal_(self.relative_pos_encoding, std=.02)
        self.register_buffer('decay_v', 1 - 2 ** (-5 - torch.arange(heads - 1, -1, -1, dtype=torch.float32)))


class _ResBlock(nn.Module):            # trajectory.py:165-221
    def __init__(self, c):
        super().__init__()
        self.conv1 = nn.Conv2d(c, c, 3, 1, 1)
        self.conv2 = nn.Conv2d(c, c, 3, 1, 1)


class _ResBlocks(nn.Module):           # trajectory.py:16-52
    def __init__(self, cin, c, n):
        super().__init__()
        self.main = nn.Sequential(nn.Conv2d(cin, c, 3, 1, 1), nn.LeakyReLU(0.1, inplace=True),
                                  nn.Sequential(*[_ResBlock(c) for _ in range(n)]))


class _Trajectory(nn.Module):          # trajectory.py:226-272
    def __init__(self, dim, n_blocks, heads, twins):
        super().__init__()
        self.LTAM = _LTAM(dim, heads, twins)
        self.resblocks = _ResBlocks(2 * dim, dim, n_blocks)
        self.fusion = nn.Conv2d(3 * dim, dim, 1, 1, 0)


class _Stage(nn.Module):               # Mlp_encoder, function.py:1267-1395
    def __init__(self, dim, depth, chunk_h, chunk_w, ratio, groups, qkv_bias, channel_mixer, local_fuse,
                 aligned, res_n, heads, twins, key_stride):
        super().__init__()
        self.dim, self.chunk_h, self.chunk_w = dim, chunk_h, chunk_w
        self.key_stride, self.heads = key_stride, heads
        if local_fuse:
            self.local_cnn = nn.Conv2d(dim, dim, 3, 1, 1)
        self.mlp_blocks = nn.ModuleList([_TAB(dim, chunk_h, chunk_w, ratio, groups, qkv_bias, channel_mixer)
                                         for _ in range(depth)])
        self.has_traj = aligned is False
        self.traj_mixing = _Trajectory(dim, res_n, heads, twins) if self.has_traj else nn.Identity()


class _Resample(nn.Module):            # UpdownkeepSampling, layers.py:761-775
    def __init__(self, cin, cout, mode):
        super().__init__()
        self.mode = mode
        f = cin * 4 if mode == 'down' else cin // 4
        self.norm = nn.LayerNorm(f)
        self.linear = nn.Linear(f, cout)


# ---------------------------------------------------------------------------------------------
class VMG(nn.Module):
    """See module docstring.  Extra keyword (not in the reference): `precision` = 'bf16' (tcgen05 path,
    bf16 activations, fp32 statistics/flows) or 'fp32' (strict fp32 CUDA-core path)."""

    def __init__(self, in_chans=3, embed_dim=[112, 224, 224, 448, 448, 224, 224, 112],
                 depths=[8, 8, 8, 8, 8, 8, 8, 8], num_heads=[2, 4, 8, 16, 16, 8, 4, 2], num_frames=7,
                 window_sizes=[(4, 4)] * 8, mdsc=True, if_concat=False, mlp_ratio=4., n_groups=1,
                 qkv_bias=True, qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.1,
                 norm_layer=nn.LayerNorm, patch_norm=True, back_RBs=0, spynet_pretrained=None,
                 image_size=[64, 112], is_train=True, if_print=False, ltam=True,
                 traj_win=[16, None, None, None], traj_keyframes_n=[3, None, None, None],
                 traj_heads=[4, None, None, None], temporal_type=[False, None, None, None],
                 temporal_empty=True, traj_res_n=[1, 0, 0, 0, 0, 0, 1], deform_groups=[8, 16, 16, 32],
                 max_residual_scale=[1, 2, 2, 4], spatial_type=[False, False, False, False],
                 flow_smooth=True, smooth_region_range=4, retention_decay=True, non_linear=True,
                 gating=True, symm=True, symm_act=nn.Tanh, relu_scale=True, relu_scale_norm=False,
                 ffn_type='vanilla', mixer_type=['mbconv', 'mbconv', 'mlps', 'mlps'], mixer_n=[2, 3, None, None],
                 r_scaling=1., chunk_ratios=[1 / 4, 1 / 4, 3 / 16, 1 / 8], traj_mode='wins', twins=[2, 2],
                 traj_scale=True, traj_refine=None, m_scaling=1., if_local_fuse=False, channel_mixer='vanilla',
                 precision='bf16'):
        super().__init__()
        self.num_layers = len(depths)
        self.num_enc_layers = self.num_layers // 2 + 1
        self.num_dec_layers = self.num_layers // 2
        self.scale = 2 ** (self.num_enc_layers - 1)
        self.embed_dim = embed_dim
        self.num_in_frames = num_frames
        self.num_out_frames = num_frames
        self.is_train = is_train
        self.init_H, self.init_W = image_size
        self.frames_mirror = False
        self.precision = precision
        if precision not in ('bf16', 'fp32'):
            raise ValueError("precision must be 'bf16' or 'fp32'")

        # ---- value set of the two shipped model families; anything else is refused ----
        def refuse(cond, what):
            if cond:
                raise NotImplementedError(f'vmg_b200.VMG: {what} is outside the supported hot path')
        n_e = self.num_enc_layers
        refuse(self.num_layers % 2 == 0 or self.num_layers < 3, 'an even number of stages')
        refuse(spynet_pretrained is None, 'spynet=None (the trajectory module indexes the flows unconditionally)')
        refuse(any(m != 'mlps' for m in mixer_type[:n_e]), "mixer_type != 'mlps'")
        refuse(any(t for t in temporal_type[:n_e]), 'temporal_type=True (DCN alignment)')
        refuse(not temporal_empty, 'temporal_empty=False (3D window attention)')
        refuse(relu_scale_norm or not relu_scale, 'relu_scale_norm=True / relu_scale=False')
        refuse(ffn_type != 'ffn_cnn', "ffn_type != 'ffn_cnn'")
        refuse(traj_mode != 'wins' or list(twins) != [2, 2], "traj_mode != 'wins' or twins != [2,2]")
        refuse(if_concat, 'if_concat=True')
        refuse(back_RBs > 0, 'back_RBs > 0')
        refuse(not (retention_decay and non_linear and gating and symm), 'retention_decay/non_linear/gating/symm=False')
        refuse(symm_act not in ('tanh', nn.Tanh), "symm_act != 'tanh'")
        refuse(not ltam or not traj_scale or traj_refine is not None, 'ltam=False / traj_scale=False / traj_refine')
        refuse(channel_mixer not in ('vanilla', 'rcab'), f'channel_mixer={channel_mixer!r}')
        refuse(temporal_type[0] is not False, 'a first stage without trajectory alignment')
        # shapes the kernels take (csrc/ltam.cu: 4 heads, C/heads/4 in {4,7,8,9}; multi-source convs: C % 16 == 0)
        refuse(traj_heads[0] != 4, 'traj_heads[0] != 4')
        refuse(embed_dim[0] % 16 != 0 or embed_dim[0] // 16 not in (4, 7, 8, 9), 'embed_dim[0] outside {64, 112, 128, 144}')
        refuse(embed_dim[-1] != embed_dim[0], 'embed_dim[-1] != embed_dim[0]')
        refuse(any(d % 8 for d in embed_dim), 'embed_dim not a multiple of 8')
        self.flow_smooth, self.region = flow_smooth, smooth_region_range
        self.r_scaling, self.m_scaling = float(r_scaling), float(m_scaling)
        self.n_groups, self.twins = n_groups, list(twins)
        self.channel_mixer = channel_mixer

        self.chunk_ratio = [float(Fraction(r)) for r in chunk_ratios]
        self.chunk_h = [int(self.init_H * r) for r in self.chunk_ratio]
        self.chunk_w = [int(self.init_W * r) for r in self.chunk_ratio]

        self.spynet = _SPyNet()
        # `spynet_pretrained` is a URL/path in the reference (mmcv load_checkpoint); there is no network
        # here -- SPyNet weights arrive through load_state_dict like every other parameter.
        self.local_fuse = if_local_fuse
        if if_local_fuse:
            self.local_cnn = nn.Conv2d(embed_dim[0], embed_dim[0], 3, 1, 1)
        self.input_proj = _InputProj(in_chans, embed_dim[0])

        def make_stage(li, j):
            return _Stage(embed_dim[li], depths[li], self.chunk_h[j], self.chunk_w[j], mlp_ratio, n_groups, qkv_bias,
                          channel_mixer, if_local_fuse, temporal_type[j], traj_res_n[li], traj_heads[j], self.twins,
                          traj_keyframes_n[j])
        self.encoder_layers = nn.ModuleList()
        self.upsample = nn.ModuleList()
        self.downsample = nn.ModuleList()
        for i in range(n_e):
            self.encoder_layers.append(make_stage(i, i))
            if i != n_e - 1:
                self.downsample.append(_Resample(embed_dim[i], embed_dim[i + 1], 'down'))
            else:
                self.upsample.append(_Resample(embed_dim[i], embed_dim[i + 1], 'up'))
        self.decoder_layers = nn.ModuleList()
        for i in range(self.num_dec_layers):
            self.decoder_layers.append(make_stage(i + n_e, -i - 2))
            if i != self.num_dec_layers - 1:
                self.upsample.append(_Resample(embed_dim[i + n_e], embed_dim[i + n_e + 1], 'up'))

        self.upconv1 = nn.Conv2d(embed_dim[-1], embed_dim[-1] * 4, 3, 1, 1)
        self.upconv2 = nn.Conv2d(embed_dim[-1], 64 * 4, 3, 1, 1)
        self.HRconv = nn.Conv2d(64, 64, 3, 1, 1)
        self.conv_last = nn.Conv2d(64, 3, 3, 1, 1)
        self.mdsc = bool(mdsc) and self.num_layers > 3
        refuse(bool(mdsc) and self.num_layers == 5, 'mdsc with 5 stages')
        if self.mdsc:
            self.sc_64_16 = nn.Sequential(nn.Conv2d(embed_dim[0], embed_dim[2], 1), nn.GroupNorm(1, embed_dim[2]), nn.ReLU())
            self.sc_32_8 = nn.Sequential(nn.Conv2d(embed_dim[1], embed_dim[3], 1), nn.GroupNorm(1, embed_dim[3]), nn.ReLU())

        self.mlp_wd_param = [p for name, p in self.named_parameters() if '.mlp_blocks.' in name]
        self.apply(self._init_weights)
        self._packed = {}
        self._graphs, self._graph_stamp = {}, None
        self.use_cuda_graph = os.environ.get('VMGB_NO_GRAPH', '0') != '1'
        self.recon_chunk = 8     # frames per reconstruction-head chunk (bounds the 4x-resolution temporaries)
        # rounding of the flow_warp coordinate normalisation (include/vmgb.h VMGB_COORD_*): 'cuda' = what the reference's
        # own forward computes on a GPU (ATen CUDA divides by a Python scalar as a * (1/b)), 'cpu' = ATen CPU (true division)
        self.coord_semantics = 'cuda'

    @staticmethod
    def _init_weights(m):               # vmg.py:415-424
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    # -------------------------------------------------------------------------------------------
    # weights
    # -------------------------------------------------------------------------------------------
    @property
    def _adt(self):
        return torch.bfloat16 if self.precision == 'bf16' else torch.float32

    @property
    def _sem(self):
        if self.coord_semantics not in ('cuda', 'cpu'):
            raise ValueError("coord_semantics must be 'cuda' or 'cpu'")
        return _ops.COORD_ATEN_CUDA if self.coord_semantics == 'cuda' else _ops.COORD_ATEN_CPU

    def _pc(self, mod, groups=1, tc=None, force=False):
        """Packed kernel-side weights of a Conv2d/Linear holder, re-packed when the parameter changed."""
        w = mod.weight
        tc = (self.precision == 'bf16') if tc is None else tc
        key = id(mod)
        ver = (w._version, w.data_ptr(), mod.bias._version if mod.bias is not None else -1, tc)
        ent = self._packed.get(key)
        if force or ent is None or ent[0] != ver:
            ent = (ver, _ops.pack_conv(w, mod.bias, groups, tc=tc))
            self._packed[key] = ent
        return ent[1]

    def _pc_nobias(self, mod):
        """Packed weights of a Linear WITHOUT its bias (LTAM value projection, see _trajectory)."""
        w = mod.weight
        key = ('nobias', id(mod))
        ver = (w._version, w.data_ptr(), self.precision)
        ent = self._packed.get(key)
        if ent is None or ent[0] != ver:
            ent = (ver, _ops.pack_conv(w, None, 1, tc=self.precision == 'bf16'))
            self._packed[key] = ent
        return ent[1]

    # -------------------------------------------------------------------------------------------
    # SPyNet (vmg.py:39-123, 435-464) -- fp32 end to end; only the level-0 flows are consumed
    # -------------------------------------------------------------------------------------------
    def _compute_flow(self, xp, B, T, mirror):
        o = _ops
        N, Hp, Wp, _ = xp.shape
        hu = Hp if Hp % 32 == 0 else 32 * (Hp // 32 + 1)
        wu = Wp if Wp % 32 == 0 else 32 * (Wp // 32 + 1)
        if not hasattr(self, '_meanstd') or self._meanstd[0] != (self.spynet.mean._version, self.spynet.std._version):
            self._meanstd = ((self.spynet.mean._version, self.spynet.std._version),
                             self.spynet.mean.flatten().tolist(), self.spynet.std.flatten().tolist())
        pyr = [o.spynet_prep(xp, self._meanstd[1], self._meanstd[2], hu, wu)]
        for _ in range(5):
            pyr.append(o.avgpool2(pyr[-1]))
        pyr = pyr[::-1]
        n_pairs = (T - 1) * B
        flows = []
        tc = self.precision == 'bf16'     # bf16 operands on the tcgen05 path; flows, warps and pyramids stay fp32
        for direction in ((0,) if mirror else (0, 1)):
            flow = None
            for level in range(6):
                y, fup = o.spynet_level_input(pyr[level], flow, n_pairs, B, direction, out_dtype=self._adt, semantics=self._sem)
                convs = self.spynet.basic_module[level].basic_module
                for m in range(4):
                    y = o.conv2d(y, self._pc(convs[m].conv, tc=tc), act=o.ACT_RELU, use_tc=tc)
                flow = o.conv2d(y, self._pc(convs[4].conv, tc=tc), epi=o.EPI_RESADD, res=fup, use_tc=tc,
                                out_dtype=torch.float32)
            f = o.flow_resize(flow, Hp, Wp)
            if self.flow_smooth:
                f = o.flow_smooth(f, self.region)
            flows.append(f.view(T - 1, B, Hp, Wp, 2))
        if mirror:                      # vmg.py:461-462 (smoothing commutes with the temporal flip)
            flows.append(flows[0].flip(0).contiguous())
        return flows[0], flows[1]

    # -------------------------------------------------------------------------------------------
    # TAB (function.py:1212-1217): Morph-FC mixing (function.py:743-805) + conv FFN (function.py:67-79)
    # -------------------------------------------------------------------------------------------
    def _tab(self, blk, st, x, B, T):
        o = _ops
        N, H, W, C = x.shape
        sm = blk.spatial_mixing
        # in-place decay of the two token-mixer weights, once per forward (function.py:765-768 / 778-781), then their packed copies
        for lin, gamma in ((sm.mlp_h[0], sm.gamma_h), (sm.mlp_w[0], sm.gamma_w)):
            o.decay_weights_(lin.weight.data, gamma)
        pc_h, pc_w = self._pc(sm.mlp_h[0], force=True), self._pc(sm.mlp_w[0], force=True)
        if self.channel_mixer == 'vanilla' and o.tab_mix_supported(x, C, sm.chunk_h, sm.chunk_w):
            # fused mixing half (csrc/tab_mix.cu): LayerNorm, both Morph-FC token mixers, the channel mixer, the global
            # mean, the 3-way mix, the gated projection and the residual in two passes over x
            pcs = (pc_h, pc_w, self._pc(sm.mlp_c[0]), self._pc(sm.proj))
            acc = o.tab_mix(x, blk.norm2.weight, blk.norm2.bias, blk.norm2.eps, pcs, 0, B)
            if acc is not None:                                          # None: the library does not take this shape on this device
                a = o.fixsum_finalize(acc, 1.0 / (T * H * W))            # function.py:791
                a = o.linear(a, self._pc(sm.reweight.fc1, tc=False), act=o.ACT_GELU, use_tc=False)
                a = o.softmax3(o.linear(a, self._pc(sm.reweight.fc2, tc=False), use_tc=False))
                y = o.tab_mix(x, blk.norm2.weight, blk.norm2.bias, blk.norm2.eps, pcs, 1, B, a=a, m_scaling=self.m_scaling)
                if y is not None:
                    return self._tab_ffn(blk, y)
        xn = o.layernorm(x, blk.norm2.weight, blk.norm2.bias)
        branches = []
        for axis, (pcl, chunk, Cp) in enumerate(((pc_h, sm.chunk_h, sm.Ch), (pc_w, sm.chunk_w, sm.Cw))):
            A = o.morph_gather(xn, chunk, axis)
            Y = torch.empty_strided(A.shape, A.stride(), device=A.device, dtype=A.dtype)   # same (16-byte padded) row pitch
            o.linear(A, pcl, act=o.ACT_RELU, out_scale=1.0 / Cp, out=Y)
            branches.append(o.morph_scatter(Y, N, H, W, C, chunk, axis))
            del A, Y
        h, w = branches
        if self.channel_mixer == 'rcab':                                 # function.py:542-583
            body = sm.mlp_c.body
            r = o.conv2d(xn, self._pc(body[0]), act=o.ACT_RELU)
            r = o.conv2d(r, self._pc(body[2]))
            y = o.channel_sum(r, None, None, 1, N, H * W, 1.0 / (H * W))
            du = body[3].conv_du
            y = o.linear(y, self._pc(du[0], tc=False), act=o.ACT_RELU, use_tc=False)
            y = o.linear(y, self._pc(du[2], tc=False), act=o.ACT_SIGMOID, use_tc=False)
            c = o.ca_apply(r, xn, y, H * W, 1.0 / C)
            del r
        else:
            c = o.linear(xn, self._pc(sm.mlp_c[0]), act=o.ACT_RELU, out_scale=1.0 / C)
        a = o.channel_sum(h, w, c, T, B, H * W, 1.0 / (T * H * W))       # function.py:791
        a = o.linear(a, self._pc(sm.reweight.fc1, tc=False), act=o.ACT_GELU, use_tc=False)
        a = o.softmax3(o.linear(a, self._pc(sm.reweight.fc2, tc=False), use_tc=False))
        mix = o.mix3(h, w, c, a, T, B, H * W)
        del h, w, c, branches
        x = o.linear(mix, self._pc(sm.proj), epi=o.EPI_GATE, res=x, aux=xn, res_scale=self.m_scaling)
        del mix, xn
        return self._tab_ffn(blk, x)

    def _tab_ffn(self, blk, x):
        o = _ops
        xn = o.layernorm(x, blk.norm3.weight, blk.norm3.bias)
        cm = blk.channel_mixing
        u = o.conv2d(xn, self._pc(cm.fc1, groups=self.n_groups), act=o.ACT_GELU)
        x = o.linear(u, self._pc(cm.fc2), epi=o.EPI_RESADD, res=x, res_scale=self.m_scaling)
        return x

    # -------------------------------------------------------------------------------------------
    # Trajectory_multi_head (trajectory.py:300-490) + LTAM wins (trajectory.py:672-795)
    # -------------------------------------------------------------------------------------------
    def _pchain(self, blocks):
        """Packed weights of the whole residual-block chain (ops.PackedChain), re-packed when any conv changed."""
        layers = [self._pc(c) for blk in blocks for c in (blk.conv1, blk.conv2)]
        key = ('chain', id(blocks))
        ver = tuple(id(q) for q in layers)
        ent = self._packed.get(key)
        if ent is None or ent[0] != ver:
            ent = (ver, _ops.pack_resblock_chain(layers))
            self._packed[key] = ent
        return ent[1]

    def _resblocks(self, rb, srcs, out, flags=None, step=0, scratch=None):
        o = _ops
        main = rb.main
        y = o.conv2d(srcs, self._pc(main[0]), act=o.ACT_LRELU, alpha=0.1)
        blocks = main[2]
        if len(blocks) == 0:
            out.copy_(y)
            return out
        fl = flags if flags is not None else o.chain_flags(y)
        f0, f1 = (fl[step & 1], fl[(step + 1) & 1]) if fl is not None else (None, None)
        return o.resblock_chain(y, self._pchain(blocks), self.r_scaling, out, f0, f1, scratch)

    def _trajectory(self, st, x, B, T, flows_f, flows_b):
        """Both sweeps of Trajectory_multi_head share their weights and are independent until the final
        fusion (trajectory.py:476-478 only reads outputs_back[i]), so they advance in lockstep as ONE image
        batch: step j = [backward sweep at frame T-1-j | forward sweep at frame j].  Key frames fall on the
        same steps (j % stride == 0) in both sweeps, so key-frame k of either sweep is step k*stride."""
        o = _ops
        tm = st.traj_mixing
        N, H, W, C = x.shape
        n = B
        # paired, step-major copies (pure data movement): [T, 2n, H, W, C]
        xp = o.pair_sweeps(x, T)
        qp = o.pair_sweeps(o.l2norm(x), T)              # queries; a key frame's keys are its own normalised input
        # step j >= 1: backward sweep uses flows_backward[:, T-1-j], forward sweep flows_forward[:, j-1]
        fl = torch.stack([flows_b.flip(0), flows_f], 1).reshape(T - 1, 2 * n, H, W, 2) if T > 1 else None
        stride = st.key_stride
        kcap = (T + stride - 1) // stride
        step_elems = 2 * n * H * W * C
        # LTAM's output projection is applied to the VALUES when a key frame is stored (proj is linear and the softmax weights
        # sum to 1: proj(sum_j p_j v_j) + b = sum_j p_j proj(v_j) + b): T/stride launches instead of T-1, and the attention
        # kernel adds the bias and the warped anchor itself (trajectory.py:791-795)
        # (bf16 path only: the strict-fp32 path keeps the reference's operation order)
        pre_proj = self.precision == 'bf16'
        proj = self._pc_nobias(tm.LTAM.proj) if pre_proj else self._pc(tm.LTAM.proj)
        pbias = tm.LTAM.proj.bias.detach().float().contiguous() if tm.LTAM.proj.bias is not None else None
        rel = tm.LTAM.relative_pos_encoding.detach().contiguous()
        decay = tm.LTAM.decay_v
        fbuf = torch.empty((T, 2 * n, H, W, C), device=x.device, dtype=x.dtype)
        vproj = torch.empty_like(fbuf) if pre_proj else None     # projected values; only the key-frame steps are written (same stride as the keys)
        loc = [torch.empty((2 * n, H, W, 2 * kcap), device=x.device, dtype=torch.float32) for _ in range(2)]
        o.loc_identity_(loc[0], 0)
        cur_loc, K = 0, 0
        prev = None
        flags = o.chain_flags(fbuf[0])                  # per-pair dependency counters of the persistent resblock-chain kernel
        scratch = (torch.empty_like(fbuf[0]), torch.empty_like(fbuf[0]))
        for j in range(T):
            is_key = j % stride == 0
            if j == 0:
                feat_in = torch.zeros((2 * n, H, W, C), device=x.device, dtype=x.dtype)
            else:
                flow = fl[j - 1]
                warped = o.flow_warp_fused(prev, flow, loc[cur_loc], 2 * K, loc[1 - cur_loc], semantics=self._sem)
                cur_loc = 1 - cur_loc
                if pre_proj:
                    feat_in = o.ltam_wins(qp[j], qp[0], vproj[0], stride * step_elems, K, loc[cur_loc], rel, decay, st.heads,
                                          out_bias=pbias, anchor=warped)
                else:
                    att = o.ltam_wins(qp[j], qp[0], fbuf[0], stride * step_elems, K, loc[cur_loc], rel, decay, st.heads)
                    feat_in = o.linear(att, proj, epi=o.EPI_RESADD, res=warped)
                if is_key:
                    o.loc_identity_(loc[cur_loc], 2 * K)
            if is_key:
                K += 1                                  # step j becomes key frame K-1 once fbuf[j] is written
            prev = self._resblocks(tm.resblocks, (xp[j], feat_in), fbuf[j], flags, j, scratch)
            if pre_proj and is_key and j + 1 < T:
                o.linear(prev, proj, out=vproj[j])
        back, fwd = o.unpair_sweeps(fbuf)               # outputs_back[i] = backward sweep at step T-1-i
        return o.conv2d((back, x, fwd), self._pc(tm.fusion), act=o.ACT_LRELU, alpha=0.1)

    # -------------------------------------------------------------------------------------------
    def _stage(self, st, x, B, T, flows):
        o = _ops
        shortcut = x
        for blk in st.mlp_blocks:
            x = self._tab(blk, st, x, B, T)
        if self.local_fuse:                                               # function.py:1503-1505
            x = o.conv2d(x, self._pc(st.local_cnn), epi=o.EPI_RESADD, res=shortcut)
        if st.has_traj:
            x = self._trajectory(st, x, B, T, flows[0], flows[1])
        return x

    def _resample(self, rs, x, add=None):                                 # layers.py:777-798
        o = _ops
        y = o.resample_ln(x, add, rs.norm.weight, rs.norm.bias, up=(rs.mode == 'up'))
        return o.linear(y, self._pc(rs.linear))

    def _shortcut(self, sc, x):                                           # vmg.py:389-400,519,525
        o = _ops
        u = o.maxpool4(x)
        u = o.linear(u, self._pc(sc[0]))
        return o.groupnorm1_relu(u, sc[1].weight, sc[1].bias)

    def _weights_stamp(self):
        """Changes whenever a parameter/buffer is re-assigned or written through torch (load_state_dict,
        .to(), optimiser steps); the in-kernel decay writes do not bump it by design."""
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    @torch.no_grad()
    def forward(self, x, flow_pretrained=None, config_amp=None):
        """models/vmg.py:585-637.  The launch sequence of one (shape, mirror) signature is captured into a
        CUDA graph on its second call and replayed afterwards (the reference issues ~4 300 ATen ops per
        frame; the sequential trajectory sweeps are launch-latency bound otherwise)."""
        o = _ops
        o.require_device()
        if self.training:
            raise NotImplementedError('vmg_b200.VMG implements the inference forward only; call .eval()')
        o.check_same_device(x, self.input_proj.proj[0].weight)
        B, T, Cin, H, W = x.shape
        assert H >= 64 and W >= 64, 'The height and width must larger than 64.'
        if T != self.num_out_frames:
            raise RuntimeError(f'shape mismatch: T={T} but num_out_frames={self.num_out_frames} (models/vmg.py:635)')
        x = x.contiguous().float()
        self.frames_mirror = o.mirror_check(x)                            # vmg.py:426-432
        if not (self.use_cuda_graph and x.is_cuda and hasattr(torch.cuda, 'CUDAGraph')):
            return self._forward_impl(x, self.frames_mirror)
        stamp = self._weights_stamp()
        if stamp != self._graph_stamp:
            self._graphs.clear()
            self._graph_stamp = stamp
        key = (tuple(x.shape), self.frames_mirror, self.precision, x.device.index, self.coord_semantics)
        ent = self._graphs.get(key)
        if ent is None:                       # first call: eager (also packs the weights)
            self._graphs[key] = 'warm'
            return self._forward_impl(x, self.frames_mirror)
        if ent == 'warm':                     # second call: capture
            xs = x.clone()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = o.launch_count()
            with torch.cuda.graph(g):
                ys = self._forward_impl(xs, self.frames_mirror)
            ent = (g, xs, ys, o.launch_count() - n0)
            self._graphs[key] = ent
        g, xs, ys, n_kernels = ent
        xs.copy_(x)
        g.replay()
        self.last_graph_launches = n_kernels      # libvmgb kernel nodes replayed by this call
        return ys.clone()

    def _forward_impl(self, x, mirror):
        o = _ops
        B, T, Cin, H, W = x.shape
        Hp = int(math.ceil(H / self.scale)) * self.scale
        Wp = int(math.ceil(W / self.scale)) * self.scale
        xp = o.pad_input(x, Hp, Wp)                                       # [T*B,Hp,Wp,3] fp32
        flows = self._compute_flow(xp, B, T, mirror)
        adt = self._adt
        E, D = self.encoder_layers, self.decoder_layers
        x0 = o.conv2d(xp, self._pc(self.input_proj.proj[0], tc=False), act=o.ACT_LRELU, alpha=0.01, out_dtype=adt,
                      use_tc=False)
        if self.num_layers > 3:                                           # vmg.py:466-567
            x1 = self._stage(E[0], x0, B, T, flows)
            x1_3 = self._shortcut(self.sc_64_16, x1) if self.mdsc else None
            x2 = self._stage(E[1], self._resample(self.downsample[0], x1), B, T, None)
            x2_4 = self._shortcut(self.sc_32_8, x2) if self.mdsc else None
            x3 = self._stage(E[2], self._resample(self.downsample[1], x2), B, T, None)
            x4 = self._stage(E[3], self._resample(self.downsample[2], x3, x1_3), B, T, None)
            x5 = self._stage(D[0], self._resample(self.upsample[0], x4, x2_4), B, T, None)
            x6 = self._stage(D[1], self._resample(self.upsample[1], x5, x3), B, T, None)
            x7 = self._stage(D[2], self._resample(self.upsample[2], x6, x2), B, T, flows)
            y = o.add_crop(x7, x1, H, W)                                  # skip connection + the crop of vmg.py:619 in one pass
        else:                                                             # vmg.py:569-582
            x1 = self._stage(E[0], x0, B, T, flows)
            x2 = self._stage(E[1], self._resample(self.downsample[0], x1), B, T, None)
            x3 = self._stage(D[0], self._resample(self.upsample[0], x2), B, T, flows)
            y = o.add(x3, x1)
        if self.local_fuse:                                               # vmg.py:616-617
            y = o.conv2d(y, self._pc(self.local_cnn), epi=o.EPI_RESADD, res=x0)
        if y.shape[1] != H or y.shape[2] != W:
            y = y[:, :H, :W, :].contiguous()                              # vmg.py:619
        out = torch.empty((B, T, 3, 4 * H, 4 * W), device=x.device, dtype=torch.float32)
        C = y.shape[-1]
        yv = y.view(T, B, H, W, C)
        tc = max(1, self.recon_chunk // B)
        for t0 in range(0, T, tc):                                        # vmg.py:629-636
            z = yv[t0:t0 + tc].reshape(-1, H, W, C)
            z = o.conv2d(z, self._pc(self.upconv1), act=o.ACT_LRELU, alpha=0.1, pixel_shuffle=2)
            z = o.conv2d(z, self._pc(self.upconv2), act=o.ACT_LRELU, alpha=0.1, pixel_shuffle=2)
            z = o.conv2d(z, self._pc(self.HRconv), act=o.ACT_LRELU, alpha=0.1)
            z = o.conv2d(z, self._pc(self.conv_last), out_dtype=torch.float32)
            o.final_output(z, x, out, t0)
        return out
