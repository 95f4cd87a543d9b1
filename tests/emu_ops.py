"""TEST INFRASTRUCTURE -- plain-PyTorch fp32 reference of every op in vmg_b200/ops.py.

Two uses, both inside tests/ only:
  * `-m gpu` tests compare each CUDA kernel with the function of the same name here;
  * `-m "not gpu"` tests monkeypatch `vmg_b200.model._ops` with this module to exercise the HOST logic
    of the drop-in (layouts, key-frame bookkeeping, weight packing order, state_dict schema) on CPU
    against the oracle.  The product never imports this file and has no CPU path.
"""
import torch
import torch.nn.functional as F

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_GELU, ACT_SIGMOID = 0, 1, 2, 3, 4
EPI_STORE, EPI_RESADD, EPI_GATE = 0, 1, 2
COORD_ATEN_CPU, COORD_ATEN_CUDA = 0, 1      # ignored here: torch follows the device the tensors live on


def require_device():
    pass


def check_same_device(x, param):
    pass


class PackedConv:
    pass


def pack_conv(weight, bias, groups=1, tc=False):
    pc = PackedConv()
    w = weight.detach().float()
    if w.dim() == 2:
        w = w[:, :, None, None]
    pc.w = w.clone()
    pc.bias = bias.detach().float().clone() if bias is not None else None
    pc.groups = groups
    pc.Cout, pc.Cin, pc.R, pc.S = w.shape[0], w.shape[1] * groups, w.shape[2], w.shape[3]
    return pc


def _act(v, act, alpha):
    if act == ACT_RELU:
        return F.relu(v)
    if act == ACT_LRELU:
        return F.leaky_relu(v, alpha)
    if act == ACT_GELU:
        return F.gelu(v)
    if act == ACT_SIGMOID:
        return torch.sigmoid(v)
    return v


def conv2d(x, pc, act=ACT_NONE, alpha=0.0, out_scale=1.0, epi=EPI_STORE, res=None, res_scale=1.0, aux=None,
           pixel_shuffle=0, out=None, out_dtype=None, use_tc=True, tc_variant=0):
    xs = x if isinstance(x, (tuple, list)) else (x,)
    odt = out.dtype if out is not None else (out_dtype or xs[0].dtype)
    xin = torch.cat([t.float() for t in xs], -1).permute(0, 3, 1, 2)
    v = F.conv2d(xin, pc.w, pc.bias, 1, (pc.R // 2, pc.S // 2), 1, pc.groups)
    v = _act(v, act, alpha) * out_scale
    if pixel_shuffle == 2:
        v = F.pixel_shuffle(v, 2)
    v = v.permute(0, 2, 3, 1)
    if epi == EPI_RESADD:
        v = res.float() + res_scale * v
    elif epi == EPI_GATE:
        v = res.float() + res_scale * ((aux.float() + v) * torch.tanh(v))
    v = v.to(odt)
    if out is not None:
        out.copy_(v)
        return out
    return v.contiguous()


def linear(x, pc, **kw):
    lead = x.shape[:-1]
    rows = 1
    for s in lead:
        rows *= s
    xv = x.reshape(1, 1, rows, x.shape[-1])
    out = kw.pop('out', None)
    for k in ('res', 'aux'):
        if kw.get(k) is not None:
            kw[k] = kw[k].reshape(1, 1, rows, kw[k].shape[-1])
    y = conv2d(xv, pc, **kw).reshape(*lead, pc.Cout)
    if out is not None:
        out.copy_(y)
        return out
    return y


def layernorm(x, gamma, beta, out_dtype=None, eps=1e-5):
    return F.layer_norm(x.float(), (x.shape[-1],), gamma.float(), beta.float(), eps).to(out_dtype or x.dtype)


def l2norm(x, out_dtype=None):
    return F.normalize(x.float(), dim=-1).to(out_dtype or x.dtype).contiguous()


def _morph_index(N, H, W, C, chunk, axis):
    """row/feature index of every (n,h,w,c) under the SURVEY A.2 map."""
    Cp = -(-C // chunk) * chunk
    S = Cp // chunk
    n, h, w, c = torch.meshgrid(torch.arange(N), torch.arange(H), torch.arange(W), torch.arange(C), indexing='ij')
    l, o_, O, L = (h, w, W, H) if axis == 0 else (w, h, H, W)
    Lg = -(-L // chunk)
    row = ((n * O + o_) * Lg + l // chunk) * chunk + c // S
    f = (l % chunk) * S + c % S
    return row, f, Cp, N * Lg * chunk * O


def morph_gather(x, chunk, axis):
    N, H, W, C = x.shape
    row, f, Cp, rows = _morph_index(N, H, W, C, chunk, axis)
    A = torch.zeros((rows, Cp), dtype=x.dtype)
    A[row.reshape(-1), f.reshape(-1)] = x.reshape(-1)
    return A


def morph_scatter(Y, N, H, W, C, chunk, axis):
    row, f, Cp, rows = _morph_index(N, H, W, C, chunk, axis)
    return Y[row.reshape(-1), f.reshape(-1)].reshape(N, H, W, C).contiguous()


def channel_sum(a, b, c, n_seg, groups, seg_rows, scale):
    C = a.shape[-1]
    s = a.float()
    if b is not None:
        s = s + b.float()
    if c is not None:
        s = s + c.float()
    return s.reshape(n_seg, groups, seg_rows, C).sum((0, 2)) * scale


def softmax3(logits):
    B, C3 = logits.shape
    return logits.reshape(B, C3 // 3, 3).softmax(-1).permute(0, 2, 1).contiguous()


def mix3(h, w, c, a, n_seg, B, seg_rows):
    C = h.shape[-1]
    sh = h.shape
    hv, wv, cv = (t.float().reshape(n_seg, B, seg_rows, C) for t in (h, w, c))
    out = hv * a[None, :, 0, None, :] + wv * a[None, :, 1, None, :] + cv * a[None, :, 2, None, :]
    return out.reshape(sh).to(h.dtype)


def ca_apply(u, x, y, rows_per_frame, scale):
    C = u.shape[-1]
    uv = u.float().reshape(-1, rows_per_frame, C)
    return ((uv * y[:, None, :] + x.float().reshape(-1, rows_per_frame, C)) * scale).reshape(u.shape).to(u.dtype)


def pair_sweeps(x, T):
    n = x.shape[0] // T
    xs = x.view((T, n) + tuple(x.shape[1:]))
    return torch.stack([xs.flip(0), xs], 1).reshape((T, 2 * n) + tuple(x.shape[1:])).contiguous()


def unpair_sweeps(paired):
    T, n = paired.shape[0], paired.shape[1] // 2
    back = paired[:, :n].flip(0).reshape((T * n,) + tuple(paired.shape[2:])).contiguous()
    fwd = paired[:, n:].reshape((T * n,) + tuple(paired.shape[2:])).contiguous()
    return back, fwd


def add(a, b):
    return (a.float() + b.float()).to(a.dtype)


def add_crop(a, b, H, W):
    return (a.float() + b.float()).to(a.dtype)[:, :H, :W, :].contiguous()


def tab_mix_supported(x, C, chunk_h, chunk_w):
    """Same predicate as vmg_b200/ops.py (the fused kernel is bf16-only, so the strict-fp32 CPU runs take the separate ops)."""
    return x.dtype == torch.bfloat16 and x.is_contiguous() and chunk_h == 8 and chunk_w == 8 and C in (64, 112) and x.shape[-1] == C


def tab_mix(x, ln_g, ln_b, eps, pcs, phase, B, acc=None, a=None, m_scaling=1.0):
    """fp32 emulation of csrc/tab_mix.cu out of the separate emulated ops: phase 0 -> per-sample channel sums (as fp32, the
    fixed-point accumulator of the kernel divided by 2^24), phase 1 -> the new activation."""
    N, H, W, C = x.shape
    T = N // B
    xn = layernorm(x.float(), ln_g, ln_b, eps=eps)
    br = []
    for axis, pc in ((0, pcs[0]), (1, pcs[1])):
        A = morph_gather(xn, 8, axis)
        Y = linear(A, pc, act=ACT_RELU, out_scale=1.0 / C)
        br.append(morph_scatter(Y, N, H, W, C, 8, axis))
    cb = linear(xn, pcs[2], act=ACT_RELU, out_scale=1.0 / C)
    if phase == 0:
        return (br[0] + br[1] + cb).view(T, B, H * W, C).sum((0, 2))
    mix = mix3(br[0], br[1], cb, a, T, B, H * W)
    return linear(mix, pcs[3], epi=EPI_GATE, res=x.float(), aux=xn, res_scale=m_scaling).to(x.dtype)


def fixsum_finalize(acc, scale):
    return acc.float() * scale


def decay_weights_(w, gamma):
    w.mul_(gamma)
    return w


def resample_ln(x, add_, gamma, beta, up, eps=1e-5):
    N, H, W, C = x.shape
    v = x.float() if add_ is None else x.float() + add_.float()
    if not up:   # 'n (h neih) (w neiw) c -> n h w (neiw neih c)'
        v = v.view(N, H // 2, 2, W // 2, 2, C).permute(0, 1, 3, 4, 2, 5).reshape(N, H // 2, W // 2, 4 * C)
    else:        # 'n h w (neiw neih c) -> n (h neih) (w neiw) c'
        v = v.view(N, H, W, 2, 2, C // 4).permute(0, 1, 4, 2, 3, 5).reshape(N, 2 * H, 2 * W, C // 4)
    return F.layer_norm(v, (v.shape[-1],), gamma.float(), beta.float(), eps).to(x.dtype).contiguous()


def maxpool4(x):
    return F.max_pool2d(x.float().permute(0, 3, 1, 2), 4).permute(0, 2, 3, 1).to(x.dtype).contiguous()


def groupnorm1_relu(x, gamma, beta, eps=1e-5):
    v = F.group_norm(x.float().permute(0, 3, 1, 2), 1, gamma.float(), beta.float(), eps)
    return F.relu(v).permute(0, 2, 3, 1).to(x.dtype).contiguous()


def _warp(x_nchw, flow_nhw2, mode):
    n, c, h, w = x_nchw.shape
    gy, gx = torch.meshgrid(torch.arange(0, h, device=x_nchw.device), torch.arange(0, w, device=x_nchw.device), indexing='ij')
    g = torch.stack((gx, gy), 2).float() + flow_nhw2
    gxn = 2.0 * g[..., 0] / max(w - 1, 1) - 1.0
    gyn = 2.0 * g[..., 1] / max(h - 1, 1) - 1.0
    return F.grid_sample(x_nchw, torch.stack((gxn, gyn), 3), mode=mode, padding_mode='border', align_corners=True)


def flow_warp_bilinear(x, flow, out=None, semantics=0):
    v = _warp(x.float().permute(0, 3, 1, 2), flow, 'bilinear').permute(0, 2, 3, 1).to(x.dtype)
    if out is not None:
        out.copy_(v)
        return out
    return v.contiguous()


def flow_warp_nearest(loc, flow, K2, out, semantics=0):
    v = _warp(loc[..., :K2].permute(0, 3, 1, 2), flow, 'nearest').permute(0, 2, 3, 1)
    out[..., :K2] = v
    return out


def flow_warp_fused(x, flow, loc, K2, loc_out, out=None, semantics=0):
    flow_warp_nearest(loc, flow, K2, loc_out)
    return flow_warp_bilinear(x, flow, out=out)


def loc_identity_(loc, c0):
    n, h, w, _ = loc.shape
    gy, gx = torch.meshgrid(torch.arange(0, h, device=loc.device), torch.arange(0, w, device=loc.device), indexing='ij')
    loc[..., c0] = gx.float()
    loc[..., c0 + 1] = gy.float()
    return loc


def flow_smooth(flow, r):
    n, h, w, _ = flow.shape
    f = flow.permute(0, 3, 1, 2)
    hf, wf = -(-h // r) * r, -(-w // r) * r
    f = F.pad(f, (0, wf - w, 0, hf - h), mode='reflect')
    f = F.adaptive_avg_pool2d(f, (hf // r, wf // r))
    f = F.interpolate(f, scale_factor=r, mode='nearest')[..., :h, :w]
    return f.permute(0, 2, 3, 1).contiguous()


def ltam_wins(q, keys0, vals0, kv_stride, K, loc, rel, decay, heads, out_bias=None, anchor=None):
    """q: [n,h,w,C]; keys0/vals0 are views whose storage offset marks key-frame 0 (see model.py)."""
    n, h, w, C = q.shape
    d = C // heads
    fe = n * h * w * C

    def frames(t0):
        base = t0.storage_offset()
        flat = t0.untyped_storage()
        full = torch.empty(0, dtype=t0.dtype).set_(flat)
        return torch.stack([full[base + k * kv_stride: base + k * kv_stride + fe].view(n, h, w, C) for k in range(K)], 0)
    keys, vals = frames(keys0).float(), frames(vals0).float()
    lx = loc[..., 0:2 * K:2].round().long()              # [n,h,w,K]
    ly = loc[..., 1:2 * K:2].round().long()
    bi = torch.arange(n)[:, None, None, None].expand(n, h, w, K)
    ki = torch.arange(K)[None, None, None, :].expand(n, h, w, K)
    inb = ((lx >= 0) & (lx < w) & (ly >= 0) & (ly < h)).float()[..., None]     # padding_mode='zeros': a zero key / value row
    lx, ly = lx.clamp(0, w - 1), ly.clamp(0, h - 1)
    kg = keys[ki, bi, ly, lx] * inb                       # [n,h,w,K,C]
    vg = vals[ki, bi, ly, lx] * inb
    qq = (q.float() * (d ** -0.5)).view(n, h // 2, 2, w // 2, 2, heads, d).permute(0, 1, 3, 5, 2, 4, 6).reshape(n, h // 2, w // 2, heads, 4, d)
    kk = kg.view(n, h // 2, 2, w // 2, 2, K, heads, d).permute(0, 1, 3, 6, 5, 2, 4, 7).reshape(n, h // 2, w // 2, heads, 4 * K, d)
    vv = vg.view(n, h // 2, 2, w // 2, 2, K, heads, d).permute(0, 1, 3, 6, 5, 2, 4, 7).reshape(n, h // 2, w // 2, heads, 4 * K, d)
    logits = torch.einsum('nyxhqd,nyxhkd->nyxhqk', qq, kk)
    pw = torch.stack([decay.float() ** (K - k) for k in range(K)], 1)          # [heads,K]
    pe = pw[:, None, :, None] * rel.float()[:, :, None, :]                      # [heads,4,K,4]
    attn = (logits + pe.reshape(heads, 4, 4 * K)[None, None, None]).softmax(-1)
    o_ = torch.einsum('nyxhqk,nyxhkd->nyxhqd', attn, vv)                        # [n,h/2,w/2,heads,4,d]
    o_ = o_.view(n, h // 2, w // 2, heads, 2, 2, d).permute(0, 1, 4, 2, 5, 3, 6).reshape(n, h, w, C)
    if out_bias is not None:
        o_ = o_ + out_bias.float()
    if anchor is not None:
        o_ = o_ + anchor.float()
    return o_.to(q.dtype).contiguous()


def spynet_prep(frames, mean, std, h_up, w_up):
    v = F.interpolate(frames.permute(0, 3, 1, 2), size=(h_up, w_up), mode='bilinear', align_corners=False)
    v = (v - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)
    return v.permute(0, 2, 3, 1).contiguous()


def avgpool2(x):
    return F.avg_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).contiguous()


def spynet_level_input(pyr, flow_coarse, n_pairs, B, direction, out_dtype=torch.float32, semantics=0):
    _, h, w, _ = pyr.shape
    ref = pyr[B:B + n_pairs] if direction == 0 else pyr[:n_pairs]
    sup = pyr[:n_pairs] if direction == 0 else pyr[B:B + n_pairs]
    if flow_coarse is None:
        fup = torch.zeros((n_pairs, h, w, 2))
    else:
        fup = (F.interpolate(flow_coarse.permute(0, 3, 1, 2), scale_factor=2, mode='bilinear', align_corners=True) * 2.0).permute(0, 2, 3, 1)
    warped = _warp(sup.permute(0, 3, 1, 2), fup, 'bilinear').permute(0, 2, 3, 1)
    return torch.cat([ref, warped, fup], -1).to(out_dtype).contiguous(), fup.contiguous()


def flow_resize(flow, h, w):
    n, hu, wu, _ = flow.shape
    f = F.interpolate(flow.permute(0, 3, 1, 2), size=(h, w), mode='bilinear', align_corners=False).clone()
    f[:, 0] *= float(w) / float(wu)
    f[:, 1] *= float(h) / float(hu)
    return f.permute(0, 2, 3, 1).contiguous()


def pad_input(x, Hp, Wp):
    B, T, C, H, W = x.shape
    v = F.pad(x, (0, Wp - W, 0, Hp - H, 0, 0), mode='replicate')
    return v.permute(1, 0, 3, 4, 2).reshape(T * B, Hp, Wp, C).contiguous()


def mirror_check(x):
    if x.size(1) % 2:
        return False
    a, b = torch.chunk(x, 2, dim=1)
    return bool(torch.linalg.norm(a - b.flip(1)) == 0)


def final_output(y, x, out, t0):
    B, T, _, H, W = x.shape
    Tc = y.shape[0] // B
    up = F.interpolate(x[:, t0:t0 + Tc].reshape(-1, 3, H, W), size=(4 * H, 4 * W), mode='bilinear', align_corners=False)
    up = up.view(B, Tc, 3, 4 * H, 4 * W)
    out[:, t0:t0 + Tc] = y.float().view(Tc, B, 4 * H, 4 * W, 3).permute(1, 0, 4, 2, 3) + up
    return out


# ---- callers either side of the forward (vmg_b200/tester.py host logic on CPU) -----------------------------------
def frames_u8_to_f32(frames_u8, reverse_channels=False):
    x = frames_u8.flip(-1) if reverse_channels else frames_u8
    return (x.float() / 255.0).permute(0, 3, 1, 2).contiguous().unsqueeze(0)


def frames_f32_to_u8(frames, reverse_channels=False, out=None):
    if frames.dim() == 5:
        frames = frames[0]
    v = torch.round(frames.clamp(0, 1) * 255.0).to(torch.uint8).permute(0, 2, 3, 1)
    return (v.flip(-1) if reverse_channels else v).contiguous()


def blend_accumulate_(E, count, patch, t0, h0, w0, zero_t=(0, 0), zero_h=(0, 0), zero_w=(0, 0)):
    B, Tp, C, ph, pw = patch.shape
    m = torch.zeros((Tp, ph, pw))
    m[zero_t[0]:Tp - zero_t[1], zero_h[0]:ph - zero_h[1], zero_w[0]:pw - zero_w[1]] = 1.0
    E[:, t0:t0 + Tp, :, h0:h0 + ph, w0:w0 + pw] += patch * m[None, :, None]
    if count is not None:
        count[t0:t0 + Tp, h0:h0 + ph, w0:w0 + pw] += m
    return E


def blend_normalize_(E, count):
    E.div_(count[None, :, None])
    return E


def augment(x, hflip, vflip, transpose):
    if hflip:
        x = x.flip(-1)
    if vflip:
        x = x.flip(-2)
    if transpose:
        x = x.transpose(-1, -2)
    return x.contiguous()


def sse_f32(a, b, clamp01=False):
    if clamp01:
        a, b = a.clamp(0, 1), b.clamp(0, 1)
    return ((a.double() - b.double()) ** 2).sum().reshape(1)


def blend_finalize_u8(E, count=None, reverse_channels=False):
    v = E if count is None else E / count[None, :, None]
    return frames_f32_to_u8(v, reverse_channels)


# ---- residual-block chain (same host-side contract as vmg_b200.ops) ---------------------------------------------
class PackedChain:
    pass


def pack_resblock_chain(layers):
    pc = PackedChain()
    pc.layers = list(layers)
    pc.n_blocks = len(pc.layers) // 2
    return pc


def chain_flags(y):
    return None


def resblock_chain(y, pc, res_scale, out, flags, flags_next, scratch=None):
    nb = len(pc.layers) // 2
    for k in range(nb):
        t = conv2d(y, pc.layers[2 * k], act=ACT_RELU)
        y = conv2d(t, pc.layers[2 * k + 1], epi=EPI_RESADD, res=y, res_scale=res_scale, out=out if k == nb - 1 else None)
    return y
