"""ctypes binding of libvmgb.so (the C ABI in include/vmgb.h) + thin tensor-level wrappers.

PyTorch is used here for device memory, streams and dtype bookkeeping only: every arithmetic step of
the VMG forward goes through a vmgb_* kernel.  There is NO fallback: importing works without a GPU
(so the ABI can be inspected), but any op call without the library or without an sm_100 device raises.

Internal activation layout: channels-last, frames time-major -- tensors are [T*B, H, W, C] (or row
matrices [rows, C]).  See include/vmgb.h for the per-kernel contracts and reference citations.
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), 'include', 'vmgb.h')
LIB_PATH = os.environ.get('VMGB_LIB') or os.path.join(_HERE, 'libvmgb.so')      # VMGB_LIB: A/B builds (tools only)

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_GELU, ACT_SIGMOID = 0, 1, 2, 3, 4
EPI_STORE, EPI_RESADD, EPI_GATE = 0, 1, 2
COORD_ATEN_CPU, COORD_ATEN_CUDA = 0, 1      # see include/vmgb.h

_CTYPES = {
    'int': ctypes.c_int, 'float': ctypes.c_float, 'long long': ctypes.c_longlong,
    'void*': ctypes.c_void_p, 'const void*': ctypes.c_void_p, 'float*': ctypes.c_void_p,
    'const float*': ctypes.c_void_p, 'double*': ctypes.c_void_p, 'const double*': ctypes.c_void_p, 'long long*': ctypes.c_void_p, 'const long long*': ctypes.c_void_p, 'const char*': ctypes.c_char_p, 'void': None,
}


def parse_header(path=HEADER):
    """Returns {symbol: (restype, [argtypes])} for every prototype declared in include/vmgb.h."""
    src = open(path).read()
    src = re.sub(r'/\*.*?\*/', ' ', src, flags=re.S)
    protos = {}
    for m in re.finditer(r'\b(int|long long|const char\*)\s+(vmgb_\w+)\s*\(([^)]*)\)\s*;', src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        argt = []
        if args and args != 'void':
            for a in args.split(','):
                a = ' '.join(a.split())
                t = re.sub(r'\s*\w+$', '', a).replace(' *', '*')
                argt.append(_CTYPES[t])
        protos[name] = (_CTYPES[ret], argt)
    return protos


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(f'{LIB_PATH} is missing: build it with vmg_b200/csrc/build.sh '
                               '(python -c "import __graft_entry__ as g; g.build()"). There is no fallback path.')
        L = ctypes.CDLL(LIB_PATH)
        for name, (ret, argt) in parse_header().items():
            fn = getattr(L, name)
            fn.restype = ret
            fn.argtypes = argt
        _lib = L
    return _lib


_checked = None      # ordinal of the device the library was first used on (one process drives one GPU)


def require_device():
    global _checked
    if not torch.cuda.is_available():
        raise RuntimeError('vmg_b200 needs a CUDA device (sm_100a); there is no CPU fallback.')
    cur = torch.cuda.current_device()
    if _checked is not None:
        if cur != _checked:
            raise RuntimeError(f'vmg_b200 was initialised on cuda:{_checked} and now runs on cuda:{cur}: one process drives '
                               'one GPU (launch one process per GPU with torchrun)')
        return
    rc = lib().vmgb_device_check()
    if rc != 0:
        raise RuntimeError(f'vmg_b200 kernels are sm_100a-only; vmgb_device_check() = {rc}')
    _checked = cur


def check_same_device(x, param):
    """Input, parameters and the current CUDA device must be the same GPU: the kernels take raw pointers and launch
    on the current device's stream (a CPU tensor would become an illegal-address fault)."""
    if not x.is_cuda or x.device != param.device or x.device.index != torch.cuda.current_device():
        raise RuntimeError(f'vmg_b200.VMG: input on {x.device}, parameters on {param.device}, current device '
                           f'cuda:{torch.cuda.current_device()} -- all three must be the same sm_100 GPU (there is no CPU path)')


def _ck(rc, name):
    if rc != 0:
        msg = ''
        if rc == -3:
            msg = ' cuda: ' + (lib().vmgb_last_cuda_error_string() or b'').decode()
        raise RuntimeError(f'{name} failed with code {rc}.{msg}')


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _dt(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f'unsupported dtype {t.dtype}')


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _pix(t):
    """(ptr, pitch) of a channels-last tensor whose pixels are uniformly strided (slices of a wider
    contiguous buffer along the channel dim are allowed)."""
    assert t.stride(-1) == 1, 'channel dim must be contiguous'
    pitch = t.stride(-2) if t.dim() >= 2 else t.shape[-1]
    exp = pitch
    for d in range(t.dim() - 2, -1, -1):
        if t.shape[d] != 1:
            assert t.stride(d) == exp, f'non-uniform pixel stride {t.stride()} for shape {tuple(t.shape)}'
        exp *= t.shape[d]
    return t.data_ptr(), pitch


def launch_count():
    return int(lib().vmgb_launch_count())


# ---------------------------------------------------------------------------------------------
# weights
# ---------------------------------------------------------------------------------------------
class PackedConv:
    """Kernel-side copy of one Conv2d / Linear weight (fp32 [G][taps][Cin_g][Cout_g] and, for the
    tcgen05 path, bf16 [G][taps][Cout_g][Kp])."""
    __slots__ = ('w32', 'wtc', 'Kp', 'bias', 'Cout', 'Cin', 'R', 'S', 'groups')


def pack_conv(weight, bias, groups=1, tc=False):
    require_device()
    w = weight.detach()
    if w.dim() == 2:
        w = w[:, :, None, None]
    w = w.contiguous().float()
    Cout, Cin_g, R, S = w.shape
    pc = PackedConv()
    pc.Cout, pc.Cin, pc.R, pc.S, pc.groups = Cout, Cin_g * groups, R, S, groups
    pc.bias = bias.detach().contiguous().float() if bias is not None else None
    pc.w32 = torch.empty(groups * R * S * Cin_g * (Cout // groups), device=w.device, dtype=torch.float32)
    _ck(lib().vmgb_pack_weight(w.data_ptr(), pc.w32.data_ptr(), Cout, Cin_g, R, S, groups, _stream()), 'vmgb_pack_weight')
    pc.wtc, pc.Kp = None, 0
    if tc:
        koff = max((g * Cin_g) % 8 for g in range(groups))     # 16-byte alignment shift of grouped convs
        pc.Kp = (Cin_g + koff + 15) // 16 * 16
        pc.wtc = torch.empty(groups * R * S * (Cout // groups) * pc.Kp, device=w.device, dtype=torch.bfloat16)
        _ck(lib().vmgb_pack_weight_tc(w.data_ptr(), pc.wtc.data_ptr(), Cout, Cin_g, pc.Kp, R, S, groups, _stream()),
            'vmgb_pack_weight_tc')
    return pc



class PackedChain:
    """Kernel-side weights of one residual-block chain: the tcgen05 layouts of conv1, conv2 of block 0, 1, ... back to back
    (bf16 [2n][9][C][Kp]) and the biases (fp32 [2n][C]); `layers` keeps the per-layer PackedConv for the per-layer path."""
    __slots__ = ('wtc', 'bias', 'Kp', 'C', 'n_blocks', 'layers')


def pack_resblock_chain(layers):
    """layers: PackedConv of conv1, conv2 of every block in order (all 3x3, C -> C, packed with tc=True)."""
    pc = PackedChain()
    pc.layers = list(layers)
    pc.n_blocks = len(pc.layers) // 2
    l0 = pc.layers[0]
    pc.C, pc.Kp = l0.Cout, l0.Kp
    pc.wtc, pc.bias = None, None
    ok = all(q.wtc is not None and q.R == 3 and q.S == 3 and q.groups == 1 and q.Cin == pc.C and q.Cout == pc.C and q.Kp == pc.Kp and
             q.bias is not None for q in pc.layers)
    if ok and len(pc.layers) == 2 * pc.n_blocks and pc.n_blocks >= 1:
        pc.wtc = torch.cat([q.wtc for q in pc.layers])
        pc.bias = torch.stack([q.bias for q in pc.layers]).contiguous()
    return pc


def resblock_chain_supported(pc, y):
    return (pc.wtc is not None and y.dtype == torch.bfloat16 and y.is_contiguous() and pc.C % 16 == 0 and (pc.C <= 128 or pc.C in (144, 160)) and
            y.shape[2] % 16 == 0 and y.shape[1] >= 8 and y.shape[-1] == pc.C)


def chain_flags(y):
    """Two zeroed dependency-counter arrays for vmgb_resblock_chain on tensors shaped like y [n,H,W,C]."""
    n, H, W, _ = y.shape
    return torch.zeros((2, n * ((H + 15) // 16) * (W // 16)), device=y.device, dtype=torch.int32)


def resblock_chain(y, pc, res_scale, out, flags, flags_next, scratch=None):
    """y <- y + res_scale * conv2(relu(conv1(y))) for every block of `pc`; y [n,H,W,C] bf16 is consumed (overwritten),
    the result is written to `out`.  ONE persistent kernel when the shape is supported, else 2 launches per block."""
    require_device()
    if resblock_chain_supported(pc, y):
        n, H, W, C = y.shape
        y_alt, t_tmp = scratch if scratch is not None else (torch.empty_like(y), torch.empty_like(y))
        assert out.is_contiguous() and out.shape == y.shape and out.dtype == y.dtype
        rc = lib().vmgb_resblock_chain(y.data_ptr(), y_alt.data_ptr(), t_tmp.data_ptr(), out.data_ptr(), pc.wtc.data_ptr(), pc.Kp,
                                       pc.bias.data_ptr(), flags.data_ptr(), flags_next.data_ptr(), n, H, W, C, pc.n_blocks,
                                       float(res_scale), _stream())
        if rc == 0:
            return out
        if rc != -2:
            _ck(rc, 'vmgb_resblock_chain')
    nb = len(pc.layers) // 2
    for k in range(nb):
        t = conv2d(y, pc.layers[2 * k], act=ACT_RELU)
        y = conv2d(t, pc.layers[2 * k + 1], epi=EPI_RESADD, res=y, res_scale=res_scale, out=out if k == nb - 1 else None)
    return y


FALLBACK_LOG = []


def tc_supported(pc, xs, n_src):
    """Shapes the tcgen05 kernel takes (see csrc/conv_tc.cu)."""
    if pc.wtc is None or xs[0].dtype != torch.bfloat16:
        return False
    if n_src > 1 and (pc.Cin // n_src) % 16:
        return False
    if pc.R > 1 and (xs[0].shape[1] < 8 or xs[0].shape[2] < 8):
        return False        # images smaller than a tile (coarsest SPyNet levels): TMA boxes would be mostly out of bounds
    return all(_pix(x)[1] % 8 == 0 and x.data_ptr() % 16 == 0 for x in xs)


def conv2d(x, pc, act=ACT_NONE, alpha=0.0, out_scale=1.0, epi=EPI_STORE, res=None, res_scale=1.0, aux=None,
           pixel_shuffle=0, out=None, out_dtype=None, use_tc=True, tc_variant=0):
    """x: [N,H,W,Cin] tensor or tuple of up to 3 equally wide tensors (virtual concat).  Returns
    [N,H,W,Cout] (or [N,2H,2W,Cout/4] with pixel_shuffle=2)."""
    require_device()
    xs = x if isinstance(x, (tuple, list)) else (x,)
    n_src = len(xs)
    N, H, W, _ = xs[0].shape
    ptrs = [_pix(t) for t in xs]
    for t, (p, pitch) in zip(xs, ptrs):
        assert t.shape == xs[0].shape and t.dtype == xs[0].dtype and pitch == ptrs[0][1]
    assert xs[0].shape[-1] * n_src == pc.Cin, (xs[0].shape, n_src, pc.Cin)
    if out is None:
        shape = (N, 2 * H, 2 * W, pc.Cout // 4) if pixel_shuffle == 2 else (N, H, W, pc.Cout)
        out = torch.empty(shape, device=xs[0].device, dtype=out_dtype or xs[0].dtype)
    optr, opitch = _pix(out)
    rptr, rpitch, rdt = (None, 0, 0)
    if res is not None:
        rptr, rpitch = _pix(res)
        rdt = _dt(res)
        assert res.shape == out.shape
    aptr, apitch, adt = (None, 0, 0)
    if aux is not None:
        aptr, apitch = _pix(aux)
        adt = _dt(aux)
        assert aux.shape == out.shape
    p1 = ptrs[1][0] if n_src > 1 else None
    p2 = ptrs[2][0] if n_src > 2 else None
    bias = _ptr(pc.bias)
    if use_tc and tc_supported(pc, xs, n_src):
        _ck(lib().vmgb_conv2d_nhwc_tc(ptrs[0][0], p1, p2, n_src, ptrs[0][1], pc.wtc.data_ptr(), pc.Kp, bias,
                                      optr, _dt(out), opitch, rptr, rdt, rpitch, aptr, adt, apitch,
                                      N, H, W, pc.Cin, pc.Cout, pc.R, pc.S, pc.groups,
                                      act, alpha, out_scale, epi, res_scale, pixel_shuffle, tc_variant, _stream()),
            'vmgb_conv2d_nhwc_tc')
    else:
        if xs[0].dtype == torch.bfloat16 and len(FALLBACK_LOG) < 64:      # diagnostics: bf16 launches that missed the tcgen05 path
            FALLBACK_LOG.append((tuple(xs[0].shape), ptrs[0][1], pc.Cin, pc.Cout, pc.R, pc.groups, n_src, use_tc, pc.wtc is not None))
        _ck(lib().vmgb_conv2d_nhwc(ptrs[0][0], p1, p2, n_src, _dt(xs[0]), ptrs[0][1], pc.w32.data_ptr(), bias,
                                   optr, _dt(out), opitch, rptr, rdt, rpitch, aptr, adt, apitch,
                                   N, H, W, pc.Cin, pc.Cout, pc.R, pc.S, pc.groups,
                                   act, alpha, out_scale, epi, res_scale, pixel_shuffle, _stream()),
            'vmgb_conv2d_nhwc')
    return out


def _rows_view(t):
    """[..., C] tensor whose rows are uniformly strided -> [1, 1, rows, C] view sharing its memory."""
    rows = 1
    for s in t.shape[:-1]:
        rows *= s
    if t.is_contiguous():
        return t.reshape(1, 1, rows, t.shape[-1])
    _, pitch = _pix(t)
    return t.as_strided((1, 1, rows, t.shape[-1]), (0, 0, pitch, 1))


def linear(x, pc, **kw):
    """x [..., K] -> [..., N]; a Linear is the 1x1 'conv' over a [1,1,rows,K] image.  x / res / aux / out may be
    channel slices of wider (pitched) buffers."""
    lead = x.shape[:-1]
    out = kw.get('out')
    for k in ('res', 'aux', 'out'):
        if kw.get(k) is not None:
            kw[k] = _rows_view(kw[k])
    y = conv2d(_rows_view(x), pc, **kw)
    return out if out is not None else y.reshape(*lead, pc.Cout)


# ---------------------------------------------------------------------------------------------
# row kernels
# ---------------------------------------------------------------------------------------------
def layernorm(x, gamma, beta, out_dtype=None, eps=1e-5):
    require_device()
    assert x.is_contiguous()
    C = x.shape[-1]
    out = torch.empty(x.shape, device=x.device, dtype=out_dtype or x.dtype)
    _ck(lib().vmgb_layernorm(x.data_ptr(), _dt(x), out.data_ptr(), _dt(out), gamma.data_ptr(), beta.data_ptr(),
                             x.numel() // C, C, eps, _stream()), 'vmgb_layernorm')
    return out


def l2norm(x, out_dtype=None):
    require_device()
    ptr, pitch = _pix(x)
    C = x.shape[-1]
    out = torch.empty(x.shape, device=x.device, dtype=out_dtype or x.dtype)
    _ck(lib().vmgb_l2norm(ptr, _dt(x), pitch, out.data_ptr(), _dt(out), x.numel() // C, C, _stream()), 'vmgb_l2norm')
    return out


def morph_gather(x, chunk, axis):
    require_device()
    N, H, W, C = x.shape
    assert x.is_contiguous()
    Cp = -(-C // chunk) * chunk
    L, O = (H, W) if axis == 0 else (W, H)
    rows = N * (-(-L // chunk) * chunk) * O
    pitch = -(-Cp // 8) * 8                      # 16-byte rows (bf16) so the Linear can read them through TMA
    A = torch.empty((rows, pitch), device=x.device, dtype=x.dtype)[:, :Cp]
    _ck(lib().vmgb_morph_permute(x.data_ptr(), _dt(x), A.data_ptr(), _dt(A), N, H, W, C, chunk, axis, 0, pitch, _stream()),
        'vmgb_morph_permute')
    return A


def morph_scatter(Y, N, H, W, C, chunk, axis):
    require_device()
    assert Y.stride(-1) == 1
    out = torch.empty((N, H, W, C), device=Y.device, dtype=Y.dtype)
    _ck(lib().vmgb_morph_permute(Y.data_ptr(), _dt(Y), out.data_ptr(), _dt(out), N, H, W, C, chunk, axis, 1, Y.stride(0), _stream()),
        'vmgb_morph_permute')
    return out


def channel_sum(a, b, c, n_seg, groups, seg_rows, scale):
    require_device()
    C = a.shape[-1]
    for t in (a, b, c):
        assert t is None or (t.is_contiguous() and t.dtype == a.dtype)
    out = torch.empty((groups, C), device=a.device, dtype=torch.float32)
    acc = torch.zeros((groups, C), device=a.device, dtype=torch.int64)        # fixed-point accumulators (order-independent atomics)
    _ck(lib().vmgb_channel_sum(a.data_ptr(), _ptr(b), _ptr(c), _dt(a), out.data_ptr(), acc.data_ptr(), n_seg, groups, seg_rows, C,
                               scale, _stream()), 'vmgb_channel_sum')
    return out


def tab_mix_supported(x, C, chunk_h, chunk_w):
    """Shapes the fused Morph-FC mixing kernel takes (csrc/tab_mix.cu)."""
    return (x.dtype == torch.bfloat16 and x.is_contiguous() and chunk_h == 8 and chunk_w == 8 and C in (64, 112) and x.shape[-1] == C and
            os.environ.get('VMGB_NO_TAB_MIX', '0') != '1')


def tab_mix(x, ln_g, ln_b, eps, pcs, phase, B, acc=None, a=None, m_scaling=1.0):
    """One phase of the fused mixing half of a TAB (include/vmgb.h vmgb_tab_mix).  pcs = PackedConv of (mlp_h, mlp_w, mlp_c, proj).
    phase 0 returns the fixed-point accumulator [B, C]; phase 1 returns the new activation.  None if the shape is unsupported."""
    require_device()
    N, H, W, C = x.shape
    ph, pw, pc, pp = pcs
    assert all(q.wtc is not None and q.Kp == ph.Kp and q.Cin == C and q.Cout == C for q in pcs)
    out = None
    if phase == 0:
        acc = torch.zeros((B, C), device=x.device, dtype=torch.int64)
    else:
        out = torch.empty_like(x)
        assert a.dtype == torch.float32 and a.is_contiguous() and tuple(a.shape) == (B, 3, C)
    rc = lib().vmgb_tab_mix(x.data_ptr(), ln_g.data_ptr(), ln_b.data_ptr(), float(eps), ph.wtc.data_ptr(), pw.wtc.data_ptr(), pc.wtc.data_ptr(),
                            pp.wtc.data_ptr(), ph.Kp, _ptr(ph.bias), _ptr(pw.bias), _ptr(pc.bias), _ptr(pp.bias), phase, _ptr(acc), _ptr(a),
                            _ptr(out), float(m_scaling), N, B, H, W, C, 8, _stream())
    if rc == -2:
        return None
    _ck(rc, 'vmgb_tab_mix')
    return acc if phase == 0 else out


def fixsum_finalize(acc, scale):
    require_device()
    out = torch.empty(acc.shape, device=acc.device, dtype=torch.float32)
    _ck(lib().vmgb_fixsum_finalize(acc.data_ptr(), out.data_ptr(), acc.numel(), float(scale), _stream()), 'vmgb_fixsum_finalize')
    return out


def softmax3(logits):
    require_device()
    B, C3 = logits.shape
    assert logits.dtype == torch.float32 and logits.is_contiguous()
    a = torch.empty((B, 3, C3 // 3), device=logits.device, dtype=torch.float32)
    _ck(lib().vmgb_softmax3(logits.data_ptr(), a.data_ptr(), B, C3 // 3, _stream()), 'vmgb_softmax3')
    return a


def mix3(h, w, c, a, n_seg, B, seg_rows):
    require_device()
    C = h.shape[-1]
    out = torch.empty_like(h)
    _ck(lib().vmgb_mix3(h.data_ptr(), w.data_ptr(), c.data_ptr(), _dt(h), a.data_ptr(), out.data_ptr(), _dt(out),
                        n_seg, B, seg_rows, C, _stream()), 'vmgb_mix3')
    return out


def ca_apply(u, x, y, rows_per_frame, scale):
    require_device()
    C = u.shape[-1]
    frames = u.numel() // C // rows_per_frame
    out = torch.empty_like(u)
    _ck(lib().vmgb_ca_apply(u.data_ptr(), x.data_ptr(), _dt(u), y.data_ptr(), out.data_ptr(), _dt(out), frames,
                            rows_per_frame, C, scale, _stream()), 'vmgb_ca_apply')
    return out


def pair_sweeps(x, T):
    """x [T*n, ...] time-major -> [T, 2n, ...]: step j = [x[T-1-j] | x[j]] (both trajectory sweeps in lockstep)."""
    require_device()
    assert x.is_contiguous() and x.shape[0] % T == 0
    n = x.shape[0] // T
    out = torch.empty((T, 2 * n) + tuple(x.shape[1:]), device=x.device, dtype=x.dtype)
    _ck(lib().vmgb_pair_sweeps(x.data_ptr(), out.data_ptr(), T, x.numel() // T * x.element_size(), _stream()), 'vmgb_pair_sweeps')
    return out


def unpair_sweeps(paired):
    """paired [T, 2n, ...] -> (back, fwd), each [T*n, ...]: back[i] = paired[T-1-i, :n], fwd[i] = paired[i, n:]."""
    require_device()
    assert paired.is_contiguous() and paired.shape[1] % 2 == 0
    T, n = paired.shape[0], paired.shape[1] // 2
    back = torch.empty((T * n,) + tuple(paired.shape[2:]), device=paired.device, dtype=paired.dtype)
    fwd = torch.empty_like(back)
    fb = paired.numel() // (2 * T) * paired.element_size()
    _ck(lib().vmgb_unpair_sweeps(paired.data_ptr(), back.data_ptr(), fwd.data_ptr(), T, fb, _stream()), 'vmgb_unpair_sweeps')
    return back, fwd


def add(a, b):
    require_device()
    assert a.shape == b.shape and a.dtype == b.dtype and a.is_contiguous() and b.is_contiguous()
    out = torch.empty_like(a)
    _ck(lib().vmgb_add(a.data_ptr(), b.data_ptr(), _dt(a), out.data_ptr(), _dt(out), a.numel(), _stream()), 'vmgb_add')
    return out


def add_crop(a, b, H, W):
    """(a + b)[:, :H, :W, :] for [N,Hp,Wp,C] tensors in one pass (bf16, C % 8 == 0), else add + slice."""
    require_device()
    N, Hp, Wp, C = a.shape
    assert a.shape == b.shape and a.dtype == b.dtype and a.is_contiguous() and b.is_contiguous()
    if a.dtype == torch.bfloat16 and C % 8 == 0:
        out = torch.empty((N, H, W, C), device=a.device, dtype=a.dtype)
        rc = lib().vmgb_add_crop(a.data_ptr(), b.data_ptr(), out.data_ptr(), N, Hp, Wp, H, W, C, _stream())
        if rc == 0:
            return out
        if rc != -2:
            _ck(rc, 'vmgb_add_crop')
    y = add(a, b)
    return y if (Hp == H and Wp == W) else y[:, :H, :W, :].contiguous()


def decay_weights_(w, gamma):
    require_device()
    assert w.dtype == torch.float32 and gamma.dtype == torch.float32 and w.is_contiguous() and gamma.is_contiguous()
    assert w.shape == gamma.shape
    _ck(lib().vmgb_decay_weights(w.data_ptr(), gamma.data_ptr(), w.numel(), _stream()), 'vmgb_decay_weights')
    return w


def resample_ln(x, add_, gamma, beta, up, eps=1e-5):
    require_device()
    N, H, W, C = x.shape
    assert x.is_contiguous() and (add_ is None or (add_.is_contiguous() and add_.shape == x.shape))
    shape = (N, 2 * H, 2 * W, C // 4) if up else (N, H // 2, W // 2, 4 * C)
    F_ = shape[-1]
    pitch = -(-F_ // 8) * 8                      # rows padded to 16 bytes (bf16) so the next Linear can use TMA
    buf = torch.empty(shape[:-1] + (pitch,), device=x.device, dtype=x.dtype)
    out = buf[..., :F_]
    _ck(lib().vmgb_resample_ln(x.data_ptr(), _ptr(add_), _dt(x), out.data_ptr(), _dt(out), pitch, gamma.data_ptr(),
                               beta.data_ptr(), N, H, W, C, int(up), eps, _stream()), 'vmgb_resample_ln')
    return out


def maxpool4(x):
    require_device()
    N, H, W, C = x.shape
    out = torch.empty((N, H // 4, W // 4, C), device=x.device, dtype=x.dtype)
    _ck(lib().vmgb_maxpool4(x.data_ptr(), _dt(x), out.data_ptr(), _dt(out), N, H, W, C, _stream()), 'vmgb_maxpool4')
    return out


def groupnorm1_relu(x, gamma, beta, eps=1e-5):
    """GroupNorm(1, C) per frame followed by ReLU; x [N,H,W,C]."""
    require_device()
    N, H, W, C = x.shape
    stats = torch.zeros((N, 2), device=x.device, dtype=torch.int64)          # fixed-point (sum, sum of squares) per frame
    _ck(lib().vmgb_frame_stats(x.data_ptr(), _dt(x), stats.data_ptr(), N, H * W * C, _stream()), 'vmgb_frame_stats')
    out = torch.empty_like(x)
    _ck(lib().vmgb_groupnorm1_relu(x.data_ptr(), _dt(x), stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                   out.data_ptr(), _dt(out), N, H * W, C, eps, _stream()), 'vmgb_groupnorm1_relu')
    return out


# ---------------------------------------------------------------------------------------------
# warps / trajectory
# ---------------------------------------------------------------------------------------------
def flow_warp_bilinear(x, flow, out=None, semantics=COORD_ATEN_CUDA):
    require_device()
    n, h, w, C = x.shape
    xp, xpitch = _pix(x)
    if out is None:
        out = torch.empty((n, h, w, C), device=x.device, dtype=x.dtype)
    op, opitch = _pix(out)
    assert flow.dtype == torch.float32 and flow.is_contiguous() and flow.shape == (n, h, w, 2)
    _ck(lib().vmgb_flow_warp_bilinear(xp, _dt(x), xpitch, flow.data_ptr(), op, _dt(out), opitch, n, h, w, C, semantics, _stream()),
        'vmgb_flow_warp_bilinear')
    return out


def flow_warp_nearest(loc, flow, K2, out, semantics=COORD_ATEN_CUDA):
    require_device()
    n, h, w, pitch = loc.shape
    assert loc.is_contiguous() and out.is_contiguous() and out.shape == loc.shape and loc.dtype == torch.float32
    _ck(lib().vmgb_flow_warp_nearest(loc.data_ptr(), pitch, flow.data_ptr(), out.data_ptr(), pitch, n, h, w, K2,
                                     semantics, _stream()), 'vmgb_flow_warp_nearest')
    return out


def flow_warp_fused(x, flow, loc, K2, loc_out, out=None, semantics=COORD_ATEN_CUDA):
    """Both warps of one trajectory step in one launch: returns the bilinear-warped features; the nearest-warped
    first K2 location channels land in `loc_out`.  Shapes the fused kernel does not take (fp32 features, odd
    pitches) go through the two separate launches."""
    require_device()
    n, h, w, C = x.shape
    assert flow.dtype == torch.float32 and flow.is_contiguous() and flow.shape == (n, h, w, 2)
    assert loc.is_contiguous() and loc_out.is_contiguous() and loc_out.shape == loc.shape and loc.dtype == torch.float32
    if out is None:
        out = torch.empty((n, h, w, C), device=x.device, dtype=x.dtype)
    if x.dtype == torch.bfloat16 and out.dtype == torch.bfloat16:
        xp, xpitch = _pix(x)
        op, opitch = _pix(out)
        rc = lib().vmgb_flow_warp_fused(xp, xpitch, flow.data_ptr(), op, opitch, loc.data_ptr(), loc.shape[-1], loc_out.data_ptr(),
                                        loc_out.shape[-1], K2, n, h, w, C, semantics, _stream())
        if rc == 0:
            return out
        if rc != -2:
            _ck(rc, 'vmgb_flow_warp_fused')
    flow_warp_bilinear(x, flow, out=out, semantics=semantics)
    flow_warp_nearest(loc, flow, K2, loc_out, semantics=semantics)
    return out


def loc_identity_(loc, c0):
    require_device()
    n, h, w, pitch = loc.shape
    _ck(lib().vmgb_loc_identity(loc.data_ptr(), pitch, c0, n, h, w, _stream()), 'vmgb_loc_identity')
    return loc


def flow_smooth(flow, r):
    require_device()
    n, h, w, _ = flow.shape
    out = torch.empty_like(flow)
    _ck(lib().vmgb_flow_smooth(flow.data_ptr(), out.data_ptr(), n, h, w, r, _stream()), 'vmgb_flow_smooth')
    return out


def ltam_wins(q, keys0, vals0, kv_stride, K, loc, rel, decay, heads, out_bias=None, anchor=None):
    """keys0 / vals0: tensors whose data_ptr is key-frame 0; key-frame k lives kv_stride elements later.
    out_bias / anchor: see include/vmgb.h (values stored already projected)."""
    require_device()
    n, h, w, C = q.shape
    assert q.is_contiguous() and keys0.dtype == q.dtype and vals0.dtype == q.dtype
    assert anchor is None or (anchor.is_contiguous() and anchor.shape == q.shape and anchor.dtype == q.dtype)
    assert out_bias is None or (out_bias.dtype == torch.float32 and out_bias.is_contiguous() and out_bias.numel() == C)
    out = torch.empty_like(q)
    _ck(lib().vmgb_ltam_wins(q.data_ptr(), keys0.data_ptr(), vals0.data_ptr(), _dt(q), kv_stride, loc.data_ptr(),
                             loc.shape[-1], rel.data_ptr(), decay.data_ptr(), out.data_ptr(), _dt(out), n, h, w, C,
                             heads, K, _ptr(out_bias), _ptr(anchor), _stream()), 'vmgb_ltam_wins')
    return out


# ---------------------------------------------------------------------------------------------
# SPyNet helpers, input / output
# ---------------------------------------------------------------------------------------------
def spynet_prep(frames, mean, std, h_up, w_up):
    require_device()
    n, H, W, _ = frames.shape
    out = torch.empty((n, h_up, w_up, 3), device=frames.device, dtype=torch.float32)
    _ck(lib().vmgb_spynet_prep(frames.data_ptr(), out.data_ptr(), *mean, *std, n, H, W, h_up, w_up, _stream()),
        'vmgb_spynet_prep')
    return out


def avgpool2(x):
    require_device()
    n, H, W, C = x.shape
    out = torch.empty((n, H // 2, W // 2, C), device=x.device, dtype=torch.float32)
    _ck(lib().vmgb_avgpool2(x.data_ptr(), out.data_ptr(), n, H, W, C, _stream()), 'vmgb_avgpool2')
    return out


def spynet_level_input(pyr, flow_coarse, n_pairs, B, direction, out_dtype=torch.float32, semantics=COORD_ATEN_CUDA):
    require_device()
    _, h, w, _ = pyr.shape
    out8 = torch.empty((n_pairs, h, w, 8), device=pyr.device, dtype=out_dtype)
    fup = torch.empty((n_pairs, h, w, 2), device=pyr.device, dtype=torch.float32)
    _ck(lib().vmgb_spynet_level_input(pyr.data_ptr(), _ptr(flow_coarse), out8.data_ptr(), _dt(out8), fup.data_ptr(), n_pairs, B,
                                      direction, h, w, semantics, _stream()), 'vmgb_spynet_level_input')
    return out8, fup


def flow_resize(flow, h, w):
    require_device()
    n, hu, wu, _ = flow.shape
    out = torch.empty((n, h, w, 2), device=flow.device, dtype=torch.float32)
    _ck(lib().vmgb_flow_resize(flow.data_ptr(), out.data_ptr(), n, hu, wu, h, w, _stream()), 'vmgb_flow_resize')
    return out


def pad_input(x, Hp, Wp):
    """x [B,T,3,H,W] fp32 -> time-major channels-last [T*B,Hp,Wp,3]."""
    require_device()
    B, T, _, H, W = x.shape
    out = torch.empty((T * B, Hp, Wp, 3), device=x.device, dtype=torch.float32)
    _ck(lib().vmgb_pad_input(x.data_ptr(), out.data_ptr(), B, T, H, W, Hp, Wp, _stream()), 'vmgb_pad_input')
    return out


def mirror_check(x):
    """vmg.py:426-432 -- True iff the clip is a palindrome (host sync, even T only)."""
    require_device()
    B, T = x.shape[:2]
    if T % 2:
        return False
    s = torch.zeros(1, device=x.device, dtype=torch.float32)
    _ck(lib().vmgb_mirror_check(x.data_ptr(), s.data_ptr(), B, T, x[0, 0].numel(), _stream()), 'vmgb_mirror_check')
    return bool(s.item() == 0.0)


def final_output(y, x, out, t0):
    """y [Tc*B,4H,4W,3] (time-major chunk starting at frame t0) ; x [B,T,3,H,W] ; out [B,T,3,4H,4W]."""
    require_device()
    B, T, _, H, W = x.shape
    Tc = y.shape[0] // B
    yp, ypitch = _pix(y)
    _ck(lib().vmgb_final_output(yp, _dt(y), ypitch, x.data_ptr(), out.data_ptr(), B, T, t0, Tc, H, W, _stream()),
        'vmgb_final_output')
    return out


# ---------------------------------------------------------------------------------------------
# callers either side of the forward: frame I/O, blending, augmentation, metrics
# ---------------------------------------------------------------------------------------------
def frames_u8_to_f32(frames_u8, reverse_channels=False):
    """uint8 [T,H,W,3] (device) -> fp32 [1,T,3,H,W] in [0,1] (Tester.evaluate: astype(float32)/255, permute, unsqueeze)."""
    require_device()
    assert frames_u8.dtype == torch.uint8 and frames_u8.is_contiguous() and frames_u8.dim() == 4 and frames_u8.shape[-1] == 3
    T, H, W, _ = frames_u8.shape
    out = torch.empty((1, T, 3, H, W), device=frames_u8.device, dtype=torch.float32)
    _ck(lib().vmgb_frames_u8_to_f32(frames_u8.data_ptr(), out.data_ptr(), T, H, W, int(reverse_channels), _stream()), 'vmgb_frames_u8_to_f32')
    return out


def frames_f32_to_u8(frames, reverse_channels=False, out=None):
    """fp32 [T,3,H,W] (or [1,T,3,H,W]) -> uint8 [T,H,W,3]: clamp, *255, round half to even (Tester.py:249-250)."""
    require_device()
    if frames.dim() == 5:
        assert frames.shape[0] == 1
        frames = frames[0]
    assert frames.dtype == torch.float32 and frames.is_contiguous() and frames.shape[1] == 3
    T, _, H, W = frames.shape
    if out is None:
        out = torch.empty((T, H, W, 3), device=frames.device, dtype=torch.uint8)
    _ck(lib().vmgb_frames_f32_to_u8(frames.data_ptr(), out.data_ptr(), T, H, W, int(reverse_channels), _stream()), 'vmgb_frames_f32_to_u8')
    return out


def blend_accumulate_(E, count, patch, t0, h0, w0, zero_t=(0, 0), zero_h=(0, 0), zero_w=(0, 0)):
    require_device()
    B, Tp, C, ph, pw = patch.shape
    _, Te, _, He, We = E.shape
    assert patch.dtype == torch.float32 and E.dtype == torch.float32 and patch.is_contiguous() and E.is_contiguous()
    assert E.shape[0] == B and E.shape[2] == C
    assert count is None or (count.dtype == torch.float32 and count.is_contiguous() and tuple(count.shape) == (Te, He, We))
    _ck(lib().vmgb_blend_accumulate(patch.data_ptr(), E.data_ptr(), _ptr(count), B, Tp, C, ph, pw, Te, He, We, t0, h0, w0,
                                    zero_t[0], zero_t[1], zero_h[0], zero_h[1], zero_w[0], zero_w[1], _stream()), 'vmgb_blend_accumulate')
    return E


def blend_normalize_(E, count):
    require_device()
    B, Te, C, He, We = E.shape
    assert tuple(count.shape) == (Te, He, We) and count.dtype == torch.float32 and count.is_contiguous() and E.is_contiguous()
    _ck(lib().vmgb_blend_normalize(E.data_ptr(), count.data_ptr(), B, Te, C, He, We, _stream()), 'vmgb_blend_normalize')
    return E


def blend_finalize_u8(E, count=None, reverse_channels=False):
    require_device()
    B, T, C, H, W = E.shape
    assert B == 1 and C == 3 and E.dtype == torch.float32 and E.is_contiguous()
    assert count is None or (tuple(count.shape) == (T, H, W) and count.dtype == torch.float32 and count.is_contiguous())
    out = torch.empty((T, H, W, 3), device=E.device, dtype=torch.uint8)
    _ck(lib().vmgb_blend_finalize_u8(E.data_ptr(), _ptr(count), out.data_ptr(), T, H, W, int(reverse_channels), _stream()),
        'vmgb_blend_finalize_u8')
    return out


def augment(x, hflip, vflip, transpose):
    """[..., H, W] fp32 -> flips / transpose of Tester.augment (applied in the reference's order)."""
    require_device()
    assert x.dtype == torch.float32 and x.is_contiguous()
    H, W = x.shape[-2:]
    planes = x.numel() // (H * W)
    out = torch.empty(tuple(x.shape[:-2]) + ((W, H) if transpose else (H, W)), device=x.device, dtype=torch.float32)
    _ck(lib().vmgb_augment(x.data_ptr(), out.data_ptr(), planes, H, W, int(hflip), int(vflip), int(transpose), _stream()), 'vmgb_augment')
    return out


def _scratch(n_img, H, W, n_planes, device):
    n = int(lib().vmgb_metrics_scratch_doubles(n_img, H, W, n_planes))
    return torch.empty(n, device=device, dtype=torch.float64)


def sse_u8(a, b):
    """Exact integer sum of squared differences of two uint8 tensors (python int; host sync)."""
    require_device()
    assert a.dtype == torch.uint8 and b.dtype == torch.uint8 and a.shape == b.shape and a.is_contiguous() and b.is_contiguous()
    acc = torch.zeros(1, device=a.device, dtype=torch.int64)
    _ck(lib().vmgb_sse_u8(a.data_ptr(), b.data_ptr(), a.numel(), acc.data_ptr(), _stream()), 'vmgb_sse_u8')
    return int(acc.item())


def sse_f32(a, b, clamp01=False):
    require_device()
    assert a.dtype == torch.float32 and b.dtype == torch.float32 and a.shape == b.shape and a.is_contiguous() and b.is_contiguous()
    out = torch.empty(1, device=a.device, dtype=torch.float64)
    _ck(lib().vmgb_sse_f32(a.data_ptr(), b.data_ptr(), a.numel(), int(clamp01), _scratch(1, 11, 11, 1, a.device).data_ptr(), out.data_ptr(),
                           _stream()), 'vmgb_sse_f32')
    return out


def sse_f64(a, b):
    require_device()
    assert a.dtype == torch.float64 and b.dtype == torch.float64 and a.shape == b.shape and a.is_contiguous() and b.is_contiguous()
    out = torch.empty(1, device=a.device, dtype=torch.float64)
    _ck(lib().vmgb_sse_f64(a.data_ptr(), b.data_ptr(), a.numel(), _scratch(1, 11, 11, 1, a.device).data_ptr(), out.data_ptr(), _stream()),
        'vmgb_sse_f64')
    return out


def rgb_to_y(rgb_u8):
    """uint8 [..., 3] RGB -> float64 [...] luma of skimage.color.rgb2ycbcr."""
    require_device()
    assert rgb_u8.dtype == torch.uint8 and rgb_u8.is_contiguous() and rgb_u8.shape[-1] == 3
    y = torch.empty(rgb_u8.shape[:-1], device=rgb_u8.device, dtype=torch.float64)
    _ck(lib().vmgb_rgb_to_y(rgb_u8.data_ptr(), y.data_ptr(), y.numel(), _stream()), 'vmgb_rgb_to_y')
    return y


def ssim(a, b, window11):
    """a, b: uint8 [n,H,W,3] (mean over channels) or float64 [n,H,W]; window11: 11 python floats.  -> fp64 [n] (device)."""
    require_device()
    assert a.shape == b.shape and a.dtype == b.dtype and a.is_contiguous() and b.is_contiguous()
    if a.dtype == torch.uint8:
        n, H, W, P = a.shape
        dt = 0
    else:
        assert a.dtype == torch.float64
        n, H, W = a.shape
        P, dt = 1, 1
    win = (ctypes.c_double * 11)(*[float(v) for v in window11])
    out = torch.empty(n, device=a.device, dtype=torch.float64)
    _ck(lib().vmgb_ssim(a.data_ptr(), b.data_ptr(), dt, n, H, W, P, ctypes.cast(win, ctypes.c_void_p), _scratch(n, H, W, P, a.device).data_ptr(),
                        out.data_ptr(), _stream()), 'vmgb_ssim')
    return out
