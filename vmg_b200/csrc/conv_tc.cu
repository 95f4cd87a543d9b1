// conv_tc.cu -- tcgen05 / TMEM / TMA implicit-GEMM convolution and linear (bf16 x bf16 -> fp32) with the
// fused epilogue menu.  One persistent, warp-specialised kernel:
//
//   warp 0 (1 lane)  TMA producer : per K block one 4-D box of the activation tensor (the shifted
//                                   128-pixel tile of one tap; TMA zero-fills the conv padding and the
//                                   channel tail) + one 2-D box of the packed weights -> smem ring
//   warp 1 (1 lane)  MMA issuer   : tcgen05.mma cta_group::1 kind::f16, M=128 x N=NT (16..256), K=16 per
//                                   instruction, SWIZZLE_128B K-major operands, fp32 accumulators in TMEM
//                                   (2 x NT columns, double buffered across tiles)
//   warp 2           TMEM alloc / dealloc
//   warps 4..7       epilogue     : tcgen05.ld 32x32b.x16 -> bias / act / scale / residual / gate /
//                                   pixel-shuffle -> 128-bit global stores
//
// GEMM view: M = pixels (tile = BH x BW = 128 pixels of one image), N = output channels of one group,
// K = taps x input channels.  A Linear is the R=S=1 case over a [1,1,rows,K] "image" (tile 1 x 128).
// Replaces the cuDNN / cuBLAS calls behind nn.Conv2d / nn.Linear on the VMG path.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <mutex>
#include <cstdlib>
#include <type_traits>

namespace vmgb {

struct TcP {
  int n_img, H, W, Cin_g, Cout_g, groups, R, S, n_src, Cs;
  int BH, BW, tiles_h, tiles_w, NT, nt_per_g, kb_per_src, stages;
  long long total_tiles;
  const float* bias; void* out; const void* res; const void* aux;
  int out_dt, out_pitch, res_dt, res_pitch, aux_dt, aux_pitch;
  int act; float alpha, out_scale; int epi; float res_scale; int ps;
  int b_stage_bytes;
  int nsub, ra, rb, a_unit_bytes, a_sbo;       // weight-stationary kernels: ra = activation ring depth, a_unit_bytes = unit stride
  int res_smem;                                // bytes of the residual staging buffer (0 = off)
  int out_smem;                                // bytes of the output staging tile [128][NT+8] bf16 (0 = direct stores)
  int aux_smem;                                // bytes of the gate-operand staging buffer (GATE epilogue; 0 = off)
  int wst;                                     // 1 = warp-private transposition staging epilogue (run_epilogue_w), 32 KB
  int v8;                                      // 1 = direct 32-byte-store epilogue (run_epilogue_v8: plain store, bf16 output)
  int res_tma;                                 // 1 = residual tile fetched by TMA one item ahead (run_epilogue_rtma, CTA-pair kernel)
  int zero;                                    // always 0, opaque to the compiler (forces data dependencies, see run_epilogue_rtma)
  int dbg;                                     // bottleneck probes (tools/microbench.py): 1 = no MMA, 2 = no drain, 4 = no activation loads
};

constexpr int A_STAGE_BYTES = 128 * 128;   // 128 pixels x 64 bf16
constexpr int TMEM_COLS = 512;
constexpr int ACC_STRIDE = 256;            // column offset of the second accumulator

// ---------------------------------------------------------------------------------------------
// Epilogue.  EPI_WARPS warps (4 or 8) drain one 128 x NT accumulator: warp w owns TMEM lanes 32*(w%4)..+31
// (its pixel rows) and the 16-column chunks ch = w/4, w/4 + EPI_WARPS/4, ...  Per tile a thread
//   1. prefetches the bf16 residual of all its chunks into registers BEFORE waiting for the accumulator
//      (the loads ride under the MMA main loop; the profile of the first version showed the epilogue warps
//      stalled on exactly these L2 round trips, one per chunk),
//   2. reads the bias from shared memory (staged once per (group, N tile)),
//   3. tcgen05.ld -> bias / activation / scale / residual / gate / pixel shuffle -> 16-byte stores.
// ---------------------------------------------------------------------------------------------

// exact-erf GELU with erf from Abramowitz & Stegun 7.1.28: erf(z) = 1 - (1 + a1 z + ... + a6 z^6)^-16, |error| <= 3e-7
// (far below bf16 resolution), z = |x|/sqrt(2).  One MUFU (rcp.approx) + ~14 FMA-class operations per element; the
// FFN fc1 epilogue evaluates ~4e9 GELUs per clip and is bound by exactly these two pipes (the 7.1.26 form needs a
// second MUFU for exp(-z^2), erff() is a branchy ~40 instructions).  A polynomial that overflows (|x| > ~25)
// gives rcp(inf) = 0, i.e. erf = 1, which is the correct limit.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float p = fmaf(z, 0.0000430638f, 0.0002765672f);
  p = fmaf(z, p, 0.0001520143f);
  p = fmaf(z, p, 0.0092705272f);
  p = fmaf(z, p, 0.0422820123f);
  p = fmaf(z, p, 0.0705230784f);
  p = fmaf(z, p, 1.f);
  p *= p; p *= p; p *= p; p *= p;                            // ^16
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p));
  const float erf_abs = 1.f - r;                             // erf(|x|/sqrt2)
  const float hx = 0.5f * x;
  return fmaf(fabsf(hx), erf_abs, hx);                      // 0.5*x*(1 + sign(x)*erf_abs) = hx + |hx|*erf_abs
}

// Packed form (Blackwell FFMA2 / FMUL2: one instruction per TWO fp32 lanes, same rounding as the scalar ops): the
// polynomial, its 16th power and the final scale run on pairs, only |x| (source modifier) and the MUFU stay scalar --
// ~9 instead of ~17 issue slots per element in the issue-bound FFN fc1 epilogue.  Bit-identical to gelu_erf_fast.
__device__ __forceinline__ unsigned long long pack_f2(float a, float b) {
  return (unsigned long long)__float_as_uint(a) | ((unsigned long long)__float_as_uint(b) << 32);
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ void gelu_erf_fast2(float& x0, float& x1) {
  const unsigned long long z = pack_f2(fabsf(x0) * 0.70710678118654752440f, fabsf(x1) * 0.70710678118654752440f);
  unsigned long long p = fma2(z, pack_f2(0.0000430638f, 0.0000430638f), pack_f2(0.0002765672f, 0.0002765672f));
  p = fma2(z, p, pack_f2(0.0001520143f, 0.0001520143f));
  p = fma2(z, p, pack_f2(0.0092705272f, 0.0092705272f));
  p = fma2(z, p, pack_f2(0.0422820123f, 0.0422820123f));
  p = fma2(z, p, pack_f2(0.0705230784f, 0.0705230784f));
  p = fma2(z, p, pack_f2(1.f, 1.f));
  p = mul2(p, p); p = mul2(p, p); p = mul2(p, p); p = mul2(p, p);                     // ^16
  float r0, r1;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(__uint_as_float((uint32_t)p)));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(__uint_as_float((uint32_t)(p >> 32))));
  const unsigned long long e = fma2(pack_f2(r0, r1), pack_f2(-1.f, -1.f), pack_f2(1.f, 1.f));      // erf(|x|/sqrt2) = 1 - r
  const unsigned long long h = mul2(pack_f2(x0, x1), pack_f2(0.5f, 0.5f));
  const float h0 = __uint_as_float((uint32_t)h), h1 = __uint_as_float((uint32_t)(h >> 32));
  x0 = fmaf(fabsf(h0), __uint_as_float((uint32_t)e), h0);
  x1 = fmaf(fabsf(h1), __uint_as_float((uint32_t)(e >> 32)), h1);
}

// The erf GELU (function.py:63 nn.GELU()) through ONE hardware tanh: erf(x / sqrt 2) = tanh(x * P(x^2)) with an odd
// minimax fit P (tools/fit_gelu.py: |gelu - 0.5 x (1 + tanh(x P))| <= 5.6e-5 over the reals) and MUFU.TANH (relative
// error 2^-11), i.e. an absolute error far below the bf16 rounding of the hidden tensor it feeds (2^-9 relative).
// 6 packed FP32x2 instructions + 2 MUFU per PAIR of elements -- half the issue slots of gelu_erf_fast2 in the
// issue-bound FFN fc1 epilogue.  The strict-fp32 path (conv_generic.cu) keeps erff.
__device__ __forceinline__ void gelu_tanh2(float& x0, float& x1) {
  const unsigned long long x = pack_f2(x0, x1);
  const unsigned long long x2 = mul2(x, x);
  unsigned long long p = fma2(x2, pack_f2(-3.82825920e-4f, -3.82825920e-4f), pack_f2(3.72235221e-2f, 3.72235221e-2f));
  p = fma2(x2, p, pack_f2(7.97223840e-1f, 7.97223840e-1f));
  const unsigned long long u = mul2(x, p);
  float t0, t1;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(__uint_as_float((uint32_t)u)));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(__uint_as_float((uint32_t)(u >> 32))));
  const unsigned long long h = mul2(x, pack_f2(0.5f, 0.5f));
  const unsigned long long y = fma2(h, pack_f2(t0, t1), h);
  x0 = __uint_as_float((uint32_t)y);
  x1 = __uint_as_float((uint32_t)(y >> 32));
}

// whether chunk `nb` (first column inside the group) of this pixel can use the vector bf16 residual prefetch
__device__ __forceinline__ bool pre_ok(const TcP& p, bool pvalid, int g, int nb) {
  return pvalid && p.epi != VMGB_EPI_STORE && p.res_dt == VMGB_BF16 && p.ps == 0 && nb + 16 <= p.Cout_g &&
         (((g * p.Cout_g + nb) | p.res_pitch) & 7) == 0;
}

// ACT / EPI / PS are compile-time for the combinations the model uses (-1 = read from the parameters): the fully
// generic body is ~2 000 instructions per chunk and thrashed the instruction cache (stall_no_inst); a
// specialised instance is ~150.
template <int ACT, int EPI, int PS, bool BIASED = false>
__device__ __forceinline__ void epilogue_chunk(const TcP& p, const uint32_t* raw, bool pvalid, int img, int hh, int ww,
                                               long long pix, int g, int nt, int ch, const float* bias_s, bool have_pre,
                                               const uint4& pre0, const uint4& pre1, uint32_t sdst0 = 0u, uint32_t sdst1 = 0u,
                                               bool have_aux = false, const uint4& ax0 = uint4{0, 0, 0, 0}, const uint4& ax1 = uint4{0, 0, 0, 0}) {
  const int act = ACT < 0 ? p.act : ACT;
  const int epi = EPI < 0 ? p.epi : EPI;
  const int ps = PS < 0 ? p.ps : PS;
  const int nb = nt * p.NT + ch * 16;            // first column of the chunk inside the group
  if (!pvalid || nb >= p.Cout_g) return;
  const int nvalid = min(16, p.Cout_g - nb);
  const int co0 = g * p.Cout_g + nb;
  float v[16];
  if (BIASED) {                                  // the caller added the bias while it moved the chunk out of the TMEM load registers
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]);
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 b4 = *reinterpret_cast<const float4*>(bias_s + ch * 16 + 4 * q);     // smem broadcast
      v[4 * q] = __uint_as_float(raw[4 * q]) + b4.x; v[4 * q + 1] = __uint_as_float(raw[4 * q + 1]) + b4.y;
      v[4 * q + 2] = __uint_as_float(raw[4 * q + 2]) + b4.z; v[4 * q + 3] = __uint_as_float(raw[4 * q + 3]) + b4.w;
    }
  }
  // one warp-uniform switch per chunk (a per-element switch compiles to 16 jump-table branches)
  switch (act) {
    case VMGB_ACT_RELU:
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
      break;
    case VMGB_ACT_LRELU:
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * p.alpha;
      break;
    case VMGB_ACT_GELU:
#pragma unroll
#ifdef VMGB_GELU_ERF      // A/B build only: ONE GELU flavour in the chunk body (two doubled it past the instruction cache: fc1 +40 %)
      for (int j = 0; j < 16; j += 2) gelu_erf_fast2(v[j], v[j + 1]);
#else
      for (int j = 0; j < 16; j += 2) gelu_tanh2(v[j], v[j + 1]);
#endif
      break;
    case VMGB_ACT_SIGMOID:
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 1.f / (1.f + __expf(-v[j]));
      break;
    default: break;
  }
  if (p.out_scale != 1.f) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] *= p.out_scale;
  }
  if (ps == 2) {
    // co = 4*c + 2*dy + dx  ->  out[(img, 2h+dy, 2w+dx), c]; 16 columns = 4 channels x 4 sub-pixels
    const int c0 = co0 >> 2;
#pragma unroll
    for (int sp = 0; sp < 4; ++sp) {
      const long long op = ((long long)img * (2 * p.H) + (2 * hh + (sp >> 1))) * (2 * p.W) + (2 * ww + (sp & 1));
      if (nvalid == 16) st4(p.out, p.out_dt, op * p.out_pitch + c0, make_float4(v[sp], v[4 + sp], v[8 + sp], v[12 + sp]));
      else
        for (int q = 0; q < 4; ++q) if (4 * q + sp < nvalid) stf(p.out, p.out_dt, op * p.out_pitch + c0 + q, v[4 * q + sp]);
    }
    return;
  }
  const long long ob = pix * p.out_pitch + co0;
  if (epi != VMGB_EPI_STORE) {
    const long long rb = pix * p.res_pitch + co0;
    float r[16], a[16];
    if (have_pre) {
      unpack8(pre0, r); unpack8(pre1, r + 8);
    } else if (nvalid == 16 && ((p.res_pitch | co0) & 3) == 0) {
#pragma unroll
      for (int q = 0; q < 4; ++q) { const float4 t = ld4(p.res, p.res_dt, rb + 4 * q); r[4 * q] = t.x; r[4 * q + 1] = t.y; r[4 * q + 2] = t.z; r[4 * q + 3] = t.w; }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = j < nvalid ? ldf(p.res, p.res_dt, rb + j) : 0.f;
    }
    if (epi == VMGB_EPI_GATE) {
      const long long ab = pix * p.aux_pitch + co0;
      if (have_aux) {
        unpack8(ax0, a); unpack8(ax1, a + 8);
      } else if (nvalid == 16 && ((p.aux_pitch | co0) & 3) == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) { const float4 t = ld4(p.aux, p.aux_dt, ab + 4 * q); a[4 * q] = t.x; a[4 * q + 1] = t.y; a[4 * q + 2] = t.z; a[4 * q + 3] = t.w; }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = j < nvalid ? ldf(p.aux, p.aux_dt, ab + j) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float th;
        asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(v[j]));
        v[j] = r[j] + p.res_scale * ((a[j] + v[j]) * th);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = r[j] + p.res_scale * v[j];
    }
  }
  if ((p.dbg & 8) && !sdst0) { if (v[0] + v[5] + v[10] + v[15] == 123.456f) stf(p.out, p.out_dt, ob, v[3]); return; }     // probe: no stores
  if (sdst0) {      // staged in shared memory (two 16-byte pieces), copied out with coalesced 16-byte stores after the drain
    const uint4 lo = pack8(v), hi = pack8(v + 8);
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sdst0), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w) : "memory");
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sdst1), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
  } else if (nvalid == 16 && p.out_dt == VMGB_BF16 && ((p.out_pitch | co0) & 7) == 0) {
    uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + ob);
    o[0] = pack8(v); o[1] = pack8(v + 8);
  } else if (nvalid == 16 && ((p.out_pitch | co0) & 3) == 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) st4(p.out, p.out_dt, ob + 4 * q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) if (j < nvalid) stf(p.out, p.out_dt, ob + j, v[j]);
  }
}

// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// Epilogue driver of the 384-thread kernels (8 epilogue warps).  `decode(t, row, combo, img, hh, ww, pvalid)`
// maps work item t and accumulator row to the output pixel.  Residual rows are staged global -> shared with
// cp.async ONE ITEM AHEAD (double buffer [chunk][128 rows][16 bf16], each thread touches only its own cells),
// so the L2 round trip of the residual read rides under the previous drain instead of stalling this one.
// ---------------------------------------------------------------------------------------------
template <int ACT, int EPI, int PS, class Decode>
__device__ __forceinline__ void run_epilogue_t(const TcP& p, float* bias_s, uint32_t tmem_base, uint32_t tfull0, uint32_t tempty0,
                                               long long first, long long step, long long total, int warp, int lane, Decode decode,
                                               uint32_t tempty_cluster0) {
  const int ew = warp - 4, quarter = warp & 3, half = ew >> 2;
  const int row = quarter * 32 + lane;
  const int nchunks = p.NT >> 4;
  uint32_t acc_phase0 = 0, acc_phase1 = 0;
  int it = 0, bias_combo = -1;
  const int epi = EPI < 0 ? p.epi : EPI;
  const int ps_ = PS < 0 ? p.ps : PS;
  const bool stage_res = p.res_smem > 0 && epi != VMGB_EPI_STORE;
  // without smem staging (weight-stationary kernel short of shared memory): same one-item-ahead idea in registers,
  // for up to 2 chunks per thread (NT <= 64)
  const bool reg_res = !stage_res && epi != VMGB_EPI_STORE && p.res_dt == VMGB_BF16 && ps_ == 0 && nchunks <= 4;
  uint4 cur[2][2], nxt[2][2];
  auto load_regs = [&](long long t, uint4 (&dst)[2][2]) {
#pragma unroll
    for (int i = 0; i < 2; ++i) { dst[i][0] = make_uint4(0, 0, 0, 0); dst[i][1] = dst[i][0]; }
    if (t < total) {
      int combo, img, hh, ww; bool pvalid;
      decode(t, row, combo, img, hh, ww, pvalid);
      const int nt = combo % p.nt_per_g, g = combo / p.nt_per_g;
      const long long pix = ((long long)img * p.H + hh) * p.W + ww;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int ch = half + 2 * i;
        if (ch < nchunks && pre_ok(p, pvalid, g, nt * p.NT + ch * 16)) {
          const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.res) + pix * p.res_pitch + g * p.Cout_g + nt * p.NT + ch * 16;
          dst[i][0] = ldg_nc_v4(rp); dst[i][1] = ldg_nc_v4(rp + 8);
        }
      }
    }
  };
  if (reg_res) load_regs(first, cur);
  const uint32_t res_s0 = smem_u32(bias_s) + 1024;          // ONE staging buffer of p.res_smem bytes after the bias
  const uint32_t out_s = p.out_smem > 0 ? res_s0 + p.res_smem : 0u;   // output staging tile after it
  const uint32_t aux_s0 = res_s0 + p.res_smem + p.out_smem;            // gate-operand staging after that
  const bool stage_aux = stage_res && p.aux_smem > 0 && epi == VMGB_EPI_GATE && p.aux_dt == VMGB_BF16 && p.aux_pitch == p.res_pitch;
  // cp.async of one 32-byte cell (chunk ch of this thread's row) of item t; every thread reads a cell of the
  // current item BEFORE it re-fills the same cell for the next item, so a single buffer is enough
  auto issue_cell = [&](bool ok, long long pix, int g, int nt, int ch) {
    if (ok) {
      const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.res) + pix * p.res_pitch + g * p.Cout_g + nt * p.NT + ch * 16;
      const uint32_t dst = res_s0 + (ch * 128 + row) * 32;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(rp) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16), "l"(rp + 8) : "memory");
      if (stage_aux) {      // same cell of the gate operand (same pitch and alignment as the residual)
        const __nv_bfloat16* ap = reinterpret_cast<const __nv_bfloat16*>(p.aux) + pix * p.aux_pitch + g * p.Cout_g + nt * p.NT + ch * 16;
        const uint32_t da = aux_s0 + (ch * 128 + row) * 32;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(da), "l"(ap) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(da + 16), "l"(ap + 8) : "memory");
      }
    }
  };
  if (stage_res && first < total) {
    int combo, img, hh, ww; bool pvalid;
    decode(first, row, combo, img, hh, ww, pvalid);
    const int nt = combo % p.nt_per_g, g = combo / p.nt_per_g;
    const long long pix = ((long long)img * p.H + hh) * p.W + ww;
    for (int ch = half; ch < nchunks; ch += 2) issue_cell(pre_ok(p, pvalid, g, nt * p.NT + ch * 16), pix, g, nt, ch);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (long long t = first; t < total; t += step, ++it) {
    const int as = it & 1;
    int combo, img, hh, ww; bool pvalid;
    decode(t, row, combo, img, hh, ww, pvalid);
    const int nt = combo % p.nt_per_g, g = combo / p.nt_per_g;
    const long long pix = ((long long)img * p.H + hh) * p.W + ww;
    // next item (for the residual refill)
    const bool has_next = stage_res && t + step < total;
    int ncombo_ = 0, nimg = 0, nhh = 0, nww = 0; bool npvalid = false;
    if (has_next) decode(t + step, row, ncombo_, nimg, nhh, nww, npvalid);
    const int nnt = ncombo_ % p.nt_per_g, ng = ncombo_ / p.nt_per_g;
    const long long npix = ((long long)nimg * p.H + nhh) * p.W + nww;
    const uint32_t aph = as ? acc_phase1 : acc_phase0;
    if (as) acc_phase1 ^= 1; else acc_phase0 ^= 1;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * ACC_STRIDE;
    if (stage_res) asm volatile("cp.async.wait_all;" ::: "memory");       // this item's residual (issued one item ago) has landed
    if (reg_res) load_regs(t + step, nxt);
    if (combo != bias_combo) {                                            // bias of this (group, N tile) -> smem
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int c = ew * 32 + lane; c < p.NT; c += 256) {
        const int nb = nt * p.NT + c;
        bias_s[c] = (p.bias && nb < p.Cout_g) ? __ldg(p.bias + g * p.Cout_g + nb) : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      bias_combo = combo;
    }
    mbar_wait(tfull0 + 8 * as, aph);
    tc_fence_after();
    if (p.dbg & 2) {
      tc_fence_before();
      if (tempty_cluster0) mbar_arrive_cluster(tempty_cluster0 + 8 * as); else mbar_arrive(tempty0 + 8 * as);
      continue;
    }
    // drain, software-pipelined over chunks: the tcgen05.ld of chunk c+2 is in flight while chunk c is processed
    auto process = [&](int ch, const uint32_t* raw) {
      uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0, a0 = r0, a1 = r0;
      const bool hp = (stage_res || reg_res) && pre_ok(p, pvalid, g, nt * p.NT + ch * 16);
      if (stage_res) {
        if (hp) {
          const uint32_t src = res_s0 + (ch * 128 + row) * 32;
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r0.x), "=r"(r0.y), "=r"(r0.z), "=r"(r0.w) : "r"(src) : "memory");
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r1.x), "=r"(r1.y), "=r"(r1.z), "=r"(r1.w) : "r"(src + 16) : "memory");
          if (stage_aux) {
            const uint32_t sa = aux_s0 + (ch * 128 + row) * 32;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a0.x), "=r"(a0.y), "=r"(a0.z), "=r"(a0.w) : "r"(sa) : "memory");
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a1.x), "=r"(a1.y), "=r"(a1.z), "=r"(a1.w) : "r"(sa + 16) : "memory");
          }
        }
        if (has_next) issue_cell(pre_ok(p, npvalid, ng, nnt * p.NT + ch * 16), npix, ng, nnt, ch);   // refill for the next item
      } else if (hp) {
        const bool second = ch >= half + 2;
        r0 = second ? cur[1][0] : cur[0][0]; r1 = second ? cur[1][1] : cur[0][1];
      }
      const uint32_t sd = out_s ? out_s + (uint32_t)(row * (p.NT + 8) + ch * 16) * 2u : 0u;      // staged tile [128 rows][NT + 8] bf16
      epilogue_chunk<ACT, EPI, PS>(p, raw, pvalid, img, hh, ww, pix, g, nt, ch, bias_s, hp, r0, r1, sd, sd + 16, hp && stage_aux, a0, a1);
    };
    {
      // one copy of the chunk body (the two-buffer ping-pong doubled a ~700-instruction body past the L0 instruction
      // cache: 16 % of the epilogue's stall samples were no_inst): the landed chunk moves to `work`, the next load
      // re-uses `raw` while `work` is processed
      uint32_t raw[16], work[16];
      int ch = half;
      __syncwarp();
      if (ch < nchunks) tmem_ld16_issue(taddr + ch * 16, raw);
#pragma unroll 1
      while (ch < nchunks) {
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) work[j] = raw[j];
        __syncwarp();
        if (ch + 2 < nchunks) tmem_ld16_issue(taddr + (ch + 2) * 16, raw);
        process(ch, work);
        ch += 2;
      }
    }
    if (stage_res) asm volatile("cp.async.commit_group;" ::: "memory");
    tc_fence_before();
    if (tempty_cluster0) mbar_arrive_cluster(tempty_cluster0 + 8 * as);     // CTA pair: the leader's MMA warp waits for both epilogues
    else mbar_arrive(tempty0 + 8 * as);
    if (out_s) {
      // cooperative copy-out: consecutive threads write consecutive 16-byte pieces of a pixel's channel run
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int vec_per_row = p.NT >> 3;
      const int tid = ew * 32 + lane;
      for (int idx = tid; idx < 128 * vec_per_row; idx += 256) {
        const int r = idx / vec_per_row, vcol = (idx - r * vec_per_row) * 8;
        int c2, i2, h2, w2; bool ok2;
        decode(t, r, c2, i2, h2, w2, ok2);
        if (ok2 && nt * p.NT + vcol + 8 <= p.Cout_g) {
          uint4 val;
          const uint32_t src = out_s + (uint32_t)(r * (p.NT + 8) + vcol) * 2u;
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w) : "r"(src) : "memory");
          const long long o = (((long long)i2 * p.H + h2) * p.W + w2) * p.out_pitch + g * p.Cout_g + nt * p.NT + vcol;
          *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + o) = val;
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");          // staging tile free for the next item
    }
    if (reg_res) {
#pragma unroll
      for (int i = 0; i < 2; ++i) { cur[i][0] = nxt[i][0]; cur[i][1] = nxt[i][1]; }
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
}

// one warp-uniform dispatch per launch onto the specialised epilogue instances
template <class Decode>
__device__ __forceinline__ void run_epilogue(const TcP& p, float* bias_s, uint32_t tmem_base, uint32_t tfull0, uint32_t tempty0,
                                             long long first, long long step, long long total, int warp, int lane, Decode decode,
                                             uint32_t tempty_cluster0 = 0) {
#define VMGB_EPI_CASE(A, E, P)                                                                                          \
  if (p.act == (A) && p.epi == (E) && p.ps == (P)) {                                                                   \
    run_epilogue_t<A, E, P>(p, bias_s, tmem_base, tfull0, tempty0, first, step, total, warp, lane, decode, tempty_cluster0); \
    return;                                                                                                            \
  }
  VMGB_EPI_CASE(VMGB_ACT_NONE, VMGB_EPI_RESADD, 0)     // resblock conv2, FFN fc2, LTAM proj
  VMGB_EPI_CASE(VMGB_ACT_RELU, VMGB_EPI_STORE, 0)      // resblock conv1, Morph-FC linears
  VMGB_EPI_CASE(VMGB_ACT_LRELU, VMGB_EPI_STORE, 0)     // trajectory input conv / fusion, HRconv
  VMGB_EPI_CASE(VMGB_ACT_GELU, VMGB_EPI_STORE, 0)      // FFN fc1
  VMGB_EPI_CASE(VMGB_ACT_NONE, VMGB_EPI_STORE, 0)      // resampling linears, RCAB conv2, conv_last
  VMGB_EPI_CASE(VMGB_ACT_NONE, VMGB_EPI_GATE, 0)       // TAB proj + gate
  VMGB_EPI_CASE(VMGB_ACT_LRELU, VMGB_EPI_STORE, 2)     // upconv + pixel shuffle
#undef VMGB_EPI_CASE
  run_epilogue_t<-1, -1, -1>(p, bias_s, tmem_base, tfull0, tempty0, first, step, total, warp, lane, decode, tempty_cluster0);
}

// ---------------------------------------------------------------------------------------------
// Epilogue with WARP-PRIVATE transposition staging (bf16 output, no pixel shuffle; STORE, or RESADD with NT <= 128).
// tcgen05.ld 32x32b hands every thread ONE accumulator row, so a direct epilogue reads the residual and writes
// the output as 32-byte pieces of 32 different pixels per warp instruction: 32 L1 wavefronts for 1 KB.  Measured
// on the trajectory resblock conv (probes in tools/microbench.py): 1.9 us (residual) + 1.3 us (stores) per
// 128 x 112 tile against 0.5 us for TMEM loads + math - the epilogue, not the tensor pipe, paced the kernel.
// Here each epilogue warp owns a 4 KB buffer = its 32 rows x 64 columns (one "group") of bf16, 16-byte pieces
// XOR-swizzled by (row & 7) so that both access patterns are bank-conflict free:
//   coalesced side : lane -> (row = lane/8 + 4i, piece = lane%8); a warp instruction moves 4 rows x 128 contiguous
//                    bytes (4 wavefronts per 512 B): cp.async of the NEXT item's residual, and the copy-out
//   row side       : lane -> its own row; reads the residual pieces of a 16-column chunk, writes the result in place
// No CTA-wide barrier is needed (__syncwarp only) and the buffer is 32 KB per CTA for any NT.
// `origin(t, combo, img, h0, w0, valid)` maps a work item to its tile; rows map to pixels by the power-of-two BW.
// ---------------------------------------------------------------------------------------------
template <int ACT, int EPI, class Origin>
__device__ __forceinline__ void run_epilogue_w_t(const TcP& p, float* bias_s, uint32_t tmem_base, uint32_t tfull0, uint32_t tempty0,
                                                 long long first, long long step, long long total, int warp, int lane, Origin origin,
                                                 uint32_t tempty_cluster0) {
  const int ew = warp - 4, quarter = warp & 3, half = ew >> 2;
  const int row = quarter * 32 + lane;
  const int nchunks = p.NT >> 4;
  // this warp's 16-column chunks: the two column halves split the accumulator evenly; a pass covers <= 4 chunks (64 columns)
  const int c_lo = half ? (nchunks + 1) >> 1 : 0, c_hi = half ? nchunks : (nchunks + 1) >> 1;
  const int epi = EPI < 0 ? p.epi : EPI;
  const bool has_res = epi != VMGB_EPI_STORE;          // host: bf16 residual, 16-byte aligned rows, NT <= 128 (one pass per warp)
  const bool has_aux = epi == VMGB_EPI_GATE;           // gate operand: a second buffer per warp, same layout, same pitch rules
  const uint32_t buf = smem_u32(bias_s) + 1024 + (uint32_t)ew * 4096u;
  const __nv_bfloat16* auxp = reinterpret_cast<const __nv_bfloat16*>(p.aux);
  const int bw_shift = 31 - __clz(p.BW), bw_mask = p.BW - 1;
  const int crow0 = lane >> 3, piece = lane & 7;
  const __nv_bfloat16* resp = reinterpret_cast<const __nv_bfloat16*>(p.res);
  __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(p.out);
  uint32_t acc_phase0 = 0, acc_phase1 = 0;
  int it = 0, bias_combo = -1;

  // coalesced side of one pass starting at chunk ch0: column of this lane's 16-byte piece and whether it exists
  auto prefetch_res = [&](long long t, int ch0, int chn) {
    int combo, img, h0, w0; bool tvalid;
    origin(t, combo, img, h0, w0, tvalid);
    const int nt = combo % p.nt_per_g, g = combo / p.nt_per_g;
    const int col = ch0 * 16 + piece * 8;
    const bool cok = tvalid && piece < 2 * chn && nt * p.NT + col + 8 <= p.Cout_g;
    const long long pbase = ((long long)img * p.H + h0) * p.W + w0;
    const __nv_bfloat16* rp0 = resp + g * p.Cout_g + nt * p.NT + col;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = crow0 + 4 * i, rr = quarter * 32 + r;
      const int dh = rr >> bw_shift, dw = rr & bw_mask;
      if (cok && h0 + dh < p.H && w0 + dw < p.W) {
        const __nv_bfloat16* rp = rp0 + (pbase + (long long)dh * p.W + dw) * p.res_pitch;
        const uint32_t dst = buf + (uint32_t)r * 128u + (uint32_t)((piece ^ (r & 7)) << 4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(rp) : "memory");
        if (has_aux) {
          const __nv_bfloat16* ap = auxp + g * p.Cout_g + nt * p.NT + col + (pbase + (long long)dh * p.W + dw) * p.aux_pitch;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 8u * 4096u), "l"(ap) : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  if (has_res && c_lo < c_hi && first < total) prefetch_res(first, c_lo, c_hi - c_lo);
  for (long long t = first; t < total; t += step, ++it) {
    const int as = it & 1;
    int combo, img, h0, w0; bool tvalid;
    origin(t, combo, img, h0, w0, tvalid);
    const int nt = combo % p.nt_per_g, g = combo / p.nt_per_g;
    const int hh = h0 + (row >> bw_shift), ww = w0 + (row & bw_mask);
    const bool pvalid = tvalid && hh < p.H && ww < p.W;
    const long long pbase = ((long long)img * p.H + h0) * p.W + w0;
    const long long pix = ((long long)img * p.H + hh) * p.W + ww;
    const uint32_t aph = as ? acc_phase1 : acc_phase0;
    if (as) acc_phase1 ^= 1; else acc_phase0 ^= 1;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * ACC_STRIDE;
    if (combo != bias_combo) {                                            // bias of this (group, N tile) -> smem
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int c = ew * 32 + lane; c < p.NT; c += 256) {
        const int nb = nt * p.NT + c;
        bias_s[c] = (p.bias && nb < p.Cout_g) ? __ldg(p.bias + g * p.Cout_g + nb) : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      bias_combo = combo;
    }
    if (has_res) { asm volatile("cp.async.wait_all;" ::: "memory"); __syncwarp(); }     // this item's residual pass has landed
    mbar_wait(tfull0 + 8 * as, aph);
    tc_fence_after();
    if (p.dbg & 2) {
      tc_fence_before();
      if (tempty_cluster0) mbar_arrive_cluster(tempty_cluster0 + 8 * as); else mbar_arrive(tempty0 + 8 * as);
      continue;
    }
    const uint32_t rowbase = buf + (uint32_t)lane * 128u;
    const uint32_t sw = (uint32_t)(lane & 7);
    __nv_bfloat16* op0 = outp + g * p.Cout_g + nt * p.NT;
    for (int ch0 = c_lo; ch0 < c_hi; ch0 += 4) {
      const int chn = min(4, c_hi - ch0);
      auto process = [&](int c, const uint32_t* raw) {
        const uint32_t c0 = rowbase + (((uint32_t)(2 * c) ^ sw) << 4), c1 = rowbase + (((uint32_t)(2 * c + 1) ^ sw) << 4);
        uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0, a0 = r0, a1 = r0;
        const bool hp = has_res && pvalid && nt * p.NT + (ch0 + c) * 16 + 16 <= p.Cout_g;
        if (hp) {
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r0.x), "=r"(r0.y), "=r"(r0.z), "=r"(r0.w) : "r"(c0) : "memory");
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r1.x), "=r"(r1.y), "=r"(r1.z), "=r"(r1.w) : "r"(c1) : "memory");
          if (has_aux) {
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a0.x), "=r"(a0.y), "=r"(a0.z), "=r"(a0.w) : "r"(c0 + 8u * 4096u) : "memory");
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a1.x), "=r"(a1.y), "=r"(a1.z), "=r"(a1.w) : "r"(c1 + 8u * 4096u) : "memory");
          }
        }
        epilogue_chunk<ACT, EPI, 0, true>(p, raw, pvalid, img, hh, ww, pix, g, nt, ch0 + c, bias_s, hp, r0, r1, c0, c1, hp && has_aux, a0, a1);
      };
      uint32_t raw[16], work[16];              // single chunk body, see run_epilogue_t
      __syncwarp();
      tmem_ld16_issue(taddr + ch0 * 16, raw);
#pragma unroll 1
      for (int c = 0; c < chn; ++c) {
        tmem_ld_wait();
        // the chunk leaves the TMEM load registers THROUGH the bias add (one FADD per element instead of a MOV plus the FADD
        // later), so the next tcgen05.ld can be issued before the activation math
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias_s + (ch0 + c) * 16 + 4 * q);
          work[4 * q] = __float_as_uint(__uint_as_float(raw[4 * q]) + b4.x); work[4 * q + 1] = __float_as_uint(__uint_as_float(raw[4 * q + 1]) + b4.y);
          work[4 * q + 2] = __float_as_uint(__uint_as_float(raw[4 * q + 2]) + b4.z); work[4 * q + 3] = __float_as_uint(__uint_as_float(raw[4 * q + 3]) + b4.w);
        }
        __syncwarp();
        if (c + 1 < chn) tmem_ld16_issue(taddr + (ch0 + c + 1) * 16, raw);
        process(c, work);
      }
      if (ch0 + 4 >= c_hi) {              // last TMEM read of this item: hand the accumulator back before the copy-out
        tc_fence_before();
        if (tempty_cluster0) mbar_arrive_cluster(tempty_cluster0 + 8 * as); else mbar_arrive(tempty0 + 8 * as);
      }
      __syncwarp();
      // coalesced copy-out of the pass: 4 rows x (up to) 128 contiguous bytes per warp instruction
      const int col = ch0 * 16 + piece * 8;
      const bool cok = tvalid && piece < 2 * chn && nt * p.NT + col + 8 <= p.Cout_g && !(p.dbg & 8);
      __nv_bfloat16* otile = op0 + pbase * p.out_pitch + col;      // 64-bit once per pass; the offsets inside a tile fit 32 bits
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = crow0 + 4 * i, rr = quarter * 32 + r;
        const int dh = rr >> bw_shift, dw = rr & bw_mask;
        if (cok && h0 + dh < p.H && w0 + dw < p.W) {
          uint4 val;
          const uint32_t src = buf + (uint32_t)r * 128u + (uint32_t)((piece ^ (r & 7)) << 4);
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w) : "r"(src) : "memory");
          *reinterpret_cast<uint4*>(otile + (dh * p.W + dw) * p.out_pitch) = val;
        }
      }
      __syncwarp();                       // buffer free: next pass / next item's residual
    }
    if (c_lo >= c_hi) {                   // idle column half (NT = 16): still a party of the accumulator hand-back
      tc_fence_before();
      if (tempty_cluster0) mbar_arrive_cluster(tempty_cluster0 + 8 * as); else mbar_arrive(tempty0 + 8 * as);
    }
    if (has_res && c_lo < c_hi && t + step < total) prefetch_res(t + step, c_lo, c_hi - c_lo);
  }
}

template <class Origin>
__device__ __forceinline__ void run_epilogue_w(const TcP& p, float* bias_s, uint32_t tmem_base, uint32_t tfull0, uint32_t tempty0,
                                               long long first, long long step, long long total, int warp, int lane, Origin origin,
                                               uint32_t tempty_cluster0 = 0) {
#define VMGB_EPIW_CASE(A, E)                                                                                             \
  if (p.act == (A) && p.epi == (E)) {                                                                                    \
    run_epilogue_w_t<A, E>(p, bias_s, tmem_base, tfull0, tempty0, first, step, total, warp, lane, origin, tempty_cluster0); \
    return;                                                                                                              \
  }
  VMGB_EPIW_CASE(VMGB_ACT_NONE, VMGB_EPI_RESADD)
  VMGB_EPIW_CASE(VMGB_ACT_NONE, VMGB_EPI_GATE)
  VMGB_EPIW_CASE(VMGB_ACT_RELU, VMGB_EPI_STORE)
  VMGB_EPIW_CASE(VMGB_ACT_LRELU, VMGB_EPI_STORE)
  VMGB_EPIW_CASE(VMGB_ACT_GELU, VMGB_EPI_STORE)
  VMGB_EPIW_CASE(VMGB_ACT_NONE, VMGB_EPI_STORE)
#undef VMGB_EPIW_CASE
  run_epilogue_w_t<-1, -1>(p, bias_s, tmem_base, tfull0, tempty0, first, step, total, warp, lane, origin, tempty_cluster0);
}

// ---------------------------------------------------------------------------------------------
// Plain-store epilogue with DIRECT 32-byte stores (bf16 output, no residual, no pixel shuffle).  tcgen05.ld 32x32b hands a
// thread 16 consecutive columns of ITS row = 32 contiguous output bytes, which sm_100 stores with one STG.256: no staging
// buffer, no copy-out pass, full 32-byte sectors.  The FFN fc1 launch (7.9 GB of hidden tensor, bound by its drain and
// behind it by the 3.94 TB/s write-only HBM rate) spent 3.8 of its 19.4 instructions per element in the staged copy-out
// and its address arithmetic.  32-byte alignment: the chunk grid starts at the first column whose global channel is a
// multiple of 16 (`head` = 0 or 8: grouped convs with Cout_g = 168 start odd groups 16 bytes off); the 8-column head /
// tail pieces go out as 16-byte stores.
// ---------------------------------------------------------------------------------------------
template <int ACT>
__device__ __forceinline__ void v8_act(const TcP& p, float* v, int n) {
  if (ACT == VMGB_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) if (j < n) v[j] = fmaxf(v[j], 0.f);
  } else if (ACT == VMGB_ACT_LRELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) if (j < n) v[j] = v[j] > 0.f ? v[j] : v[j] * p.alpha;
  } else if (ACT == VMGB_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 16; j += 2) if (j < n) gelu_tanh2(v[j], v[j + 1]);
  }
  if (p.out_scale != 1.f) {
#pragma unroll
    for (int j = 0; j < 16; ++j) if (j < n) v[j] *= p.out_scale;
  }
}

template <int ACT, class Origin>
__device__ __forceinline__ void run_epilogue_v8_t(const TcP& p, float* bias_s, uint32_t tmem_base, uint32_t tfull0, uint32_t tempty0,
                                                  long long first, long long step, long long total, int warp, int lane, Origin origin,
                                                  uint32_t tempty_cluster0) {
  const int ew = warp - 4, quarter = warp & 3, half = ew >> 2;
  const int row = quarter * 32 + lane;
  const int bw_shift = 31 - __clz(p.BW), bw_mask = p.BW - 1;
  __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(p.out);
  uint32_t acc_phase0 = 0, acc_phase1 = 0;
  int it = 0, bias_combo = -1;
  for (long long t = first; t < total; t += step, ++it) {
    const int as = it & 1;
    int combo, img, h0, w0; bool tvalid;
    origin(t, combo, img, h0, w0, tvalid);
    const int nt = combo % p.nt_per_g, g = combo / p.nt_per_g;
    const int hh = h0 + (row >> bw_shift), ww = w0 + (row & bw_mask);
    const bool pvalid = tvalid && hh < p.H && ww < p.W;
    const uint32_t aph = as ? acc_phase1 : acc_phase0;
    if (as) acc_phase1 ^= 1; else acc_phase0 ^= 1;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * ACC_STRIDE;
    if (combo != bias_combo) {                                            // bias of this (group, N tile) -> smem
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int c = ew * 32 + lane; c < p.NT; c += 256) {
        const int nb = nt * p.NT + c;
        bias_s[c] = (p.bias && nb < p.Cout_g) ? __ldg(p.bias + g * p.Cout_g + nb) : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      bias_combo = combo;
    }
    // column geometry: local column 0 is global channel cbase; 32-byte pieces start where the global channel is a multiple of 16
    const int cbase = g * p.Cout_g + nt * p.NT;
    const int ncols = min(p.NT, p.Cout_g - nt * p.NT);
    const int head = (16 - (cbase & 15)) & 15;                            // 0 or 8 (host: Cout_g % 8 == 0)
    const int nfull = (ncols - head) >> 4, tail = ncols - head - 16 * nfull;
    const int n0 = (nfull + 1) >> 1;
    const int lo = half ? n0 : 0, hi = half ? nfull : n0;
    __nv_bfloat16* orow = outp + (((long long)img * p.H + hh) * p.W + ww) * p.out_pitch + cbase;
    mbar_wait(tfull0 + 8 * as, aph);
    tc_fence_after();
    if (!(p.dbg & 2)) {
      uint32_t raw[16];
      float v[16];
      int k = lo;
      __syncwarp();
      if (k < hi) tmem_ld16_issue(taddr + head + k * 16, raw);
#pragma unroll 1
      while (k < hi) {
        tmem_ld_wait();
        const float* bk = bias_s + head + k * 16;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 b4 = *reinterpret_cast<const float4*>(bk + 4 * q);
          v[4 * q] = __uint_as_float(raw[4 * q]) + b4.x; v[4 * q + 1] = __uint_as_float(raw[4 * q + 1]) + b4.y;
          v[4 * q + 2] = __uint_as_float(raw[4 * q + 2]) + b4.z; v[4 * q + 3] = __uint_as_float(raw[4 * q + 3]) + b4.w;
        }
        __syncwarp();
        if (k + 1 < hi) tmem_ld16_issue(taddr + head + (k + 1) * 16, raw);      // next chunk in flight during the activation math
        v8_act<ACT>(p, v, 16);
        if (pvalid) stg_v8(orow + head + k * 16, pack8(v), pack8(v + 8));
        ++k;
      }
      const int piece = half ? tail : head;                               // the 8-column piece of this column half (16-byte store)
      if (piece) {
        const int col = half ? head + 16 * nfull : 0;
        uint32_t r8[8];
        __syncwarp();
        tmem_ld8_issue(taddr + col, r8);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r8[j]) + bias_s[col + j];
        v8_act<ACT>(p, v, 8);
        if (pvalid) *reinterpret_cast<uint4*>(orow + col) = pack8(v);
      }
    }
    tc_fence_before();
    if (tempty_cluster0) mbar_arrive_cluster(tempty_cluster0 + 8 * as); else mbar_arrive(tempty0 + 8 * as);
  }
}

template <class Origin>
__device__ __forceinline__ void run_epilogue_v8(const TcP& p, float* bias_s, uint32_t tmem_base, uint32_t tfull0, uint32_t tempty0,
                                                long long first, long long step, long long total, int warp, int lane, Origin origin,
                                                uint32_t tempty_cluster0 = 0) {
  switch (p.act) {
    case VMGB_ACT_GELU: run_epilogue_v8_t<VMGB_ACT_GELU>(p, bias_s, tmem_base, tfull0, tempty0, first, step, total, warp, lane, origin, tempty_cluster0); break;
    case VMGB_ACT_RELU: run_epilogue_v8_t<VMGB_ACT_RELU>(p, bias_s, tmem_base, tfull0, tempty0, first, step, total, warp, lane, origin, tempty_cluster0); break;
    case VMGB_ACT_LRELU: run_epilogue_v8_t<VMGB_ACT_LRELU>(p, bias_s, tmem_base, tfull0, tempty0, first, step, total, warp, lane, origin, tempty_cluster0); break;
    default: run_epilogue_v8_t<VMGB_ACT_NONE>(p, bias_s, tmem_base, tfull0, tempty0, first, step, total, warp, lane, origin, tempty_cluster0); break;
  }
}

// ---------------------------------------------------------------------------------------------
// Residual-add epilogue with the residual tile fetched by TMA (CTA-pair kernel, resblock conv2: act = none, bf16 in/out).
// The per-thread cp.async staging of run_epilogue_t costs ~1.2 us of LSU / address work per 128 x 112 tile - as much as
// the MMAs of the tile.  Here one elected thread issues a single 4-D box {NT channels, 8 cols, 16 rows} of the residual
// tensor per tile (row r of the accumulator = row r of the box, NT*2 bytes apart, TMA zero-fills the image overhang), one
// item ahead: the box of item t+1 is requested as soon as every epilogue thread has read item t and lands while the
// MMAs of t+1 run.  Epilogue threads only do ld.shared + FMA + store.
// ---------------------------------------------------------------------------------------------
template <class Origin>
__device__ __forceinline__ void run_epilogue_rtma(const TcP& p, float* bias_s, uint32_t tmem_base, uint32_t tfull0, uint32_t tempty0,
                                                  uint32_t rfull, const CUtensorMap* mapR, long long first, long long step, long long total,
                                                  int warp, int lane, Origin origin, uint32_t tempty_cluster0) {
  const int ew = warp - 4, quarter = warp & 3, half = ew >> 2;
  const int row = quarter * 32 + lane;
  const int nchunks = p.NT >> 4;
  const uint32_t res_s = (smem_u32(bias_s) + 1024u + 127u) & ~127u;
  const uint32_t row_bytes = (uint32_t)p.NT * 2u;
  const int bw_shift = 31 - __clz(p.BW), bw_mask = p.BW - 1;
  uint32_t acc_phase0 = 0, acc_phase1 = 0, rphase = 0;
  int it = 0, bias_combo = -1;
  const bool issuer_warp = ew == 0;
  auto request = [&](long long t) {             // whole warp 4 calls, one lane issues
    int combo, img, h0, w0; bool tvalid;
    origin(t, combo, img, h0, w0, tvalid);
    const int nt = combo % p.nt_per_g, g = combo / p.nt_per_g;
    if (elect_one()) {
      mbar_expect_tx(rfull, 128u * row_bytes);
      tma_load_4d(res_s, mapR, rfull, g * p.Cout_g + nt * p.NT, w0, h0, tvalid ? img : p.n_img);
    }
    __syncwarp();
  };
  if (issuer_warp && first < total) request(first);
  for (long long t = first; t < total; t += step, ++it) {
    const int as = it & 1;
    int combo, img, h0, w0; bool tvalid;
    origin(t, combo, img, h0, w0, tvalid);
    const int nt = combo % p.nt_per_g, g = combo / p.nt_per_g;
    const int hh = h0 + (row >> bw_shift), ww = w0 + (row & bw_mask);
    const bool pvalid = tvalid && hh < p.H && ww < p.W;
    const long long pix = ((long long)img * p.H + hh) * p.W + ww;
    const uint32_t aph = as ? acc_phase1 : acc_phase0;
    if (as) acc_phase1 ^= 1; else acc_phase0 ^= 1;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * ACC_STRIDE;
    if (combo != bias_combo) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int c = ew * 32 + lane; c < p.NT; c += 256) {
        const int nb = nt * p.NT + c;
        bias_s[c] = (p.bias && nb < p.Cout_g) ? __ldg(p.bias + g * p.Cout_g + nb) : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      bias_combo = combo;
    }
    mbar_wait(rfull, rphase);                    // residual box of this item
    rphase ^= 1;
    // the thread's residual pieces move to registers BEFORE the accumulator is awaited, so that the box of the next item
    // can be requested now and lands while this item is drained (host: NT <= 128, i.e. <= 4 chunks per thread)
    const uint32_t rrow = res_s + (uint32_t)row * row_bytes;
    uint4 rr[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ch = half + 2 * i;
      rr[i][0] = make_uint4(0, 0, 0, 0); rr[i][1] = rr[i][0];
      if (ch < nchunks) {
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rr[i][0].x), "=r"(rr[i][0].y), "=r"(rr[i][0].z), "=r"(rr[i][0].w) : "r"(rrow + ch * 32) : "memory");
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rr[i][1].x), "=r"(rr[i][1].y), "=r"(rr[i][1].z), "=r"(rr[i][1].w) : "r"(rrow + ch * 32 + 16) : "memory");
      }
    }
    // the loads above must have RETURNED before the box is re-armed: a barrier orders their issue, not their completion, and
    // with a power-of-two row pitch (NT = 64 / 128: 32-way bank conflicts) they were still queued when the next box landed.
    // A data-dependent (never taken) branch makes every thread wait for its registers first.
    if ((or_all8(rr) & (uint32_t)p.zero) != 0u) __trap();
    asm volatile("bar.sync 1, 256;" ::: "memory");      // every thread holds its residual: the buffer can be refilled
    if (issuer_warp && t + step < total) request(t + step);
    mbar_wait(tfull0 + 8 * as, aph);
    tc_fence_after();
    if (!(p.dbg & 2)) {
      uint32_t raw[16], work[16];
      int ch = half, i = 0;
      __syncwarp();
      if (ch < nchunks) tmem_ld16_issue(taddr + ch * 16, raw);
#pragma unroll 1
      while (ch < nchunks) {
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) work[j] = raw[j];
        __syncwarp();
        if (ch + 2 < nchunks) tmem_ld16_issue(taddr + (ch + 2) * 16, raw);
        const uint4 r0 = i == 0 ? rr[0][0] : (i == 1 ? rr[1][0] : (i == 2 ? rr[2][0] : rr[3][0]));
        const uint4 r1 = i == 0 ? rr[0][1] : (i == 1 ? rr[1][1] : (i == 2 ? rr[2][1] : rr[3][1]));
        const bool hp = pvalid && nt * p.NT + ch * 16 + 16 <= p.Cout_g;
        epilogue_chunk<VMGB_ACT_NONE, VMGB_EPI_RESADD, 0>(p, work, pvalid, img, hh, ww, pix, g, nt, ch, bias_s, hp, r0, r1);
        ch += 2; ++i;
      }
    }
    tc_fence_before();
    if (tempty_cluster0) mbar_arrive_cluster(tempty_cluster0 + 8 * as); else mbar_arrive(tempty0 + 8 * as);
  }
}

constexpr int TC_THREADS = 384;      // warps 0-3: producer / MMA / TMEM alloc / spare ; warps 4-11: epilogue
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB, const TcP p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + p.stages * A_STAGE_BYTES;
  const uint32_t bar_base = b_base + p.stages * p.b_stage_bytes;
  // barriers: full[stages], empty[stages], tmem_full[2], tmem_empty[2]; then the TMEM base address slot
  const uint32_t full0 = bar_base, empty0 = bar_base + 8 * p.stages;
  const uint32_t tfull0 = bar_base + 16 * p.stages, tempty0 = tfull0 + 16;
  const uint32_t tmem_slot = tempty0 + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  float* bias_s = reinterpret_cast<float*>(smem_raw + (tmem_slot + 16 - smem_u32(smem_raw)));   // NT floats after the barriers

  const int taps = p.R * p.S;
  const int kblocks = taps * p.n_src * p.kb_per_src;
  const int tiles_per_img = p.tiles_h * p.tiles_w;
  const int ncombo = p.groups * p.nt_per_g;

  if (warp == 0) {
    // ===================== TMA producer (whole warp runs the loops, one elected lane issues) =====================
    int stage = 0; uint32_t phase = 0;
    const uint32_t tx_bytes = A_STAGE_BYTES + p.NT * 128;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int combo = (int)(tile % ncombo); const long long mt = tile / ncombo;
      const int nt = combo % p.nt_per_g, g = combo / p.nt_per_g;
      const int img = (int)(mt / tiles_per_img); const int rem = (int)(mt % tiles_per_img);
      const int h0 = (rem / p.tiles_w) * p.BH, w0 = (rem % p.tiles_w) * p.BW;
      for (int tap = 0; tap < taps; ++tap) {
        const int dh = tap / p.S - p.R / 2, dw = tap % p.S - p.S / 2;
        const int brow = (g * taps + tap) * p.Cout_g + nt * p.NT;
        for (int src = 0; src < p.n_src; ++src) {
          const CUtensorMap* mA = src == 0 ? &mapA0 : (src == 1 ? &mapA1 : &mapA2);
          for (int kb = 0; kb < p.kb_per_src; ++kb) {
            mbar_wait(empty0 + 8 * stage, phase ^ 1);
            // grouped conv: start the box at a 16-byte aligned channel; the packed weights of the group are
            // shifted by the same `koff` rows of zeros (vmgb_pack_weight_tc)
            const int ca = (p.n_src > 1 ? 0 : ((g * p.Cin_g) & ~7)) + kb * 64;
            if (elect_one()) {
              mbar_expect_tx(full0 + 8 * stage, tx_bytes);
              tma_load_4d(a_base + stage * A_STAGE_BYTES, mA, full0 + 8 * stage, ca, w0 + dw, h0 + dh, img);
              tma_load_2d(b_base + stage * p.b_stage_bytes, &mapB, full0 + 8 * stage, src * p.Cs + kb * 64, brow);
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp runs the loops, one elected lane issues) =====================
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int stage = 0; uint32_t phase = 0;
    uint32_t acc_phase[2] = {0, 0};
    int it = 0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      mbar_wait(tempty0 + 8 * as, acc_phase[as] ^ 1);
      acc_phase[as] ^= 1;
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * ACC_STRIDE;
      uint32_t accumulate = 0;
      const int g_mma = (int)(tile % ncombo) / p.nt_per_g;
      const int kext = p.Cs + (p.n_src > 1 ? 0 : ((g_mma * p.Cin_g) & 7));     // K extent incl. the alignment shift
      for (int kbi = 0; kbi < kblocks; ++kbi) {
        const int kb = kbi % p.kb_per_src;
        const int nks = min(4, (kext - kb * 64 + 15) >> 4);
        mbar_wait(full0 + 8 * stage, phase);
        tc_fence_after();
        const uint64_t ad = smem_desc_sw128(a_base + stage * A_STAGE_BYTES);
        const uint64_t bd = smem_desc_sw128(b_base + stage * p.b_stage_bytes);
        if (elect_one()) {
          for (int ks = 0; ks < nks; ++ks)      // +32 B per K step; the very first K step of the tile overwrites
            umma_f16(d_tmem, ad + (uint64_t)(ks * 2), bd + (uint64_t)(ks * 2), idesc, accumulate | (uint32_t)(ks > 0));
          umma_commit(empty0 + 8 * stage);
          if (kbi == kblocks - 1) umma_commit(tfull0 + 8 * as);
        }
        __syncwarp();
        accumulate = 1;
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps) =====================
    if (p.wst)
      run_epilogue_w(p, bias_s, tmem_base, tfull0, tempty0, (long long)blockIdx.x, (long long)gridDim.x, p.total_tiles, warp, lane,
                     [&](long long tile, int& combo, int& img, int& h0, int& w0, bool& valid) {
                       combo = (int)(tile % ncombo);
                       const long long mt = tile / ncombo;
                       img = (int)(mt / tiles_per_img);
                       const int rem = (int)(mt % tiles_per_img);
                       h0 = (rem / p.tiles_w) * p.BH; w0 = (rem % p.tiles_w) * p.BW;
                       valid = true;
                     });
    else
    run_epilogue(p, bias_s, tmem_base, tfull0, tempty0, (long long)blockIdx.x, (long long)gridDim.x, p.total_tiles, warp, lane,
                 [&](long long tile, int row, int& combo, int& img, int& hh, int& ww, bool& pvalid) {
                   combo = (int)(tile % ncombo);
                   const long long mt = tile / ncombo;
                   img = (int)(mt / tiles_per_img);
                   const int rem = (int)(mt % tiles_per_img);
                   hh = (rem / p.tiles_w) * p.BH + row / p.BW;
                   ww = (rem % p.tiles_w) * p.BW + row % p.BW;
                   pvalid = hh < p.H && ww < p.W;
                 });
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------
// RxR (3x3 / 7x7) convolution, WEIGHT-STATIONARY variant (the trajectory resblock convs: 63-67 % of all FLOPs; SPyNet).
// Measured on B200: the per-tap kernel is bound by the L2 -> SM fabric (~9 TB/s aggregate, ~60 GB/s per SM),
// moving 484 KB per 128-pixel tile at C = N = 112, of which 226 KB are the same weights again and again.
// Here the output channels of a group are split into `nt_per_g` slices of NT columns and every CTA keeps the
// COMPLETE weight set of one slice (9 taps x K blocks, <= ~150 KB) resident in shared memory for the whole
// launch; only activations stream, as column-shifted (16+R-1)-row "A units": one TMA box
// {64 ch, 8 cols, 18 rows} per (K block, column shift) serves the three taps of that shift.  Per tile and
// slice the fabric moves ~97 KB instead of 484 KB.
// Work split: CTA b owns combo = b % (groups*slices) for the whole launch and walks the M tiles
// b / ncombo, b / ncombo + grid / ncombo, ...
// ---------------------------------------------------------------------------------------------
// BWX = 8: one box {64 ch, 8 cols, 16+R-1 rows} per column shift (R loads per K block, row shifts are +1024 B
// descriptor offsets).  BWX > 8 ("wide"): ONE box {64 ch, BWX >= 8+R-1 cols, 16+R-1 rows} per K block; a tap
// (r, s) is the descriptor start (r*BWX + s)*128 B with 8-row groups BWX*128 B apart.  That start is not
// 1024-byte aligned, which works because the 128B swizzle is a function of the shared-memory ADDRESS bits
// (TMA writes and tcgen05 reads agree as long as the unit base is 1024-aligned) - R x less activation traffic.
template <int RR, int BWX>      // square odd kernel size: 3 or 7 (compile-time so that the MMA issue loop stays tight)
__global__ void __launch_bounds__(TC_THREADS, 1)
conv3x3_ws_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                  const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB, const TcP p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr int TAPS = RR * RR, UROWS = 16 + RR - 1, PADR = RR / 2;
  const int nboxes = p.n_src * p.kb_per_src * TAPS;
  const uint32_t a_base = smem_base;
  const uint32_t w_base = a_base + p.ra * p.a_unit_bytes;
  const uint32_t bar_base = w_base + nboxes * p.b_stage_bytes;
  const uint32_t afull0 = bar_base, aempty0 = afull0 + 8 * p.ra;
  const uint32_t wfull = aempty0 + 8 * p.ra;
  const uint32_t tfull0 = wfull + 16, tempty0 = tfull0 + 16;      // keeps the bias staging 16-byte aligned
  const uint32_t tmem_slot = tempty0 + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.ra; ++i) { mbar_init(afull0 + 8 * i, 1); mbar_init(aempty0 + 8 * i, 1); }
    mbar_init(wfull, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  float* bias_s = reinterpret_cast<float*>(smem_raw + (tmem_slot + 16 - smem_u32(smem_raw)));

  const int tiles_per_img = p.tiles_h * p.tiles_w;
  const int ncombo = p.groups * p.nt_per_g;
  const int combo = (int)(blockIdx.x % ncombo);
  const int nt = combo % p.nt_per_g, g = combo / p.nt_per_g;
  const long long first = blockIdx.x / ncombo, step = gridDim.x / ncombo;
  const long long n_mt = (long long)p.n_img * tiles_per_img;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {                 // the whole weight slice, once
      mbar_expect_tx(wfull, (uint32_t)(nboxes * p.NT * 128));
      for (int src = 0; src < p.n_src; ++src)
        for (int kb = 0; kb < p.kb_per_src; ++kb)
          for (int tap = 0; tap < TAPS; ++tap)
            tma_load_2d(w_base + ((src * p.kb_per_src + kb) * TAPS + tap) * p.b_stage_bytes, &mapB, wfull, src * p.Cs + kb * 64,
                        (g * TAPS + tap) * p.Cout_g + nt * p.NT);
    }
    __syncwarp();
    int ia = 0; uint32_t pa = 0;
    constexpr uint32_t a_bytes = (uint32_t)UROWS * BWX * 128;
    constexpr int NSFT = BWX == 8 ? RR : 1;
    for (long long mt = first; mt < n_mt; mt += step) {
      const int img = (int)(mt / tiles_per_img); const int rem = (int)(mt % tiles_per_img);
      const int h0 = (rem / p.tiles_w) * 16, w0 = (rem % p.tiles_w) * 8;
      for (int src = 0; src < p.n_src; ++src) {
        const CUtensorMap* mA = src == 0 ? &mapA0 : (src == 1 ? &mapA1 : &mapA2);
        for (int kb = 0; kb < p.kb_per_src; ++kb) {
          const int ca = (p.n_src > 1 ? 0 : ((g * p.Cin_g) & ~7)) + kb * 64;
          for (int sft = 0; sft < NSFT; ++sft) {
            mbar_wait(aempty0 + 8 * ia, pa ^ 1);
            if (elect_one()) {
              mbar_expect_tx(afull0 + 8 * ia, a_bytes);
              tma_load_4d(a_base + ia * p.a_unit_bytes, mA, afull0 + 8 * ia, ca, w0 + sft - PADR, h0 - PADR, img);
            }
            __syncwarp();
            if (++ia == p.ra) { ia = 0; pa ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const int kext = p.Cs + (p.n_src > 1 ? 0 : ((g * p.Cin_g) & 7));
    int ia = 0; uint32_t pa = 0;
    uint32_t acc_phase0 = 0, acc_phase1 = 0;
    int it = 0;
    mbar_wait(wfull, 0);
    tc_fence_after();
    for (long long mt = first; mt < n_mt; mt += step, ++it) {
      const int as = it & 1;
      const uint32_t aph = as ? acc_phase1 : acc_phase0;
      mbar_wait(tempty0 + 8 * as, aph ^ 1);
      if (as) acc_phase1 ^= 1; else acc_phase0 ^= 1;
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * ACC_STRIDE;
      uint32_t accumulate = 0;
      for (int src = 0; src < p.n_src; ++src) {
        for (int kb = 0; kb < p.kb_per_src; ++kb) {
          const int nks = min(4, (kext - kb * 64 + 15) >> 4);
          const bool last_kb = src == p.n_src - 1 && kb == p.kb_per_src - 1;
          if (BWX == 8) {
            for (int sft = 0; sft < RR; ++sft) {
              mbar_wait(afull0 + 8 * ia, pa);
              tc_fence_after();
              const uint32_t a_unit = a_base + ia * p.a_unit_bytes;
              if (elect_one()) {
                for (int r = 0; r < RR; ++r) {
                  const uint64_t ad = smem_desc_sw128(a_unit + (uint32_t)r * 1024u);
                  const uint64_t bd = smem_desc_sw128(w_base + ((src * p.kb_per_src + kb) * TAPS + r * RR + sft) * p.b_stage_bytes);
                  for (int ks = 0; ks < nks; ++ks)
                    umma_f16(d_tmem, ad + (uint64_t)(ks * 2), bd + (uint64_t)(ks * 2), idesc, accumulate | (uint32_t)((r | ks) > 0));
                }
                umma_commit(aempty0 + 8 * ia);
                if (last_kb && sft == RR - 1) umma_commit(tfull0 + 8 * as);
              }
              __syncwarp();
              accumulate = 1;
              if (++ia == p.ra) { ia = 0; pa ^= 1; }
            }
          } else {
            mbar_wait(afull0 + 8 * ia, pa);
            tc_fence_after();
            const uint32_t a_unit = a_base + ia * p.a_unit_bytes;
            if (elect_one()) {
              for (int r = 0; r < RR; ++r) {
                for (int sft = 0; sft < RR; ++sft) {
                  const uint32_t a_addr = a_unit + (uint32_t)(r * BWX + sft) * 128u;
                  uint64_t ad = smem_desc_sw128_sbo(a_addr, (uint32_t)BWX * 128u);
                  if (p.rb) ad |= (uint64_t)((a_addr >> 7) & 7u) << 49;      // experiment: matrix base offset field
                  const uint64_t bd = smem_desc_sw128(w_base + ((src * p.kb_per_src + kb) * TAPS + r * RR + sft) * p.b_stage_bytes);
                  for (int ks = 0; ks < nks; ++ks)
                    umma_f16(d_tmem, ad + (uint64_t)(ks * 2), bd + (uint64_t)(ks * 2), idesc, accumulate | (uint32_t)((r | sft | ks) > 0));
                }
              }
              umma_commit(aempty0 + 8 * ia);
              if (last_kb) umma_commit(tfull0 + 8 * as);
            }
            __syncwarp();
            accumulate = 1;
            if (++ia == p.ra) { ia = 0; pa ^= 1; }
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps) =====================
    if (p.wst)
      run_epilogue_w(p, bias_s, tmem_base, tfull0, tempty0, first, step, n_mt, warp, lane,
                     [&](long long mt, int& cmb, int& img, int& h0, int& w0, bool& valid) {
                       cmb = combo;
                       img = (int)(mt / tiles_per_img);
                       const int rem = (int)(mt % tiles_per_img);
                       h0 = (rem / p.tiles_w) * 16; w0 = (rem % p.tiles_w) * 8;
                       valid = true;
                     });
    else
    run_epilogue(p, bias_s, tmem_base, tfull0, tempty0, first, step, n_mt, warp, lane,
                 [&](long long mt, int row, int& cmb, int& img, int& hh, int& ww, bool& pvalid) {
                   cmb = combo;
                   img = (int)(mt / tiles_per_img);
                   const int rem = (int)(mt % tiles_per_img);
                   hh = (rem / p.tiles_w) * 16 + (row >> 3);
                   ww = (rem % p.tiles_w) * 8 + (row & 7);
                   pvalid = hh < p.H && ww < p.W;
                 });
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------
// 3x3 convolution, weight-stationary on a CTA PAIR (tcgen05 cta_group::2, M = 256).
// The single-CTA weight-stationary kernel above cannot hold the 9 x K x 112 weights of a trajectory resblock
// conv (258 KB) and therefore splits the output channels in two slices: every activation unit crosses the
// L2 -> SM fabric twice and the MMAs run at N = 64, where operand reads from shared memory (A: 4 KB + B: 2 KB
// per 128x64x16 instruction) cost more cycles than the math.  A CTA pair fixes both: each CTA keeps only its
// HALF of the weight rows (N/2 x 9 x K, 129 KB at C = 112), loads the wide activation unit of its OWN 16x8
// tile once, and the leader's single thread issues 256 x N x 16 MMAs that read A from both SMs and half of B
// from each.  Barriers: activation / weight TMA of both CTAs complete on the LEADER's full barriers
// (cta_group::2 TMA form), tcgen05.commit multicasts "slot free" and "accumulator full" to both CTAs, and the
// 2 x 256 epilogue threads of the pair arrive on the leader's "accumulator empty" barrier.
// Work split: cluster c owns combo = c % slices and walks the tile PAIRS c / slices, + clusters / slices, ...
// (tile 2t -> rank 0, tile 2t+1 -> rank 1; an odd tail tile is loaded fully out of bounds = zeros).
// ---------------------------------------------------------------------------------------------
// R = 7 (SPyNet): the same kernel with {64 ch, 14 cols, 22 rows} units; one weight barrier per kernel ROW (7 boxes).
constexpr int WS2_MAX_BOXES = 36;      // weight barriers: 9 taps x up to 4 K blocks (R = 3), 7 kernel rows x up to 4 K blocks (R = 7)
template <int RR>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv3x3_ws2_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                   const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB,
                   const __grid_constant__ CUtensorMap mapR, const TcP p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr int BWX = 8 + RR - 1, TAPS = RR * RR, UROWS = 16 + RR - 1, PADR = RR / 2;
  constexpr int WG = RR == 3 ? 1 : RR;              // weight boxes per barrier
  // weights: one SWIZZLE_128B box {64 k, N/2 rows} per (K block, tap).  (A dense SWIZZLE_32B layout, one {16 k, N/2}
  // box per K = 16 slice, saves 16 KB at C = 112 but measured ~25 % slower MMA issue on B200.)
  const int nkb = p.n_src * p.kb_per_src;              // K blocks of one tile: (source of the virtual concat, 64-channel block)
  const int nboxes = nkb * TAPS;
  const uint32_t a_base = smem_base;
  const uint32_t w_base = a_base + p.ra * p.a_unit_bytes;
  const uint32_t bar_base = (w_base + nboxes * p.b_stage_bytes + 1023u) & ~1023u;
  const uint32_t afull0 = bar_base, aempty0 = afull0 + 8 * p.ra;
  const uint32_t wfull0 = aempty0 + 8 * p.ra;          // one barrier per weight box (<= 36): the first tile starts on tap 0's weights
  const uint32_t tfull0 = wfull0 + 8 * WS2_MAX_BOXES, tempty0 = tfull0 + 16;
  const uint32_t rfull = tempty0 + 16;                 // residual box of the next item (run_epilogue_rtma)
  const uint32_t tmem_slot = rfull + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.ra; ++i) { mbar_init(afull0 + 8 * i, 1); mbar_init(aempty0 + 8 * i, 1); }
    for (int i = 0; i < nboxes / WG; ++i) mbar_init(wfull0 + 8 * i, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 512); }
    mbar_init(rfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();            // barriers of both CTAs initialised before any remote complete_tx / arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  float* bias_s = reinterpret_cast<float*>(smem_raw + (tmem_slot + 16 - smem_u32(smem_raw)));

  const int tiles_per_img = p.tiles_h * p.tiles_w;
  const int ncombo = p.groups * p.nt_per_g;
  const long long cl = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const int combo = (int)(cl % ncombo);
  const int nt = combo % p.nt_per_g, g = combo / p.nt_per_g;
  const int koff = p.n_src > 1 ? 0 : (g * p.Cin_g) & 7;      // grouped conv: the channel box starts 16-byte aligned, the packed weights are shifted alike
  const long long first = cl / ncombo, step = ncl / ncombo;
  const long long n_mt = (long long)p.n_img * tiles_per_img;
  const long long n_pairs = (n_mt + 1) >> 1;
  const int NH = p.NT >> 1;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    const uint32_t wfull0_l = mapa_rank(wfull0, 0), afull0_l = mapa_rank(afull0, 0);
    if (elect_one()) {                 // this CTA's half of the weight slice, once, in the order the MMAs consume it
      for (int src = 0, kbt = 0; src < p.n_src; ++src)
       for (int kb = 0; kb < p.kb_per_src; ++kb, ++kbt)
        for (int tap = 0; tap < TAPS; ++tap) {
          const int box = kbt * TAPS + tap;
          if (rank == 0 && box % WG == 0) mbar_expect_tx(wfull0 + 8 * (box / WG), (uint32_t)(2 * WG * NH * 128));
          tma_load_2d_2sm(w_base + box * p.b_stage_bytes, &mapB, wfull0_l + 8 * (box / WG), src * p.Cs + kb * 64,
                          (g * TAPS + tap) * p.Cout_g + nt * p.NT + (int)rank * NH);
        }
    }
    __syncwarp();
    pdl_wait();                        // the weights are static; the activations come from the previous kernel
    int ia = 0; uint32_t pa = 0;
    constexpr uint32_t a_bytes = (uint32_t)UROWS * BWX * 128;
    for (long long t = first; t < n_pairs; t += step) {
      const long long mt = 2 * t + rank;
      int img = p.n_img, h0 = 0, w0 = 0;             // odd tail: a box entirely outside the tensor (zero fill)
      if (mt < n_mt) {
        img = (int)(mt / tiles_per_img); const int rem = (int)(mt % tiles_per_img);
        h0 = (rem / p.tiles_w) * 16; w0 = (rem % p.tiles_w) * 8;
      }
      for (int src = 0; src < p.n_src; ++src)
      for (int kb = 0; kb < p.kb_per_src; ++kb) {
        const CUtensorMap* mA = src == 0 ? &mapA0 : (src == 1 ? &mapA1 : &mapA2);
        const int ca = (p.n_src > 1 ? 0 : ((g * p.Cin_g) & ~7)) + kb * 64;
        mbar_wait(aempty0 + 8 * ia, pa ^ 1);
        if (elect_one()) {
          if (p.dbg & 4) { if (rank == 0) mbar_arrive(afull0 + 8 * ia); }
          else {
            if (rank == 0) mbar_expect_tx(afull0 + 8 * ia, 2 * a_bytes);
            tma_load_4d_2sm(a_base + ia * p.a_unit_bytes, mA, afull0_l + 8 * ia, ca, w0 - PADR, h0 - PADR, img);
          }
        }
        __syncwarp();
        if (++ia == p.ra) { ia = 0; pa ^= 1; }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===================== MMA issuer (leader CTA only) =====================
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.NT >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    int ia = 0; uint32_t pa = 0;
    uint32_t acc_phase0 = 0, acc_phase1 = 0;
    int it = 0;
    const uint32_t issue = elect_one() ? 1u : 0u;        // one fixed lane issues every tcgen05 instruction of this warp
    for (long long t = first; t < n_pairs; t += step, ++it) {
      const int as = it & 1;
      const uint32_t aph = as ? acc_phase1 : acc_phase0;
      mbar_wait(tempty0 + 8 * as, aph ^ 1);
      if (as) acc_phase1 ^= 1; else acc_phase0 ^= 1;
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * ACC_STRIDE;
      uint32_t accumulate = 0;
      for (int kbt = 0, kb = 0; kbt < nkb; ++kbt, kb = (kb + 1 == p.kb_per_src ? 0 : kb + 1)) {
        const int nks = min(4, (p.Cs + koff - kb * 64 + 15) >> 4);
        mbar_wait(afull0 + 8 * ia, pa);
        tc_fence_after();
        const uint32_t a_unit = a_base + ia * p.a_unit_bytes;
        if (!(p.dbg & 1)) {
          // descriptors as 32-bit halves: high words constant, low words = start address >> 4 (+ LBO bit 16)
          constexpr uint32_t a_hi = ((uint32_t)(BWX * 128) >> 4) | (1u << 14) | (2u << 29);
          constexpr uint32_t b_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
          const uint32_t a_lo0 = ((a_unit & 0x3FFFFu) >> 4) | (1u << 16);
          const uint32_t bs16 = (uint32_t)p.b_stage_bytes >> 4;
          uint32_t b_lo = (((w_base + (uint32_t)(kbt * TAPS) * (uint32_t)p.b_stage_bytes) & 0x3FFFFu) >> 4) | (1u << 16);
          if (it == 0) {                                   // first tile only: the weights of this K block have landed
            for (int wb = 0; wb < TAPS / WG; ++wb) mbar_wait(wfull0 + 8 * (kbt * (TAPS / WG) + wb), 0);
            tc_fence_after();
          }
          // one fully unrolled instance per K-slice count (the tail K block of C = 112 has 3 slices, SPyNet 1 or 2)
          auto issue_block = [&](auto nk_c) {
            constexpr int NK = decltype(nk_c)::value;
#pragma unroll
            for (int tap = 0; tap < TAPS; ++tap) {
              const uint32_t a_lo = a_lo0 + (uint32_t)((tap / RR) * BWX + (tap % RR)) * 8u;
#pragma unroll
              for (int ks = 0; ks < NK; ++ks)
                umma_f16_2sm_if_lh(issue, d_tmem, a_lo + 2u * ks, a_hi, b_lo + 2u * ks, b_hi, idesc, accumulate | (uint32_t)((tap | ks) > 0));
              b_lo += bs16;
            }
          };
          switch (nks) {
            case 4: issue_block(std::integral_constant<int, 4>{}); break;
            case 3: issue_block(std::integral_constant<int, 3>{}); break;
            case 2: issue_block(std::integral_constant<int, 2>{}); break;
            default: issue_block(std::integral_constant<int, 1>{}); break;
          }
        }
        umma_commit_2sm_if(issue, aempty0 + 8 * ia, (uint16_t)3);
        if (kbt == nkb - 1) umma_commit_2sm_if(issue, tfull0 + 8 * as, (uint16_t)3);
        __syncwarp();
        accumulate = 1;
        if (++ia == p.ra) { ia = 0; pa ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps per CTA, each CTA drains its own 128 accumulator rows) =====================
    pdl_wait();                        // residual reads and output writes wait for the predecessors
    if (p.res_tma)
      run_epilogue_rtma(p, bias_s, tmem_base, tfull0, tempty0, rfull, &mapR, first, step, n_pairs, warp, lane,
                        [&](long long t, int& cmb, int& img, int& h0, int& w0, bool& valid) {
                          const long long mt = 2 * t + rank;
                          cmb = combo;
                          valid = mt < n_mt;
                          img = valid ? (int)(mt / tiles_per_img) : 0;
                          const int rem = valid ? (int)(mt % tiles_per_img) : 0;
                          h0 = (rem / p.tiles_w) * 16; w0 = (rem % p.tiles_w) * 8;
                        },
                        mapa_rank(tempty0, 0));
    else if (p.v8)
      run_epilogue_v8(p, bias_s, tmem_base, tfull0, tempty0, first, step, n_pairs, warp, lane,
                      [&](long long t, int& cmb, int& img, int& h0, int& w0, bool& valid) {
                        const long long mt = 2 * t + rank;
                        cmb = combo;
                        valid = mt < n_mt;
                        img = valid ? (int)(mt / tiles_per_img) : 0;
                        const int rem = valid ? (int)(mt % tiles_per_img) : 0;
                        h0 = (rem / p.tiles_w) * 16; w0 = (rem % p.tiles_w) * 8;
                      },
                      mapa_rank(tempty0, 0));
    else if (p.wst)
      run_epilogue_w(p, bias_s, tmem_base, tfull0, tempty0, first, step, n_pairs, warp, lane,
                     [&](long long t, int& cmb, int& img, int& h0, int& w0, bool& valid) {
                       const long long mt = 2 * t + rank;
                       cmb = combo;
                       valid = mt < n_mt;
                       img = valid ? (int)(mt / tiles_per_img) : 0;
                       const int rem = valid ? (int)(mt % tiles_per_img) : 0;
                       h0 = (rem / p.tiles_w) * 16; w0 = (rem % p.tiles_w) * 8;
                     },
                     mapa_rank(tempty0, 0));
    else
    run_epilogue(p, bias_s, tmem_base, tfull0, tempty0, first, step, n_pairs, warp, lane,
                 [&](long long t, int row, int& cmb, int& img, int& hh, int& ww, bool& pvalid) {
                   const long long mt = 2 * t + rank;
                   cmb = combo;
                   img = (int)(mt / tiles_per_img);
                   const int rem = (int)(mt % tiles_per_img);
                   hh = (rem / p.tiles_w) * 16 + (row >> 3);
                   ww = (rem % p.tiles_w) * 8 + (row & 7);
                   pvalid = mt < n_mt && hh < p.H && ww < p.W;
                   if (mt >= n_mt) { img = 0; hh = 0; ww = 0; }
                 },
                 mapa_rank(tempty0, 0));
  }

  tc_fence_before();
  cluster_sync_all();            // nobody exits (or frees TMEM) while the peer can still read this CTA's smem / signal its barriers
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static int g_num_sms = 0;
static int g_smem_optin = 0;
static std::once_flag g_once;
static int g_pdl = 1;         // VMGB_PDL=0 turns programmatic dependent launch off
static int g_split_n = 1;     // VMGB_SPLIT_N=0: keep N <= 256 linears in one N tile (A/B switch)
static int g_res_tma = 1;     // VMGB_RES_TMA=0: residual of the pair kernel by per-thread cp.async instead of TMA (A/B switch)
static int g_ws2_multi = 1;   // VMGB_WS2_MULTI=0: multi-source convs stay on the per-tap kernel (A/B switch)
static int g_ws2_lin = 1;     // VMGB_WS2_LIN=0: Linears stay on the per-tap kernel (A/B switch)
static int g_v8 = 1;          // VMGB_V8=0: staged copy-out instead of the direct 32-byte-store epilogue (A/B switch)
static int g_wst_gate = 1;    // VMGB_WST_GATE=0: gate epilogue on the per-thread staging path (A/B switch)
static int g_no_wst = 0;      // VMGB_NO_WST=1: A/B switch (tools/microbench.py) for the warp-private staging epilogue
static int g_device = -1;     // the device tc_init ran on: attributes / SM count are per device, other devices are refused

static void tc_init() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
    g_encode = (EncodeTiledFn)fn;
  if (const char* e = getenv("VMGB_NO_WST")) g_no_wst = atoi(e);
  if (const char* e = getenv("VMGB_PDL")) g_pdl = atoi(e);
  if (const char* e = getenv("VMGB_WS2_MULTI")) g_ws2_multi = atoi(e);
  if (const char* e = getenv("VMGB_RES_TMA")) g_res_tma = atoi(e);
  if (const char* e = getenv("VMGB_SPLIT_N")) g_split_n = atoi(e);
  if (const char* e = getenv("VMGB_V8")) g_v8 = atoi(e);
  if (const char* e = getenv("VMGB_WS2_LIN")) g_ws2_lin = atoi(e);
  if (const char* e = getenv("VMGB_WST_GATE")) g_wst_gate = atoi(e);      // A/B switch for tools/microbench.py
  int dev = 0;
  cudaGetDevice(&dev);
  g_device = dev;
  cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&g_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (g_smem_optin > 0) {
    cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin);
    cudaFuncSetAttribute(conv3x3_ws_kernel<3, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin);
    cudaFuncSetAttribute(conv3x3_ws_kernel<7, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin);
    cudaFuncSetAttribute(conv3x3_ws_kernel<3, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin);
    cudaFuncSetAttribute(conv3x3_ws_kernel<3, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin);
    cudaFuncSetAttribute(conv3x3_ws2_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin);
    cudaFuncSetAttribute(conv3x3_ws2_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin);
    cudaFuncSetAttribute(conv3x3_ws2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin);
  }
}

static bool encode_bf16(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                        const cuuint32_t* box, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides_bytes, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

constexpr int WS2_MAX_BOXES_HOST = 36;
constexpr int WST_BYTES = 8 * 4096;     // run_epilogue_w: one 32-row x 64-column bf16 buffer per epilogue warp

// whether the warp-private transposition epilogue applies (see run_epilogue_w_t)
static bool wst_ok(int NT, int Cout_g, const void* out, int out_dt, int out_pitch, int ps, int epi, const void* res, int res_dt, int res_pitch,
                   const void* aux = nullptr, int aux_dt = 0, int aux_pitch = 0) {
  if (g_no_wst) return false;
  if (out_dt != VMGB_BF16 || ps != 0 || (out_pitch & 7) || (Cout_g & 7) || (reinterpret_cast<uintptr_t>(out) & 15)) return false;
  if (epi == VMGB_EPI_STORE) return true;
  const bool res_ok = res_dt == VMGB_BF16 && !(res_pitch & 7) && !(reinterpret_cast<uintptr_t>(res) & 15) && NT <= 128;      // one pass per warp
  if (epi == VMGB_EPI_RESADD) return res_ok;
  return epi == VMGB_EPI_GATE && g_wst_gate && res_ok && aux && aux_dt == VMGB_BF16 && !(aux_pitch & 7) && !(reinterpret_cast<uintptr_t>(aux) & 15);
}

}  // namespace vmgb

using namespace vmgb;

extern "C" int vmgb_conv2d_nhwc_tc(const void* x, const void* x1, const void* x2, int n_src, int x_pitch, const void* w, int Kp,
                                   const void* bias, void* out, int out_dt, int out_pitch, const void* res, int res_dt,
                                   int res_pitch, const void* aux, int aux_dt, int aux_pitch, int n_img, int H, int W, int Cin,
                                   int Cout, int R, int S, int groups, int act, float act_alpha, float out_scale, int epi,
                                   float res_scale, int pixel_shuffle, int variant, void* stream) {
  if (!x || !w || !out || !dt_ok(out_dt)) return VMGB_EINVAL;
  if (n_img <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || groups <= 0 || Cin % groups || Cout % groups) return VMGB_EINVAL;
  if (!(R & 1) || !(S & 1) || (x_pitch & 7) || (Kp & 15) || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w) & 15))
    return VMGB_EINVAL;
  if (n_src < 1 || n_src > 3 || (n_src > 1 && (!x1 || groups != 1 || Cin % n_src || (Cin / n_src) % 16))) return VMGB_EINVAL;
  if (n_src > 2 && !x2) return VMGB_EINVAL;
  if (epi != VMGB_EPI_STORE && (!res || !dt_ok(res_dt))) return VMGB_EINVAL;
  if (epi == VMGB_EPI_GATE && (!aux || !dt_ok(aux_dt))) return VMGB_EINVAL;
  if (pixel_shuffle != 0 && (pixel_shuffle != 2 || (Cout & 15) || groups != 1 || epi != VMGB_EPI_STORE || (out_pitch & 3)))
    return VMGB_EUNSUPPORTED;
  const int Cin_g = Cin / groups, Cout_g = Cout / groups;
  if (Kp < Cin_g) return VMGB_EINVAL;
  std::call_once(g_once, tc_init);
  if (!g_encode || g_num_sms <= 0) return VMGB_EUNSUPPORTED;
  {   // one process drives one GPU (torchrun): the opt-in shared-memory attributes and the SM count were set for g_device only
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != g_device) return VMGB_EUNSUPPORTED;
  }

  TcP p;
  p.n_img = n_img; p.H = H; p.W = W; p.Cin_g = Cin_g; p.Cout_g = Cout_g; p.groups = groups; p.R = R; p.S = S; p.n_src = n_src;
  p.Cs = n_src > 1 ? Cin / n_src : Cin_g;
  p.nt_per_g = (Cout_g + 255) / 256;
  // HBM-bound 1x1 layers with a residual / gate epilogue and 128 < N <= 256 (level-1 proj and fc2-like linears): two N
  // tiles of <= 128 columns keep them on the coalesced warp-private staging epilogue (one pass per warp); the second
  // read of the activations comes from L2
  if (R == 1 && S == 1 && epi != VMGB_EPI_STORE && g_split_n && Cout_g > 128 && Cout_g <= 256 && (Cout_g & 31) == 0 && Cin <= 256) p.nt_per_g = 2;
  p.NT = (((Cout_g + p.nt_per_g - 1) / p.nt_per_g) + 15) / 16 * 16;
  int max_koff = 0;
  for (int g = 0; g < groups; ++g) max_koff = max(max_koff, (g * Cin_g) & 7);
  if (Kp < Cin_g + max_koff) return VMGB_EINVAL;
  p.kb_per_src = (p.Cs + max_koff + 63) / 64;
  p.b_stage_bytes = (p.NT * 128 + 1023) / 1024 * 1024;
  p.bias = (const float*)bias; p.out = out; p.res = res; p.aux = aux;
  p.out_dt = out_dt; p.out_pitch = out_pitch; p.res_dt = res_dt; p.res_pitch = res_pitch; p.aux_dt = aux_dt; p.aux_pitch = aux_pitch;
  p.act = act; p.alpha = act_alpha; p.out_scale = out_scale; p.epi = epi; p.res_scale = res_scale; p.ps = pixel_shuffle;
  p.dbg = 0; p.zero = 0; p.wst = 0; p.v8 = 0; p.res_tma = 0;
  p.nsub = 0; p.ra = 0; p.rb = 0; p.a_unit_bytes = 0; p.a_sbo = 0; p.res_smem = 0; p.out_smem = 0; p.aux_smem = 0;
  const int budget = g_smem_optin - 1024 - 512 - 1024;     // alignment slack, barriers, bias staging

  // ---- CTA-pair weight-stationary 3x3 kernel (variant 13, or auto): grouped single-source convs, or ungrouped with up
  //      to 3 equally wide sources (virtual concat)
  if (variant >= 32) { p.dbg = (variant >> 5) & 15; variant = 13; }      // probes: 13 + 32*dbg
  // A Linear ([1,1,rows,K] "image", rows % 8 == 0) takes the same kernel with R = 1: the rows are read as an [rows/8, 8] image,
  // a 16 x 8 tile = 128 consecutive rows.  The per-tap kernel re-streams the weights with every 128-row tile (C = 224:
  // 100 KB of weights per 57 KB of activations through the L2 -> SM fabric, 0.63 of the HBM rate); here they are resident.
  const bool lin1 = R == 1 && S == 1 && H == 1 && n_img == 1 && (W & 7) == 0 && W >= 128 && n_src == 1 && groups == 1 && pixel_shuffle == 0 &&
                    (epi == VMGB_EPI_STORE || (epi == VMGB_EPI_RESADD && Cout <= 128) || variant == 13) && g_ws2_lin;
  const int Hc = lin1 ? W / 8 : H, Wc = lin1 ? 8 : W;
  if (R == S && (R == 3 || R == 7 || lin1) && Hc >= 8 && Wc >= 8 && (n_src == 1 || (groups == 1 && (g_ws2_multi || variant == 13))) &&
      (variant == 0 || variant == 13)) {
    const int ubytes = ((16 + R - 1) * (8 + R - 1) * 128 + 1023) / 1024 * 1024;
    const int nboxes = R * R * n_src * p.kb_per_src;      // SWIZZLE_128B weight boxes {64 k, NT/2 rows}
    const int nbars = nboxes / (R == 3 ? 1 : R);
    const long long n_mt = (long long)n_img * ((Hc + 15) / 16) * ((Wc + 7) / 8);
    const long long n_pairs = (n_mt + 1) / 2;
    int best_slices = 0, best_nt = 0, best_ra = 0, best_res = 0, best_out = 0, best_wst = 0;
    for (int sl = 1; sl <= 8 && !best_slices; ++sl) {
      const int nts = (((Cout_g + sl - 1) / sl) + 15) / 16 * 16;
      if (nts > 256) continue;
      const int wbytes = (nboxes * (nts / 2) * 128 + 1023) / 1024 * 1024;
      const int old_resb = (epi != VMGB_EPI_STORE && res_dt == VMGB_BF16 && pixel_shuffle == 0) ? nts * 256 : 0;
      // the warp-private staging epilogue only where it does not cost an activation ring slot
      // (wide outputs were copied out through a whole-tile staging buffer whose index arithmetic cost more than the GELU)
      const bool wst = wst_ok(nts, Cout_g, out, out_dt, out_pitch, pixel_shuffle, epi, res, res_dt, res_pitch) &&
                       (budget - wbytes - WST_BYTES) / ubytes >= 3 &&
                       (Cout >= 256 || (budget - wbytes - WST_BYTES) / ubytes >= (budget - wbytes - old_resb) / ubytes);
      const int resb = wst ? WST_BYTES : ((epi != VMGB_EPI_STORE && res_dt == VMGB_BF16 && pixel_shuffle == 0) ? nts * 256 : 0);
      int left = budget - wbytes - resb;
      int outb = 0;
      if (!wst && Cout >= 256 && out_dt == VMGB_BF16 && pixel_shuffle == 0 && (out_pitch & 7) == 0 && (Cout_g & 7) == 0 &&
          !(reinterpret_cast<uintptr_t>(out) & 15) && (left - 128 * (nts + 8) * 2) / ubytes >= 3)
        outb = 128 * (nts + 8) * 2;
      left -= outb;
      const int ra = left / ubytes;
      if (ra >= 2) { best_slices = sl; best_nt = nts; best_ra = ra > 6 ? 6 : ra; best_res = resb; best_out = outb; best_wst = wst; }
    }
    const long long clusters_max = (g_num_sms / 2) / (groups * (best_slices ? best_slices : 1));
    // (Linears: measured wins with the whole weight matrix resident in one slice, K <= 768: 112 / 224 / 448-wide layers 0.84 -> 0.90
    //  of the HBM rate, FFN fc2 672 -> 112 -9 %; 1344 -> 224 needs two slices and loses)
    const bool worth = best_slices && clusters_max >= 1 && n_pairs >= 3 * clusters_max &&
                       (lin1 ? (best_slices == 1 && p.kb_per_src <= 12) : n_src * p.kb_per_src <= 4);
    if (best_slices && clusters_max >= 1 && nbars <= WS2_MAX_BOXES_HOST && (variant == 13 || worth)) {       // weight-box barriers in the kernel
      p.nt_per_g = best_slices; p.NT = best_nt; p.b_stage_bytes = (best_nt / 2) * 128; p.ra = best_ra; p.out_smem = best_out;
      p.wst = best_wst; p.res_smem = best_wst ? 0 : best_res;
      // (3x3 only: the 7x7 SPyNet layers with 16-64 output channels measured 15 % slower on it than on the staged copy-out)
      p.v8 = g_v8 && R != 7 && epi == VMGB_EPI_STORE && out_dt == VMGB_BF16 && pixel_shuffle == 0 && (out_pitch & 15) == 0 && (Cout_g & 7) == 0 &&
             !(reinterpret_cast<uintptr_t>(out) & 31) && (act == VMGB_ACT_NONE || act == VMGB_ACT_RELU || act == VMGB_ACT_LRELU || act == VMGB_ACT_GELU);
      p.BH = 16; p.BW = 8; p.a_unit_bytes = ubytes; p.stages = 0;
      p.H = Hc; p.W = Wc;
      p.tiles_h = (Hc + 15) / 16; p.tiles_w = (Wc + 7) / 8;
      p.total_tiles = n_mt * groups * p.nt_per_g;
      CUtensorMap mA[3], mB, mR;
      // resblock conv2 (no activation, bf16 residual and output): the residual tile comes by TMA into the same 128 x NT
      // staging buffer the cp.async path would use
      p.res_tma = g_res_tma && !best_wst && best_res > 0 && epi == VMGB_EPI_RESADD && act == VMGB_ACT_NONE && groups == 1 &&
                  out_dt == VMGB_BF16 && res_dt == VMGB_BF16 && pixel_shuffle == 0 && (Cout_g & 15) == 0 && (res_pitch & 7) == 0 &&
                  (out_pitch & 7) == 0 && !(reinterpret_cast<uintptr_t>(res) & 15) && !(reinterpret_cast<uintptr_t>(out) & 15) && best_nt <= 128;
      if (p.res_tma) {
        cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)Wc, (cuuint64_t)Hc, (cuuint64_t)n_img};
        cuuint64_t str[3] = {(cuuint64_t)res_pitch * 2, (cuuint64_t)Wc * res_pitch * 2, (cuuint64_t)Hc * Wc * res_pitch * 2};
        cuuint32_t box[4] = {(cuuint32_t)best_nt, 8, 16, 1};
        if (!encode_bf16(&mR, res, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE)) p.res_tma = 0;
      }
      {
        const void* srcs[3] = {x, x1, x2};
        const int c_extent = n_src > 1 ? p.Cs : Cin;
        for (int s = 0; s < 3; ++s) {
          const void* ptr = s < n_src ? srcs[s] : x;
          if (reinterpret_cast<uintptr_t>(ptr) & 15) return VMGB_EINVAL;
          cuuint64_t dims[4] = {(cuuint64_t)c_extent, (cuuint64_t)Wc, (cuuint64_t)Hc, (cuuint64_t)n_img};
          cuuint64_t str[3] = {(cuuint64_t)x_pitch * 2, (cuuint64_t)Wc * x_pitch * 2, (cuuint64_t)Hc * Wc * x_pitch * 2};
          cuuint32_t box[4] = {64, (cuuint32_t)(8 + R - 1), (cuuint32_t)(16 + R - 1), 1};
          if (!encode_bf16(&mA[s], ptr, 4, dims, str, box)) return VMGB_EINVAL;
        }
      }
      {
        cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)groups * R * R * Cout_g};
        cuuint64_t str[1] = {(cuuint64_t)Kp * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)(p.NT / 2)};
        if (!encode_bf16(&mB, w, 2, dims, str, box)) return VMGB_EINVAL;
      }
      long long per = n_pairs < clusters_max ? n_pairs : clusters_max;
      const int smem = p.ra * p.a_unit_bytes + (nboxes * p.b_stage_bytes + 1023) / 1024 * 1024 + 1024 + 512 + 1024 + best_res + p.out_smem +
                       (p.res_tma ? 256 : 0);
      if (!p.res_tma) mR = mB;            // unused, but a valid descriptor
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(2 * per * groups * best_slices)); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = as_stream(stream);
      cudaLaunchAttribute at[2];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
      // the kernel loads its weight slice BEFORE griddepcontrol.wait (weights are static); a launch that directly follows a
      // weight (re)pack on this stream must therefore not overlap its predecessor
      cfg.attrs = at; cfg.numAttrs = (g_pdl && !g_pack_pending) ? 2 : 1;
      g_pack_pending = 0;
      cudaError_t e = R == 3 ? cudaLaunchKernelEx(&cfg, conv3x3_ws2_kernel<3>, mA[0], mA[1], mA[2], mB, mR, p)
                    : R == 7 ? cudaLaunchKernelEx(&cfg, conv3x3_ws2_kernel<7>, mA[0], mA[1], mA[2], mB, mR, p)
                             : cudaLaunchKernelEx(&cfg, conv3x3_ws2_kernel<1>, mA[0], mA[1], mA[2], mB, mR, p);
      if (e != cudaSuccess) { g_last_cuda_error = (int)e; return VMGB_ECUDA; }
      return check_launch();
    }
  }
  if (variant == 13) return VMGB_EUNSUPPORTED;

  // ---- weight-stationary 3x3 kernel (variant 8, or auto when the launch is large enough to amortise the
  //      one-off weight load): find the smallest number of output-channel slices whose weights fit
  if (R == S && (R == 3 || R == 7) && H >= 8 && W >= 8 && (variant == 0 || (variant >= 8 && variant <= 12))) {
    // variants 9..12 (experiments, 3x3 only): wide activation boxes, 9/10 = 10 columns, 11/12 = 16 columns; even = base-offset field set
    const int bwx = (R == 3 && variant >= 9) ? (variant <= 10 ? 10 : 16) : 8;
    const int urows = 16 + R - 1, ubytes = (urows * bwx * 128 + 1023) / 1024 * 1024;
    const int nboxes = n_src * p.kb_per_src * R * R;
    const long long n_mt = (long long)n_img * ((H + 15) / 16) * ((W + 7) / 8);
    int best_slices = 0, best_nt = 0, best_ra = 0, best_res = 0, best_out = 0, best_wst = 0;
    for (int sl = 1; sl <= 8 && !best_slices; ++sl) {
      const int nts = (((Cout_g + sl - 1) / sl) + 15) / 16 * 16;
      if (nts > 256) continue;
      const bool wst = wst_ok(nts, Cout_g, out, out_dt, out_pitch, pixel_shuffle, epi, res, res_dt, res_pitch);
      const int resb = wst ? WST_BYTES : ((epi != VMGB_EPI_STORE && res_dt == VMGB_BF16 && pixel_shuffle == 0) ? nts * 256 : 0);
      int left = budget - nboxes * nts * 128 - resb;
      // coalesced copy-out for wide outputs (FFN hidden tensor): only if the activation ring keeps >= 3 units
      int outb = 0;
      if (!wst && Cout >= 256 && out_dt == VMGB_BF16 && pixel_shuffle == 0 && (out_pitch & 7) == 0 && (Cout_g & 7) == 0 &&
          !(reinterpret_cast<uintptr_t>(out) & 15) && (left - 128 * (nts + 8) * 2) / ubytes >= 3)
        outb = 128 * (nts + 8) * 2;
      left -= outb;
      const int ra = left / ubytes;
      if (ra >= 2) { best_slices = sl; best_nt = nts; best_ra = ra > 6 ? 6 : ra; best_res = resb; best_out = outb; best_wst = wst; }
    }
    const long long ctas = g_num_sms / (groups * (best_slices ? best_slices : 1));
    // auto: measured wins (tools/microbench.py) are the single-source, ungrouped convs with <= 2 K blocks
    // (trajectory resblocks at C=112, recon head); at least ~4 M tiles per CTA amortise the resident weight load
    const bool worth = best_slices && ctas >= 1 && n_mt >= 4 * ctas && n_src == 1 && groups == 1 && p.kb_per_src <= 2 &&
                       (R == 3 || Cin <= 32);      // 7x7: measured wins only where the 64-wide K block is mostly padding
    if (best_slices && ctas >= 1 && (variant >= 8 || worth)) {
      p.rb = (variant == 10 || variant == 12) ? 1 : 0;
      p.nt_per_g = best_slices; p.NT = best_nt; p.b_stage_bytes = best_nt * 128; p.ra = best_ra; p.out_smem = best_out;
      p.wst = best_wst; p.res_smem = best_wst ? 0 : best_res;
      p.BH = 16; p.BW = 8; p.a_unit_bytes = ubytes; p.stages = 0;
      p.tiles_h = (H + 15) / 16; p.tiles_w = (W + 7) / 8;
      p.total_tiles = n_mt * groups * p.nt_per_g;
      CUtensorMap mA[3], mB, mR;
      // resblock conv2 (no activation, bf16 residual and output): the residual tile comes by TMA into the same 128 x NT
      // staging buffer the cp.async path would use
      p.res_tma = g_res_tma && !best_wst && best_res > 0 && epi == VMGB_EPI_RESADD && act == VMGB_ACT_NONE && groups == 1 &&
                  out_dt == VMGB_BF16 && res_dt == VMGB_BF16 && pixel_shuffle == 0 && (Cout_g & 15) == 0 && (res_pitch & 7) == 0 &&
                  (out_pitch & 7) == 0 && !(reinterpret_cast<uintptr_t>(res) & 15) && !(reinterpret_cast<uintptr_t>(out) & 15) && best_nt <= 128;
      if (p.res_tma) {
        cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_img};
        cuuint64_t str[3] = {(cuuint64_t)res_pitch * 2, (cuuint64_t)W * res_pitch * 2, (cuuint64_t)H * W * res_pitch * 2};
        cuuint32_t box[4] = {(cuuint32_t)best_nt, 8, 16, 1};
        if (!encode_bf16(&mR, res, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE)) p.res_tma = 0;
      }
      const void* srcs[3] = {x, x1, x2};
      const int c_extent = n_src > 1 ? p.Cs : Cin;
      for (int s = 0; s < 3; ++s) {
        const void* ptr = s < n_src ? srcs[s] : x;
        if (reinterpret_cast<uintptr_t>(ptr) & 15) return VMGB_EINVAL;
        cuuint64_t dims[4] = {(cuuint64_t)c_extent, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_img};
        cuuint64_t str[3] = {(cuuint64_t)x_pitch * 2, (cuuint64_t)W * x_pitch * 2, (cuuint64_t)H * W * x_pitch * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)bwx, (cuuint32_t)urows, 1};
        if (!encode_bf16(&mA[s], ptr, 4, dims, str, box)) return VMGB_EINVAL;
      }
      {
        cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)groups * R * R * Cout_g};
        cuuint64_t str[1] = {(cuuint64_t)Kp * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)p.NT};
        if (!encode_bf16(&mB, w, 2, dims, str, box)) return VMGB_EINVAL;
      }
      const int ncombo = groups * p.nt_per_g;
      long long per = n_mt < ctas ? n_mt : ctas;
      const int smem = p.ra * p.a_unit_bytes + nboxes * p.b_stage_bytes + 1024 + 512 + 1024 + best_res + p.out_smem;
      const dim3 grid((unsigned)(per * ncombo));
      if (R == 7) conv3x3_ws_kernel<7, 8><<<grid, TC_THREADS, smem, as_stream(stream)>>>(mA[0], mA[1], mA[2], mB, p);
      else if (bwx == 10) conv3x3_ws_kernel<3, 10><<<grid, TC_THREADS, smem, as_stream(stream)>>>(mA[0], mA[1], mA[2], mB, p);
      else if (bwx == 16) conv3x3_ws_kernel<3, 16><<<grid, TC_THREADS, smem, as_stream(stream)>>>(mA[0], mA[1], mA[2], mB, p);
      else conv3x3_ws_kernel<3, 8><<<grid, TC_THREADS, smem, as_stream(stream)>>>(mA[0], mA[1], mA[2], mB, p);
      return check_launch();
    }
    if (variant >= 8) return VMGB_EUNSUPPORTED;
  }

  // ---- per-tap kernel (variant 1, or auto for everything else)
  if (variant != 0 && variant != 1) return VMGB_EINVAL;
  if (H == 1) { p.BW = 128; p.BH = 1; }
  else { p.BW = W >= 16 ? 16 : (W >= 8 ? 8 : 4); p.BH = 128 / p.BW; }
  const int stage_bytes = A_STAGE_BYTES + p.b_stage_bytes;
  p.res_smem = 0; p.out_smem = 0; p.aux_smem = 0;
  int epi_smem = 0;
  const int wst_bytes = epi == VMGB_EPI_GATE ? 2 * WST_BYTES : WST_BYTES;
  if (wst_ok(p.NT, Cout_g, out, out_dt, out_pitch, pixel_shuffle, epi, res, res_dt, res_pitch, aux, aux_dt, aux_pitch) &&
      (budget - wst_bytes) / stage_bytes >= 3) {
    p.wst = 1; epi_smem = wst_bytes;      // warp-private transposition staging: coalesced residual / gate reads and output writes
  } else {
    // residual staging (128 rows x NT bf16, one item ahead) when it leaves at least 4 pipeline stages
    if (epi != VMGB_EPI_STORE && res_dt == VMGB_BF16 && pixel_shuffle == 0 && (budget - p.NT * 256) / stage_bytes >= 4)
      p.res_smem = p.NT * 256;
    // output staging (coalesced copy-out): bf16 output, 16-byte aligned channel runs, whole-vector validity
    if (out_dt == VMGB_BF16 && pixel_shuffle == 0 && (out_pitch & 7) == 0 && (Cout_g & 7) == 0 && !(reinterpret_cast<uintptr_t>(out) & 15) &&
        (budget - p.res_smem - 128 * (p.NT + 8) * 2) / stage_bytes >= 4)
      p.out_smem = 128 * (p.NT + 8) * 2;
    if (epi == VMGB_EPI_GATE && p.res_smem && aux_dt == VMGB_BF16 && aux_pitch == res_pitch && !(reinterpret_cast<uintptr_t>(aux) & 15) &&
        (budget - p.res_smem - p.out_smem - p.NT * 256) / stage_bytes >= 3)
      p.aux_smem = p.NT * 256;
    epi_smem = p.res_smem + p.out_smem + p.aux_smem;
  }
  p.stages = (budget - epi_smem) / stage_bytes;
  if (p.stages > 8) p.stages = 8;
  if (p.stages < 2) return VMGB_EUNSUPPORTED;
  p.tiles_h = (H + p.BH - 1) / p.BH; p.tiles_w = (W + p.BW - 1) / p.BW;
  p.total_tiles = (long long)n_img * p.tiles_h * p.tiles_w * groups * p.nt_per_g;

  // activation maps: dims (C, W, H, N) ; channel extent = the source's own channel count so that the
  // K tail and the virtual-concat boundary are zero-filled by TMA
  CUtensorMap mA[3], mB;
  const void* srcs[3] = {x, x1, x2};
  const int c_extent = n_src > 1 ? p.Cs : Cin;
  for (int s = 0; s < 3; ++s) {
    const void* ptr = s < n_src ? srcs[s] : x;
    if (reinterpret_cast<uintptr_t>(ptr) & 15) return VMGB_EINVAL;
    cuuint64_t dims[4] = {(cuuint64_t)c_extent, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_img};
    cuuint64_t str[3] = {(cuuint64_t)x_pitch * 2, (cuuint64_t)W * x_pitch * 2, (cuuint64_t)H * W * x_pitch * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)p.BW, (cuuint32_t)p.BH, 1};
    if (!encode_bf16(&mA[s], ptr, 4, dims, str, box)) return VMGB_EINVAL;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)groups * R * S * Cout_g};
    cuuint64_t str[1] = {(cuuint64_t)Kp * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)p.NT};
    if (!encode_bf16(&mB, w, 2, dims, str, box)) return VMGB_EINVAL;
  }
  const long long grid = p.total_tiles < g_num_sms ? p.total_tiles : g_num_sms;
  const int smem = p.stages * stage_bytes + 1024 + 512 + 1024 + epi_smem;
  conv_tc_kernel<<<(unsigned)grid, TC_THREADS, smem, as_stream(stream)>>>(mA[0], mA[1], mA[2], mB, p);
  return check_launch();
}
