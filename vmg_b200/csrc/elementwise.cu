// elementwise.cu -- the HBM-bound row kernels of the TAB / resampling / shortcut path.
// All kernels: channels-last rows, fp32 math, f32 or bf16 storage.  References per entry point are in
// include/vmgb.h.
#include "common.cuh"

namespace vmgb {

constexpr int MAXV = 32;  // per-lane register tile: rows up to 32*32 = 1024 channels

// ---------------------------------------------------------------------------------------------
// LayerNorm: one warp per row
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_kernel(const void* x, int x_dt, void* out, int out_dt,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        long long rows, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[MAXV];
  const long long base = row * C;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) { const int c = lane + 32 * i; v[i] = c < C ? ldf(x, x_dt, base + c) : 0.f; s += v[i]; }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) { const int c = lane + 32 * i; const float d = c < C ? v[i] - mean : 0.f; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) { const int c = lane + 32 * i; if (c < C) stf(out, out_dt, base + c, (v[i] - mean) * rstd * gamma[c] + beta[c]); }
}

// vectorised variants: lane owns channels [4*lane + 128*i, +4), C % 4 == 0, C <= 1024
constexpr int MAXQ = 8;
__global__ void __launch_bounds__(256) layernorm_vec_kernel(const void* x, int x_dt, void* out, int out_dt,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            long long rows, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float4 v[MAXQ];
  const long long base = row * C;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXQ; ++i) {
    const int c = 4 * lane + 128 * i;
    v[i] = c < C ? ld4(x, x_dt, base + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXQ; ++i) {
    if (4 * lane + 128 * i < C) {
      const float a = v[i].x - mean, b = v[i].y - mean, c2 = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c2 * c2 + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
  for (int i = 0; i < MAXQ; ++i) {
    const int c = 4 * lane + 128 * i;
    if (c < C) {
      const float4 g = *reinterpret_cast<const float4*>(gamma + c), b = *reinterpret_cast<const float4*>(beta + c);
      st4(out, out_dt, base + c, make_float4((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y,
                                             (v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w));
    }
  }
}

__global__ void __launch_bounds__(256) l2norm_vec_kernel(const void* x, int x_dt, int x_pitch, void* out, int out_dt, long long rows, int C) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float4 v[MAXQ];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXQ; ++i) {
    const int c = 4 * lane + 128 * i;
    v[i] = c < C ? ld4(x, x_dt, row * x_pitch + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    s += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
  const float inv = 1.f / fmaxf(sqrtf(warp_sum(s)), 1e-12f);
#pragma unroll
  for (int i = 0; i < MAXQ; ++i) {
    const int c = 4 * lane + 128 * i;
    if (c < C) st4(out, out_dt, row * C + c, make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv));
  }
}

// narrow rows (C <= 128, C % 4 == 0): one vector per lane, ROWS independent rows per warp iteration so that
// several 8..16-byte loads per lane are in flight (a single 224-byte row per warp leaves HBM latency exposed)
template <int ROWS>
__global__ void __launch_bounds__(256) layernorm_narrow_kernel(const void* x, int x_dt, void* out, int out_dt,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               long long rows, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const long long warp_id = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int c = 4 * lane;
  const bool act = c < C;
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f), b = g;
  if (act) { g = *reinterpret_cast<const float4*>(gamma + c); b = *reinterpret_cast<const float4*>(beta + c); }
  for (long long r0 = warp_id * ROWS; r0 < rows; r0 += nwarps * ROWS) {
    float4 v[ROWS];
#pragma unroll
    for (int k = 0; k < ROWS; ++k)
      v[k] = (act && r0 + k < rows) ? ld4(x, x_dt, (r0 + k) * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
      const float mean = warp_sum((v[k].x + v[k].y) + (v[k].z + v[k].w)) / (float)C;
      const float d0 = v[k].x - mean, d1 = v[k].y - mean, d2 = v[k].z - mean, d3 = v[k].w - mean;
      const float q = act ? (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3) : 0.f;
      const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
      if (act && r0 + k < rows)
        st4(out, out_dt, (r0 + k) * C + c, make_float4(d0 * rstd * g.x + b.x, d1 * rstd * g.y + b.y, d2 * rstd * g.z + b.z, d3 * rstd * g.w + b.w));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Lockstep sweep pairing of the trajectory (model.py _trajectory): pure 16-byte data movement.
//   pair  : out[j][0] = x[T-1-j], out[j][1] = x[j]            (one read, two writes)
//   unpair: back[i] = paired[T-1-i][0], fwd[i] = paired[i][1]
// A "frame" is the n*H*W*C block of one time step.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pair_sweeps_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int T, long long fv) {
  const long long total = (long long)T * fv;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 4 * stride) {
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { const long long i = i0 + k * stride; if (i < total) v[k] = __ldg(x + i); }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long i = i0 + k * stride;
      if (i < total) {
        const long long j = i / fv, r = i - j * fv;
        out[(2 * j + 1) * fv + r] = v[k];
        out[(2 * (T - 1 - j)) * fv + r] = v[k];
      }
    }
  }
}

__global__ void __launch_bounds__(256) unpair_sweeps_kernel(const uint4* __restrict__ paired, uint4* __restrict__ back,
                                                            uint4* __restrict__ fwd, int T, long long fv) {
  const long long total = 2LL * T * fv;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 4 * stride) {
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { const long long i = i0 + k * stride; if (i < total) v[k] = __ldg(paired + i); }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long i = i0 + k * stride;
      if (i < total) {
        const long long h = i / fv, r = i - h * fv;      // h = 2*j + half
        const long long j = h >> 1;
        if (h & 1) fwd[j * fv + r] = v[k];
        else back[(T - 1 - j) * fv + r] = v[k];
      }
    }
  }
}

// bf16 rows, C % 8 == 0: LPR lanes own a row (LPR = 16: C <= 128, a half-warp; LPR = 32: C <= 256), 16 bytes per lane, ROWS
// independent rows in flight per lane group (the 8-byte-per-lane kernels above leave most of the HBM latency exposed:
// measured 1.1 TB/s at C = 112 against 4.7 TB/s here).  MODE 0 = LayerNorm, MODE 1 = L2 normalise (F.normalize, eps 1e-12).
template <int LPR, int ROWS, int MODE>
__global__ void __launch_bounds__(256) rownorm16_kernel(const __nv_bfloat16* __restrict__ x, int x_pitch, __nv_bfloat16* __restrict__ out,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        long long rows, int C, float eps) {
  constexpr int GPW = 32 / LPR;                                   // row groups per warp
  const int hl = threadIdx.x & (LPR - 1), hsel = (threadIdx.x & 31) / LPR;
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int c = 8 * hl;
  const bool act = c < C;
  float g[8], b[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { g[k] = (MODE == 0 && act) ? gamma[c + k] : 0.f; b[k] = (MODE == 0 && act) ? beta[c + k] : 0.f; }
  const float inv_c = 1.f / (float)C;
  // the loop bound is uniform over the warp (every lane takes part in every full-mask shuffle)
  for (long long w0 = warp_id * GPW * ROWS; w0 < rows; w0 += n_warps * GPW * ROWS) {
    const long long r0 = w0 + hsel * ROWS;
    uint4 raw[ROWS];
#pragma unroll
    for (int k = 0; k < ROWS; ++k)
      raw[k] = (act && r0 + k < rows) ? __ldg(reinterpret_cast<const uint4*>(x + (r0 + k) * x_pitch + c)) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
      float v[8];
      const uint32_t w[4] = {raw[k].x, raw[k].y, raw[k].z, raw[k].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) { v[2 * q] = __uint_as_float(w[q] << 16); v[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u); }
      uint32_t o[4];
      if (MODE == 0) {
        float s = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
#pragma unroll
        for (int m = LPR / 2; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
        const float mean = s * inv_c;
        float q2 = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) { v[e] -= mean; q2 += v[e] * v[e]; }
        q2 = act ? q2 : 0.f;
#pragma unroll
        for (int m = LPR / 2; m >= 1; m >>= 1) q2 += __shfl_xor_sync(0xffffffffu, q2, m);
        const float rstd = rsqrtf(q2 * inv_c + eps);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * q] * rstd * g[2 * q] + b[2 * q], v[2 * q + 1] * rstd * g[2 * q + 1] + b[2 * q + 1]);
          o[q] = *reinterpret_cast<const uint32_t*>(&h);
        }
      } else {
        float q2 = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) q2 += v[e] * v[e];
#pragma unroll
        for (int m = LPR / 2; m >= 1; m >>= 1) q2 += __shfl_xor_sync(0xffffffffu, q2, m);
        const float inv = 1.f / fmaxf(sqrtf(q2), 1e-12f);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * q] * inv, v[2 * q + 1] * inv);
          o[q] = *reinterpret_cast<const uint32_t*>(&h);
        }
      }
      if (act && r0 + k < rows) *reinterpret_cast<uint4*>(out + (r0 + k) * C + c) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

template <int MODE>
static void launch_rownorm16(const void* x, int x_pitch, void* out, const float* gamma, const float* beta, long long rows, int C, float eps,
                             cudaStream_t st) {
  const int lpr = C <= 128 ? 16 : 32, rows_per_warp_iter = (32 / lpr) * (lpr == 16 ? 8 : 4);
  const long long wanted = (rows + 8LL * rows_per_warp_iter - 1) / (8LL * rows_per_warp_iter);
  const unsigned nb = (unsigned)(wanted < 148 * 6 ? (wanted < 1 ? 1 : wanted) : 148 * 6);
  if (lpr == 16) rownorm16_kernel<16, 8, MODE><<<nb, 256, 0, st>>>((const __nv_bfloat16*)x, x_pitch, (__nv_bfloat16*)out, gamma, beta, rows, C, eps);
  else rownorm16_kernel<32, 4, MODE><<<nb, 256, 0, st>>>((const __nv_bfloat16*)x, x_pitch, (__nv_bfloat16*)out, gamma, beta, rows, C, eps);
}

// L2 normalise (F.normalize, eps 1e-12)
__global__ void __launch_bounds__(256) l2norm_kernel(const void* x, int x_dt, int x_pitch, void* out, int out_dt,
                                                     long long rows, int C) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) { const int c = lane + 32 * i; v[i] = c < C ? ldf(x, x_dt, row * x_pitch + c) : 0.f; s += v[i] * v[i]; }
  const float inv = 1.f / fmaxf(sqrtf(warp_sum(s)), 1e-12f);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) { const int c = lane + 32 * i; if (c < C) stf(out, out_dt, row * C + c, v[i] * inv); }
}

// ---------------------------------------------------------------------------------------------
// Morph-FC chunk permutation.  Token/feature maps (function.py:763-785):
//   axis 0: row = ((bt*W + w)*(Hp/ch) + hg)*ch + g ; f = r*S + s ; src x[bt, hg*ch + r, w, g*S + s]
//   axis 1: row = ((bt*H + h)*(Wp/ch) + wg)*ch + g ; f = r*S + s ; src x[bt, h, wg*ch + r, g*S + s]
// ---------------------------------------------------------------------------------------------
__global__ void morph_gather_kernel(const void* x, int x_dt, void* A, int a_dt, int BT, int H, int W, int C,
                                    int ch, int axis, long long total) {
  const int Cp = (C + ch - 1) / ch * ch, S = Cp / ch;
  const int L = axis == 0 ? H : W, Lg = (L + ch - 1) / ch, O = axis == 0 ? W : H;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % Cp); long long row = i / Cp;
    const int g = (int)(row % ch); row /= ch;
    const int lg = (int)(row % Lg); row /= Lg;
    const int o = (int)(row % O); const int bt = (int)(row / O);
    const int r = f / S, s = f - r * S;
    const int l = lg * ch + r, c = g * S + s;
    float v = 0.f;
    if (l < L && c < C) {
      const int h = axis == 0 ? l : o, w = axis == 0 ? o : l;
      v = ldf(x, x_dt, (((long long)bt * H + h) * W + w) * C + c);
    }
    stf(A, a_dt, i, v);
  }
}

__global__ void morph_scatter_kernel(const void* Y, int y_dt, void* out, int out_dt, int BT, int H, int W, int C,
                                     int ch, int axis, long long total) {
  const int Cp = (C + ch - 1) / ch * ch, S = Cp / ch;
  const int L = axis == 0 ? H : W, Lg = (L + ch - 1) / ch, O = axis == 0 ? W : H;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C); long long pix = i / C;
    const int w = (int)(pix % W); pix /= W;
    const int h = (int)(pix % H); const int bt = (int)(pix / H);
    const int l = axis == 0 ? h : w, o = axis == 0 ? w : h;
    const int lg = l / ch, r = l - lg * ch, g = c / S, s = c - g * S;
    const long long row = (((long long)bt * O + o) * Lg + lg) * ch + g;
    stf(out, out_dt, i, ldf(Y, y_dt, row * Cp + r * S + s));
  }
}


// ---------------------------------------------------------------------------------------------
// Morph-FC permutation, block-staged version.  For a fixed (bt, other-axis position o, chunk group lg)
// the ch x C source elements map onto ONE contiguous run of ch x Cp destination elements
// (rows (..)*ch + g, g = 0..ch-1), so a warp stages the block in shared memory: coalesced reads of
// C-element pixel rows, in-smem (r, g, s) -> (g, r, s) transpose, coalesced writes.  inverse = scatter.
// ---------------------------------------------------------------------------------------------
constexpr int MORPH_WARPS = 4;
template <bool VEC>
__global__ void __launch_bounds__(32 * MORPH_WARPS) morph_block_kernel(const void* src, int src_dt, void* dst, int dst_dt, int BT, int H,
                                                                      int W, int C, int ch, int axis, int inverse, long long n_blocks, int a_pitch) {
  extern __shared__ float sm_all[];
  const int Cp = (C + ch - 1) / ch * ch, S = Cp / ch;
  const int L = axis == 0 ? H : W, Lg = (L + ch - 1) / ch, O = axis == 0 ? W : H;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float* sm = sm_all + wid * (ch * Cp);
  const long long blk = (long long)blockIdx.x * (blockDim.x >> 5) + wid;
  if (blk >= n_blocks) return;
  // block index = (bt*O + o)*Lg + lg  (same order as the row index / ch)
  const int lg = (int)(blk % Lg); const long long t = blk / Lg; const int o = (int)(t % O); const int bt = (int)(t / O);
  const long long rows0 = blk * ch;                     // first A row of this block
  constexpr int V = VEC ? 4 : 1;
  const int Cv = C / V, Cpv = Cp / V;
  if (!inverse) {
    for (int i = lane; i < ch * Cp; i += 32) sm[i] = 0.f;
    __syncwarp();
    for (int i = lane; i < ch * Cv; i += 32) {
      const int r = i / Cv, c = (i - r * Cv) * V;
      const int l = lg * ch + r;
      if (l < L) {
        const int h = axis == 0 ? l : o, w = axis == 0 ? o : l;
        const long long gi = (((long long)bt * H + h) * W + w) * C + c;
        if (VEC) {
          const float4 v = ld4(src, src_dt, gi);
          const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) { const int g = (c + k) / S, s2 = (c + k) - g * S; sm[g * Cp + r * S + s2] = vv[k]; }
        } else {
          const int g = c / S, s2 = c - g * S;
          sm[g * Cp + r * S + s2] = ldf(src, src_dt, gi);
        }
      }
    }
    __syncwarp();
    for (int i = lane; i < ch * Cpv; i += 32) {
      const int rr = (V * i) / Cp, cc = (V * i) - rr * Cp;
      if (VEC) st4(dst, dst_dt, (rows0 + rr) * a_pitch + cc, *reinterpret_cast<const float4*>(sm + 4 * i));
      else stf(dst, dst_dt, (rows0 + rr) * a_pitch + cc, sm[i]);
    }
  } else {
    for (int i = lane; i < ch * Cpv; i += 32) {
      const int rr = (V * i) / Cp, cc = (V * i) - rr * Cp;
      if (VEC) *reinterpret_cast<float4*>(sm + 4 * i) = ld4(src, src_dt, (rows0 + rr) * a_pitch + cc);
      else sm[i] = ldf(src, src_dt, (rows0 + rr) * a_pitch + cc);
    }
    __syncwarp();
    for (int i = lane; i < ch * Cv; i += 32) {
      const int r = i / Cv, c = (i - r * Cv) * V;
      const int l = lg * ch + r;
      if (l < L) {
        const int h = axis == 0 ? l : o, w = axis == 0 ? o : l;
        const long long gi = (((long long)bt * H + h) * W + w) * C + c;
        if (VEC) {
          float vv[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) { const int g = (c + k) / S, s2 = (c + k) - g * S; vv[k] = sm[g * Cp + r * S + s2]; }
          st4(dst, dst_dt, gi, make_float4(vv[0], vv[1], vv[2], vv[3]));
        } else {
          const int g = c / S, s2 = c - g * S;
          stf(dst, dst_dt, gi, sm[g * Cp + r * S + s2]);
        }
      }
    }
  }
}

// bf16 fast path of the block-staged permutation (C, Cp, a_pitch multiples of 8; 16-byte aligned tensors): both
// global sides move 16 bytes per lane, the (r, g, s) <-> (g, r, s) shuffle happens on 2-byte elements in a bf16
// shared-memory copy of the block, and the index arithmetic is incremental (the fp32-staged kernel above spends
// ~4 integer divisions per 8 bytes and reached 0.6-1.2 TB/s).  Warps walk the blocks grid-stride.
__global__ void __launch_bounds__(32 * MORPH_WARPS) morph_block16_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                                        int BT, int H, int W, int C, int ch, int axis, int inverse,
                                                                        long long n_blocks, int a_pitch) {
  extern __shared__ uint4 sm_vec_all[];
  const int Cp = (C + ch - 1) / ch * ch, S = Cp / ch;
  const int L = axis == 0 ? H : W, Lg = (L + ch - 1) / ch, O = axis == 0 ? W : H;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int Cv = C >> 3, Cpv = Cp >> 3;
  const int per_warp_vec = ch * (Cv > Cpv ? Cv : Cpv);
  uint4* smv = sm_vec_all + wid * per_warp_vec;
  const unsigned short* sme = reinterpret_cast<const unsigned short*>(smv);
  const uint32_t* smw32 = reinterpret_cast<const uint32_t*>(smv);
  const bool s_even = (S & 1) == 0 && (C & 1) == 0;
  const float inv_S = 1.f / (float)S, inv_Cv = 1.f / (float)Cv, inv_Cpv = 1.f / (float)Cpv;
  const long long wstride = (long long)gridDim.x * MORPH_WARPS;
  for (long long wi = (long long)blockIdx.x * MORPH_WARPS + wid; wi < n_blocks; wi += wstride) {
    // block index = (bt*O + o)*Lg + lg  (same order as the row index / ch).  Work order: for axis 0 the pixel rows of a
    // block are W*C apart, so neighbouring warps take neighbouring o (= w) and the pixel side of the traffic is one
    // contiguous run per image row (measured neutral on B200: the kernel is paced by the instructions of the 2-byte
    // element shuffle, not by DRAM locality; rotating the reads to dodge their 4-way bank conflicts made it slower).
    int lg, o, bt;
    const unsigned wu = (unsigned)wi;                  // the host caps n_blocks below 2^31: 32-bit divisions
    if (axis == 0) { o = (int)(wu % (unsigned)O); const unsigned t = wu / (unsigned)O; lg = (int)(t % (unsigned)Lg); bt = (int)(t / (unsigned)Lg); }
    else { lg = (int)(wu % (unsigned)Lg); const unsigned t = wu / (unsigned)Lg; o = (int)(t % (unsigned)O); bt = (int)(t / (unsigned)O); }
    const long long blk = ((long long)bt * O + o) * Lg + lg;
    const long long rows0 = blk * ch;
    if (!inverse) {
      // 1. ch pixel rows -> sm[r*C + c]
      for (int i = lane; i < ch * Cv; i += 32) {
        const int r = (int)(((float)i + 0.5f) * inv_Cv), c8 = i - r * Cv;
        const int l = lg * ch + r;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (l < L) {
          const int h = axis == 0 ? l : o, w = axis == 0 ? o : l;
          v = __ldg(reinterpret_cast<const uint4*>(src + (((long long)bt * H + h) * W + w) * C + 8 * c8));
        }
        smv[i] = v;
      }
      __syncwarp();
      // 2. output vectors: element d = g*Cp + r*S + s  <-  sm[r*C + g*S + s]
      for (int i = lane; i < ch * Cpv; i += 32) {
        const int g = (int)(((float)i + 0.5f) * inv_Cpv), f = 8 * (i - g * Cpv);
        int r = (int)(((float)f + 0.5f) * inv_S), s2 = f - r * S;
        uint4 v;
        if (s_even) {            // S even: element pairs never straddle a segment -> move 32-bit words
          uint32_t wv[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int c = g * S + s2;
            wv[k] = c < C ? smw32[(r * C + c) >> 1] : 0u;
            s2 += 2;
            if (s2 == S) { s2 = 0; ++r; }
          }
          v = make_uint4(wv[0], wv[1], wv[2], wv[3]);
        } else {
          unsigned short e[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int c = g * S + s2;
            e[k] = c < C ? sme[r * C + c] : (unsigned short)0;
            if (++s2 == S) { s2 = 0; ++r; }
          }
          v = make_uint4((uint32_t)e[0] | ((uint32_t)e[1] << 16), (uint32_t)e[2] | ((uint32_t)e[3] << 16),
                         (uint32_t)e[4] | ((uint32_t)e[5] << 16), (uint32_t)e[6] | ((uint32_t)e[7] << 16));
        }
        *reinterpret_cast<uint4*>(dst + (rows0 + g) * a_pitch + f) = v;
      }
      __syncwarp();
    } else {
      // 1. ch token rows -> sm[g*Cp + f]
      for (int i = lane; i < ch * Cpv; i += 32) {
        const int g = (int)(((float)i + 0.5f) * inv_Cpv), f8 = i - g * Cpv;
        smv[i] = __ldg(reinterpret_cast<const uint4*>(src + (rows0 + g) * a_pitch + 8 * f8));
      }
      __syncwarp();
      // 2. pixel-row vectors: element (r, c), c = g*S + s  <-  sm[g*Cp + r*S + s]
      for (int i = lane; i < ch * Cv; i += 32) {
        const int r = (int)(((float)i + 0.5f) * inv_Cv), c0 = 8 * (i - r * Cv);
        const int l = lg * ch + r;
        if (l < L) {
          int g = (int)(((float)c0 + 0.5f) * inv_S), s2 = c0 - g * S;
          uint4 v;
          if (s_even) {
            uint32_t wv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              wv[k] = smw32[(g * Cp + r * S + s2) >> 1];
              s2 += 2;
              if (s2 == S) { s2 = 0; ++g; }
            }
            v = make_uint4(wv[0], wv[1], wv[2], wv[3]);
          } else {
            unsigned short e[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              e[k] = sme[g * Cp + r * S + s2];
              if (++s2 == S) { s2 = 0; ++g; }
            }
            v = make_uint4((uint32_t)e[0] | ((uint32_t)e[1] << 16), (uint32_t)e[2] | ((uint32_t)e[3] << 16),
                           (uint32_t)e[4] | ((uint32_t)e[5] << 16), (uint32_t)e[6] | ((uint32_t)e[7] << 16));
          }
          const int h = axis == 0 ? l : o, w = axis == 0 ? o : l;
          *reinterpret_cast<uint4*>(dst + (((long long)bt * H + h) * W + w) * C + c0) = v;
        }
      }
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// grouped channel sums (atomic accumulation of per-block partials into a zeroed fp32 buffer)
// grid = (row chunks, groups); block = 256 threads = 8 row lanes x 32 channel lanes
// ---------------------------------------------------------------------------------------------
// Partials are accumulated as 2^-24 fixed point in 64-bit integers: integer addition is associative, so the result does
// not depend on the order in which the blocks' atomics land (float atomics made the whole forward differ in the last
// bits from run to run).
constexpr float CSUM_FIX = 16777216.f;
__global__ void __launch_bounds__(256) channel_sum_kernel(const void* a, const void* b, const void* c, int dt, unsigned long long* acc,
                                                          int n_groups, long long rows_per_group, int C, int rows_per_block) {
  __shared__ float part[8][33];
  // blockIdx.y = segment*n_groups + group ; rows_per_group here = rows of ONE segment
  const int g = blockIdx.y % n_groups;
  const long long seg_base = (long long)blockIdx.y * rows_per_group;
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(rows_per_group, r0 + rows_per_block);
  for (int c0 = 0; c0 < C; c0 += 32) {
    const int ch = c0 + cl;
    float s = 0.f;
    if (ch < C) {
      for (long long r = r0 + rl; r < r1; r += 8) {
        const long long idx = (seg_base + r) * C + ch;
        float v = ldf(a, dt, idx);
        if (b) v += ldf(b, dt, idx);
        if (c) v += ldf(c, dt, idx);
        s += v;
      }
    }
    part[rl][cl] = s;
    __syncthreads();
    if (rl == 0 && ch < C) {
      float tot = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) tot += part[k][cl];
      atomicAdd(acc + (long long)g * C + ch, (unsigned long long)__float2ll_rn(tot * CSUM_FIX));
    }
    __syncthreads();
  }
}

// bf16, C % 8 == 0: 16 bytes (8 channels) per thread and row, C/8 channel lanes x 256/(C/8) row lanes per block.  The scalar
// kernel above issues one 2-byte load per element and reached 1.4 TB/s (profiles/r2_suite_ncu.md); same fixed-point
// accumulation, so the result stays independent of the order of the atomics.
__global__ void __launch_bounds__(256) channel_sum_vec_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, const uint4* __restrict__ c,
                                                              unsigned long long* acc, int n_groups, long long rows_per_group, int C8,
                                                              int rows_per_block) {
  extern __shared__ float part_v[];             // [row lanes][C]
  const int g = blockIdx.y % n_groups;
  const long long seg_base = (long long)blockIdx.y * rows_per_group;
  const int rlanes = 256 / C8;
  const int cl = threadIdx.x % C8, rl = threadIdx.x / C8;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(rows_per_group, r0 + rows_per_block);
  float s[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] = 0.f;
  if (rl < rlanes) {
    for (long long r = r0 + rl; r < r1; r += rlanes) {
      const long long idx = (seg_base + r) * C8 + cl;
      const uint4 va = __ldg(a + idx);
      uint4 vb = make_uint4(0, 0, 0, 0), vc = vb;
      if (b) vb = __ldg(b + idx);
      if (c) vc = __ldg(c + idx);
      const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w}, wc[4] = {vc.x, vc.y, vc.z, vc.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        s[2 * k] += (__uint_as_float(wa[k] << 16) + __uint_as_float(wb[k] << 16)) + __uint_as_float(wc[k] << 16);
        s[2 * k + 1] += (__uint_as_float(wa[k] & 0xffff0000u) + __uint_as_float(wb[k] & 0xffff0000u)) + __uint_as_float(wc[k] & 0xffff0000u);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) part_v[rl * (C8 * 8) + cl * 8 + k] = s[k];
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < C8 * 8; ch += 256) {
    float tot = 0.f;
    for (int k = 0; k < rlanes; ++k) tot += part_v[k * (C8 * 8) + ch];
    atomicAdd(acc + (long long)g * (C8 * 8) + ch, (unsigned long long)__float2ll_rn(tot * CSUM_FIX));
  }
}

__global__ void channel_sum_finalize_kernel(const unsigned long long* __restrict__ acc, float* __restrict__ out, int n, float scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)((double)(long long)acc[i] * (1.0 / 16777216.0) * (double)scale);
}

__global__ void softmax3_kernel(const float* __restrict__ logits, float* __restrict__ a, int B, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i - b * C;
  const float* l = logits + (long long)b * 3 * C + 3 * c;
  const float m = fmaxf(l[0], fmaxf(l[1], l[2]));
  const float e0 = expf(l[0] - m), e1 = expf(l[1] - m), e2 = expf(l[2] - m);
  const float inv = 1.f / (e0 + e1 + e2);
  float* o = a + (long long)b * 3 * C + c;
  o[0] = e0 * inv; o[C] = e1 * inv; o[2 * C] = e2 * inv;
}

__global__ void mix3_kernel(const void* h, const void* w, const void* c, int dt, const float* __restrict__ a, void* out, int out_dt,
                            int B, long long rows_per_b, int C, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % C);
    const long long row = i / C;
    const int b = (int)((row / rows_per_b) % B);
    const float* ab = a + (long long)b * 3 * C + ch;
    stf(out, out_dt, i, ldf(h, dt, i) * ab[0] + ldf(w, dt, i) * ab[C] + ldf(c, dt, i) * ab[2 * C]);
  }
}

__global__ void mix3_vec_kernel(const void* h, const void* w, const void* c, int dt, const float* __restrict__ a, void* out, int out_dt,
                                int B, long long rows_per_b, int C4, long long total4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    const long long row = i / C4;
    const int b = (int)((row / rows_per_b) % B);
    const float* ab = a + (long long)b * 12 * C4 + 4 * c4;
    const float4 a0 = *reinterpret_cast<const float4*>(ab), a1 = *reinterpret_cast<const float4*>(ab + 4 * C4),
                 a2 = *reinterpret_cast<const float4*>(ab + 8 * C4);
    const float4 hv = ld4(h, dt, i * 4), wv = ld4(w, dt, i * 4), cv = ld4(c, dt, i * 4);
    st4(out, out_dt, i * 4, make_float4(hv.x * a0.x + wv.x * a1.x + cv.x * a2.x, hv.y * a0.y + wv.y * a1.y + cv.y * a2.y,
                                        hv.z * a0.z + wv.z * a1.z + cv.z * a2.z, hv.w * a0.w + wv.w * a1.w + cv.w * a2.w));
  }
}

__global__ void ca_apply_vec_kernel(const void* u, const void* x, int dt, const float* __restrict__ y, void* out, int out_dt,
                                    long long rows_per_frame, int C4, float scale, long long total4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    const long long fr = (i / C4) / rows_per_frame;
    const float4 yv = *reinterpret_cast<const float4*>(y + (fr * C4 + c4) * 4);
    const float4 uv = ld4(u, dt, i * 4), xv = ld4(x, dt, i * 4);
    st4(out, out_dt, i * 4, make_float4((uv.x * yv.x + xv.x) * scale, (uv.y * yv.y + xv.y) * scale, (uv.z * yv.z + xv.z) * scale,
                                        (uv.w * yv.w + xv.w) * scale));
  }
}

__global__ void ca_apply_kernel(const void* u, const void* x, int dt, const float* __restrict__ y, void* out, int out_dt,
                                long long rows_per_frame, int C, float scale, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % C);
    const long long fr = (i / C) / rows_per_frame;
    stf(out, out_dt, i, (ldf(u, dt, i) * y[fr * C + ch] + ldf(x, dt, i)) * scale);
  }
}

__global__ void add_kernel(const void* a, const void* b, int dt, void* out, int out_dt, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    stf(out, out_dt, i, ldf(a, dt, i) + ldf(b, dt, i));
}

// out[n, h < H, w < W, :] = a + b over [n, Hp, Wp, C] bf16 operands (the last skip connection of the U-Net, vmg.py:547, and the
// crop back to the input size, vmg.py:619, in one pass; 16 bytes per thread)
__global__ void __launch_bounds__(256) add_crop_bf16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ out,
                                                            int Hp, int Wp, int H, int W, int C8, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8); long long t = i / C8;
    const int w = (int)(t % W); t /= W;
    const int h = (int)(t % H); const long long n = t / H;
    const long long src = ((n * Hp + h) * Wp + w) * C8 + c;
    const uint4 va = __ldg(a + src), vb = __ldg(b + src);
    const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __nv_bfloat162 hsum = __floats2bfloat162_rn(__uint_as_float(wa[k] << 16) + __uint_as_float(wb[k] << 16),
                                                        __uint_as_float(wa[k] & 0xffff0000u) + __uint_as_float(wb[k] & 0xffff0000u));
      o[k] = *reinterpret_cast<const uint32_t*>(&hsum);
    }
    out[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void decay_kernel(float* __restrict__ w, const float* __restrict__ g, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    w[i] = w[i] * g[i];
}

// ---------------------------------------------------------------------------------------------
// resampling gather + LayerNorm (layers.py:777-798): one warp per OUTPUT row
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) resample_ln_kernel(const void* x, const void* add, int dt, void* out, int out_dt, int out_pitch,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          int BT, int H, int W, int C, int up, float eps) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int Ho = up ? 2 * H : H / 2, Wo = up ? 2 * W : W / 2, F = up ? C / 4 : 4 * C;
  if (row >= (long long)BT * Ho * Wo) return;
  const int wo = (int)(row % Wo); const long long t = row / Wo; const int ho = (int)(t % Ho); const int bt = (int)(t / Ho);
  float v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int f = lane + 32 * i;
    v[i] = 0.f;
    if (f >= F) continue;
    long long idx;
    if (!up) {   // f = (neiw*2 + neih)*C + c
      const int q = f / C, c = f - q * C, neiw = q >> 1, neih = q & 1;
      idx = (((long long)bt * H + (2 * ho + neih)) * W + (2 * wo + neiw)) * C + c;
    } else {     // source channel (neiw*2 + neih)*(C/4) + f
      const int neih = ho & 1, neiw = wo & 1;
      idx = (((long long)bt * H + (ho >> 1)) * W + (wo >> 1)) * C + (neiw * 2 + neih) * F + f;
    }
    float val = ldf(x, dt, idx);
    if (add) val += ldf(add, dt, idx);
    v[i] = val; s += val;
  }
  const float mean = warp_sum(s) / (float)F;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) { const int f = lane + 32 * i; const float d = f < F ? v[i] - mean : 0.f; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) / (float)F + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) { const int f = lane + 32 * i; if (f < F) stf(out, out_dt, row * out_pitch + f, (v[i] - mean) * rstd * gamma[f] + beta[f]); }
}

// vectorised variant (C % 4 == 0, F % 4 == 0): lane owns features [4*lane + 128*i, +4)
__global__ void __launch_bounds__(256) resample_ln_vec_kernel(const void* x, const void* add, int dt, void* out, int out_dt, int out_pitch,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              int BT, int H, int W, int C, int up, float eps) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int Ho = up ? 2 * H : H / 2, Wo = up ? 2 * W : W / 2, F = up ? C / 4 : 4 * C;
  if (row >= (long long)BT * Ho * Wo) return;
  const int wo = (int)(row % Wo); const long long t = row / Wo; const int ho = (int)(t % Ho); const int bt = (int)(t / Ho);
  float4 v[MAXQ];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXQ; ++i) {
    const int f = 4 * lane + 128 * i;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (f >= F) continue;
    long long idx;
    if (!up) {
      const int q = f / C, c = f - q * C, neiw = q >> 1, neih = q & 1;
      idx = (((long long)bt * H + (2 * ho + neih)) * W + (2 * wo + neiw)) * C + c;
    } else {
      const int neih = ho & 1, neiw = wo & 1;
      idx = (((long long)bt * H + (ho >> 1)) * W + (wo >> 1)) * C + (neiw * 2 + neih) * F + f;
    }
    float4 val = ld4(x, dt, idx);
    if (add) { const float4 a = ld4(add, dt, idx); val.x += a.x; val.y += a.y; val.z += a.z; val.w += a.w; }
    v[i] = val;
    s += (val.x + val.y) + (val.z + val.w);
  }
  const float mean = warp_sum(s) / (float)F;
  float q2 = 0.f;
#pragma unroll
  for (int i = 0; i < MAXQ; ++i) {
    if (4 * lane + 128 * i < F) {
      const float a = v[i].x - mean, b = v[i].y - mean, c2 = v[i].z - mean, d = v[i].w - mean;
      q2 += (a * a + b * b) + (c2 * c2 + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q2) / (float)F + eps);
#pragma unroll
  for (int i = 0; i < MAXQ; ++i) {
    const int f = 4 * lane + 128 * i;
    if (f < F) {
      const float4 g = *reinterpret_cast<const float4*>(gamma + f), b = *reinterpret_cast<const float4*>(beta + f);
      st4(out, out_dt, row * out_pitch + f, make_float4((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y,
                                                        (v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w));
    }
  }
}

// up-sampling rows are short (F = C/4 = 56 or 112 features): LPR lanes own an output row, 16 bytes (8 channels) per lane, so a
// warp handles 32/LPR rows per pass instead of one row on 14 of its 32 lanes (resample_ln_vec_kernel: 0.2 of the HBM peak)
template <int LPR>
__global__ void __launch_bounds__(256) resample_ln_up16_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ add,
                                                               __nv_bfloat16* __restrict__ out, int out_pitch, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, int BT, int H, int W, int C, float eps) {
  const int F = C / 4, Ho = 2 * H, Wo = 2 * W;
  const int hl = threadIdx.x & (LPR - 1);
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const long long rows = (long long)BT * Ho * Wo;
  const bool rok = row < rows;                     // whole LPR-lane groups share it; every lane takes part in the shuffles
  const long long rc = rok ? row : rows - 1;
  const int wo = (int)(rc % Wo); const long long t = rc / Wo; const int ho = (int)(t % Ho); const int bt = (int)(t / Ho);
  const int f = hl * 8;
  const bool act = f < F;
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = 0.f;
  if (act) {
    const long long idx = (((long long)bt * H + (ho >> 1)) * W + (wo >> 1)) * C + ((wo & 1) * 2 + (ho & 1)) * F + f;
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(x + idx));
    const uint32_t w4[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[2 * k] = __uint_as_float(w4[k] << 16); v[2 * k + 1] = __uint_as_float(w4[k] & 0xffff0000u); }
    if (add) {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(add + idx));
      const uint32_t a4[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) { v[2 * k] += __uint_as_float(a4[k] << 16); v[2 * k + 1] += __uint_as_float(a4[k] & 0xffff0000u); }
    }
  }
  float s = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
#pragma unroll
  for (int m = LPR / 2; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  const float mean = s / (float)F;
  float q2 = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) { v[k] -= mean; q2 += v[k] * v[k]; }
  q2 = act ? q2 : 0.f;
#pragma unroll
  for (int m = LPR / 2; m >= 1; m >>= 1) q2 += __shfl_xor_sync(0xffffffffu, q2, m);
  const float rstd = rsqrtf(q2 / (float)F + eps);
  if (act && rok) {
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + f), g1 = *reinterpret_cast<const float4*>(gamma + f + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(beta + f), b1 = *reinterpret_cast<const float4*>(beta + f + 4);
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * k] * rstd * g[2 * k] + b[2 * k], v[2 * k + 1] * rstd * g[2 * k + 1] + b[2 * k + 1]);
      o[k] = *reinterpret_cast<const uint32_t*>(&h2);
    }
    *reinterpret_cast<uint4*>(out + row * out_pitch + f) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void maxpool4_kernel(const void* x, int dt, void* out, int out_dt, int BT, int H, int W, int C, long long total) {
  const int Ho = H / 4, Wo = W / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C); long long p = i / C;
    const int wo = (int)(p % Wo); p /= Wo; const int ho = (int)(p % Ho); const int bt = (int)(p / Ho);
    float m = -INFINITY;
    for (int dy = 0; dy < 4; ++dy)
      for (int dx = 0; dx < 4; ++dx)
        m = fmaxf(m, ldf(x, dt, (((long long)bt * H + 4 * ho + dy) * W + 4 * wo + dx) * C + c));
    stf(out, out_dt, i, m);
  }
}

// per-frame (sum, sum of squares) -> stats[frame][2]; grid = (chunks, frames)
// (2^-20 fixed point in 64-bit integers, see channel_sum_kernel: order-independent, hence run-to-run identical)
__global__ void __launch_bounds__(256) frame_stats_kernel(const void* x, int dt, unsigned long long* stats, long long elems, int elems_per_block) {
  __shared__ float sh[2][8];
  const int fr = blockIdx.y;
  const long long e0 = (long long)blockIdx.x * elems_per_block, e1 = min(elems, e0 + elems_per_block);
  float s = 0.f, q = 0.f;
  for (long long e = e0 + threadIdx.x; e < e1; e += blockDim.x) { const float v = ldf(x, dt, (long long)fr * elems + e); s += v; q += v * v; }
  s = warp_sum(s); q = warp_sum(q);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ts = 0.f, tq = 0.f;
    for (int k = 0; k < 8; ++k) { ts += sh[0][k]; tq += sh[1][k]; }
    atomicAdd(stats + 2 * fr, (unsigned long long)__float2ll_rn(ts * 1048576.f));
    atomicAdd(stats + 2 * fr + 1, (unsigned long long)__float2ll_rn(tq * 1048576.f));
  }
}

__global__ void groupnorm1_relu_kernel(const void* x, int dt, const long long* __restrict__ stats, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, void* out, int out_dt, long long rows_per_frame, int C,
                                       float eps, long long total) {
  const float cnt = (float)(rows_per_frame * C);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long fr = (i / C) / rows_per_frame;
    const float mean = (float)((double)stats[2 * fr] * (1.0 / 1048576.0)) / cnt;
    const float var = fmaxf((float)((double)stats[2 * fr + 1] * (1.0 / 1048576.0)) / cnt - mean * mean, 0.f);
    const float v = (ldf(x, dt, i) - mean) * rsqrtf(var + eps) * gamma[c] + beta[c];
    stf(out, out_dt, i, fmaxf(v, 0.f));
  }
}

inline int grid_for(long long total, int threads = 256, int cap = 148 * 16) {
  long long b = (total + threads - 1) / threads;
  if (b < 1) b = 1;
  return (int)(b > cap ? cap : b);
}

}  // namespace vmgb

using namespace vmgb;

extern "C" int vmgb_layernorm(const void* x, int x_dt, void* out, int out_dt, const void* gamma, const void* beta,
                              long long rows, int C, float eps, void* stream) {
  if (!x || !out || !gamma || !beta || !dt_ok(x_dt) || !dt_ok(out_dt) || C <= 0 || C > 32 * MAXV || rows <= 0) return VMGB_EINVAL;
  long long blocks = (rows + 7) / 8;
  const bool vec = (C & 3) == 0 && !((uintptr_t)x & 15) && !((uintptr_t)out & 15) && !((uintptr_t)gamma & 15) && !((uintptr_t)beta & 15);
  if (vec && C <= 256 && (C & 7) == 0 && x_dt == VMGB_BF16 && out_dt == VMGB_BF16) {
    launch_rownorm16<0>(x, C, out, (const float*)gamma, (const float*)beta, rows, C, eps, as_stream(stream));
  } else if (vec && C <= 128) {
    const long long wanted = (rows + 8 * 4 - 1) / (8 * 4);
    const unsigned nb = (unsigned)(wanted < 148 * 8 ? (wanted < 1 ? 1 : wanted) : 148 * 8);
    layernorm_narrow_kernel<4><<<nb, 256, 0, as_stream(stream)>>>(x, x_dt, out, out_dt, (const float*)gamma, (const float*)beta, rows, C, eps);
  } else if (vec) layernorm_vec_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, x_dt, out, out_dt, (const float*)gamma, (const float*)beta, rows, C, eps);
  else
  layernorm_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, x_dt, out, out_dt, (const float*)gamma, (const float*)beta, rows, C, eps);
  return check_launch();
}

extern "C" int vmgb_l2norm(const void* x, int x_dt, int x_pitch, void* out, int out_dt, long long rows, int C, void* stream) {
  if (!x || !out || !dt_ok(x_dt) || !dt_ok(out_dt) || C <= 0 || C > 32 * MAXV || rows <= 0) return VMGB_EINVAL;
  long long blocks = (rows + 7) / 8;
  const bool vec = (C & 3) == 0 && (x_pitch & 3) == 0 && !((uintptr_t)x & 15) && !((uintptr_t)out & 15);
  if (vec && C <= 256 && (C & 7) == 0 && (x_pitch & 7) == 0 && x_dt == VMGB_BF16 && out_dt == VMGB_BF16)
    launch_rownorm16<1>(x, x_pitch, out, nullptr, nullptr, rows, C, 0.f, as_stream(stream));
  else if (vec) l2norm_vec_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, x_dt, x_pitch, out, out_dt, rows, C);
  else
  l2norm_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, x_dt, x_pitch, out, out_dt, rows, C);
  return check_launch();
}

extern "C" int vmgb_morph_permute(const void* src, int src_dt, void* dst, int dst_dt, int BT, int H, int W, int C, int chunk,
                                  int axis, int inverse, int a_pitch, void* stream) {
  if (!src || !dst || !dt_ok(src_dt) || !dt_ok(dst_dt) || chunk <= 0 || (axis != 0 && axis != 1)) return VMGB_EINVAL;
  const int Cp = (C + chunk - 1) / chunk * chunk;
  if (a_pitch < Cp) return VMGB_EINVAL;
  const int L = axis == 0 ? H : W, Lp = (L + chunk - 1) / chunk * chunk, O = axis == 0 ? W : H;
  const long long n_blocks_all = (long long)BT * O * (Lp / chunk);
  if (src_dt == VMGB_BF16 && dst_dt == VMGB_BF16 && (C & 7) == 0 && (Cp & 7) == 0 && (a_pitch & 7) == 0 &&
      !(((uintptr_t)src | (uintptr_t)dst) & 15) && (size_t)chunk * Cp * 2 * MORPH_WARPS <= 48 * 1024 && Cp < 4096 && n_blocks_all < 2147483647LL) {
    const size_t smem = (size_t)chunk * Cp * 2 * MORPH_WARPS;        // C <= Cp
    const long long wanted = (n_blocks_all + MORPH_WARPS - 1) / MORPH_WARPS;
    const unsigned grid = (unsigned)(wanted < 148 * 16 ? wanted : 148 * 16);
    morph_block16_kernel<<<grid, 32 * MORPH_WARPS, smem, as_stream(stream)>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, BT, H, W, C, chunk,
                                                                            axis, inverse, n_blocks_all, a_pitch);
    return check_launch();
  }
  const size_t per_warp = (size_t)chunk * Cp * sizeof(float);
  int warps = (int)((48 * 1024) / per_warp);
  if (warps > MORPH_WARPS) warps = MORPH_WARPS;
  if (warps >= 1) {
    const size_t smem = per_warp * warps;
    const long long n_blocks = (long long)BT * O * (Lp / chunk);
    const long long grid = (n_blocks + warps - 1) / warps;
    if (grid > 2147483647LL) return VMGB_EINVAL;
    const bool vec = (C & 3) == 0 && (Cp & 3) == 0 && (a_pitch & 3) == 0 && !(((uintptr_t)src | (uintptr_t)dst) & 15);
    if (vec) morph_block_kernel<true><<<(unsigned)grid, 32 * warps, smem, as_stream(stream)>>>(src, src_dt, dst, dst_dt, BT, H, W, C, chunk, axis,
                                                                                               inverse, n_blocks, a_pitch);
    else
    morph_block_kernel<false><<<(unsigned)grid, 32 * warps, smem, as_stream(stream)>>>(src, src_dt, dst, dst_dt, BT, H, W, C, chunk, axis,
                                                                                      inverse, n_blocks, a_pitch);
    return check_launch();
  }
  if (a_pitch != Cp) return VMGB_EUNSUPPORTED;
  if (!inverse) {
    long long total = (long long)BT * Lp * O * Cp;
    morph_gather_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(src, src_dt, dst, dst_dt, BT, H, W, C, chunk, axis, total);
  } else {
    long long total = (long long)BT * H * W * C;
    morph_scatter_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(src, src_dt, dst, dst_dt, BT, H, W, C, chunk, axis, total);
  }
  return check_launch();
}

extern "C" int vmgb_channel_sum(const void* a, const void* b, const void* c, int dt, float* out, long long* acc, int n_seg, int groups,
                                long long seg_rows, int C, float scale, void* stream) {
  if (!a || !out || !acc || !dt_ok(dt) || groups <= 0 || n_seg <= 0 || seg_rows <= 0 || C <= 0) return VMGB_EINVAL;
  if ((long long)n_seg * groups > 65535) return VMGB_EINVAL;
  int rpb = 256;
  long long chunks = (seg_rows + rpb - 1) / rpb;
  while (chunks * groups * n_seg > 148LL * 32 && rpb < (1 << 20)) { rpb *= 2; chunks = (seg_rows + rpb - 1) / rpb; }
  dim3 grid((unsigned)chunks, n_seg * groups);
  const bool vec = dt == VMGB_BF16 && (C & 7) == 0 && C <= 2048 && !((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15);
  if (vec) {
    const int rlanes = 256 / (C / 8);
    channel_sum_vec_kernel<<<grid, 256, (size_t)rlanes * C * sizeof(float), as_stream(stream)>>>((const uint4*)a, (const uint4*)b, (const uint4*)c,
                                                                                                  (unsigned long long*)acc, groups, seg_rows, C / 8, rpb);
  } else
  channel_sum_kernel<<<grid, 256, 0, as_stream(stream)>>>(a, b, c, dt, (unsigned long long*)acc, groups, seg_rows, C, rpb);
  channel_sum_finalize_kernel<<<(groups * C + 255) / 256, 256, 0, as_stream(stream)>>>((const unsigned long long*)acc, out, groups * C, scale);
  return check_launch();
}

extern "C" int vmgb_softmax3(const float* logits, float* a, int B, int C, void* stream) {
  if (!logits || !a || B <= 0 || C <= 0) return VMGB_EINVAL;
  softmax3_kernel<<<(B * C + 255) / 256, 256, 0, as_stream(stream)>>>(logits, a, B, C);
  return check_launch();
}

extern "C" int vmgb_mix3(const void* h, const void* w, const void* c, int dt, const float* a, void* out, int out_dt, int n_seg,
                         int B, long long seg_rows, int C, void* stream) {
  if (!h || !w || !c || !a || !out || !dt_ok(dt) || !dt_ok(out_dt) || n_seg <= 0 || B <= 0) return VMGB_EINVAL;
  long long total = (long long)n_seg * B * seg_rows * C;
  if ((C & 3) == 0 && !(((uintptr_t)h | (uintptr_t)w | (uintptr_t)c | (uintptr_t)out | (uintptr_t)a) & 15))
    mix3_vec_kernel<<<grid_for(total / 4, 256, 148 * 32), 256, 0, as_stream(stream)>>>(h, w, c, dt, a, out, out_dt, B, seg_rows, C / 4, total / 4);
  else
  mix3_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(h, w, c, dt, a, out, out_dt, B, seg_rows, C, total);
  return check_launch();
}

extern "C" int vmgb_ca_apply(const void* u, const void* x, int dt, const float* y, void* out, int out_dt, int frames,
                             long long rows_per_frame, int C, float scale, void* stream) {
  if (!u || !x || !y || !out || !dt_ok(dt) || !dt_ok(out_dt)) return VMGB_EINVAL;
  long long total = (long long)frames * rows_per_frame * C;
  if ((C & 3) == 0 && !(((uintptr_t)u | (uintptr_t)x | (uintptr_t)y | (uintptr_t)out) & 15))
    ca_apply_vec_kernel<<<grid_for(total / 4, 256, 148 * 32), 256, 0, as_stream(stream)>>>(u, x, dt, y, out, out_dt, rows_per_frame, C / 4, scale, total / 4);
  else
  ca_apply_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(u, x, dt, y, out, out_dt, rows_per_frame, C, scale, total);
  return check_launch();
}

extern "C" int vmgb_pair_sweeps(const void* x, void* out, int T, long long frame_bytes, void* stream) {
  if (!x || !out || T <= 0 || frame_bytes <= 0 || (frame_bytes & 15) || ((uintptr_t)x & 15) || ((uintptr_t)out & 15)) return VMGB_EINVAL;
  const long long fv = frame_bytes / 16, total = (long long)T * fv;
  const long long wanted = (total + 1023) / 1024;
  pair_sweeps_kernel<<<(unsigned)(wanted < 148 * 8 ? wanted : 148 * 8), 256, 0, as_stream(stream)>>>((const uint4*)x, (uint4*)out, T, fv);
  return check_launch();
}

extern "C" int vmgb_unpair_sweeps(const void* paired, void* back, void* fwd, int T, long long frame_bytes, void* stream) {
  if (!paired || !back || !fwd || T <= 0 || frame_bytes <= 0 || (frame_bytes & 15) || ((uintptr_t)paired & 15) || ((uintptr_t)back & 15) ||
      ((uintptr_t)fwd & 15))
    return VMGB_EINVAL;
  const long long fv = frame_bytes / 16, total = 2LL * T * fv;
  const long long wanted = (total + 1023) / 1024;
  unpair_sweeps_kernel<<<(unsigned)(wanted < 148 * 8 ? wanted : 148 * 8), 256, 0, as_stream(stream)>>>((const uint4*)paired, (uint4*)back,
                                                                                                     (uint4*)fwd, T, fv);
  return check_launch();
}

extern "C" int vmgb_add(const void* a, const void* b, int dt, void* out, int out_dt, long long n, void* stream) {
  if (!a || !b || !out || !dt_ok(dt) || !dt_ok(out_dt) || n <= 0) return VMGB_EINVAL;
  add_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(a, b, dt, out, out_dt, n);
  return check_launch();
}

extern "C" int vmgb_add_crop(const void* a, const void* b, void* out, int n_img, int Hp, int Wp, int H, int W, int C, void* stream) {
  if (!a || !b || !out || n_img < 1 || H < 1 || W < 1 || H > Hp || W > Wp || C < 8) return VMGB_EINVAL;
  if ((C & 7) || ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15)) return VMGB_EUNSUPPORTED;
  const long long total = (long long)n_img * H * W * (C / 8);
  add_crop_bf16_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>((const uint4*)a, (const uint4*)b, (uint4*)out, Hp, Wp, H, W, C / 8, total);
  return check_launch();
}

extern "C" int vmgb_decay_weights(float* w, const float* gamma, long long n, void* stream) {
  if (!w || !gamma || n <= 0) return VMGB_EINVAL;
  decay_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(w, gamma, n);
  return check_launch();
}

extern "C" int vmgb_resample_ln(const void* x, const void* add, int dt, void* out, int out_dt, int out_pitch, const void* gamma,
                                const void* beta, int BT, int H, int W, int C, int up, float eps, void* stream) {
  if (!x || !out || !gamma || !beta || !dt_ok(dt) || !dt_ok(out_dt)) return VMGB_EINVAL;
  if (up ? (C & 3) : ((H & 1) || (W & 1))) return VMGB_EINVAL;
  const int F = up ? C / 4 : 4 * C;
  if (F > 32 * MAXV || out_pitch < F) return VMGB_EINVAL;
  const long long rows = up ? (long long)BT * 4 * H * W : (long long)BT * (H / 2) * (W / 2);
  const bool rvec = (C & 3) == 0 && (F & 3) == 0 && (out_pitch & 3) == 0 && F <= 128 * MAXQ &&
                    !(((uintptr_t)x | (uintptr_t)out | (uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)add) & 15);
  if (rvec && up && dt == VMGB_BF16 && out_dt == VMGB_BF16 && (F & 7) == 0 && F <= 128 && (out_pitch & 7) == 0 && (C & 7) == 0) {
    const int lpr = F <= 64 ? 8 : 16;
    const long long blocks = (rows * lpr + 255) / 256;
    if (lpr == 8) resample_ln_up16_kernel<8><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)add,
        (__nv_bfloat16*)out, out_pitch, (const float*)gamma, (const float*)beta, BT, H, W, C, eps);
    else resample_ln_up16_kernel<16><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)add,
        (__nv_bfloat16*)out, out_pitch, (const float*)gamma, (const float*)beta, BT, H, W, C, eps);
    return check_launch();
  }
  if (rvec) resample_ln_vec_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, as_stream(stream)>>>(x, add, dt, out, out_dt, out_pitch, (const float*)gamma,
                                                                                   (const float*)beta, BT, H, W, C, up, eps);
  else
  resample_ln_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, as_stream(stream)>>>(x, add, dt, out, out_dt, out_pitch, (const float*)gamma,
                                                                               (const float*)beta, BT, H, W, C, up, eps);
  return check_launch();
}

extern "C" int vmgb_maxpool4(const void* x, int dt, void* out, int out_dt, int BT, int H, int W, int C, void* stream) {
  if (!x || !out || !dt_ok(dt) || !dt_ok(out_dt) || (H & 3) || (W & 3)) return VMGB_EINVAL;
  long long total = (long long)BT * (H / 4) * (W / 4) * C;
  maxpool4_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(x, dt, out, out_dt, BT, H, W, C, total);
  return check_launch();
}

extern "C" int vmgb_frame_stats(const void* x, int dt, long long* stats, int frames, long long elems_per_frame, void* stream) {
  if (!x || !stats || !dt_ok(dt) || frames <= 0 || elems_per_frame <= 0) return VMGB_EINVAL;
  int epb = 8192;
  dim3 grid((unsigned)((elems_per_frame + epb - 1) / epb), frames);
  frame_stats_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, dt, (unsigned long long*)stats, elems_per_frame, epb);
  return check_launch();
}

extern "C" int vmgb_groupnorm1_relu(const void* x, int dt, const long long* stats, const void* gamma, const void* beta, void* out,
                                    int out_dt, int frames, long long rows_per_frame, int C, float eps, void* stream) {
  if (!x || !stats || !gamma || !beta || !out || !dt_ok(dt) || !dt_ok(out_dt)) return VMGB_EINVAL;
  long long total = (long long)frames * rows_per_frame * C;
  groupnorm1_relu_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(x, dt, stats, (const float*)gamma, (const float*)beta, out,
                                                                        out_dt, rows_per_frame, C, eps, total);
  return check_launch();
}
