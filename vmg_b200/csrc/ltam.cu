// ltam.cu -- LTAM 'wins' trajectory attention (trajectory.py:672-795, cal_pe :534-547; per-pixel
// restatement SURVEY.md A.5).  HBM/L2-bound gather kernel: O(K) key-frame reads per step.
//
// Work split: one thread per (2x2 window, head, query position).  The 16 threads of a window sit in
// one half-warp; the 4 query threads of a head read the SAME key/value segment (hardware broadcast).
// Keys arrive L2-normalised (normalising at store time commutes with the integer gather), queries
// arrive L2-normalised; logits, softmax (online) and the weighted sum are fp32.
#include "common.cuh"

namespace vmgb {

// packed FP32x2 (Blackwell FFMA2: one issue slot per TWO fp32 lanes, the same rounding as scalar fma.rn)
__device__ __forceinline__ unsigned long long lt_pack(float a, float b) {
  return (unsigned long long)__float_as_uint(a) | ((unsigned long long)__float_as_uint(b) << 32);
}
__device__ __forceinline__ unsigned long long lt_fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ unsigned long long lt_mul2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// a 32-bit word holding two bf16 -> the packed pair (element 2k, element 2k+1) as fp32
__device__ __forceinline__ unsigned long long lt_unpack(uint32_t w) {
  return (unsigned long long)(w << 16) | ((unsigned long long)(w & 0xffff0000u) << 32);
}

template <int D4>
__global__ void __launch_bounds__(128) ltam_wins_kernel(const void* q, const void* keys, const void* vals, int dt, long long kv_stride,
                                                        const float* __restrict__ loc, int loc_pitch, const float* __restrict__ rel,
                                                        const float* __restrict__ decay, void* out, int out_dt, int n, int h, int w,
                                                        int C, int heads, int K, long long n_windows, const float* __restrict__ obias,
                                                        const void* anchor) {
  constexpr int D = D4 * 4;
  const long long win = (long long)blockIdx.x * 8 + (threadIdx.x >> 4);
  if (win >= n_windows) return;
  const int sub = threadIdx.x & 15;
  const int head = sub >> 2, qi = sub & 3;
  const int wh = h >> 1, ww = w >> 1;
  const int wx = (int)(win % ww); const long long t = win / ww; const int wy = (int)(t % wh); const int img = (int)(t / wh);
  const long long pix0 = ((long long)img * h + 2 * wy) * w + 2 * wx;    // top-left pixel of the window
  const long long qpix = pix0 + (qi >> 1) * w + (qi & 1);
  const float scale = rsqrtf((float)D);

  float qv[D], acc[D];
#pragma unroll
  for (int i = 0; i < D4; ++i) {
    const float4 v = ld4(q, dt, qpix * C + head * D + i * 4);
    qv[4 * i] = v.x * scale; qv[4 * i + 1] = v.y * scale; qv[4 * i + 2] = v.z * scale; qv[4 * i + 3] = v.w * scale;
    acc[4 * i] = acc[4 * i + 1] = acc[4 * i + 2] = acc[4 * i + 3] = 0.f;
  }
  const float dh = decay[head];
  float relq[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) relq[j] = rel[(head * 4 + qi) * 4 + j];

  float m = -INFINITY, l = 0.f, pe = 1.f;
  for (int k = K - 1; k >= 0; --k) {
    pe *= dh;                                   // decay^(K-k), built by the same multiply chain as cal_pe
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long pj = pix0 + (j >> 1) * w + (j & 1);
      const float2 lc = *reinterpret_cast<const float2*>(loc + pj * loc_pitch + 2 * k);
      const int sx = (int)rintf(lc.x), sy = (int)rintf(lc.y);
      const bool inb = sx >= 0 && sx < w && sy >= 0 && sy < h;    // padding_mode='zeros'
      const long long src = (long long)k * kv_stride + (((long long)img * h + sy) * w + sx) * C + head * D;
      float dot = 0.f;
      if (inb) {
#pragma unroll
        for (int i = 0; i < D4; ++i) {
          const float4 kv = ld4(keys, dt, src + i * 4);
          dot += kv.x * qv[4 * i] + kv.y * qv[4 * i + 1] + kv.z * qv[4 * i + 2] + kv.w * qv[4 * i + 3];
        }
      }
      const float logit = dot + pe * relq[j];
      const float mn = fmaxf(m, logit);
      const float corr = __expf(m - mn), p = __expf(logit - mn);
      l = l * corr + p;
      m = mn;
#pragma unroll
      for (int i = 0; i < D4; ++i) {
        float4 vv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (inb) vv = ld4(vals, dt, src + i * 4);
        acc[4 * i] = acc[4 * i] * corr + p * vv.x;
        acc[4 * i + 1] = acc[4 * i + 1] * corr + p * vv.y;
        acc[4 * i + 2] = acc[4 * i + 2] * corr + p * vv.z;
        acc[4 * i + 3] = acc[4 * i + 3] * corr + p * vv.w;
      }
    }
  }
  const float inv = 1.f / l;
#pragma unroll
  for (int i = 0; i < D4; ++i) {
    float4 o = make_float4(acc[4 * i] * inv, acc[4 * i + 1] * inv, acc[4 * i + 2] * inv, acc[4 * i + 3] * inv);
    if (obias) { const float4 b4 = *reinterpret_cast<const float4*>(obias + head * D + i * 4); o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w; }
    if (anchor) { const float4 a4 = ld4(anchor, out_dt, qpix * C + head * D + i * 4); o.x += a4.x; o.y += a4.y; o.z += a4.z; o.w += a4.w; }
    st4(out, out_dt, qpix * C + head * D + i * 4, o);
  }
}


// ---------------------------------------------------------------------------------------------
// bf16 variant with shared-memory staging.  A block owns 8 windows; per key frame the 32 gathered key rows
// and 32 value rows (C bf16 each) are fetched with 16-byte cp.async (coalesced inside a row, double buffered
// across key frames) and every (head, query) thread then reads its head segment from shared memory.  The
// direct version above issues 8-byte loads per thread and re-reads each row through L1 for the 4 queries.
// ---------------------------------------------------------------------------------------------
constexpr int LTAM_STAGES = 3;      // measured: 3 stages (47 KB, 4 blocks per SM) beat 4 stages (62 KB, 3 blocks) by ~8 %
template <int D4>
__global__ void __launch_bounds__(128) ltam_wins_smem_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ keys,
                                                             const __nv_bfloat16* __restrict__ vals, long long kv_stride,
                                                             const float* __restrict__ loc, int loc_pitch, const float* __restrict__ rel,
                                                             const float* __restrict__ decay, void* out, int out_dt, int n, int h, int w,
                                                             int K, long long n_windows, const float* __restrict__ obias, const void* anchor) {
  constexpr int D = D4 * 4, C = 4 * D, ROWB = C * 2, CH = ROWB / 16;     // bytes / 16-byte chunks per row
  constexpr int NST = LTAM_STAGES;
  extern __shared__ __align__(16) uint8_t sm[];                           // [NST stages][2 (k,v)][32 rows][ROWB], then src_s[32][K]
  int* src_s = reinterpret_cast<int*>(sm + NST * (2 * 32 * ROWB));        // gathered source pixel of (row, key frame), -1 = outside
  const int tid = threadIdx.x;
  const long long win0 = (long long)blockIdx.x * 8;
  const int wl = tid >> 4, sub = tid & 15, head = sub >> 2, qi = sub & 3;
  const long long win = win0 + wl;
  const bool active = win < n_windows;
  const int wh = h >> 1, ww = w >> 1;
  const long long wc = active ? win : n_windows - 1;
  const int wx = (int)(wc % ww); const long long t0 = wc / ww; const int wy = (int)(t0 % wh); const int img = (int)(t0 / wh);
  const long long pix0 = ((long long)img * h + 2 * wy) * w + 2 * wx;
  const long long qpix = pix0 + (qi >> 1) * w + (qi & 1);
  const float scale = rsqrtf((float)D);

  // the sampling locations of all key frames first (one dependent global round trip per block instead of one per
  // key frame): src_s[r*K + k] = pixel index inside the key frame, or -1 (padding_mode='zeros')
  for (int idx = tid; idx < 32 * K; idx += 128) {
    const int r = idx / K, k = idx - r * K;                  // r = window-local row: (window r>>2, position r&3)
    long long wn = win0 + (r >> 2);
    if (wn >= n_windows) wn = n_windows - 1;
    const int x2 = (int)(wn % ww); const long long t2 = wn / ww; const int y2 = (int)(t2 % wh); const int i2 = (int)(t2 / wh);
    const long long pj = ((long long)i2 * h + 2 * y2 + ((r & 3) >> 1)) * w + 2 * x2 + (r & 1);
    const float2 lc = *reinterpret_cast<const float2*>(loc + pj * loc_pitch + 2 * k);
    const int sx = (int)rintf(lc.x), sy = (int)rintf(lc.y);
    src_s[idx] = (sx >= 0 && sx < w && sy >= 0 && sy < h) ? (int)(((long long)i2 * h + sy) * w + sx) : -1;
  }

  // queries and accumulators as packed fp32 pairs (element 2k, 2k+1)
  unsigned long long qv2[D / 2], acc2[D / 2];
#pragma unroll
  for (int i = 0; i < D4; ++i) {
    const float4 v = ld4(q, VMGB_BF16, qpix * C + head * D + i * 4);
    qv2[2 * i] = lt_pack(v.x * scale, v.y * scale);
    qv2[2 * i + 1] = lt_pack(v.z * scale, v.w * scale);
    acc2[2 * i] = 0ull; acc2[2 * i + 1] = 0ull;
  }
  const float dh = decay[head];
  float relq[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) relq[j] = rel[(head * 4 + qi) * 4 + j];
  __syncthreads();

  // stage one key frame: 32 rows x CH chunks for keys and the same for values
  auto stage = [&](int k, int buf) {
    uint8_t* base = sm + buf * (2 * 32 * ROWB);
    for (int c = tid; c < 32 * CH; c += 128) {
      const int r = c / CH, cc = c - r * CH;
      const int sp = src_s[r * K + k];
      const uint32_t dk = (uint32_t)__cvta_generic_to_shared(base + r * ROWB + cc * 16);
      const uint32_t dv = dk + 32 * ROWB;
      if (sp >= 0) {
        const long long src = (long long)k * kv_stride + (long long)sp * C + cc * 8;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dk), "l"(keys + src) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dv), "l"(vals + src) : "memory");
      } else {
        asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(dk), "r"(0) : "memory");
        asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(dv), "r"(0) : "memory");
      }
    }
  };

  float m = -INFINITY, l = 0.f, pe = 1.f;
  // NST-deep ring over the key frames (newest first); one commit group per iteration, empty groups at the tail
#pragma unroll
  for (int sidx = 0; sidx < NST - 1; ++sidx) {
    if (K - 1 - sidx >= 0) stage(K - 1 - sidx, sidx);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  int buf = 0;
  for (int k = K - 1; k >= 0; --k) {
    asm volatile("cp.async.wait_group %0;" ::"n"(NST - 2) : "memory");      // key frame k has landed (this thread's part)
    __syncthreads();                                        // ... everyone's part; and everyone is done with the buffer of k+1
    {
      const int kn = k - (NST - 1);
      int bn = buf + NST - 1; if (bn >= NST) bn -= NST;
      if (kn >= 0) stage(kn, bn);                            // re-fills the buffer consumed in the previous iteration
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    pe *= dh;
    const uint32_t rowk = (uint32_t)__cvta_generic_to_shared(sm + buf * (2 * 32 * ROWB) + (wl * 4) * ROWB + head * D * 2);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t kr = rowk + j * ROWB, vr = kr + 32 * ROWB;
      // dot product on packed pairs: 8-byte shared loads (4 bf16), two fused multiply-adds per load
      unsigned long long d2 = 0ull;
#pragma unroll
      for (int i = 0; i < D4; ++i) {
        uint32_t w0, w1;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w0), "=r"(w1) : "r"(kr + i * 8));
        d2 = lt_fma2(lt_unpack(w0), qv2[2 * i], d2);
        d2 = lt_fma2(lt_unpack(w1), qv2[2 * i + 1], d2);
      }
      const float logit = __uint_as_float((uint32_t)d2) + __uint_as_float((uint32_t)(d2 >> 32)) + pe * relq[j];
      // online softmax with a LAZY reference point: keys and queries are L2-normalised over the full channel vector, so the
      // logits of one pixel span a few units at most; the running reference m is only moved (and the accumulators rescaled)
      // when a logit exceeds it by more than 16 -- in practice once, at the first key (m = -inf).  exp(logit - m) / sum is
      // the same softmax as with the true maximum.
      if (logit > m + 16.f) {
        const float corr = __expf(m - logit);
        const unsigned long long c2 = lt_pack(corr, corr);
        l *= corr;
#pragma unroll
        for (int i = 0; i < D / 2; ++i) acc2[i] = lt_mul2(acc2[i], c2);
        m = logit;
      }
      const float pw = __expf(logit - m);
      l += pw;
      const unsigned long long p2 = lt_pack(pw, pw);
#pragma unroll
      for (int i = 0; i < D4; ++i) {
        uint32_t w0, w1;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w0), "=r"(w1) : "r"(vr + i * 8));
        acc2[2 * i] = lt_fma2(lt_unpack(w0), p2, acc2[2 * i]);
        acc2[2 * i + 1] = lt_fma2(lt_unpack(w1), p2, acc2[2 * i + 1]);
      }
    }
    if (++buf == NST) buf = 0;
  }
  if (!active) return;
  const float inv = 1.f / l;
#pragma unroll
  for (int i = 0; i < D4; ++i) {
    float4 o = make_float4(__uint_as_float((uint32_t)acc2[2 * i]) * inv, __uint_as_float((uint32_t)(acc2[2 * i] >> 32)) * inv,
                           __uint_as_float((uint32_t)acc2[2 * i + 1]) * inv, __uint_as_float((uint32_t)(acc2[2 * i + 1] >> 32)) * inv);
    // values that were projected at store time (proj is linear and the softmax weights sum to 1): bias and anchor here
    if (obias) { const float4 b4 = *reinterpret_cast<const float4*>(obias + head * D + i * 4); o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w; }
    if (anchor) { const float4 a4 = ld4(anchor, out_dt, qpix * C + head * D + i * 4); o.x += a4.x; o.y += a4.y; o.z += a4.z; o.w += a4.w; }
    st4(out, out_dt, qpix * C + head * D + i * 4, o);
  }
}

}  // namespace vmgb

using namespace vmgb;

extern "C" int vmgb_ltam_wins(const void* q, const void* keys, const void* vals, int dt, long long kv_stride, const float* loc,
                              int loc_pitch, const float* rel, const float* decay, void* out, int out_dt, int n, int h, int w, int C,
                              int heads, int K, const float* out_bias, const void* anchor, void* stream) {
  if (!q || !keys || !vals || !loc || !rel || !decay || !out || !dt_ok(dt) || !dt_ok(out_dt)) return VMGB_EINVAL;
  if (heads != 4 || K <= 0 || (h & 1) || (w & 1) || C % (4 * heads) || 2 * K > loc_pitch) return VMGB_EINVAL;
  if (out_bias && (reinterpret_cast<uintptr_t>(out_bias) & 15)) return VMGB_EINVAL;
  const long long n_windows = (long long)n * (h / 2) * (w / 2);
  const unsigned blocks = (unsigned)((n_windows + 7) / 8);
  const int D4 = C / heads / 4;
  if (dt == VMGB_BF16 && !(((uintptr_t)keys | (uintptr_t)vals) & 15) && (kv_stride & 7) == 0 && (D4 == 7 || D4 == 9)) {
    const int smem = LTAM_STAGES * 2 * 32 * C * 2 + 32 * K * 4;
    if (smem > 200 * 1024) return VMGB_EUNSUPPORTED;
#define LAUNCH_S(DD)                                                                                                          \
    if (smem > 48 * 1024) cudaFuncSetAttribute(ltam_wins_smem_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);         \
    ltam_wins_smem_kernel<DD><<<blocks, 128, smem, as_stream(stream)>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)keys,  \
        (const __nv_bfloat16*)vals, kv_stride, loc, loc_pitch, rel, decay, out, out_dt, n, h, w, K, n_windows, out_bias, anchor)
    if (D4 == 7) { LAUNCH_S(7); } else { LAUNCH_S(9); }
#undef LAUNCH_S
    return check_launch();
  }
#define LAUNCH(DD) ltam_wins_kernel<DD><<<blocks, 128, 0, as_stream(stream)>>>(q, keys, vals, dt, kv_stride, loc, loc_pitch, rel, decay, out, out_dt, n, h, w, C, heads, K, n_windows, out_bias, anchor)
  switch (D4) {
    case 4: LAUNCH(4); break;    // C = 64
    case 7: LAUNCH(7); break;    // C = 112 (VMG-REDS)
    case 8: LAUNCH(8); break;    // C = 128
    case 9: LAUNCH(9); break;    // C = 144 (few_levels)
    default: return VMGB_EUNSUPPORTED;
  }
#undef LAUNCH
  return check_launch();
}
