// ltam.cu -- LTAM 'wins' trajectory attention (trajectory.py:672-795, cal_pe :534-547; per-pixel
// restatement SURVEY.md A.5).  HBM/L2-bound gather kernel: O(K) key-frame reads per step.
//
// Work split: one thread per (2x2 window, head, query position).  The 16 threads of a window sit in
// one half-warp; the 4 query threads of a head read the SAME key/value segment (hardware broadcast).
// Keys arrive L2-normalised (normalising at store time commutes with the integer gather), queries
// arrive L2-normalised; logits, softmax (online) and the weighted sum are fp32.
#include "common.cuh"

namespace vmgb {

template <int D4>
__global__ void __launch_bounds__(128) ltam_wins_kernel(const void* q, const void* keys, const void* vals, int dt, long long kv_stride,
                                                        const float* __restrict__ loc, int loc_pitch, const float* __restrict__ rel,
                                                        const float* __restrict__ decay, void* out, int out_dt, int n, int h, int w,
                                                        int C, int heads, int K, long long n_windows) {
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
  for (int i = 0; i < D4; ++i)
    st4(out, out_dt, qpix * C + head * D + i * 4,
        make_float4(acc[4 * i] * inv, acc[4 * i + 1] * inv, acc[4 * i + 2] * inv, acc[4 * i + 3] * inv));
}

}  // namespace vmgb

using namespace vmgb;

extern "C" int vmgb_ltam_wins(const void* q, const void* keys, const void* vals, int dt, long long kv_stride, const float* loc,
                              int loc_pitch, const float* rel, const float* decay, void* out, int out_dt, int n, int h, int w, int C,
                              int heads, int K, void* stream) {
  if (!q || !keys || !vals || !loc || !rel || !decay || !out || !dt_ok(dt) || !dt_ok(out_dt)) return VMGB_EINVAL;
  if (heads != 4 || K <= 0 || (h & 1) || (w & 1) || C % (4 * heads) || 2 * K > loc_pitch) return VMGB_EINVAL;
  const long long n_windows = (long long)n * (h / 2) * (w / 2);
  const unsigned blocks = (unsigned)((n_windows + 7) / 8);
  const int D4 = C / heads / 4;
#define LAUNCH(DD) ltam_wins_kernel<DD><<<blocks, 128, 0, as_stream(stream)>>>(q, keys, vals, dt, kv_stride, loc, loc_pitch, rel, decay, out, out_dt, n, h, w, C, heads, K, n_windows)
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
