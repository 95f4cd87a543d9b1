// evalpath.cu -- the callers either side of the model forward, on device:
//   frame I/O      uint8 THWC frames <-> fp32 [T,3,H,W] in [0,1]          (utils/eval_utils.py:8-36, tools/Tester.py:217-222,249-250)
//   tile blending  overlapped spatial tiles / temporal windows of Tester   (tools/Tester.py:107-177)
//   metrics        PSNR, RGB -> Y, 11x11 Gaussian SSIM                      (utils/metrics.py:11-70, tools/test_reds4.py:194-218)
//   augmentation   flips / transpose of the self-ensemble route            (tools/Tester.py:387-445)
// All of it is HBM-bound byte / fp32 / fp64 work: coalesced grid-stride kernels, fp64 where the reference computes in
// float64 (numpy), integer accumulation where the reference's sums are exact integers.
#include "common.cuh"

namespace vmgb {

static inline int grid_for_e(long long total, int threads = 256, int cap = 148 * 16) {
  long long b = (total + threads - 1) / threads;
  if (b < 1) b = 1;
  return (int)(b > cap ? cap : b);
}

// ---------------------------------------------------------------------------------------------
// frame I/O
// ---------------------------------------------------------------------------------------------
// u8 [T,H,W,3] -> f32 [T,3,H,W] = u8 / 255 (numpy: astype(float32) / 255., one correctly rounded fp32 division).
// reverse = 1 swaps channel order (cv2 BGR <-> RGB, utils/eval_utils.py:19).  One thread per pixel: 3 bytes in, 3 planes out.
__global__ void frames_u8_to_f32_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, long long hw, int reverse, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long t = i / hw, p = i - t * hw;
    const uint8_t* s = src + i * 3;
    float* d = dst + t * 3 * hw + p;
#pragma unroll
    for (int c = 0; c < 3; ++c) d[c * hw] = __fdiv_rn((float)s[reverse ? 2 - c : c], 255.0f);
  }
}

// quantisation of Tester.evaluate (tools/Tester.py:249-250): clamp(0,1) -> fp32 * 255.0 -> np.round (half to even) -> uint8
__device__ __forceinline__ uint8_t quant_u8(float v) {
  v = fminf(fmaxf(v, 0.f), 1.f);
  return (uint8_t)__float2int_rn(__fmul_rn(v, 255.0f));
}

// f32 [T,3,H,W] -> u8 [T,H,W,3]
__global__ void frames_f32_to_u8_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, long long hw, int reverse, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long t = i / hw, p = i - t * hw;
    const float* s = src + t * 3 * hw + p;
    uint8_t* d = dst + i * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) d[reverse ? 2 - c : c] = quant_u8(s[c * hw]);
  }
}

// ---------------------------------------------------------------------------------------------
// tile / window blending (Tester.test_image / test_clips): E[region] += patch * mask, Wm[region] += mask
// patch [B,Tp,C,ph,pw] (contiguous); E [B,Te,C,He,We]; Wm is a count plane [Te,He,We] (the reference keeps a full
// tensor of ones whose channels and batch entries are all equal).  The mask zeroes `zt0` leading / `zt1` trailing frames,
// `zh0`/`zh1` rows and `zw0`/`zw1` columns of the patch.  The reference multiplies the masked-out patch samples by 0 and
// adds them (E + 0*v = E unless v is inf/nan); here they are skipped.
// ---------------------------------------------------------------------------------------------
struct BlendP {
  int B, Tp, C, ph, pw, Te, He, We, t0, h0, w0, zt0, zt1, zh0, zh1, zw0, zw1;
};
__global__ void blend_accumulate_kernel(const float* __restrict__ patch, float* __restrict__ E, float* __restrict__ Wm, BlendP q, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % q.pw); long long r = i / q.pw;
    const int y = (int)(r % q.ph); r /= q.ph;
    const int c = (int)(r % q.C); r /= q.C;
    const int t = (int)(r % q.Tp); const int b = (int)(r / q.Tp);
    const bool on = t >= q.zt0 && t < q.Tp - q.zt1 && y >= q.zh0 && y < q.ph - q.zh1 && x >= q.zw0 && x < q.pw - q.zw1;
    if (!on) continue;
    const long long e = ((((long long)b * q.Te + (q.t0 + t)) * q.C + c) * q.He + (q.h0 + y)) * q.We + (q.w0 + x);
    E[e] += patch[i];
    if (Wm && b == 0 && c == 0) Wm[((long long)(q.t0 + t) * q.He + (q.h0 + y)) * q.We + (q.w0 + x)] += 1.0f;
  }
}

// E /= Wm (broadcast over batch and channels); fp32 division as torch's E.div_(W)
__global__ void blend_normalize_kernel(float* __restrict__ E, const float* __restrict__ Wm, int C, int Te, long long hw, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i % hw; long long r = i / hw;
    r /= C;
    const int t = (int)(r % Te);
    E[i] = __fdiv_rn(E[i], Wm[(long long)t * hw + p]);
  }
}

// fused tail of Tester.evaluate: (E / Wm) -> clamp -> *255 -> round -> uint8 THWC (B = 1); Wm may be NULL
__global__ void blend_finalize_u8_kernel(const float* __restrict__ E, const float* __restrict__ Wm, uint8_t* __restrict__ dst, long long hw,
                                         int reverse, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long t = i / hw, p = i - t * hw;
    const float* s = E + t * 3 * hw + p;
    const float wv = Wm ? Wm[i] : 1.0f;
    uint8_t* d = dst + i * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) d[reverse ? 2 - c : c] = quant_u8(Wm ? __fdiv_rn(s[c * hw], wv) : s[c * hw]);
  }
}

// ---------------------------------------------------------------------------------------------
// augmentation (Tester.augment / augment_inverse): hflip -> vflip -> transpose(H, W), in that order, on [N,H,W] planes
// ---------------------------------------------------------------------------------------------
__global__ void augment_kernel(const float* __restrict__ src, float* __restrict__ dst, int H, int W, int hflip, int vflip, int tr, long long total) {
  const int Ho = tr ? W : H, Wo = tr ? H : W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xo = (int)(i % Wo); long long r = i / Wo;
    const int yo = (int)(r % Ho); const long long pl = r / Ho;
    int y = tr ? xo : yo, x = tr ? yo : xo;            // coordinates in the flipped (pre-transpose) image
    if (vflip) y = H - 1 - y;
    if (hflip) x = W - 1 - x;
    dst[i] = src[(pl * H + y) * W + x];
  }
}

// ---------------------------------------------------------------------------------------------
// metrics
// ---------------------------------------------------------------------------------------------
// sum of squared differences of two uint8 arrays, exact in 64-bit integers (calculate_psnr: the float64 mean of integer
// squares is exact, so PSNR = 20 log10(255 / sqrt(sse / n)) is bit-reproducible on the host)
__global__ void __launch_bounds__(256) sse_u8_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, long long n, unsigned long long* out) {
  unsigned long long s = 0;
  const long long n16 = n / 16;
  const uint4* a4 = reinterpret_cast<const uint4*>(a);
  const uint4* b4 = reinterpret_cast<const uint4*>(b);
  const bool vec = ((((uintptr_t)a) | ((uintptr_t)b)) & 15) == 0;
  if (vec) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) {
      const uint4 x = a4[i], y = b4[i];
      const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
      uint32_t acc = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int d = (int)((xs[k] >> (8 * j)) & 255u) - (int)((ys[k] >> (8 * j)) & 255u);
          acc += (uint32_t)(d * d);
        }
      }
      s += acc;
    }
  }
  for (long long i = (vec ? n16 * 16 : 0) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)a[i] - (int)b[i];
    s += (unsigned long long)(d * d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

// sum of squared differences of two fp32 arrays in fp64 (test_clips_max picks the window with the best PSNR per frame)
__global__ void __launch_bounds__(256) sse_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, int clamp01, double* partial) {
  __shared__ double sh[8];
  double s = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float x = a[i], y = b[i];
    if (clamp01) { x = fminf(fmaxf(x, 0.f), 1.f); y = fminf(fmaxf(y, 0.f), 1.f); }
    const double d = (double)x - (double)y;
    s += d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0.0; for (int k = 0; k < 8; ++k) t += sh[k]; partial[blockIdx.x] = t; }
}

// deterministic final reduction of per-block partials (fixed order)
__global__ void reduce_partials_kernel(const double* __restrict__ partial, int n, double scale, double* out) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) *out = sh[0] * scale;
}

// Y of skimage.color.rgb2ycbcr on uint8 RGB (tools/test_reds4.py:208-209): float64
//   Y = 16 + 65.481 r + 128.553 g + 24.966 b,  (r, g, b) = uint8 / 255
__global__ void rgb_to_y_kernel(const uint8_t* __restrict__ rgb, double* __restrict__ y, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const uint8_t* s = rgb + i * 3;
    const double inv = 1.0 / 255.0;      // skimage img_as_float: uint8 * (1 / 255)
    const double r = (double)s[0] * inv, g = (double)s[1] * inv, b = (double)s[2] * inv;
    y[i] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(r, 65.481), __dmul_rn(g, 128.553)), __dmul_rn(b, 24.966)), 16.0);
  }
}

__global__ void __launch_bounds__(256) sse_f64_kernel(const double* __restrict__ a, const double* __restrict__ b, long long n, double* partial) {
  __shared__ double sh[8];
  double s = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double d = a[i] - b[i];
    s += d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0.0; for (int k = 0; k < 8; ++k) t += sh[k]; partial[blockIdx.x] = t; }
}

// SSIM (utils/metrics.py:33-70): 11x11 Gaussian window (sigma 1.5), valid region only, float64.
// One block = a 16 x 16 tile of the (H-10) x (W-10) SSIM map of one plane: the 26 x 26 input patches of both images go to
// shared memory, the window is applied separably (rows, then columns) to x, y, x^2, y^2, xy.
// Planes are addressed as base + pixel * pix_stride (+ channel offset by the caller): interleaved uint8 RGB (stride 3) or a
// float64 Y plane (stride 1).
struct SsimP {
  double w[11];
  int H, W, tiles_x, tiles_y;
};
template <typename T>
__global__ void __launch_bounds__(256) ssim_kernel(const T* __restrict__ a, const T* __restrict__ b, int pix_stride, int n_planes, int plane_step,
                                                   long long img_stride, SsimP q, double* partial) {
  __shared__ double sa[26][27], sb[26][27];
  __shared__ double h[5][26][17];
  __shared__ double red[8];
  const int tile = blockIdx.x % (q.tiles_x * q.tiles_y), plane = blockIdx.x / (q.tiles_x * q.tiles_y);
  const int img = plane / n_planes, ch = plane % n_planes;
  const int ty0 = (tile / q.tiles_x) * 16, tx0 = (tile % q.tiles_x) * 16;
  const T* pa = a + (long long)img * img_stride + (long long)ch * plane_step;
  const T* pb = b + (long long)img * img_stride + (long long)ch * plane_step;
  for (int i = threadIdx.x; i < 26 * 26; i += 256) {
    const int r = i / 26, c = i % 26;
    const int y = min(ty0 + r, q.H - 1), x = min(tx0 + c, q.W - 1);
    sa[r][c] = (double)pa[((long long)y * q.W + x) * pix_stride];
    sb[r][c] = (double)pb[((long long)y * q.W + x) * pix_stride];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 26 * 16; i += 256) {
    const int r = i / 16, c = i % 16;
    double m1 = 0, m2 = 0, s11 = 0, s22 = 0, s12 = 0;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const double x = sa[r][c + k], y = sb[r][c + k], wk = q.w[k];
      m1 += wk * x; m2 += wk * y; s11 += wk * (x * x); s22 += wk * (y * y); s12 += wk * (x * y);
    }
    h[0][r][c] = m1; h[1][r][c] = m2; h[2][r][c] = s11; h[3][r][c] = s22; h[4][r][c] = s12;
  }
  __syncthreads();
  double val = 0.0;
  {
    const int r = threadIdx.x / 16, c = threadIdx.x % 16;
    if (ty0 + r < q.H - 10 && tx0 + c < q.W - 10) {
      double m1 = 0, m2 = 0, s11 = 0, s22 = 0, s12 = 0;
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        const double wk = q.w[k];
        m1 += wk * h[0][r + k][c]; m2 += wk * h[1][r + k][c]; s11 += wk * h[2][r + k][c]; s22 += wk * h[3][r + k][c]; s12 += wk * h[4][r + k][c];
      }
      const double C1 = (0.01 * 255) * (0.01 * 255), C2 = (0.03 * 255) * (0.03 * 255);
      const double m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
      const double v1 = s11 - m11, v2 = s22 - m22, v12 = s12 - m12;
      val = ((2 * m12 + C1) * (2 * v12 + C2)) / ((m11 + m22 + C1) * (v1 + v2 + C2));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = val;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0.0; for (int k = 0; k < 8; ++k) t += red[k]; partial[blockIdx.x] = t; }
}

// per-image mean over the planes of one image: out[img] = sum(partial of that image) / (n_planes * valid pixels)
__global__ void ssim_finish_kernel(const double* __restrict__ partial, int blocks_per_img, double scale, double* out) {
  __shared__ double sh[256];
  const double* p = partial + (long long)blockIdx.x * blocks_per_img;
  double s = 0.0;
  for (int i = threadIdx.x; i < blocks_per_img; i += 256) s += p[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) out[blockIdx.x] = sh[0] * scale;
}

}  // namespace vmgb

using namespace vmgb;

extern "C" int vmgb_frames_u8_to_f32(const void* frames_u8, float* out, int T, int H, int W, int reverse_channels, void* stream) {
  if (!frames_u8 || !out || T < 1 || H < 1 || W < 1) return VMGB_EINVAL;
  const long long total = (long long)T * H * W;
  frames_u8_to_f32_kernel<<<grid_for_e(total), 256, 0, as_stream(stream)>>>((const uint8_t*)frames_u8, out, (long long)H * W, reverse_channels, total);
  return check_launch();
}

extern "C" int vmgb_frames_f32_to_u8(const float* frames, void* out_u8, int T, int H, int W, int reverse_channels, void* stream) {
  if (!frames || !out_u8 || T < 1 || H < 1 || W < 1) return VMGB_EINVAL;
  const long long total = (long long)T * H * W;
  frames_f32_to_u8_kernel<<<grid_for_e(total), 256, 0, as_stream(stream)>>>(frames, (uint8_t*)out_u8, (long long)H * W, reverse_channels, total);
  return check_launch();
}

extern "C" int vmgb_blend_accumulate(const float* patch, float* E, float* count, int B, int Tp, int C, int ph, int pw, int Te, int He, int We,
                                     int t0, int h0, int w0, int zt0, int zt1, int zh0, int zh1, int zw0, int zw1, void* stream) {
  if (!patch || !E || B < 1 || Tp < 1 || C < 1 || ph < 1 || pw < 1) return VMGB_EINVAL;
  if (t0 < 0 || h0 < 0 || w0 < 0 || t0 + Tp > Te || h0 + ph > He || w0 + pw > We) return VMGB_EINVAL;
  if (zt0 < 0 || zt1 < 0 || zh0 < 0 || zh1 < 0 || zw0 < 0 || zw1 < 0) return VMGB_EINVAL;
  BlendP q{B, Tp, C, ph, pw, Te, He, We, t0, h0, w0, zt0, zt1, zh0, zh1, zw0, zw1};
  const long long total = (long long)B * Tp * C * ph * pw;
  blend_accumulate_kernel<<<grid_for_e(total, 256, 148 * 32), 256, 0, as_stream(stream)>>>(patch, E, count, q, total);
  return check_launch();
}

extern "C" int vmgb_blend_normalize(float* E, const float* count, int B, int Te, int C, int He, int We, void* stream) {
  if (!E || !count || B < 1 || Te < 1 || C < 1) return VMGB_EINVAL;
  const long long total = (long long)B * Te * C * He * We;
  blend_normalize_kernel<<<grid_for_e(total, 256, 148 * 32), 256, 0, as_stream(stream)>>>(E, count, C, Te, (long long)He * We, total);
  return check_launch();
}

extern "C" int vmgb_blend_finalize_u8(const float* E, const float* count, void* out_u8, int T, int H, int W, int reverse_channels, void* stream) {
  if (!E || !out_u8 || T < 1) return VMGB_EINVAL;
  const long long total = (long long)T * H * W;
  blend_finalize_u8_kernel<<<grid_for_e(total, 256, 148 * 32), 256, 0, as_stream(stream)>>>(E, count, (uint8_t*)out_u8, (long long)H * W,
                                                                                           reverse_channels, total);
  return check_launch();
}

extern "C" int vmgb_augment(const float* src, float* dst, long long planes, int H, int W, int hflip, int vflip, int transpose, void* stream) {
  if (!src || !dst || src == dst || planes < 1 || H < 1 || W < 1) return VMGB_EINVAL;
  const long long total = planes * H * W;
  augment_kernel<<<grid_for_e(total, 256, 148 * 32), 256, 0, as_stream(stream)>>>(src, dst, H, W, hflip, vflip, transpose, total);
  return check_launch();
}

extern "C" int vmgb_sse_u8(const void* a, const void* b, long long n, long long* out_sum, void* stream) {
  if (!a || !b || !out_sum || n < 1) return VMGB_EINVAL;
  sse_u8_kernel<<<grid_for_e((n + 15) / 16, 256, 148 * 8), 256, 0, as_stream(stream)>>>((const uint8_t*)a, (const uint8_t*)b, n,
                                                                                       reinterpret_cast<unsigned long long*>(out_sum));
  return check_launch();
}

static const int SSE_BLOCKS = 148 * 4;
extern "C" long long vmgb_metrics_scratch_doubles(int n_img, int H, int W, int n_planes) {
  const long long tiles = (long long)((H - 10 + 15) / 16) * ((W - 10 + 15) / 16);
  const long long ssim = (long long)n_img * n_planes * (tiles > 0 ? tiles : 0);
  return ssim > SSE_BLOCKS ? ssim : SSE_BLOCKS;
}

extern "C" int vmgb_sse_f32(const float* a, const float* b, long long n, int clamp01, double* scratch, double* out_sum, void* stream) {
  if (!a || !b || !scratch || !out_sum || n < 1) return VMGB_EINVAL;
  sse_f32_kernel<<<SSE_BLOCKS, 256, 0, as_stream(stream)>>>(a, b, n, clamp01, scratch);
  int rc = check_launch();
  if (rc) return rc;
  reduce_partials_kernel<<<1, 256, 0, as_stream(stream)>>>(scratch, SSE_BLOCKS, 1.0, out_sum);
  return check_launch();
}

extern "C" int vmgb_rgb_to_y(const void* rgb_u8, double* y, long long pixels, void* stream) {
  if (!rgb_u8 || !y || pixels < 1) return VMGB_EINVAL;
  rgb_to_y_kernel<<<grid_for_e(pixels), 256, 0, as_stream(stream)>>>((const uint8_t*)rgb_u8, y, pixels);
  return check_launch();
}

extern "C" int vmgb_sse_f64(const double* a, const double* b, long long n, double* scratch, double* out_sum, void* stream) {
  if (!a || !b || !scratch || !out_sum || n < 1) return VMGB_EINVAL;
  sse_f64_kernel<<<SSE_BLOCKS, 256, 0, as_stream(stream)>>>(a, b, n, scratch);
  int rc = check_launch();
  if (rc) return rc;
  reduce_partials_kernel<<<1, 256, 0, as_stream(stream)>>>(scratch, SSE_BLOCKS, 1.0, out_sum);
  return check_launch();
}

// images: n_img x [H, W, n_planes] interleaved (dt 0: uint8 RGB, n_planes = 3) or n_img x [H, W] float64 (dt 1, n_planes = 1)
extern "C" int vmgb_ssim(const void* a, const void* b, int dt, int n_img, int H, int W, int n_planes, const double* window11, double* scratch,
                         double* out_per_img, void* stream) {
  if (!a || !b || !window11 || !scratch || !out_per_img || n_img < 1 || H < 11 || W < 11 || n_planes < 1) return VMGB_EINVAL;
  if (dt != 0 && dt != 1) return VMGB_EINVAL;
  SsimP q;
  for (int k = 0; k < 11; ++k) q.w[k] = window11[k];      // host pointer: 11 weights of cv2.getGaussianKernel(11, 1.5)
  q.H = H; q.W = W; q.tiles_y = (H - 10 + 15) / 16; q.tiles_x = (W - 10 + 15) / 16;
  const int blocks_per_img = n_planes * q.tiles_x * q.tiles_y;
  const long long blocks = (long long)n_img * blocks_per_img;
  if (blocks > 0x7fffffffLL) return VMGB_EINVAL;
  if (dt == 0)
    ssim_kernel<uint8_t><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>((const uint8_t*)a, (const uint8_t*)b, n_planes, n_planes, 1,
                                                                          (long long)H * W * n_planes, q, scratch);
  else
    ssim_kernel<double><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>((const double*)a, (const double*)b, 1, 1, 0, (long long)H * W, q, scratch);
  int rc = check_launch();
  if (rc) return rc;
  const double scale = 1.0 / ((double)n_planes * (double)(H - 10) * (double)(W - 10));
  ssim_finish_kernel<<<n_img, 256, 0, as_stream(stream)>>>(scratch, blocks_per_img, scale, out_per_img);
  return check_launch();
}
