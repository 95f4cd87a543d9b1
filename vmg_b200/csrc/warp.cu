// warp.cu -- gather / resize kernels: flow warps, flow smoothing, SPyNet pyramid helpers, input
// padding and the final bilinear-residual output.  References per entry point are in include/vmgb.h.
#include "common.cuh"

namespace vmgb {

inline int grid_for2(long long total, int threads = 256, int cap = 148 * 16) {
  long long b = (total + threads - 1) / threads;
  if (b < 1) b = 1;
  return (int)(b > cap ? cap : b);
}

// ATen area_pixel_compute_source_index (upsample_bilinear2d), align_corners False / True
__device__ __forceinline__ void src_index(int dst, float scale, bool align, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float src = align ? scale * (float)dst : fmaxf(scale * ((float)dst + 0.5f) - 0.5f, 0.f);
  i0 = min((int)src, in_size - 1);
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - (float)i0;
  l0 = 1.f - l1;
}
__device__ __forceinline__ float scale_of(int in_size, int out_size, bool align) {
  if (align) return out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
  return (float)in_size / (float)out_size;
}

// ---------------------------------------------------------------------------------------------
// feature warp: bilinear, border, align_corners=True.  One thread per (pixel, 4 channels).
// ---------------------------------------------------------------------------------------------
__global__ void flow_warp_bilinear_kernel(const void* x, int x_dt, int x_pitch, const float* __restrict__ flow, void* out, int out_dt,
                                          int out_pitch, int n, int h, int w, int C, int sem, long long total) {
  const int C4 = C >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4); const long long pix = i / C4;
    const int px = (int)(pix % w); const long long t = pix / w; const int py = (int)(t % h); const int img = (int)(t / h);
    const float2 f = *reinterpret_cast<const float2*>(flow + pix * 2);
    const float ix = warp_coord((float)px, f.x, w, sem), iy = warp_coord((float)py, f.y, h, sem);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
    const float wx1 = ix - fx0, wx0 = (fx0 + 1.f) - ix, wy1 = iy - fy0, wy0 = (fy0 + 1.f) - iy;
    const long long ib = (long long)img * h;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    auto tap = [&](int yy, int xx, float wgt) {
      if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
        const float4 v = ld4(x, x_dt, ((ib + yy) * w + xx) * x_pitch + c4 * 4);
        acc.x += v.x * wgt; acc.y += v.y * wgt; acc.z += v.z * wgt; acc.w += v.w * wgt;
      }
    };
    tap(y0, x0, wx0 * wy0); tap(y0, x1, wx1 * wy0); tap(y1, x0, wx0 * wy1); tap(y1, x1, wx1 * wy1);
    st4(out, out_dt, pix * out_pitch + c4 * 4, acc);
  }
}

// bf16 fast path: one thread per (pixel, 8 channels), 16-byte loads and stores (same tap order and fp32 accumulation as the
// generic kernel above, so the results are bit-identical); half the threads and half the redundant coordinate work
__global__ void __launch_bounds__(256) flow_warp_bilinear8_kernel(const __nv_bfloat16* __restrict__ x, int x_pitch, const float* __restrict__ flow,
                                                                 __nv_bfloat16* __restrict__ out, int out_pitch, int n, int h, int w, int C,
                                                                 int sem, long long total) {
  const int C8 = C >> 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8); const long long pix = i / C8;
    const int px = (int)(pix % w); const long long t = pix / w; const int py = (int)(t % h); const int img = (int)(t / h);
    const float2 f = *reinterpret_cast<const float2*>(flow + pix * 2);
    const float ix = warp_coord((float)px, f.x, w, sem), iy = warp_coord((float)py, f.y, h, sem);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
    const float wx1 = ix - fx0, wx0 = (fx0 + 1.f) - ix, wy1 = iy - fy0, wy0 = (fy0 + 1.f) - iy;
    const long long ib = (long long)img * h;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    const int ys[4] = {y0, y0, y1, y1}, xs[4] = {x0, x1, x0, x1};
    const float ws[4] = {wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1};
    uint4 v[4];
    bool ok[4];
#pragma unroll
    for (int tp = 0; tp < 4; ++tp) {
      ok[tp] = ys[tp] >= 0 && ys[tp] < h && xs[tp] >= 0 && xs[tp] < w;
      v[tp] = ok[tp] ? __ldg(reinterpret_cast<const uint4*>(x + ((ib + ys[tp]) * w + xs[tp]) * x_pitch + c8 * 8)) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int tp = 0; tp < 4; ++tp) {
      if (ok[tp]) {
        const uint32_t wd[4] = {v[tp].x, v[tp].y, v[tp].z, v[tp].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          acc[2 * q] += __uint_as_float(wd[q] << 16) * ws[tp];
          acc[2 * q + 1] += __uint_as_float(wd[q] & 0xffff0000u) * ws[tp];
        }
      }
    }
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[2 * q], acc[2 * q + 1]);
      o[q] = *reinterpret_cast<const uint32_t*>(&h2);
    }
    *reinterpret_cast<uint4*>(out + pix * out_pitch + c8 * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// location warp: nearest, border (locations stay exact in-range integers)
__global__ void flow_warp_nearest_kernel(const float* __restrict__ loc, int in_pitch, const float* __restrict__ flow,
                                         float* __restrict__ out, int out_pitch, int n, int h, int w, int K2, int sem, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % K2); const long long pix = i / K2;
    const int px = (int)(pix % w); const long long t = pix / w; const int py = (int)(t % h); const int img = (int)(t / h);
    const float2 f = *reinterpret_cast<const float2*>(flow + pix * 2);
    const int sx = (int)rintf(warp_coord((float)px, f.x, w, sem)), sy = (int)rintf(warp_coord((float)py, f.y, h, sem));
    out[pix * out_pitch + c] = loc[(((long long)img * h + sy) * w + sx) * in_pitch + c];
  }
}

// one thread per (pixel, tracked location = 2 floats): half the threads and coordinate work of the scalar kernel
__global__ void __launch_bounds__(256) flow_warp_nearest2_kernel(const float* __restrict__ loc, int in_pitch, const float* __restrict__ flow,
                                                                float* __restrict__ out, int out_pitch, int n, int h, int w, int K,
                                                                int sem, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K); const long long pix = i / K;
    const int px = (int)(pix % w); const long long t = pix / w; const int py = (int)(t % h); const int img = (int)(t / h);
    const float2 f = *reinterpret_cast<const float2*>(flow + pix * 2);
    const int sx = (int)rintf(warp_coord((float)px, f.x, w, sem)), sy = (int)rintf(warp_coord((float)py, f.y, h, sem));
    *reinterpret_cast<float2*>(out + pix * out_pitch + 2 * k) =
        *reinterpret_cast<const float2*>(loc + (((long long)img * h + sy) * w + sx) * in_pitch + 2 * k);
  }
}

// ---------------------------------------------------------------------------------------------
// flow_warp_fused: ONE launch per trajectory step for both warps that share a flow field (trajectory.py:330-333):
// the bilinear feature warp (bf16, 8 channels per thread) and the nearest warp of all K tracked locations (one (x, y)
// pair per thread).  Roles are BLOCK-uniform (blocks [0, nb_feat) warp features, the rest locations): mixing both kinds of
// work items inside a warp made every warp execute both bodies (measured 54 us against 44 us for the two launches).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) flow_warp_fused_kernel(const __nv_bfloat16* __restrict__ x, int x_pitch, const float* __restrict__ flow,
                                                             __nv_bfloat16* __restrict__ out, int out_pitch, const float* __restrict__ loc,
                                                             int loc_in_pitch, float* __restrict__ loc_out, int loc_out_pitch, int K, int n, int h,
                                                             int w, int C, int sem, int nb_feat) {
  const long long npix = (long long)n * h * w;
  if ((int)blockIdx.x >= nb_feat) {
    const long long total = npix * K, stride = (long long)(gridDim.x - nb_feat) * blockDim.x;
    for (long long i = (long long)(blockIdx.x - nb_feat) * blockDim.x + threadIdx.x; i < total; i += stride) {
      const int k = (int)(i % K); const long long pix = i / K;
      const int px = (int)(pix % w); const long long t = pix / w; const int py = (int)(t % h); const int img = (int)(t / h);
      const float2 f = *reinterpret_cast<const float2*>(flow + pix * 2);
      const int sx = (int)rintf(warp_coord((float)px, f.x, w, sem)), sy = (int)rintf(warp_coord((float)py, f.y, h, sem));
      *reinterpret_cast<float2*>(loc_out + pix * loc_out_pitch + 2 * k) =
          *reinterpret_cast<const float2*>(loc + (((long long)img * h + sy) * w + sx) * loc_in_pitch + 2 * k);
    }
    return;
  }
  const int C8 = C >> 3;
  const long long total = npix * C8, stride = (long long)nb_feat * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c8 = (int)(i % C8); const long long pix = i / C8;
    const int px = (int)(pix % w); const long long t = pix / w; const int py = (int)(t % h); const int img = (int)(t / h);
    const float2 f = *reinterpret_cast<const float2*>(flow + pix * 2);
    const float ix = warp_coord((float)px, f.x, w, sem), iy = warp_coord((float)py, f.y, h, sem);
    const long long ib = (long long)img * h;
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
    const float wx1 = ix - fx0, wx0 = (fx0 + 1.f) - ix, wy1 = iy - fy0, wy0 = (fy0 + 1.f) - iy;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    const int ys[4] = {y0, y0, y1, y1}, xs[4] = {x0, x1, x0, x1};
    const float ws[4] = {wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1};
    uint4 v[4];
    bool ok[4];
#pragma unroll
    for (int tp = 0; tp < 4; ++tp) {
      ok[tp] = ys[tp] >= 0 && ys[tp] < h && xs[tp] >= 0 && xs[tp] < w;
      v[tp] = ok[tp] ? __ldg(reinterpret_cast<const uint4*>(x + ((ib + ys[tp]) * w + xs[tp]) * x_pitch + c8 * 8)) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int tp = 0; tp < 4; ++tp) {
      if (ok[tp]) {
        const uint32_t wd[4] = {v[tp].x, v[tp].y, v[tp].z, v[tp].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          acc[2 * q] += __uint_as_float(wd[q] << 16) * ws[tp];
          acc[2 * q + 1] += __uint_as_float(wd[q] & 0xffff0000u) * ws[tp];
        }
      }
    }
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[2 * q], acc[2 * q + 1]);
      o[q] = *reinterpret_cast<const uint32_t*>(&h2);
    }
    *reinterpret_cast<uint4*>(out + pix * out_pitch + c8 * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void loc_identity_kernel(float* __restrict__ loc, int pitch, int c0, int n, int h, int w, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int px = (int)(i % w); const int py = (int)((i / w) % h);
    loc[i * pitch + c0] = (float)px;
    loc[i * pitch + c0 + 1] = (float)py;
  }
}

// flow smoothing: reflect pad (right/bottom) to a multiple of r, r x r mean, nearest x r, crop.  NHWC2 fp32.
__global__ void flow_smooth_kernel(const float* __restrict__ in, float* __restrict__ out, int n, int h, int w, int r, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int px = (int)(i % w); const long long t = i / w; const int py = (int)(t % h); const int img = (int)(t / h);
    const int by = py / r * r, bx = px / r * r;
    float sx = 0.f, sy = 0.f;
    for (int dy = 0; dy < r; ++dy) {
      int yy = by + dy; if (yy >= h) yy = 2 * (h - 1) - yy;
      for (int dx = 0; dx < r; ++dx) {
        int xx = bx + dx; if (xx >= w) xx = 2 * (w - 1) - xx;
        const float2 v = *reinterpret_cast<const float2*>(in + (((long long)img * h + yy) * w + xx) * 2);
        sx += v.x; sy += v.y;
      }
    }
    const float inv = (float)(r * r);
    *reinterpret_cast<float2*>(out + i * 2) = make_float2(sx / inv, sy / inv);
  }
}

// ---------------------------------------------------------------------------------------------
// SPyNet helpers (fp32, channels-last, 3 / 8 / 2 channels)
// ---------------------------------------------------------------------------------------------
__global__ void spynet_prep_kernel(const float* __restrict__ fr, float* __restrict__ out, float m0, float m1, float m2, float s0,
                                   float s1, float s2, int n, int H, int W, int hu, int wu, long long total) {
  const float sch = scale_of(H, hu, false), scw = scale_of(W, wu, false);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int px = (int)(i % wu); const long long t = i / wu; const int py = (int)(t % hu); const int img = (int)(t / hu);
    int y0, y1, x0, x1; float ly0, ly1, lx0, lx1;
    src_index(py, sch, false, H, y0, y1, ly0, ly1);
    src_index(px, scw, false, W, x0, x1, lx0, lx1);
    const float* b = fr + (long long)img * H * W * 3;
    const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = ly0 * (lx0 * b[((long long)y0 * W + x0) * 3 + c] + lx1 * b[((long long)y0 * W + x1) * 3 + c]) +
                      ly1 * (lx0 * b[((long long)y1 * W + x0) * 3 + c] + lx1 * b[((long long)y1 * W + x1) * 3 + c]);
      out[i * 3 + c] = (v - mean[c]) / sd[c];
    }
  }
}

__global__ void avgpool2_kernel(const float* __restrict__ x, float* __restrict__ out, int n, int H, int W, int C, long long total) {
  const int Ho = H / 2, Wo = W / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C); long long p = i / C;
    const int wo = (int)(p % Wo); p /= Wo; const int ho = (int)(p % Ho); const int img = (int)(p / Ho);
    const float* b = x + (((long long)img * H + 2 * ho) * W + 2 * wo) * C + c;
    out[i] = (b[0] + b[C] + b[(long long)W * C] + b[(long long)W * C + C]) / 4.f;
  }
}

// one thread per output pixel of one (ref, supp) pair
__global__ void spynet_level_input_kernel(const float* __restrict__ pyr, const float* __restrict__ fc, void* out8, int out8_dt,
                                          float* __restrict__ fup, int B, int dir, int h, int w, int sem, long long total) {
  const int hc = h / 2, wc = w / 2;
  const float sch = scale_of(hc, h, true), scw = scale_of(wc, w, true);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int px = (int)(i % w); const long long t = i / w; const int py = (int)(t % h); const int pair = (int)(t / h);
    const int fref = pair + (dir == 0 ? B : 0), fsup = pair + (dir == 0 ? 0 : B);
    float fx = 0.f, fy = 0.f;
    if (fc) {
      int y0, y1, x0, x1; float ly0, ly1, lx0, lx1;
      src_index(py, sch, true, hc, y0, y1, ly0, ly1);
      src_index(px, scw, true, wc, x0, x1, lx0, lx1);
      const float* q = fc + (long long)pair * hc * wc * 2;
      const float2 v00 = *reinterpret_cast<const float2*>(q + ((long long)y0 * wc + x0) * 2);
      const float2 v01 = *reinterpret_cast<const float2*>(q + ((long long)y0 * wc + x1) * 2);
      const float2 v10 = *reinterpret_cast<const float2*>(q + ((long long)y1 * wc + x0) * 2);
      const float2 v11 = *reinterpret_cast<const float2*>(q + ((long long)y1 * wc + x1) * 2);
      fx = (ly0 * (lx0 * v00.x + lx1 * v01.x) + ly1 * (lx0 * v10.x + lx1 * v11.x)) * 2.0f;
      fy = (ly0 * (lx0 * v00.y + lx1 * v01.y) + ly1 * (lx0 * v10.y + lx1 * v11.y)) * 2.0f;
    }
    const float ix = warp_coord((float)px, fx, w, sem), iy = warp_coord((float)py, fy, h, sem);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
    const float wx1 = ix - fx0, wx0 = (fx0 + 1.f) - ix, wy1 = iy - fy0, wy0 = (fy0 + 1.f) - iy;
    const float* s = pyr + (long long)fsup * h * w * 3;
    float wv[3] = {0.f, 0.f, 0.f};
    auto tap = [&](int yy, int xx, float wgt) {
      if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
        const float* q = s + ((long long)yy * w + xx) * 3;
        wv[0] += q[0] * wgt; wv[1] += q[1] * wgt; wv[2] += q[2] * wgt;
      }
    };
    tap(y0, x0, wx0 * wy0); tap(y0, x1, wx1 * wy0); tap(y1, x0, wx0 * wy1); tap(y1, x1, wx1 * wy1);
    const float* r = pyr + ((long long)fref * h * w + (long long)py * w + px) * 3;
    st4(out8, out8_dt, i * 8, make_float4(r[0], r[1], r[2], wv[0]));
    st4(out8, out8_dt, i * 8 + 4, make_float4(wv[1], wv[2], fx, fy));
    fup[i * 2] = fx; fup[i * 2 + 1] = fy;
  }
}

__global__ void flow_resize_kernel(const float* __restrict__ flow, float* __restrict__ out, int n, int hu, int wu, int h, int w,
                                   long long total) {
  const float sch = scale_of(hu, h, false), scw = scale_of(wu, w, false);
  const float mx = (float)w / (float)wu, my = (float)h / (float)hu;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int px = (int)(i % w); const long long t = i / w; const int py = (int)(t % h); const int img = (int)(t / h);
    int y0, y1, x0, x1; float ly0, ly1, lx0, lx1;
    src_index(py, sch, false, hu, y0, y1, ly0, ly1);
    src_index(px, scw, false, wu, x0, x1, lx0, lx1);
    const float* q = flow + (long long)img * hu * wu * 2;
    const float2 v00 = *reinterpret_cast<const float2*>(q + ((long long)y0 * wu + x0) * 2);
    const float2 v01 = *reinterpret_cast<const float2*>(q + ((long long)y0 * wu + x1) * 2);
    const float2 v10 = *reinterpret_cast<const float2*>(q + ((long long)y1 * wu + x0) * 2);
    const float2 v11 = *reinterpret_cast<const float2*>(q + ((long long)y1 * wu + x1) * 2);
    out[i * 2] = (ly0 * (lx0 * v00.x + lx1 * v01.x) + ly1 * (lx0 * v10.x + lx1 * v11.x)) * mx;
    out[i * 2 + 1] = (ly0 * (lx0 * v00.y + lx1 * v01.y) + ly1 * (lx0 * v10.y + lx1 * v11.y)) * my;
  }
}

// ---------------------------------------------------------------------------------------------
// input / output layout kernels
// ---------------------------------------------------------------------------------------------
__global__ void pad_input_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int T, int H, int W, int Hp, int Wp, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % 3); long long p = i / 3;
    const int px = (int)(p % Wp); p /= Wp; const int py = (int)(p % Hp); const int tb = (int)(p / Hp);
    const int img = (tb % B) * T + tb / B;      // time-major destination <- [B,T] source
    out[i] = x[(((long long)img * 3 + c) * H + min(py, H - 1)) * W + min(px, W - 1)];
  }
}

__global__ void __launch_bounds__(256) mirror_check_kernel(const float* __restrict__ x, float* sumsq, int B, int T, long long fe, long long total) {
  __shared__ float sh[8];
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i % fe; long long r = i / fe;
    const int t = (int)(r % (T / 2)); const int b = (int)(r / (T / 2));
    const float d = x[((long long)b * T + t) * fe + e] - x[((long long)b * T + (T - 1 - t)) * fe + e];
    s += d * d;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { float tsum = 0.f; for (int k = 0; k < 8; ++k) tsum += sh[k]; if (tsum != 0.f) atomicAdd(sumsq, tsum); }
}

__global__ void final_output_kernel(const void* y, int y_dt, int y_pitch, const float* __restrict__ x, float* __restrict__ out, int B,
                                    int T, int t0, int H, int W, long long total) {
  const int Ho = 4 * H, Wo = 4 * W;
  const float sch = scale_of(H, Ho, false), scw = scale_of(W, Wo, false);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int px = (int)(i % Wo); const long long t = i / Wo; const int py = (int)(t % Ho); const int tb = (int)(t / Ho);
    const int img = (tb % B) * T + t0 + tb / B;   // [B,T] frame index of time-major chunk frame tb
    int y0, y1, x0, x1; float ly0, ly1, lx0, lx1;
    src_index(py, sch, false, H, y0, y1, ly0, ly1);
    src_index(px, scw, false, W, x0, x1, lx0, lx1);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* b = x + ((long long)img * 3 + c) * H * W;
      const float up = ly0 * (lx0 * b[(long long)y0 * W + x0] + lx1 * b[(long long)y0 * W + x1]) +
                       ly1 * (lx0 * b[(long long)y1 * W + x0] + lx1 * b[(long long)y1 * W + x1]);
      out[(((long long)img * 3 + c) * Ho + py) * Wo + px] = ldf(y, y_dt, i * y_pitch + c) + up;
    }
  }
}

}  // namespace vmgb

using namespace vmgb;

static bool sem_ok(int sem) { return sem == VMGB_COORD_ATEN_CPU || sem == VMGB_COORD_ATEN_CUDA; }

extern "C" int vmgb_flow_warp_bilinear(const void* x, int x_dt, int x_pitch, const float* flow, void* out, int out_dt, int out_pitch,
                                       int n, int h, int w, int C, int semantics, void* stream) {
  if (!x || !flow || !out || !dt_ok(x_dt) || !dt_ok(out_dt) || (C & 3) || (x_pitch & 3) || (out_pitch & 3) || !sem_ok(semantics)) return VMGB_EINVAL;
  if (x_dt == VMGB_BF16 && out_dt == VMGB_BF16 && (C & 7) == 0 && (x_pitch & 7) == 0 && (out_pitch & 7) == 0 &&
      !(((uintptr_t)x | (uintptr_t)out) & 15)) {
    const long long total8 = (long long)n * h * w * (C / 8);
    flow_warp_bilinear8_kernel<<<grid_for2(total8, 256, 148 * 32), 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)x, x_pitch, flow,
                                                                                                 (__nv_bfloat16*)out, out_pitch, n, h, w, C, semantics, total8);
    return check_launch();
  }
  long long total = (long long)n * h * w * (C / 4);
  flow_warp_bilinear_kernel<<<grid_for2(total, 256, 148 * 32), 256, 0, as_stream(stream)>>>(x, x_dt, x_pitch, flow, out, out_dt, out_pitch, n, h, w, C, semantics, total);
  return check_launch();
}

extern "C" int vmgb_flow_warp_nearest(const float* loc, int in_pitch, const float* flow, float* out, int out_pitch, int n, int h,
                                      int w, int K2, int semantics, void* stream) {
  if (!loc || !flow || !out || K2 <= 0 || loc == out || !sem_ok(semantics)) return VMGB_EINVAL;
  if (!(K2 & 1) && !(in_pitch & 1) && !(out_pitch & 1) && !(((uintptr_t)loc | (uintptr_t)out) & 7)) {
    const long long total2 = (long long)n * h * w * (K2 / 2);
    flow_warp_nearest2_kernel<<<grid_for2(total2, 256, 148 * 32), 256, 0, as_stream(stream)>>>(loc, in_pitch, flow, out, out_pitch, n, h, w, K2 / 2, semantics, total2);
    return check_launch();
  }
  long long total = (long long)n * h * w * K2;
  flow_warp_nearest_kernel<<<grid_for2(total, 256, 148 * 32), 256, 0, as_stream(stream)>>>(loc, in_pitch, flow, out, out_pitch, n, h, w, K2, semantics, total);
  return check_launch();
}

extern "C" int vmgb_flow_warp_fused(const void* x, int x_pitch, const float* flow, void* out, int out_pitch, const float* loc, int loc_in_pitch,
                                    float* loc_out, int loc_out_pitch, int K2, int n, int h, int w, int C, int semantics, void* stream) {
  if (!x || !flow || !out || !loc || !loc_out || loc == loc_out || K2 <= 0 || (K2 & 1) || !sem_ok(semantics)) return VMGB_EINVAL;
  if ((C & 7) || (x_pitch & 7) || (out_pitch & 7) || (((uintptr_t)x | (uintptr_t)out) & 15) || (loc_in_pitch & 1) || (loc_out_pitch & 1) ||
      (((uintptr_t)loc | (uintptr_t)loc_out) & 7))
    return VMGB_EUNSUPPORTED;      // the caller falls back to the two separate launches
  const long long npix = (long long)n * h * w;
  const int nb_feat = grid_for2(npix * (C / 8), 256, 148 * 24), nb_loc = grid_for2(npix * (K2 / 2), 256, 148 * 8);
  flow_warp_fused_kernel<<<nb_feat + nb_loc, 256, 0, as_stream(stream)>>>(
      (const __nv_bfloat16*)x, x_pitch, flow, (__nv_bfloat16*)out, out_pitch, loc, loc_in_pitch, loc_out, loc_out_pitch, K2 / 2, n, h, w, C,
      semantics, nb_feat);
  return check_launch();
}

extern "C" int vmgb_loc_identity(float* loc, int pitch, int c0, int n, int h, int w, void* stream) {
  if (!loc || c0 < 0 || c0 + 2 > pitch) return VMGB_EINVAL;
  long long total = (long long)n * h * w;
  loc_identity_kernel<<<grid_for2(total), 256, 0, as_stream(stream)>>>(loc, pitch, c0, n, h, w, total);
  return check_launch();
}

extern "C" int vmgb_flow_smooth(const float* flow_in, float* out_nhwc, int n, int h, int w, int r, void* stream) {
  if (!flow_in || !out_nhwc || r <= 0 || flow_in == out_nhwc) return VMGB_EINVAL;
  // reflect padding needs pad < size (torch raises otherwise)
  if (((h + r - 1) / r * r - h) >= h || ((w + r - 1) / r * r - w) >= w) return VMGB_EINVAL;
  long long total = (long long)n * h * w;
  flow_smooth_kernel<<<grid_for2(total), 256, 0, as_stream(stream)>>>(flow_in, out_nhwc, n, h, w, r, total);
  return check_launch();
}

extern "C" int vmgb_spynet_prep(const float* frames, float* out, float m0, float m1, float m2, float s0, float s1, float s2,
                                int n, int H, int W, int h_up, int w_up, void* stream) {
  if (!frames || !out) return VMGB_EINVAL;
  long long total = (long long)n * h_up * w_up;
  spynet_prep_kernel<<<grid_for2(total), 256, 0, as_stream(stream)>>>(frames, out, m0, m1, m2, s0, s1, s2,
                                                                      n, H, W, h_up, w_up, total);
  return check_launch();
}

extern "C" int vmgb_avgpool2(const float* x, float* out, int n, int H, int W, int C, void* stream) {
  if (!x || !out || (H & 1) || (W & 1)) return VMGB_EINVAL;
  long long total = (long long)n * (H / 2) * (W / 2) * C;
  avgpool2_kernel<<<grid_for2(total), 256, 0, as_stream(stream)>>>(x, out, n, H, W, C, total);
  return check_launch();
}

extern "C" int vmgb_spynet_level_input(const float* pyr, const float* flow_coarse, void* out8, int out8_dt, float* flow_up, int n_pairs, int B,
                                       int dir, int h, int w, int semantics, void* stream) {
  if (!pyr || !out8 || !flow_up || !dt_ok(out8_dt) || n_pairs < 1 || B < 1 || (flow_coarse && ((h & 1) || (w & 1))) || !sem_ok(semantics)) return VMGB_EINVAL;
  long long total = (long long)n_pairs * h * w;
  spynet_level_input_kernel<<<grid_for2(total), 256, 0, as_stream(stream)>>>(pyr, flow_coarse, out8, out8_dt, flow_up, B, dir, h, w, semantics, total);
  return check_launch();
}

extern "C" int vmgb_flow_resize(const float* flow, float* out, int n, int h_up, int w_up, int h, int w, void* stream) {
  if (!flow || !out) return VMGB_EINVAL;
  long long total = (long long)n * h * w;
  flow_resize_kernel<<<grid_for2(total), 256, 0, as_stream(stream)>>>(flow, out, n, h_up, w_up, h, w, total);
  return check_launch();
}

extern "C" int vmgb_pad_input(const float* x, float* out, int B, int T, int H, int W, int Hp, int Wp, void* stream) {
  if (!x || !out || Hp < H || Wp < W || B < 1 || T < 1) return VMGB_EINVAL;
  long long total = (long long)B * T * Hp * Wp * 3;
  pad_input_kernel<<<grid_for2(total), 256, 0, as_stream(stream)>>>(x, out, B, T, H, W, Hp, Wp, total);
  return check_launch();
}

extern "C" int vmgb_mirror_check(const float* x, float* sumsq, int B, int T, long long frame_elems, void* stream) {
  if (!x || !sumsq || (T & 1)) return VMGB_EINVAL;
  long long total = (long long)B * (T / 2) * frame_elems;
  mirror_check_kernel<<<grid_for2(total), 256, 0, as_stream(stream)>>>(x, sumsq, B, T, frame_elems, total);
  return check_launch();
}

extern "C" int vmgb_final_output(const void* y, int y_dt, int y_pitch, const float* x, float* out, int B, int T, int t0, int Tc,
                                 int H, int W, void* stream) {
  if (!y || !x || !out || !dt_ok(y_dt) || t0 < 0 || Tc < 1 || t0 + Tc > T) return VMGB_EINVAL;
  long long total = (long long)B * Tc * 16 * H * W;
  final_output_kernel<<<grid_for2(total, 256, 148 * 32), 256, 0, as_stream(stream)>>>(y, y_dt, y_pitch, x, out, B, T, t0, H, W, total);
  return check_launch();
}
