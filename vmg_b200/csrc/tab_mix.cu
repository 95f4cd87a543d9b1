// tab_mix.cu -- the Morph-FC mixing half of a TAB (function.py:743-805 Enhanced_MorphFCs_decay.forward, called from
// TAB.forward function.py:1212-1215) as TWO launches of one persistent tcgen05 kernel instead of ~14 HBM passes:
//
//   x += m * (LN2(x) + z) * tanh(z),  z = proj(a0 . h + a1 . w + a2 . c),
//   h = ReLU(MorphFC_H(LN2 x)) / C,  w = ReLU(MorphFC_W(LN2 x)) / C,  c = ReLU(Linear_C(LN2 x)) / C,
//   a = softmax3(reweight(mean_{t,h,w}(h + w + c)))      <- a GLOBAL reduction between the branches and the mix
//
//   phase 0 ("stats"): reads x, LayerNorm in registers, the three branch GEMMs on the tensor core, per-channel sums of
//                      ReLU(.)/C over the valid pixels -> 64-bit fixed-point accumulators [B][C] (order independent);
//   (tiny kernels: finalize the mean, reweight MLP, softmax over the 3 branches)
//   phase 1 ("apply"): recomputes LayerNorm and the three branch GEMMs (0.44 TFLOP per level-0 TAB: cheaper than one
//                      HBM pass of h, w, c), scales by a_j / C in the drain, projects, gates, adds the residual, writes x.
// HBM traffic: 2 reads + 1 write of the activation instead of ~30 passes (LN out, 2 gathers, 3 linears in/out, 2 scatters,
// channel sums, mix, gated projection with 3 operands).
//
// Geometry (chunk 8 x 8, C <= 128, S = C / 8 even): a tile = 8 rows x 16 columns = 128 pixels = two 8 x 8 Morph blocks,
// which is self-contained for BOTH token mixers: the H tokens of a block are (column w, channel group g)
// with features (row r, sub-channel s), the W tokens (row h, g) with features (column r, s).  One tile = one M = 128
// tcgen05.mma tile for each of the three branches; the SAME 128 x C LayerNorm outputs are written three times into
// shared memory by the producer threads, each time in the K-major SWIZZLE_128B layout of its branch:
//     A_c[pixel][c]                     A_h[(b, w, g)][r * S + s]                 A_w[(b, h, g)][r * S + s]
// (b = which of the two blocks, pixel = row * 16 + column).  TMA cannot do this gather: the channel-group stride is
// S * 2 = 28 bytes, not a multiple of 16.  The drained branch outputs go back to the SAME buffers in pixel-major
// K-major layout (B_h, B_w, B_c, scaled by a_j / C), so that the projection is three accumulating MMAs with ONE resident
// weight matrix: the tensor core adds the three branches, no read-modify-write mix.  All four weight matrices
// (4 x 28 KB at C = 112) stay resident for the whole launch; accumulators: 4 x 128 TMEM columns.
// The gate needs LN2(x) again: it is recomputed from x (an L2 hit) and the per-pixel (mean, rstd) kept in shared memory.
#include "tc_ptx.cuh"
#include <mutex>

namespace vmgb {

struct TabP {
  const __nv_bfloat16* x;
  __nv_bfloat16* out;
  const float* ln_g;
  const float* ln_b;
  const float* bias[4];               // h, w, c, proj
  const float* a;                     // [B][3][C] branch weights (phase 1)
  unsigned long long* acc;            // [B][C] 2^24 fixed-point sums (phase 0)
  int n_img, B, H, W, C, S, nkb, tiles_h, tiles_w, phase;
  long long n_tiles;
  float eps, m_scaling;
};

constexpr int TM_THREADS = 512;
constexpr int TM_TILE_KB = 16384;     // one 64-channel K block of a 128-row operand tile

__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts64(uint32_t a, uint32_t v0, uint32_t v1) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(v0), "r"(v1) : "memory");
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t bf2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// byte offset of 4-byte word `word` (K index / 2) of row `row` inside a K-major SWIZZLE_128B operand tile
__device__ __forceinline__ uint32_t sw_word(int row, int word) {
  return (uint32_t)((word >> 5) * TM_TILE_KB + row * 128 + ((((word & 31) >> 2) ^ (row & 7)) << 4) + ((word & 3) << 2));
}

template <int CC>
__global__ void __launch_bounds__(TM_THREADS, 1)
tab_mix_kernel(const __grid_constant__ CUtensorMap mapW0, const __grid_constant__ CUtensorMap mapW1, const __grid_constant__ CUtensorMap mapW2,
               const __grid_constant__ CUtensorMap mapW3, const __grid_constant__ TabP p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  constexpr int C = CC, S = CC / 8, S2 = S / 2, nkb = (CC + 63) / 64;
  constexpr uint32_t wbytes = (uint32_t)(nkb * C * 128), tbytes = (uint32_t)(nkb * TM_TILE_KB);
  const uint32_t w_s = base, t_s = base + 4u * wbytes;
  const uint32_t misc = t_s + 3u * tbytes;
  const uint32_t wbar = misc, bar1 = misc + 8, bar2 = misc + 16, tmem_slot = misc + 24;
  float* fl = reinterpret_cast<float*>(base_ptr + (misc - base) + 64);
  float* gam_s = fl;                  // [128]
  float* bet_s = fl + 128;            // [128]
  float* biasp_s = fl + 256;          // [128] projection bias
  float* biasc_s = fl + 384;          // [128] channel-branch bias
  float* ac_s = fl + 512;             // [128] a_c / C of the current batch sample
  float* mu_s = fl + 640;             // [128] per-pixel LayerNorm mean of the current tile
  float* rstd_s = fl + 768;           // [128]
  float* biashw_s = fl + 896;         // [2][8 r'][16]: token-mixer biases, S columns per r' padded to 16 (one LDS.128 per 4)
  float* ahw_s = fl + 1152;           // [2][8 g][20]: a_h / a_w / C per channel group, padded to 20 (conflict-free LDS.128)
  float* part_s = reinterpret_cast<float*>(base_ptr + 3u * wbytes);     // phase 0 only: [16 warps][8 g][16], in the (unused) proj-weight slot

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  constexpr float inv_c = 1.f / (float)C;

  if (tid == 0) {
    mbar_init(wbar, 1); mbar_init(bar1, 1); mbar_init(bar2, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // zero the operand tiles once: the K padding (channels C .. 64*nkb) is never written afterwards and must not hold NaNs
  for (uint32_t o = (uint32_t)tid * 16u; o < 3u * tbytes; o += TM_THREADS * 16u) sts128(t_s + o, make_uint4(0, 0, 0, 0));
  for (int c = tid; c < 128; c += TM_THREADS) {
    gam_s[c] = c < C ? p.ln_g[c] : 0.f;
    bet_s[c] = c < C ? p.ln_b[c] : 0.f;
    biasp_s[c] = (c < C && p.bias[3]) ? p.bias[3][c] : 0.f;
    biasc_s[c] = (c < C && p.bias[2]) ? p.bias[2][c] : 0.f;
    ac_s[c] = 0.f;
  }
  for (int c = tid; c < 256; c += TM_THREADS) {
    const int j = c >> 7, rp = (c >> 4) & 7, i = c & 15;
    biashw_s[c] = (i < S && p.bias[j]) ? p.bias[j][rp * S + i] : 0.f;
  }
  for (int c = tid; c < 320; c += TM_THREADS) ahw_s[c] = 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<const uint32_t*>(base_ptr + (tmem_slot - base));

  // resident weights: one SWIZZLE_128B box {64 k, C rows} per (matrix, K block); phase 0 does not need the projection
  const int n_mat = p.phase ? 4 : 3;
  if (warp == 0 && elect_one()) {
    mbar_expect_tx(wbar, (uint32_t)n_mat * wbytes);
    for (int j = 0; j < n_mat; ++j) {
      const CUtensorMap* m = j == 0 ? &mapW0 : (j == 1 ? &mapW1 : (j == 2 ? &mapW2 : &mapW3));
      for (int kb = 0; kb < nkb; ++kb) tma_load_2d(w_s + j * wbytes + kb * C * 128, m, wbar, kb * 64, 0);
    }
  }

  // ---- producer mapping: a half-warp owns a pixel (16 bytes = 8 channels per lane), 32 pixels per pass, 4 passes per tile
  const int pl = lane & 15, psel = lane >> 4;
  const bool pact = pl * 8 < C;
  int pg[4], pj[4];                   // channel pair (8*pl + 2i, +1) -> (channel group g, word j = s / 2 inside the group)
#pragma unroll
  for (int i = 0; i < 4; ++i) { const int c = pl * 8 + 2 * i; pg[i] = c / S; pj[i] = (c - pg[i] * S) >> 1; }
  // ---- drain mapping: warp = (TMEM lane quarter q, column quarter cq); row m = 32 q + lane
  const int q = warp & 3, cq = warp >> 2;
  const int m = q * 32 + lane;
  constexpr int cpq = C >> 2;         // columns per quarter (28 at C = 112): two r' of S columns for the token mixers
  const int mg = m & 7, mw = (m >> 3) & 7, mb = m >> 6;                 // token row (block mb, position mw, channel group mg)
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  const int tiles_per_img = p.tiles_h * p.tiles_w;
  uint32_t ph1 = 0, ph2 = 0;
  int cur_b = -1;
  long long fix_acc = 0;              // phase 0: this thread's channel (tid < C), 2^24 fixed point

  auto tile_coords = [&](long long t, int& img, int& h0, int& w0) {
    img = (int)(t / tiles_per_img);
    const int rem = (int)(t - (long long)img * tiles_per_img);
    const int th = rem / p.tiles_w;
    h0 = th * 8; w0 = (rem - th * p.tiles_w) * 16;
  };
  uint4 pre[4];
  auto prefetch = [&](long long t) {
    int img, h0, w0;
    tile_coords(t, img, h0, w0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pm = i * 32 + warp * 2 + psel, r = pm >> 4, col = pm & 15;
      const bool v = pact && t < p.n_tiles && h0 + r < p.H && w0 + col < p.W;
      pre[i] = v ? __ldg(reinterpret_cast<const uint4*>(p.x + (((long long)img * p.H + h0 + r) * p.W + w0 + col) * C + pl * 8))
                 : make_uint4(0, 0, 0, 0);
    }
  };
  pdl_wait();
  prefetch(blockIdx.x);

  for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    int img, h0, w0;
    tile_coords(tile, img, h0, w0);
    const int bsm = img % p.B;
    if (bsm != cur_b) {               // block-uniform
      if (p.phase == 0) {
        if (cur_b >= 0 && tid < C && fix_acc != 0) atomicAdd(p.acc + (long long)cur_b * C + tid, (unsigned long long)fix_acc);
        fix_acc = 0;
      } else {
        for (int c = tid; c < C; c += TM_THREADS) {
          const int g = c / S, i = c - g * S;
          ahw_s[g * 20 + i] = p.a[((long long)bsm * 3 + 0) * C + c] * inv_c;
          ahw_s[160 + g * 20 + i] = p.a[((long long)bsm * 3 + 1) * C + c] * inv_c;
          ac_s[c] = p.a[((long long)bsm * 3 + 2) * C + c] * inv_c;
        }
      }
      cur_b = bsm;
    }
    // ================= producer: LayerNorm + the three operand layouts =================
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pm = i * 32 + warp * 2 + psel, r = pm >> 4, col = pm & 15;
      const bool valid = h0 + r < p.H && w0 + col < p.W;
      float v[8];
      const uint32_t wv[4] = {pre[i].x, pre[i].y, pre[i].z, pre[i].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) { v[2 * k] = __uint_as_float(wv[k] << 16); v[2 * k + 1] = __uint_as_float(wv[k] & 0xffff0000u); }
      float s = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
#pragma unroll
      for (int sh = 8; sh >= 1; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
      const float mean = s * inv_c;
      float q2 = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) { v[e] -= mean; q2 += v[e] * v[e]; }
      q2 = pact ? q2 : 0.f;
#pragma unroll
      for (int sh = 8; sh >= 1; sh >>= 1) q2 += __shfl_xor_sync(0xffffffffu, q2, sh);
      const float rstd = rsqrtf(q2 * inv_c + p.eps);
      if (pl == 0) { mu_s[pm] = mean; rstd_s[pm] = rstd; }
      if (pact) {
        uint32_t o[4];
        const float4 g0 = *reinterpret_cast<const float4*>(gam_s + pl * 8), g1 = *reinterpret_cast<const float4*>(gam_s + pl * 8 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(bet_s + pl * 8), b1 = *reinterpret_cast<const float4*>(bet_s + pl * 8 + 4);
        const float lg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, lb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
          o[k] = valid ? bf2(v[2 * k] * rstd * lg[2 * k] + lb[2 * k], v[2 * k + 1] * rstd * lg[2 * k + 1] + lb[2 * k + 1]) : 0u;   // zero padding after the norm
        // channel branch: pixel-major
        sts128(t_s + 2u * tbytes + (uint32_t)((pl >> 3) * TM_TILE_KB + pm * 128 + (((pl & 7) ^ (pm & 7)) << 4)), make_uint4(o[0], o[1], o[2], o[3]));
        const int bb = col >> 3, wq = col & 7;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          sts32(t_s + sw_word(bb * 64 + wq * 8 + pg[k], r * S2 + pj[k]), o[k]);                    // H tokens: (column, group) x (row, s)
          sts32(t_s + tbytes + sw_word(bb * 64 + r * 8 + pg[k], wq * S2 + pj[k]), o[k]);           // W tokens: (row, group) x (column, s)
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    // ================= branch GEMMs: acc_j = A_j . W_j^T (M = 128, N = C, K = C) =================
    if (warp == 0) {
      if (tile == blockIdx.x) mbar_wait(wbar, 0);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll 1
        for (int j = 0; j < 3; ++j)
          for (int kb = 0; kb < nkb; ++kb) {
            const int nks = min(4, (C - kb * 64 + 15) >> 4);
            const uint64_t ad = smem_desc_sw128(t_s + j * tbytes + kb * TM_TILE_KB), bd = smem_desc_sw128(w_s + j * wbytes + kb * C * 128);
            for (int ks = 0; ks < nks; ++ks) umma_f16(tmem_base + j * 128, ad + (uint64_t)(ks * 2), bd + (uint64_t)(ks * 2), idesc, (uint32_t)((kb | ks) > 0));
          }
        umma_commit(bar1);
      }
      __syncwarp();
    }
    prefetch(tile + gridDim.x);       // the next tile's pixels are in flight during the drains
    mbar_wait(bar1, ph1); ph1 ^= 1;
    tc_fence_after();

    const bool rowv_h0 = w0 + mb * 8 + mw < p.W;          // H tokens: this thread's pixel column
    const bool rowv_w0 = h0 + mw < p.H;                   // W tokens: this thread's pixel row
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cq * cpq);
    if (p.phase == 0) {
      // ================= phase 0: per-channel sums of the three branches over the valid pixels =================
      // Column sums by TRANSPOSE-REDUCTION: at every butterfly level a lane keeps one half of its values and sends the
      // other half, so 32 columns over 32 lanes cost 31 shuffles (a plain xor-reduce per column: 160) -- the shuffles
      // share the MIO pipe with the shared-memory traffic, which bounds this kernel.
      float* my = part_s + warp * 128;                     // [8 g][16]
#pragma unroll 1
      for (int j = 0; j < 2; ++j) {
        uint32_t u[32];
        tmem_ld16_issue(tq + j * 128, u); tmem_ld16_issue(tq + j * 128 + 16, u + 16); tmem_ld_wait();
        float ps[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) ps[i] = 0.f;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int rp = 2 * cq + rr;                      // r' of these S columns
          const bool pv = j == 0 ? (rowv_h0 && h0 + rp < p.H) : (rowv_w0 && w0 + mb * 8 + rp < p.W);
          float bq[16];
#pragma unroll
          for (int k = 0; k < 4; ++k) *reinterpret_cast<float4*>(bq + 4 * k) = *reinterpret_cast<const float4*>(biashw_s + j * 128 + rp * 16 + 4 * k);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (i < S) {
              const float val = fmaxf(__uint_as_float(u[rr * S + i]) + bq[i], 0.f);
              ps[i] += pv ? val : 0.f;
            }
        }
        // lanes that share the channel group g = lane & 7 differ in lane bits 3 and 4: 16 -> 8 -> 4 values
        {
          const bool up = (lane & 16) != 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float send = up ? ps[i] : ps[i + 8], keep = up ? ps[i + 8] : ps[i];
            ps[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
          }
        }
        {
          const bool up = (lane & 8) != 0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float send = up ? ps[i] : ps[i + 4], keep = up ? ps[i + 4] : ps[i];
            ps[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
          }
        }
        // this lane now holds s' = 4 * (lane >> 3) .. + 3 of group g (s' >= S: zeros)
        float4* dst = reinterpret_cast<float4*>(my + (lane & 7) * 16 + (lane >> 3) * 4);
        if (j == 0) *dst = make_float4(ps[0], ps[1], ps[2], ps[3]);
        else { const float4 o = *dst; *dst = make_float4(o.x + ps[0], o.y + ps[1], o.z + ps[2], o.w + ps[3]); }
        __syncwarp();
      }
      {
        uint32_t u[32];
        tmem_ld16_issue(tq + 256, u); tmem_ld16_issue(tq + 256 + 16, u + 16); tmem_ld_wait();
        const bool pv = h0 + (m >> 4) < p.H && w0 + (m & 15) < p.W;
        float t[32];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (4 * k < cpq) b4 = *reinterpret_cast<const float4*>(biasc_s + cq * cpq + 4 * k);
          const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) t[4 * k + e] = (pv && 4 * k + e < cpq) ? fmaxf(__uint_as_float(u[4 * k + e]) + bb[e], 0.f) : 0.f;
        }
#pragma unroll
        for (int half = 16; half >= 1; half >>= 1) {
          const bool up = (lane & half) != 0;
#pragma unroll
          for (int i = 0; i < half; ++i) {
            const float send = up ? t[i] : t[i + half], keep = up ? t[i + half] : t[i];
            t[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
          }
        }
        // lane L holds the total of column L of this quarter
        if (lane < cpq) { const int c = cq * cpq + lane, g = c / S; my[g * 16 + (c - g * S)] += t[0]; }
      }
      tc_fence_before();
      __syncthreads();
      if (tid < C) {
        float tot = 0.f;
#pragma unroll
        for (int w2 = 0; w2 < 16; ++w2) tot += part_s[w2 * 128 + (tid / S) * 16 + (tid % S)];
        fix_acc += __float2ll_rn(tot * inv_c * 16777216.f);
      }
      __syncthreads();
      continue;
    }
    // ================= phase 1: B_j = a_j / C . ReLU(acc_j + b_j), pixel-major, into the operand buffers =================
#pragma unroll 1
    for (int j = 0; j < 2; ++j) {
      uint32_t u[32];
      tmem_ld16_issue(tq + j * 128, u); tmem_ld16_issue(tq + j * 128 + 16, u + 16); tmem_ld_wait();
      float aq[16];
#pragma unroll
      for (int k = 0; k < 4; ++k) *reinterpret_cast<float4*>(aq + 4 * k) = *reinterpret_cast<const float4*>(ahw_s + j * 160 + mg * 20 + 4 * k);
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int rp = 2 * cq + rr;
        float bq[16];
#pragma unroll
        for (int k = 0; k < 4; ++k) *reinterpret_cast<float4*>(bq + 4 * k) = *reinterpret_cast<const float4*>(biashw_s + j * 128 + rp * 16 + 4 * k);
        // H: pixel (row rp, column mb*8 + mw) ; W: pixel (row mw, column mb*8 + rp) -- channels mg*S + i
        const int prow = j == 0 ? rp * 16 + mb * 8 + mw : mw * 16 + mb * 8 + rp;
#pragma unroll
        for (int i = 0; i < 16; i += 2)
          if (i < S) {
            const float f0 = fmaxf(__uint_as_float(u[rr * S + i]) + bq[i], 0.f) * aq[i];
            const float f1 = fmaxf(__uint_as_float(u[rr * S + i + 1]) + bq[i + 1], 0.f) * aq[i + 1];
            sts32(t_s + j * tbytes + sw_word(prow, mg * S2 + (i >> 1)), bf2(f0, f1));
          }
      }
    }
    {
      uint32_t u[32];
      tmem_ld16_issue(tq + 256, u); tmem_ld16_issue(tq + 256 + 16, u + 16); tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        if (i < cpq) {
          const float4 b4 = *reinterpret_cast<const float4*>(biasc_s + cq * cpq + i), a4 = *reinterpret_cast<const float4*>(ac_s + cq * cpq + i);
          const float f0 = fmaxf(__uint_as_float(u[i]) + b4.x, 0.f) * a4.x, f1 = fmaxf(__uint_as_float(u[i + 1]) + b4.y, 0.f) * a4.y;
          const float f2 = fmaxf(__uint_as_float(u[i + 2]) + b4.z, 0.f) * a4.z, f3 = fmaxf(__uint_as_float(u[i + 3]) + b4.w, 0.f) * a4.w;
          const int word = (cq * cpq + i) >> 1;            // two words, 8-byte aligned (cpq % 4 == 0)
          sts64(t_s + 2u * tbytes + sw_word(m, word), bf2(f0, f1), bf2(f2, f3));
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    // ================= projection: z = (B_h + B_w + B_c) . Wp^T, summed by the tensor core =================
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll 1
        for (int j = 0; j < 3; ++j)
          for (int kb = 0; kb < nkb; ++kb) {
            const int nks = min(4, (C - kb * 64 + 15) >> 4);
            const uint64_t ad = smem_desc_sw128(t_s + j * tbytes + kb * TM_TILE_KB), bd = smem_desc_sw128(w_s + 3u * wbytes + kb * C * 128);
            for (int ks = 0; ks < nks; ++ks) umma_f16(tmem_base + 384, ad + (uint64_t)(ks * 2), bd + (uint64_t)(ks * 2), idesc, (uint32_t)((j | kb | ks) > 0));
          }
        umma_commit(bar2);
      }
      __syncwarp();
    }
    // the gate needs x again at (pixel = accumulator row, 28 channels): a row-per-lane global access touches 32 lines per
    // instruction (the L1 tag stage was the phase-1 bottleneck), so the x tile comes in with coalesced 16-byte loads
    // while the projection runs, is staged in the (now dead) operand buffers with a conflict-free pitch, and the
    // result leaves the same way
    constexpr int CPP = C / 8, NCH = 128 * CPP, PIT = C * 2 + 8;      // 16-byte chunks per pixel / per tile; staging pitch
    constexpr uint32_t OS_OFF = (uint32_t)((128 * PIT + 1023) / 1024 * 1024);      // output staging behind the x staging (both inside the 3 operand tiles)
    static_assert(2 * OS_OFF <= 3 * tbytes, "staging does not fit");
    uint4 xr[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int id = tid + TM_THREADS * k, pm = id / CPP, c16 = id - pm * CPP, r = pm >> 4, col = pm & 15;
      xr[k] = make_uint4(0, 0, 0, 0);
      if (id < NCH && h0 + r < p.H && w0 + col < p.W)
        xr[k] = __ldg(reinterpret_cast<const uint4*>(p.x + (((long long)img * p.H + h0 + r) * p.W + w0 + col) * C + c16 * 8));
    }
    mbar_wait(bar2, ph2); ph2 ^= 1;
    tc_fence_after();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int id = tid + TM_THREADS * k, pm = id / CPP, c16 = id - pm * CPP;
      if (id < NCH) {
        sts64(t_s + (uint32_t)(pm * PIT + c16 * 16), xr[k].x, xr[k].y);
        sts64(t_s + (uint32_t)(pm * PIT + c16 * 16 + 8), xr[k].z, xr[k].w);
      }
    }
    __syncthreads();
    // ================= gate + residual: out = x + m * (LN2(x) + z) * tanh(z) =================
    {
      uint32_t u[32];
      tmem_ld16_issue(tq + 384, u); tmem_ld16_issue(tq + 384 + 16, u + 16); tmem_ld_wait();
      const float mu = mu_s[m], rs = rstd_s[m];
      const uint32_t xs = t_s + (uint32_t)(m * PIT + cq * cpq * 2), os = xs + OS_OFF;
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        if (i < cpq) {
          const uint2 xw = lds64(xs + i * 2);
          const float xv[4] = {__uint_as_float(xw.x << 16), __uint_as_float(xw.x & 0xffff0000u), __uint_as_float(xw.y << 16),
                               __uint_as_float(xw.y & 0xffff0000u)};
          const float4 bp4 = *reinterpret_cast<const float4*>(biasp_s + cq * cpq + i), g4 = *reinterpret_cast<const float4*>(gam_s + cq * cpq + i),
                       e4 = *reinterpret_cast<const float4*>(bet_s + cq * cpq + i);
          const float bpv[4] = {bp4.x, bp4.y, bp4.z, bp4.w}, gv[4] = {g4.x, g4.y, g4.z, g4.w}, ev[4] = {e4.x, e4.y, e4.z, e4.w};
          float o[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float z = __uint_as_float(u[i + k]) + bpv[k];
            const float xn = (xv[k] - mu) * rs * gv[k] + ev[k];
            float th;
            asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(z));
            o[k] = xv[k] + p.m_scaling * ((xn + z) * th);
          }
          sts64(os + i * 2, bf2(o[0], o[1]), bf2(o[2], o[3]));
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int id = tid + TM_THREADS * k, pm = id / CPP, c16 = id - pm * CPP, r = pm >> 4, col = pm & 15;
      if (id < NCH && h0 + r < p.H && w0 + col < p.W) {
        const uint2 lo = lds64(t_s + OS_OFF + (uint32_t)(pm * PIT + c16 * 16)), hi = lds64(t_s + OS_OFF + (uint32_t)(pm * PIT + c16 * 16 + 8));
        *reinterpret_cast<uint4*>(p.out + (((long long)img * p.H + h0 + r) * p.W + w0 + col) * C + c16 * 8) = make_uint4(lo.x, lo.y, hi.x, hi.y);
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  if (p.phase == 0 && cur_b >= 0 && tid < C && fix_acc != 0) atomicAdd(p.acc + (long long)cur_b * C + tid, (unsigned long long)fix_acc);
  if (warp == 0 && blockIdx.x >= p.n_tiles) mbar_wait(wbar, 0);      // never exit with weight boxes in flight
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

__global__ void fixsum_finalize_kernel(const unsigned long long* __restrict__ acc, float* __restrict__ out, int n, float scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)((double)(long long)acc[i] * (1.0 / 16777216.0) * (double)scale);
}

typedef CUresult (*EncodeTiledFnT)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFnT g_enc_t = nullptr;
static int g_sms_t = 0, g_smem_t = 0, g_dev_t = -1;
static std::once_flag g_tab_once;

static void tab_init() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
    g_enc_t = (EncodeTiledFnT)fn;
  cudaGetDevice(&g_dev_t);
  cudaDeviceGetAttribute(&g_sms_t, cudaDevAttrMultiProcessorCount, g_dev_t);
  cudaDeviceGetAttribute(&g_smem_t, cudaDevAttrMaxSharedMemoryPerBlockOptin, g_dev_t);
  if (g_smem_t > 0) {
    cudaFuncSetAttribute(tab_mix_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_t);
    cudaFuncSetAttribute(tab_mix_kernel<112>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_t);
    cudaFuncSetAttribute(tab_mix_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_t);
  }
}

}  // namespace vmgb

using namespace vmgb;

// Shapes the fused kernel takes; everything else returns VMGB_EUNSUPPORTED and the caller issues the separate launches.
static bool tab_mix_shape_ok(int C, int chunk) {
  return chunk == 8 && (C == 64 || C == 112 || C == 128);
}

extern "C" int vmgb_tab_mix(const void* x, const float* ln_g, const float* ln_b, float eps, const void* w_h, const void* w_w, const void* w_c,
                            const void* w_p, int Kp, const float* b_h, const float* b_w, const float* b_c, const float* b_p, int phase,
                            void* acc, const float* a, void* out, float m_scaling, int n_img, int B, int H, int W, int C, int chunk,
                            void* stream) {
  if (!x || !ln_g || !ln_b || !w_h || !w_w || !w_c || n_img < 1 || B < 1 || n_img % B || H < 1 || W < 1) return VMGB_EINVAL;
  if (phase != 0 && phase != 1) return VMGB_EINVAL;
  if (phase == 0 ? !acc : (!a || !out || !w_p)) return VMGB_EINVAL;
  if (!tab_mix_shape_ok(C, chunk)) return VMGB_EUNSUPPORTED;
  if (Kp < C || (Kp & 7)) return VMGB_EINVAL;
  const void* ptrs[5] = {x, w_h, w_w, w_c, phase ? w_p : w_c};
  for (const void* q : ptrs) if (reinterpret_cast<uintptr_t>(q) & 15) return VMGB_EINVAL;
  if (phase && (reinterpret_cast<uintptr_t>(out) & 15)) return VMGB_EINVAL;
  std::call_once(g_tab_once, tab_init);
  if (!g_enc_t || g_sms_t < 1) return VMGB_EUNSUPPORTED;
  int cur = -1;
  if (cudaGetDevice(&cur) != cudaSuccess || cur != g_dev_t) return VMGB_EUNSUPPORTED;

  TabP p;
  p.x = (const __nv_bfloat16*)x; p.out = (__nv_bfloat16*)out; p.ln_g = ln_g; p.ln_b = ln_b;
  p.bias[0] = b_h; p.bias[1] = b_w; p.bias[2] = b_c; p.bias[3] = b_p;
  p.a = a; p.acc = (unsigned long long*)acc;
  p.n_img = n_img; p.B = B; p.H = H; p.W = W; p.C = C; p.S = C / 8; p.nkb = (C + 63) / 64;
  p.tiles_h = (H + 7) / 8; p.tiles_w = (W + 15) / 16; p.phase = phase;
  p.n_tiles = (long long)n_img * p.tiles_h * p.tiles_w;
  p.eps = eps; p.m_scaling = m_scaling;
  const int wbytes = p.nkb * C * 128, tbytes = p.nkb * TM_TILE_KB;
  const int smem = 1024 + 4 * wbytes + 3 * tbytes + 64 + 1472 * 4 + 64;
  if (smem > g_smem_t) return VMGB_EUNSUPPORTED;
  CUtensorMap mW[4];
  const void* ws[4] = {w_h, w_w, w_c, phase ? w_p : w_c};
  for (int j = 0; j < 4; ++j) {
    cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)C};
    cuuint64_t str[1] = {(cuuint64_t)Kp * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)C};
    cuuint32_t es[2] = {1, 1};
    if (g_enc_t(&mW[j], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ws[j]), dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return VMGB_EINVAL;
  }
  const long long grid = p.n_tiles < g_sms_t ? p.n_tiles : g_sms_t;
  g_pack_pending = 0;                   // no programmatic launch here: the (re)packed weights are complete before this kernel starts
  if (C == 64) tab_mix_kernel<64><<<(unsigned)grid, TM_THREADS, smem, as_stream(stream)>>>(mW[0], mW[1], mW[2], mW[3], p);
  else if (C == 112) tab_mix_kernel<112><<<(unsigned)grid, TM_THREADS, smem, as_stream(stream)>>>(mW[0], mW[1], mW[2], mW[3], p);
  else tab_mix_kernel<128><<<(unsigned)grid, TM_THREADS, smem, as_stream(stream)>>>(mW[0], mW[1], mW[2], mW[3], p);
  return check_launch();
}

extern "C" int vmgb_fixsum_finalize(const void* acc, float* out, int n, float scale, void* stream) {
  if (!acc || !out || n < 1) return VMGB_EINVAL;
  fixsum_finalize_kernel<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>((const unsigned long long*)acc, out, n, scale);
  return check_launch();
}
