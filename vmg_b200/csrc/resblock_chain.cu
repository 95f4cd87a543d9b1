// resblock_chain.cu -- the residual-block chain of the trajectory module as ONE persistent kernel
// (trajectory.py:16-52 ResidualBlocksWithInputConv.main[2], :165-221 ResidualBlockNoBN0:
//  y <- y + s * conv2(relu(conv1(y))), 15 blocks = 30 dependent 3x3 convolutions per time step and sweep).
//
// Why: issued as 30 launches of the CTA-pair weight-stationary kernel (conv_tc.cu), every convolution pays ~14 us of
// one-off cost (launch, barrier / TMEM / cluster prologue, pipeline fill and drain) on top of ~15 us of steady state
// (profiles/r1_bottleneck_probes.md); the recurrence is sequential, so the launches cannot be made larger.  Here the
// 30 layers run inside one kernel and never drain the pipeline:
//   * every cluster (CTA pair, tcgen05 cta_group::2, M = 256) owns a fixed, contiguous run of tile pairs for ALL layers
//     and walks it middle-out, so the tiles a layer needs from its neighbours (the 3x3 halo) were finished most of a
//     layer-time ago;
//   * instead of a grid barrier per layer, the halo dependency is tracked per tile pair: the epilogue publishes
//     "layer l of pair p is in memory" with a release increment of flags[p], the TMA producer acquires the 9 flags of
//     the neighbourhood before it requests the activation boxes of (p, l+1).  Write-after-read hazards on the three
//     ping-pong buffers are covered by the same flags (a buffer is rewritten two layers after it was read, and layer
//     l+1 of every reader is a dependency of layer l+2 of the writer);
//   * the 129 KB weight slice of the next layer streams into shared memory K block by K block: as soon as the last tile of
//     the current layer has retired its MMAs of block 0 (tcgen05.commit -> wfree0[0]) the next layer's block 0 is loaded,
//     box by box in consumption order, while block 1 of that tile is still being multiplied and the epilogue drains;
//   * the residual tile of conv2 arrives by TMA one tile ahead (first tile of a layer: during the preceding conv1 layer).
// Warp roles (14 warps): 0 activation producer, 1 MMA issuer (leader CTA), 2 TMEM allocation then DEPENDENCY CHECKER,
// 3 weight (+ bias) producer, 4-11 epilogue, 12 PUBLISHER, 13 residual requester.  The epilogue warps never meet at a
// block barrier and never execute a GPU-scope fence: each warp bumps a shared-memory counter when its stores of a tile are
// issued (pub_cnt) / when it holds its residual pieces in registers (res_cnt); the publisher turns "8 warps done" into ONE
// red.release.gpu on flags[pair] (the MEMBAR.ALL.GPU + L1 invalidation of that release took 40 % of the time of the epilogue
// warp that used to execute it, with the other seven waiting for it at the next barrier -- profiles/r2_chain_ncu.md), the
// requester re-arms the residual box, and the checker polls the 9 neighbour flags of every tile ahead of the producer and
// posts its progress in shared memory (the producer used to spend 2/3 of its time in ld.acquire.gpu round trips).
// Cross-proxy ordering: outputs are written with generic stores and re-read by TMA (async proxy) from other SMs:
// writer = __syncwarp + red.release.cta (epilogue warp) -> ld.acquire.cta + red.release.gpu (publisher);
// reader = ld.acquire.gpu (checker) -> st.release.cta / ld.acquire.cta -> fence.proxy.async (producer) -> TMA.
// Every spin loop traps after ~2 s: a lost dependency becomes an error, never a hang.
#include "tc_ptx.cuh"
#include <mutex>

namespace vmgb {

struct ChainP {
  int n_img, H, W, C;
  int tiles_h, pairs_w;               // 16 x 8 tiles; a pair = two horizontally adjacent tiles (rank 0 left, rank 1 right)
  int n_layers, nkb, ra, a_unit_bytes, b_stage_bytes;
  int slices, NT;                     // output channels in `slices` slices of NT columns (NT % 16 == 0): C <= 128 -> 1 x C, C = 144 -> 2 x 80
  int n_pairs, n_clusters, segs;      // segs > 0: row-aligned ownership (cluster = (tile row, segment)); 0: equal raster ranges
  const float* bias;                  // [n_layers][C]
  __nv_bfloat16* buf[4];              // 0 / 1: the block ping-pong (buf[0] holds the chain input), 2: conv1 output, 3: final output
  unsigned* flags;                    // this launch's per-pair counters (zero on entry)
  unsigned* flags_next;               // the next launch's counters: zeroed here
  float res_scale;
  int zero;                           // always 0, opaque to the compiler (forces a data dependency)
};

constexpr int CH_THREADS = 448;
constexpr int CH_MAX_BOXES = 27;       // 9 taps x up to 3 K blocks

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// shared-memory progress counters between the warps of one CTA
__device__ __forceinline__ unsigned lds_acquire_cta(uint32_t a) {
  unsigned v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_release_cta(uint32_t a, unsigned v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void reds_release_cta_add(uint32_t a, unsigned v) {
  asm volatile("red.release.cta.shared::cta.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
// all lanes spin until the counter at `a` reaches `target` (monotonic counters: no phase to miss)
__device__ __forceinline__ void spin_counter(uint32_t a, unsigned target) {
  const long long t0 = clock64();
  while ((int)(lds_acquire_cta(a) - target) < 0) {
    if (clock64() - t0 > (1LL << 32)) __trap();
  }
}

// first / end pair of cluster c
__device__ __forceinline__ void chain_range(const ChainP& p, int c, int& p0, int& p1) {
  if (p.segs > 0) {
    const int row = c / p.segs, seg = c - row * p.segs;
    p0 = row * p.pairs_w + (seg * p.pairs_w) / p.segs;
    p1 = row * p.pairs_w + ((seg + 1) * p.pairs_w) / p.segs;
  } else {
    p0 = (int)(((long long)c * p.n_pairs) / p.n_clusters);
    p1 = (int)(((long long)(c + 1) * p.n_pairs) / p.n_clusters);
  }
}
// middle-out walk over n items: slot 0 = middle, then alternately right / left
__device__ __forceinline__ int chain_order(int n, int s) {
  const int mid = (n - 1) >> 1;
  return (s & 1) ? mid + ((s + 1) >> 1) : mid - (s >> 1);
}

__global__ void __launch_bounds__(CH_THREADS, 1)
resblock_chain_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                      const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB,
                      const __grid_constant__ CUtensorMap mapR0, const __grid_constant__ CUtensorMap mapR1, const __grid_constant__ ChainP p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr int BWX = 10, TAPS = 9, UROWS = 18;
  const int nboxes = p.nkb * TAPS;
  const uint32_t a_base = smem_base;
  const uint32_t w_base = a_base + p.ra * p.a_unit_bytes;
  const uint32_t bar_base = (w_base + nboxes * p.b_stage_bytes + 1023u) & ~1023u;
  const uint32_t afull0 = bar_base, aempty0 = afull0 + 8 * p.ra;
  const uint32_t wfull0 = aempty0 + 8 * p.ra;
  const uint32_t tfull0 = wfull0 + 8 * CH_MAX_BOXES, tempty0 = tfull0 + 16;
  const uint32_t rfull = tempty0 + 16, wfree0 = rfull + 8;      // wfree0 + 8 * kb: the weight boxes of K block kb may be replaced
  const uint32_t tmem_slot = wfree0 + 32;      // 3 wfree barriers + pad: keeps bias_s (tmem_slot + 16) 16-byte aligned
  const uint32_t dep_ready = tmem_slot + 4, pub_cnt = tmem_slot + 8, res_cnt = tmem_slot + 12;      // progress counters
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  float* bias_s = reinterpret_cast<float*>(smem_raw + (tmem_slot + 16 - smem_u32(smem_raw)));      // [2][128]: layer parity
  const uint32_t res_s = (smem_u32(bias_s) + 1024u + 127u) & ~127u;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.ra; ++i) { mbar_init(afull0 + 8 * i, 1); mbar_init(aempty0 + 8 * i, 1); }
    for (int i = 0; i < nboxes; ++i) mbar_init(wfull0 + 8 * i, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 512); }
    mbar_init(rfull, 1);
    mbar_init(wfree0, 1); mbar_init(wfree0 + 8, 1); mbar_init(wfree0 + 16, 1);
    tmem_slot_ptr[1] = 0u; tmem_slot_ptr[2] = 0u; tmem_slot_ptr[3] = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // cluster = (run of tile pairs, output-channel slice): the slices of a pair are computed by different clusters from the
  // same activations; a pair of layer l is complete when all 2 * slices CTAs have published it
  const int cl = (blockIdx.x >> 1) / p.slices, slice = (blockIdx.x >> 1) - cl * p.slices;
  int p0, p1;
  chain_range(p, cl, p0, p1);
  const int n_own = p1 - p0;
  const int NT = p.NT, NH = p.NT >> 1;
  const int cbase = slice * NT;                       // first output channel of this slice
  const int NL = p.n_layers;

  auto tile_of = [&](int pair, int& img, int& h0, int& w0) {
    const int row = pair / p.pairs_w, pc = pair - row * p.pairs_w;
    img = row / p.tiles_h;
    h0 = (row - img * p.tiles_h) * 16;
    w0 = (2 * pc + (int)rank) * 8;
  };

  if (warp == 0) {
    // ===================== activation producer (both CTAs) =====================
    const uint32_t afull0_l = mapa_rank(afull0, 0);
    pdl_wait();                                       // the chain input (and the flag arrays) belong to the predecessors until here
    // zero the NEXT launch's counters of the pairs this cluster owns (nobody touches that array during this launch)
    if (rank == 0) for (int i = lane; i < n_own; i += 32) p.flags_next[p0 + i] = 0u;
    int ia = 0; uint32_t pa = 0;
    constexpr uint32_t a_bytes = (uint32_t)UROWS * BWX * 128;
    for (int l = 0; l < NL; ++l) {
      const CUtensorMap* mA = (l & 1) ? &mapA2 : (((l >> 1) & 1) ? &mapA1 : &mapA0);
      for (int s = 0; s < n_own; ++s) {
        const int pair = p0 + chain_order(n_own, s);
        int img, h0, w0;
        tile_of(pair, img, h0, w0);
        if (l > 0) {                                  // halo dependency, resolved ahead of time by the checker warp
          spin_counter(dep_ready, (unsigned)(l * n_own + s + 1));
          fence_proxy_async();
        }
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(aempty0 + 8 * ia, pa ^ 1);
          if (elect_one()) {
            if (rank == 0) mbar_expect_tx(afull0 + 8 * ia, 2 * a_bytes);
            tma_load_4d_2sm(a_base + ia * p.a_unit_bytes, mA, afull0_l + 8 * ia, kb * 64, w0 - 1, h0 - 1, img);
          }
          __syncwarp();
          if (++ia == p.ra) { ia = 0; pa ^= 1; }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== weight producer (both CTAs): the slice of layer l+1 replaces layer l's as soon as it is free =====================
    const uint32_t wfull0_l = mapa_rank(wfull0, 0);
    for (int l = 0; l < NL; ++l) {
      for (int kb = 0; kb < p.nkb; ++kb) {
        // the boxes of K block kb are free once the LAST tile of layer l-1 has retired its MMAs of that block: block 0 of
        // the next layer streams in while block 1 of the last tile is still being multiplied
        if (l > 0) mbar_wait(wfree0 + 8 * kb, (uint32_t)((l - 1) & 1));
        if (kb == 0) {
          // bias of layer l: its buffer was last read by the epilogue of layer l-2, and the last tile of layer l-1 (whose
          // MMAs have started) depended on the last tile of layer l-2 having been drained
          for (int c = lane; c < NT; c += 32) bias_s[(l & 1) * 128 + c] = cbase + c < p.C ? __ldg(p.bias + (long long)l * p.C + cbase + c) : 0.f;
          __threadfence_block();
          __syncwarp();
        }
        if (elect_one()) {
          for (int tap = 0; tap < TAPS; ++tap) {
            const int box = kb * TAPS + tap;
            if (rank == 0) mbar_expect_tx(wfull0 + 8 * box, (uint32_t)(2 * NH * 128));
            // rows past the slice's real channels (cbase + .. >= C) belong to the next tap: finite garbage in accumulator columns
            // that are never stored (past the tensor end the box is zero-filled)
            tma_load_2d_2sm(w_base + box * p.b_stage_bytes, &mapB, wfull0_l + 8 * box, kb * 64, (l * TAPS + tap) * p.C + cbase + (int)rank * NH);
          }
        }
        __syncwarp();
      }
    }
    for (int kb = 0; kb < p.nkb; ++kb) mbar_wait(wfree0 + 8 * kb, (uint32_t)((NL - 1) & 1));   // the last multicast arrivals have landed before this CTA may exit
  } else if (warp == 1 && rank == 0) {
    // ===================== MMA issuer (leader CTA) =====================
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    int ia = 0; uint32_t pa = 0;
    uint32_t acc_phase0 = 0, acc_phase1 = 0;
    int it = 0;
    const uint32_t issue = elect_one() ? 1u : 0u;
    for (int l = 0; l < NL; ++l) {
      for (int s = 0; s < n_own; ++s, ++it) {
        const int as = it & 1;
        const uint32_t aph = as ? acc_phase1 : acc_phase0;
        mbar_wait(tempty0 + 8 * as, aph ^ 1);
        if (as) acc_phase1 ^= 1; else acc_phase0 ^= 1;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256;
        uint32_t accumulate = 0;
        for (int kb = 0; kb < p.nkb; ++kb) {
          const int nks = min(4, (p.C - kb * 64 + 15) >> 4);
          mbar_wait(afull0 + 8 * ia, pa);
          tc_fence_after();
          const uint32_t a_unit = a_base + ia * p.a_unit_bytes;
          constexpr uint32_t a_hi = ((uint32_t)(BWX * 128) >> 4) | (1u << 14) | (2u << 29);
          constexpr uint32_t b_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
          const uint32_t a_lo0 = ((a_unit & 0x3FFFFu) >> 4) | (1u << 16);
          const uint32_t bs16 = (uint32_t)p.b_stage_bytes >> 4;
          uint32_t b_lo = (((w_base + (uint32_t)(kb * TAPS) * (uint32_t)p.b_stage_bytes) & 0x3FFFFu) >> 4) | (1u << 16);
          if (s == 0) {                                     // first tile of the layer: this K block's weight boxes have landed
            for (int tap = 0; tap < TAPS; ++tap) mbar_wait(wfull0 + 8 * (kb * TAPS + tap), (uint32_t)(l & 1));
            tc_fence_after();
          }
          auto issue_block = [&](auto nk_c) {
            constexpr int NK = decltype(nk_c)::value;
#pragma unroll
            for (int tap = 0; tap < TAPS; ++tap) {
              const uint32_t a_lo = a_lo0 + (uint32_t)((tap / 3) * BWX + (tap % 3)) * 8u;
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
          umma_commit_2sm_if(issue, aempty0 + 8 * ia, (uint16_t)3);
          if (kb == p.nkb - 1) umma_commit_2sm_if(issue, tfull0 + 8 * as, (uint16_t)3);
          if (s == n_own - 1) umma_commit_2sm_if(issue, wfree0 + 8 * kb, (uint16_t)3);     // last tile of the layer: this K block's weights may be replaced
          __syncwarp();
          accumulate = 1;
          if (++ia == p.ra) { ia = 0; pa ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== dependency checker (both CTAs): runs ahead of the producer, never blocked by the ring =====================
    pdl_wait();                                       // this launch's counters were zeroed by the previous launch
    for (int l = 1; l < NL; ++l) {
      const unsigned need = 2u * (unsigned)p.slices * (unsigned)l;   // layer l-1 of the 3 x 3 pair neighbourhood is in memory (both tiles, every slice)
      for (int s = 0; s < n_own; ++s) {
        const int pair = p0 + chain_order(n_own, s);
        const int row = pair / p.pairs_w, pc = pair - row * p.pairs_w, img = row / p.tiles_h, th = row - img * p.tiles_h;
        const int dr = lane / 3 - 1, dc = lane % 3 - 1;
        const bool act = lane < 9 && th + dr >= 0 && th + dr < p.tiles_h && pc + dc >= 0 && pc + dc < p.pairs_w;
        const unsigned* f = p.flags + (act ? (row + dr) * p.pairs_w + pc + dc : pair);
        const long long t0 = clock64();
        while (true) {
          const unsigned v = act ? ld_acquire_gpu(f) : need;
          if (__all_sync(0xffffffffu, v >= need)) break;
          if (clock64() - t0 > (1LL << 32)) __trap();
        }
        if (lane == 0) sts_release_cta(dep_ready, (unsigned)(l * n_own + s + 1));
        __syncwarp();
      }
    }
  } else if (warp == 12) {
    // ===================== publisher (both CTAs): "8 epilogue warps stored (pair, layer l)" -> one release increment of flags[pair] =====================
    pdl_wait();
    unsigned done = 0;
    for (int l = 0; l < NL; ++l)
      for (int s = 0; s < n_own; ++s) {
        done += 8u;
        spin_counter(pub_cnt, done);
        if (lane == 0) red_release_gpu_add(p.flags + p0 + chain_order(n_own, s), 1u);
        __syncwarp();
      }
  } else if (warp == 13) {
    // ===================== residual requester (both CTAs): the box of conv2 tile j+1 is requested as soon as tile j's is in registers =====================
    pdl_wait();
    unsigned consumed = 0;
    const uint32_t row_bytes = (uint32_t)NT * 2u;
    for (int l = 1; l < NL; l += 2) {
      // the residual of layer l is the block input, written by this CTA's epilogue in layer l-2 (l = 1: the chain input)
      spin_counter(pub_cnt, 8u * (unsigned)(n_own * (l - 1)));
      const CUtensorMap* mR = ((l >> 1) & 1) ? &mapR1 : &mapR0;
      for (int s = 0; s < n_own; ++s) {
        spin_counter(res_cnt, consumed);                // the previous box has been moved to registers by all 8 warps
        consumed += 8u;
        int img, h0, w0;
        tile_of(p0 + chain_order(n_own, s), img, h0, w0);
        if (elect_one()) {
          fence_proxy_async();
          mbar_expect_tx(rfull, 128u * row_bytes);
          tma_load_4d(res_s, mR, rfull, cbase, w0, h0, img);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ===================== epilogue (8 warps per CTA, each CTA drains its own 128 accumulator rows) =====================
    pdl_wait();
    const int ew = warp - 4, quarter = warp & 3, half = ew >> 2;
    const int row = quarter * 32 + lane;
    const int nchunks = NT >> 4;
    const uint32_t row_bytes = (uint32_t)NT * 2u;
    const uint32_t tempty_l = mapa_rank(tempty0, 0);
    uint32_t acc_phase0 = 0, acc_phase1 = 0, rphase = 0;
    int it = 0;
    for (int l = 0; l < NL; ++l) {
      const bool conv2 = (l & 1) != 0;
      const int k = l >> 1;
      __nv_bfloat16* outp = conv2 ? (l == NL - 1 ? p.buf[3] : p.buf[(k + 1) & 1]) : p.buf[2];
      const float* bias_l = bias_s + (l & 1) * 128;     // staged by the weight producer before this layer's first weight box
      for (int s = 0; s < n_own; ++s, ++it) {
        const int pair = p0 + chain_order(n_own, s);
        int img, h0, w0;
        tile_of(pair, img, h0, w0);
        const int as = it & 1;
        const int hh = h0 + (row >> 3), ww = w0 + (row & 7);
        const bool pvalid = hh < p.H && ww < p.W;
        const long long pix = ((long long)img * p.H + hh) * p.W + ww;
        const uint32_t aph = as ? acc_phase1 : acc_phase0;
        if (as) acc_phase1 ^= 1; else acc_phase0 ^= 1;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * 256;
        uint4 rr[4][2];
        if (conv2) {
          mbar_wait(rfull, rphase);
          rphase ^= 1;
          const uint32_t rrow = res_s + (uint32_t)row * row_bytes;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int ch = half + 2 * i;
            rr[i][0] = make_uint4(0, 0, 0, 0); rr[i][1] = rr[i][0];
            if (ch < nchunks) {
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rr[i][0].x), "=r"(rr[i][0].y), "=r"(rr[i][0].z), "=r"(rr[i][0].w) : "r"(rrow + ch * 32) : "memory");
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rr[i][1].x), "=r"(rr[i][1].y), "=r"(rr[i][1].z), "=r"(rr[i][1].w) : "r"(rrow + ch * 32 + 16) : "memory");
            }
          }
          // the loads must have RETURNED in every lane before the box is re-armed (a fence orders issue, not completion of
          // bank-conflicted shared loads against the TMA write): the warp-wide OR consumes every lane's registers
          const unsigned z = __reduce_or_sync(0xffffffffu, or_all8(rr) & (uint32_t)p.zero);
          if (lane == 0) reds_release_cta_add(res_cnt, 1u + z);   // this warp holds its pieces: 8 of these re-arm the box
        }
        mbar_wait(tfull0 + 8 * as, aph);
        tc_fence_after();
        {
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
            float v[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 b4 = *reinterpret_cast<const float4*>(bias_l + ch * 16 + 4 * q);
              v[4 * q] = __uint_as_float(work[4 * q]) + b4.x; v[4 * q + 1] = __uint_as_float(work[4 * q + 1]) + b4.y;
              v[4 * q + 2] = __uint_as_float(work[4 * q + 2]) + b4.z; v[4 * q + 3] = __uint_as_float(work[4 * q + 3]) + b4.w;
            }
            if (conv2) {
              const uint4 r0 = i == 0 ? rr[0][0] : (i == 1 ? rr[1][0] : (i == 2 ? rr[2][0] : rr[3][0]));
              const uint4 r1 = i == 0 ? rr[0][1] : (i == 1 ? rr[1][1] : (i == 2 ? rr[2][1] : rr[3][1]));
              float r[16];
              unpack8(r0, r); unpack8(r1, r + 8);
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = r[j] + p.res_scale * v[j];
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (pvalid && cbase + ch * 16 < p.C) {
              stg_v8(outp + pix * p.C + cbase + ch * 16, pack8(v), pack8(v + 8));      // one 32-byte store (STG.256): C % 16 == 0, 32-byte aligned buffers
            }
            ch += 2; ++i;
          }
        }
        tc_fence_before();
        mbar_arrive_cluster(tempty_l + 8 * as);
        // (pair, layer l): this warp's stores are issued; the publisher releases flags[pair] once all 8 warps are here
        __syncwarp();
        if (lane == 0) reds_release_cta_add(pub_cnt, 1u);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFnC)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFnC g_enc = nullptr;
static int g_sms = 0, g_smem = 0, g_dev = -1;
static std::once_flag g_chain_once;

static void chain_init() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
    g_enc = (EncodeTiledFnC)fn;
  cudaGetDevice(&g_dev);
  cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, g_dev);
  cudaDeviceGetAttribute(&g_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, g_dev);
  if (g_smem > 0) cudaFuncSetAttribute(resblock_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem);
}

static bool enc_bf16(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* str, const cuuint32_t* box,
                     CUtensorMapSwizzle swz) {
  cuuint32_t es[4] = {1, 1, 1, 1};
  return g_enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace vmgb

using namespace vmgb;

extern "C" int vmgb_resblock_chain(void* y_inout, void* y_alt, void* t_tmp, void* out, const void* w_layers, int Kp, const float* bias_layers,
                                   void* flags, void* flags_next, int n_img, int H, int W, int C, int n_blocks, float res_scale, void* stream) {
  if (!y_inout || !y_alt || !t_tmp || !out || !w_layers || !bias_layers || !flags || !flags_next || flags == flags_next) return VMGB_EINVAL;
  if (n_img < 1 || H < 8 || W < 8 || n_blocks < 1 || Kp < C || (Kp & 15)) return VMGB_EINVAL;
  if ((C & 15) || C > 160 || (W & 15)) return VMGB_EUNSUPPORTED;       // <= 2 slices of <= 128 columns; pairs must not straddle rows
  const void* ptrs[5] = {y_inout, y_alt, t_tmp, out, w_layers};
  for (const void* q : ptrs) if (reinterpret_cast<uintptr_t>(q) & 15) return VMGB_EINVAL;
  for (const void* q : ptrs) if (reinterpret_cast<uintptr_t>(q) & 31) return VMGB_EUNSUPPORTED;   // 32-byte stores: the caller takes the per-layer path
  std::call_once(g_chain_once, chain_init);
  if (!g_enc || g_sms < 2) return VMGB_EUNSUPPORTED;
  int cur = -1;
  if (cudaGetDevice(&cur) != cudaSuccess || cur != g_dev) return VMGB_EUNSUPPORTED;

  ChainP p;
  p.n_img = n_img; p.H = H; p.W = W; p.C = C;
  p.slices = C <= 128 ? 1 : 2;
  p.NT = ((C + p.slices - 1) / p.slices + 15) / 16 * 16;               // 144 -> 2 x 80
  p.tiles_h = (H + 15) / 16; p.pairs_w = W / 16;
  p.n_layers = 2 * n_blocks; p.nkb = (C + 63) / 64;
  p.a_unit_bytes = (18 * 10 * 128 + 1023) / 1024 * 1024;
  p.b_stage_bytes = (p.NT / 2) * 128;
  if (p.b_stage_bytes & 1023) return VMGB_EUNSUPPORTED;                // SWIZZLE_128B atoms: NT / 2 a multiple of 8
  p.n_pairs = n_img * p.tiles_h * p.pairs_w;
  const int max_cl = (g_sms / 2) / p.slices;
  const int rows = n_img * p.tiles_h;
  const int per = (p.n_pairs + max_cl - 1) / max_cl;
  const int segs = (p.pairs_w + per - 1) / per;
  if (rows * segs <= max_cl) { p.segs = segs; p.n_clusters = rows * segs; }
  else { p.segs = 0; p.n_clusters = p.n_pairs < max_cl ? p.n_pairs : max_cl; }
  p.bias = bias_layers;
  p.buf[0] = (__nv_bfloat16*)y_inout; p.buf[1] = (__nv_bfloat16*)y_alt; p.buf[2] = (__nv_bfloat16*)t_tmp; p.buf[3] = (__nv_bfloat16*)out;
  p.flags = (unsigned*)flags; p.flags_next = (unsigned*)flags_next;
  p.res_scale = res_scale; p.zero = 0;
  if (p.nkb * 9 > CH_MAX_BOXES) return VMGB_EUNSUPPORTED;
  const int wbytes = (p.nkb * 9 * p.b_stage_bytes + 1023) / 1024 * 1024;
  // after the 1024-byte alignment: ring, weights, barriers + counters (16 * ra + 320 <= 512), bias [2][128], residual box
  const int fixed = 1023 + 512 + 1024 + 128 + 128 * p.NT * 2;
  p.ra = (g_smem - wbytes - fixed) / p.a_unit_bytes;
  if (p.ra > 6) p.ra = 6;
  if (p.ra < 2) return VMGB_EUNSUPPORTED;
  const int smem = p.ra * p.a_unit_bytes + wbytes + fixed;

  CUtensorMap mA[3], mB, mR[2];
  for (int i = 0; i < 3; ++i) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_img};
    cuuint64_t str[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, 10, 18, 1};
    if (!enc_bf16(&mA[i], p.buf[i], 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return VMGB_EINVAL;
    if (i < 2) {
      cuuint32_t rbox[4] = {(cuuint32_t)p.NT, 8, 16, 1};
      if (!enc_bf16(&mR[i], p.buf[i], 4, dims, str, rbox, CU_TENSOR_MAP_SWIZZLE_NONE)) return VMGB_EINVAL;
    }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)p.n_layers * 9 * C};
    cuuint64_t str[1] = {(cuuint64_t)Kp * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)(p.NT / 2)};
    if (!enc_bf16(&mB, w_layers, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return VMGB_EINVAL;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * p.n_clusters * p.slices)); cfg.blockDim = dim3(CH_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = as_stream(stream);
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = g_pack_pending ? 1 : 2;
  g_pack_pending = 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, resblock_chain_kernel, mA[0], mA[1], mA[2], mB, mR[0], mR[1], p);
  if (e != cudaSuccess) { g_last_cuda_error = (int)e; return VMGB_ECUDA; }
  return check_launch();
}
