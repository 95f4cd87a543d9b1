/* vmgb.h -- C ABI of libvmgb.so, the sm_100a kernel library under the VMG model-forward hot path.
 *
 * The reference (EasyVision-Ton/VMG) is pure Python/PyTorch: it has NO FFI for this path; its seam is
 * the nn.Module contract of models/vmg.py:176-210,585-637 (kept by vmg_b200/model.py).  This header
 * is the boundary BELOW that module: one entry point per fused kernel family.  Each prototype
 * cites the reference call site(s) it replaces.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every function returns 0 on success or a negative VMGB_E* code; nothing throws, nothing
 *     allocates, nothing synchronises; `stream` is a cudaStream_t passed as void*.
 *   - all data pointers are caller-owned DEVICE pointers; activations are channels-last
 *     ("NHWC": [images, H, W, C], C contiguous).  `*_dt` arguments are VMGB_F32 / VMGB_BF16.
 *   - `pitch` arguments are the distance in ELEMENTS between consecutive pixels of a tensor
 *     (>= its channel count), so kernels can read/write channel slices of wider buffers.
 *   - arithmetic is fp32 (CUDA-core kernels) or bf16 x bf16 -> fp32 (tcgen05 kernels); statistics,
 *     softmaxes and coordinate math are always fp32.
 *
 * Prototype grammar is deliberately trivial (one line per argument type: const void*, void*, int,
 * long long, float): vmg_b200/ops.py parses THIS FILE to build its ctypes signatures, and
 * tests/test_abi.py checks that every symbol declared here is exported.
 */
#ifndef VMGB_H_
#define VMGB_H_

#ifdef __cplusplus
extern "C" {
#endif

#define VMGB_OK            0
#define VMGB_EINVAL       -1   /* bad shape / unsupported argument combination */
#define VMGB_EUNSUPPORTED -2   /* valid request that this build has no kernel for */
#define VMGB_ECUDA        -3   /* CUDA runtime error; see vmgb_last_cuda_error() */
#define VMGB_EARCH        -4   /* device is not sm_100 */

#define VMGB_F32  0
#define VMGB_BF16 1

/* activation codes for the conv / GEMM epilogue */
#define VMGB_ACT_NONE    0
#define VMGB_ACT_RELU    1
#define VMGB_ACT_LRELU   2   /* slope = act_alpha */
#define VMGB_ACT_GELU    3   /* exact erf form (torch.nn.GELU default) */
#define VMGB_ACT_SIGMOID 4

/* epilogue modes:  v = act(acc + bias) * out_scale
 *   STORE : out = v
 *   RESADD: out = res + res_scale * v
 *   GATE  : out = res + res_scale * ((aux + v) * tanh(v))       (function.py:796-802 + :1215) */
#define VMGB_EPI_STORE  0
#define VMGB_EPI_RESADD 1
#define VMGB_EPI_GATE   2

/* whose fp32 rounding the flow_warp coordinate normalisation `2*g / max(size-1, 1) - 1` (trajectory.py:101-106)
 * follows: ATen's CPU div kernel divides; ATen's CUDA div kernel multiplies by the fp32 reciprocal of a Python-scalar
 * divisor.  The two differ by one ulp now and then, which matters only to the NEAREST warp of the tracked locations
 * (a coordinate that lands on an exact .5 tie).  The model defaults to the CUDA order (what the reference's own
 * GPU forward computes); tests against CPU-minted goldens select the CPU order. */
#define VMGB_COORD_ATEN_CPU  0
#define VMGB_COORD_ATEN_CUDA 1

int         vmgb_version(void);
const char* vmgb_arch(void);                 /* "sm_100a" */
int         vmgb_last_cuda_error(void);      /* last cudaError_t seen by this library */
const char* vmgb_last_cuda_error_string(void);
int         vmgb_device_check(void);         /* VMGB_OK iff current device is compute capability 10.x */
long long   vmgb_launch_count(void);         /* kernels launched by this library since load */

/* ---- conv2d / linear, CUDA-core fp32 path -------------------------------------------------
 * Replaces every nn.Conv2d / nn.Linear call on the path: F.conv2d stride 1,
 * "same" zero padding, kernel R x S in {1,3,7}, `groups`; a Linear is R=S=1 over H=1,W=rows.
 * x [n_img,H,W,(x_pitch)] ; w fp32 packed [groups][R*S][Cin/groups][Cout/groups] ; bias fp32[Cout] or NULL.
 * pixel_shuffle=2 stores as nn.PixelShuffle(2) would (vmg.py:629-630): out is [n_img,2H,2W,Cout/4].
 * res / aux (may be NULL) are addressed like `out`.
 * "virtual concat": n_src in {1,2,3} equally wide sources x, x1, x2 (same dtype and pitch) are read as
 * torch.cat([x, x1, x2], channel) would be, without materialising it (trajectory.py:372,457,476);
 * Cin is the TOTAL; requires groups == 1 and (Cin/n_src) % 16 == 0 when n_src > 1. */
int vmgb_conv2d_nhwc(const void* x, const void* x1, const void* x2, int n_src, int x_dt, int x_pitch,
                     const void* w, const void* bias,
                     void* out, int out_dt, int out_pitch,
                     const void* res, int res_dt, int res_pitch,
                     const void* aux, int aux_dt, int aux_pitch,
                     int n_img, int H, int W, int Cin, int Cout, int R, int S, int groups,
                     int act, float act_alpha, float out_scale, int epi, float res_scale,
                     int pixel_shuffle, void* stream);

/* ---- conv2d / linear, tcgen05 + TMEM + TMA path (bf16 operands, fp32 accumulate) -----------
 * Same contract as vmgb_conv2d_nhwc; x must be bf16 with x_pitch % 8 == 0; w is bf16 packed
 * [groups][R*S][Cout/groups][Kp], Kp = roundup(Cin/groups, 16) zero padded (see vmgb_pack_weight_tc).
 * variant: 0 = automatic kernel choice (what the model uses); the others force one kernel for tests and
 * tools/microbench.py and return VMGB_EUNSUPPORTED (-2) when that kernel does not take the shape:
 *   1 = per-tap ring kernel; 8 / 9 = single-CTA weight-stationary kernel (column-shifted / wide activation units,
 *   R = 3 or 7); 13 = weight-stationary kernel on a CTA pair (cta_group::2; R = 3 or 7, up to 3 sources; or a Linear
 *   (R = 1, H = 1, W = rows % 8 == 0, one source, plain store / residual add) read as an [rows/8, 8] image);
 *   13 + 32*d = the same with bottleneck-probe flags d (tools only, results are NOT valid). */
int vmgb_conv2d_nhwc_tc(const void* x, const void* x1, const void* x2, int n_src, int x_pitch,
                        const void* w, int Kp, const void* bias,
                        void* out, int out_dt, int out_pitch,
                        const void* res, int res_dt, int res_pitch,
                        const void* aux, int aux_dt, int aux_pitch,
                        int n_img, int H, int W, int Cin, int Cout, int R, int S, int groups,
                        int act, float act_alpha, float out_scale, int epi, float res_scale,
                        int pixel_shuffle, int variant, void* stream);

/* ---- resblock_chain: the residual-block chain of the trajectory module as ONE persistent kernel -------------------
 * trajectory.py:16-52 (ResidualBlocksWithInputConv.main[2]) / :165-221 (ResidualBlockNoBN0):
 *     for each of n_blocks blocks:  y <- y + res_scale * conv2(relu(conv1(y)))        (3x3, C -> C, zero padding)
 * i.e. 2*n_blocks dependent convolutions per call (30 per trajectory step).  bf16 channels-last tensors [n_img,H,W,C]
 * with pitch C.  y_inout holds the chain input (the output of the input conv) and is OVERWRITTEN; y_alt and t_tmp are
 * scratch tensors of the same size; the result lands in `out`.
 * w_layers: bf16 [2*n_blocks][9][C][Kp] -- the vmgb_pack_weight_tc layouts of conv1, conv2 of block 0, 1, ... back to
 * back; bias_layers: fp32 [2*n_blocks][C].
 * flags / flags_next: two DIFFERENT arrays of n_img * ceil(H/16) * (W/16) 32-bit counters.  `flags` must be all zero on
 * entry; the kernel leaves it dirty and zeroes `flags_next`, so consecutive calls alternate the two arrays.
 * Cross-tile (halo) dependencies between layers are tracked through these counters instead of grid barriers.
 * Takes C % 16 == 0 with C <= 128 (one output-channel slice) or C in {144, 160} (two slices of 80 columns computed by
 * different clusters from the same activations), W % 16 == 0; VMGB_EUNSUPPORTED otherwise (the caller issues
 * vmgb_conv2d_nhwc_tc per layer). */
int vmgb_resblock_chain(void* y_inout, void* y_alt, void* t_tmp, void* out, const void* w_layers, int Kp, const float* bias_layers,
                        void* flags, void* flags_next, int n_img, int H, int W, int C, int n_blocks, float res_scale, void* stream);

/* ---- tab_mix: the Morph-FC mixing half of a TAB as two launches of one persistent tcgen05 kernel ---------------------
 * function.py:743-805 (Enhanced_MorphFCs_decay.forward after the in-place decay) inside TAB.forward (function.py:1212-1215):
 *     xn = LayerNorm(x) (norm2, eps) ; h = ReLU(MorphFC_H(xn)) / C ; w = ReLU(MorphFC_W(xn)) / C ; c = ReLU(Linear_C(xn)) / C
 *     a = softmax_3(reweight(mean_{t,h,w}(h + w + c))) ; z = proj(a0.h + a1.w + a2.c) ; out = x + m_scaling * (xn + z) * tanh(z)
 * x / out: bf16 channels-last, time-major [n_img = T*B, H, W, C] (image i belongs to batch sample i % B).
 * w_h, w_w, w_c, w_p: vmgb_pack_weight_tc layouts ([C][Kp] bf16, K-major) of mlp_h[0] (already decayed), mlp_w[0], mlp_c[0], proj;
 * b_*: fp32 [C] or NULL.
 * phase 0: accumulates sum over the valid pixels of (h + w + c) per (batch sample, channel) into acc [B][C], 64-bit 2^-24
 *          fixed point (zero it first; vmgb_fixsum_finalize turns it into the fp32 mean); w_p, a, out unused.
 * phase 1: a = fp32 [B][3][C] branch weights; writes out (must not alias x).
 * Takes chunk_h = chunk_w = chunk = 8 and C in {64, 112} (the level-0 stages of VMG-REDS); VMGB_EUNSUPPORTED otherwise
 * (the caller issues layernorm / morph_permute / linears / channel_sum / mix3 / gated projection separately). */
int vmgb_tab_mix(const void* x, const float* ln_g, const float* ln_b, float eps, const void* w_h, const void* w_w, const void* w_c,
                 const void* w_p, int Kp, const float* b_h, const float* b_w, const float* b_c, const float* b_p, int phase,
                 void* acc, const float* a, void* out, float m_scaling, int n_img, int B, int H, int W, int C, int chunk, void* stream);
/* out[i] = acc[i] * 2^-24 * scale  (acc: the signed 64-bit fixed-point sums of vmgb_tab_mix phase 0 / vmgb_channel_sum) */
int vmgb_fixsum_finalize(const void* acc, float* out, int n, float scale, void* stream);

/* OIHW fp32 conv weight [Cout][Cin/groups][R][S] (or Linear [Cout][Cin], R=S=1) -> packed layouts.
 * tc layout: input channel ci of group g is stored at k = ((g*Cin/groups) % 8) + ci, so
 * Kp >= Cin/groups + max_g((g*Cin/groups) % 8), Kp % 16 == 0. */
int vmgb_pack_weight(const void* w_oihw, void* w_packed, int Cout, int Cin_g, int R, int S, int groups, void* stream);
int vmgb_pack_weight_tc(const void* w_oihw, void* w_packed_bf16, int Cout, int Cin_g, int Kp, int R, int S, int groups, void* stream);

/* ---- LayerNorm over C (eps 1e-5, biased variance): function.py:1164,1195 ; layers.py:768-775 */
int vmgb_layernorm(const void* x, int x_dt, void* out, int out_dt, const void* gamma, const void* beta,
                   long long rows, int C, float eps, void* stream);

/* ---- Morph-FC chunk permutation (function.py:763-785, closed form SURVEY.md A.2) -------------
 * axis 0 = H branch, 1 = W branch.  inverse=0: x[BT,H,W,C] -> A[rows, Cp] (zero padded rows/channels);
 * inverse=1: Y[rows, Cp] -> y[BT,H,W,C] (cropped).  Cp = ceil(C/chunk)*chunk,
 * rows = BT * ceil(L/chunk)*chunk * other, L = H (axis 0) or W (axis 1); matrix rows are a_pitch (>= Cp) elements apart. */
int vmgb_morph_permute(const void* src, int src_dt, void* dst, int dst_dt,
                       int BT, int H, int W, int C, int chunk, int axis, int inverse, int a_pitch, void* stream);

/* ---- grouped channel sums (fp32 out; acc = groups*C 64-bit integers of scratch, ZEROED by the caller: the
 * partial sums are accumulated in 2^-24 fixed point so that the result is independent of the order of the
 * atomics, i.e. bit-identical from run to run).  function.py:791 (TAB global
 * mean over T,H,W per sample), :546-558 (CALayer per-frame pool).  Rows are laid out as
 * [n_seg][groups][seg_rows]: out[g][c] = scale * sum_{s,r} (a+b+c)[(s*groups + g)*seg_rows + r][c].
 * (frames are stored time-major [T,B,...]: per-sample sums use n_seg=T, groups=B, seg_rows=H*W.) */
int vmgb_channel_sum(const void* a, const void* b, const void* c, int dt, float* out, long long* acc,
                     int n_seg, int groups, long long seg_rows, int C, float scale, void* stream);

/* softmax over the 3 branch weights: logits fp32 [B][3C] with feature index 3*c+j (function.py:792)
 * -> a fp32 [B][3][C] */
int vmgb_softmax3(const float* logits, float* a, int B, int C, void* stream);

/* mix = h*a0 + w*a1 + c*a2 (function.py:793); a fp32 [B][3][C]; rows [n_seg][B][seg_rows] as above. */
int vmgb_mix3(const void* h, const void* w, const void* c, int dt, const float* a, void* out, int out_dt,
              int n_seg, int B, long long seg_rows, int C, void* stream);

/* RCAB tail (function.py:556-558,578-581 and scale_c): out = (u * y[frame][c] + x) * scale */
int vmgb_ca_apply(const void* u, const void* x, int dt, const float* y, void* out, int out_dt,
                  int frames, long long rows_per_frame, int C, float scale, void* stream);

/* lockstep sweep pairing of Trajectory_multi_head (trajectory.py:404-478: the backward and the forward sweep share
 * their weights and only meet in the final fusion, so step j = [backward sweep at frame T-1-j | forward sweep at
 * frame j] runs as one image batch).  x: T frames of frame_bytes; out: T steps of 2 frames.
 *   pair  : out[j][0] = x[T-1-j], out[j][1] = x[j]
 *   unpair: back[i] = paired[T-1-i][0] (= outputs_back[i]), fwd[i] = paired[i][1]
 * frame_bytes and all pointers must be multiples of 16. */
int vmgb_pair_sweeps(const void* x, void* out, int T, long long frame_bytes, void* stream);
int vmgb_unpair_sweeps(const void* paired, void* back, void* fwd, int T, long long frame_bytes, void* stream);

/* elementwise out = a + b */
int vmgb_add(const void* a, const void* b, int dt, void* out, int out_dt, long long n, void* stream);
/* out [n_img,H,W,C] = (a + b)[:, :H, :W, :] for bf16 channels-last a, b [n_img,Hp,Wp,C], C % 8 == 0: the last skip connection
 * (vmg.py:547 / :581) and the crop back to the unpadded size (vmg.py:619) in one pass.  VMGB_EUNSUPPORTED for other layouts. */
int vmgb_add_crop(const void* a, const void* b, void* out, int n_img, int Hp, int Wp, int H, int W, int C, void* stream);

/* W *= gamma in place, fp32 (function.py:765-768,778-781) */
int vmgb_decay_weights(float* w, const float* gamma, long long n, void* stream);

/* ---- flow warps (trajectory.py:71-116 + ATen grid_sampler_2d, align_corners=True, border) ----
 * feature warp: bilinear; x,out [n,h,w,C]; flow fp32 [n,h,w,2] (dx,dy).
 * location warp: nearest; loc fp32 [n,h,w,2*K] in, [n,h,w,(out_pitch)] out, `K2` = 2*K channels moved. */
int vmgb_flow_warp_bilinear(const void* x, int x_dt, int x_pitch, const float* flow, void* out, int out_dt, int out_pitch,
                            int n, int h, int w, int C, int semantics, void* stream);
int vmgb_flow_warp_nearest(const float* loc, int in_pitch, const float* flow, float* out, int out_pitch,
                           int n, int h, int w, int K2, int semantics, void* stream);
/* flow_warp_fused: both warps of one trajectory step (trajectory.py:330-333 / 410-430: feat_prop bilinear, all 2K
 * location channels nearest, same flow) in ONE launch; bf16 features with C % 8 == 0 and 16-byte aligned rows,
 * K2 even; returns VMGB_EUNSUPPORTED otherwise (the caller then issues the two launches above). */
int vmgb_flow_warp_fused(const void* x, int x_pitch, const float* flow, void* out, int out_pitch,
                         const float* loc, int loc_in_pitch, float* loc_out, int loc_out_pitch, int K2,
                         int n, int h, int w, int C, int semantics, void* stream);
/* writes the identity grid (x, y) into channels [c0, c0+2) of loc [n,h,w,pitch] (trajectory.py:316-317,365) */
int vmgb_loc_identity(float* loc, int pitch, int c0, int n, int h, int w, void* stream);

/* flow smoothing (function.py:1466-1478): reflect-pad to a multiple of r, r x r mean, nearest x r, crop.
 * in / out: fp32 channels-last [n,h,w,2] (out must not alias in). */
int vmgb_flow_smooth(const float* flow_in, float* out_nhwc, int n, int h, int w, int r, void* stream);

/* L2-normalise each pixel over C (F.normalize eps 1e-12): trajectory.py:744-749 */
int vmgb_l2norm(const void* x, int x_dt, int x_pitch, void* out, int out_dt, long long rows, int C, void* stream);

/* ---- LTAM 'wins' attention (trajectory.py:672-795,534-547) ---------------------
 * q      [n,h,w,C]            L2-normalised current features
 * keys   [Kcap][n,h,w,C]      L2-normalised key-frame stage inputs   (first K used)
 * vals   [Kcap][n,h,w,C]      key-frame propagated features
 * loc    fp32 [n,h,w,loc_pitch] tracked (x,y) per key-frame, channels 2k,2k+1
 * rel    fp32 [heads][4][4], decay fp32 [heads]
 * out    [n,h,w,C]  attention output BEFORE the proj Linear (proj + anchor add = conv2d RESADD), or -- with
 *        out_bias / anchor given -- out = attention(vals) + out_bias[c] + anchor[pixel][c]: the caller stores the values
 *        ALREADY PROJECTED (vals_k = proj.weight . sparse_k, one launch per key frame instead of one per step; exact
 *        in real arithmetic because proj is linear and the softmax weights sum to 1: trajectory.py:791-795), out_bias =
 *        proj.bias (fp32 [C], 16-byte aligned) and anchor = the warped propagated feature (dtype out_dt, pitch C). */
int vmgb_ltam_wins(const void* q, const void* keys, const void* vals, int dt, long long kv_stride,
                   const float* loc, int loc_pitch, const float* rel, const float* decay,
                   void* out, int out_dt, int n, int h, int w, int C, int heads, int K,
                   const float* out_bias, const void* anchor, void* stream);

/* ---- resampling (layers.py:777-798): pixel-(un)shuffle gather + LayerNorm fused; Linear follows.
 * down: x[BT,H,W,C] -> out[BT,H/2,W/2,4C] feature (neiw,neih,c);  up: x[BT,H,W,C] -> out[BT,2H,2W,C/4].
 * `add` (may be NULL) is summed to x first (skip adds of vmg.py:530-543).  out rows are out_pitch elements apart
 * (>= feature count; a multiple of 8 lets the following Linear read them through TMA). */
int vmgb_resample_ln(const void* x, const void* add, int dt, void* out, int out_dt, int out_pitch, const void* gamma, const void* beta,
                     int BT, int H, int W, int C, int up, float eps, void* stream);

/* ---- mdsc shortcut pieces (vmg.py:389-400,519,525) */
int vmgb_maxpool4(const void* x, int dt, void* out, int out_dt, int BT, int H, int W, int C, void* stream);
/* per-frame GroupNorm(1,C)+ReLU given per-frame sums: stats = 64-bit integers [frames][2] = (sum, sum of squares) in
 * 2^-20 fixed point (order-independent atomics), ZEROED by the caller before vmgb_frame_stats */
int vmgb_frame_stats(const void* x, int dt, long long* stats, int frames, long long elems_per_frame, void* stream);
int vmgb_groupnorm1_relu(const void* x, int dt, const long long* stats, const void* gamma, const void* beta, void* out, int out_dt,
                         int frames, long long rows_per_frame, int C, float eps, void* stream);

/* ---- SPyNet helpers (vmg.py:39-123) : all fp32 channels-last ----------------------------------
 * prep: padded frames NHWC [n,H,W,3] -> bilinear resize (align_corners=False) to [n,h_up,w_up,3], then
 *       (v - mean)/std  (vmg.py:104-110 then :43-44); every frame is prepared once and shared by all pairs */
int vmgb_spynet_prep(const float* frames, float* out, float m0, float m1, float m2, float s0, float s1, float s2,
                     int n, int H, int W, int h_up, int w_up, void* stream);
/* avg_pool2d(2,2) pyramid step (vmg.py:47-59) */
int vmgb_avgpool2(const float* x, float* out, int n, int H, int W, int C, void* stream);
/* one pyramid level (vmg.py:65-83): pyr [T*B,h,w,3] time-major; pair p = t*B+b: dir 0 -> ref = frame p+B,
 * supp = frame p (forward flow, vmg.py:458), dir 1 -> ref = p, supp = p+B (backward flow, :460).
 * flow_up = 2*bilinear_x2(flow_coarse [pairs,h/2,w/2,2], align_corners=True) (zeros when NULL);
 * out8 [pairs,h,w,8] = cat[ref, warp_border(supp, flow_up), flow_up]; flow_up [pairs,h,w,2] also stored. */
int vmgb_spynet_level_input(const float* pyr, const float* flow_coarse, void* out8, int out8_dt, float* flow_up,
                            int n_pairs, int B, int dir, int h, int w, int semantics, void* stream);
/* final: bilinear resize (align_corners=False) of flow [n,h_up,w_up,2] to [n,h,w,2] and rescale (vmg.py:113-121) */
int vmgb_flow_resize(const float* flow, float* out, int n, int h_up, int w_up, int h, int w, void* stream);

/* ---- input / output layout kernels (vmg.py:590-599,635-636) ------------------------------------
 * Inside the library frames are stored TIME-MAJOR [T,B,...] so that one recurrent step of the
 * trajectory sweep (all clips of the batch at frame t) is one contiguous image batch.
 * pad_input   : x NCHW fp32 [B,T,3,H,W] -> NHWC fp32 [T,B,Hp,Wp,3], replicate padded (vmg.py:596-599)
 * mirror_check: *sumsq += sum |x[:, :T/2] - flip(x[:, T/2:])|^2   (vmg.py:426-432; mirror iff == 0)
 * final_output: out NCHW fp32 [B, t0..t0+Tc, 3,4H,4W] = y NHWC [Tc,B,4H,4W,3 (pitch)] + bilinear_x4(x)
 *               (trilinear with unit depth scale == per-frame bilinear, align_corners=False; vmg.py:593,636) */
int vmgb_pad_input(const float* x, float* out, int B, int T, int H, int W, int Hp, int Wp, void* stream);
int vmgb_mirror_check(const float* x, float* sumsq, int B, int T, long long frame_elems, void* stream);
int vmgb_final_output(const void* y, int y_dt, int y_pitch, const float* x, float* out, int B, int T, int t0, int Tc,
                      int H, int W, void* stream);

/* ==== the callers either side of the forward: frame I/O, Tester tiling / blending, metrics, ensemble ====
 * Prototype grammar note: `double*` / `const double*` are fp64 device pointers unless stated otherwise. */

/* ---- frame I/O (utils/eval_utils.py:8-36 read_image/read_seq_images; tools/Tester.py:217-222,249-250) ----------
 * u8 -> f32: frames uint8 [T,H,W,3] -> fp32 [T,3,H,W] = u8 / 255 (numpy `astype(float32) / 255.`), bit-exact.
 * f32 -> u8: clamp(0,1) -> * 255.0 -> round half to even -> uint8 [T,H,W,3] (Tester.evaluate), bit-exact.
 * reverse_channels = 1 flips the channel order (cv2 BGR <-> RGB, eval_utils.py:19; cv2.imwrite(output[..., ::-1])). */
int vmgb_frames_u8_to_f32(const void* frames_u8, float* out, int T, int H, int W, int reverse_channels, void* stream);
int vmgb_frames_f32_to_u8(const float* frames, void* out_u8, int T, int H, int W, int reverse_channels, void* stream);

/* ---- overlapped tile / window blending (tools/Tester.py:107-143 test_image, :146-177 test_clips) ----------------
 * accumulate: E[b, t0+t, c, h0+y, w0+x] += patch[b,t,c,y,x] and count[t0+t, h0+y, w0+x] += 1 for the patch samples the
 *             reference does not zero: frames [zt0, Tp-zt1), rows [zh0, ph-zh1), columns [zw0, pw-zw1).
 *             patch fp32 [B,Tp,C,ph,pw]; E fp32 [B,Te,C,He,We]; count fp32 [Te,He,We] (may be NULL).
 *             Units must be accumulated in the reference's loop order for bit-identical sums (stream order does that).
 * normalize : E /= count (fp32 division, broadcast over batch and channels)   -- `E.div_(W)`
 * finalize  : uint8 [T,H,W,3] = round(clamp(E / count, 0, 1) * 255) for B = 1, C = 3 (count may be NULL: no division) */
int vmgb_blend_accumulate(const float* patch, float* E, float* count, int B, int Tp, int C, int ph, int pw, int Te, int He, int We,
                          int t0, int h0, int w0, int zt0, int zt1, int zh0, int zh1, int zw0, int zw1, void* stream);
int vmgb_blend_normalize(float* E, const float* count, int B, int Te, int C, int He, int We, void* stream);
int vmgb_blend_finalize_u8(const float* E, const float* count, void* out_u8, int T, int H, int W, int reverse_channels, void* stream);

/* ---- augmentation of the self-ensemble route (tools/Tester.py:387-445 augment / augment_inverse) ---------------
 * per [H,W] plane: hflip (reverse W), then vflip (reverse H), then transpose (H <-> W); dst is [planes, W, H] when
 * transpose is set.  augment_inverse applies the SAME sequence (the reference does; it is an involution for
 * hflip == vflip or transpose == 0). */
int vmgb_augment(const float* src, float* dst, long long planes, int H, int W, int hflip, int vflip, int transpose, void* stream);

/* ---- metrics (utils/metrics.py:11-70; tools/test_reds4.py:194-218) ---------------------------------------------
 * sse_u8 : *out_sum (64-bit integer, ZEROED by the caller) += sum (a-b)^2 over n uint8 samples; exact, so
 *          calculate_psnr = 20 log10(255 / sqrt(sse / n)) is bit-reproducible on the host.
 * sse_f32 / sse_f64: fp64 sum of squared differences (deterministic two-stage reduction; scratch from
 *          vmgb_metrics_scratch_doubles); clamp01 clamps both fp32 inputs to [0,1] first (Tester.test_clips_max).
 * rgb_to_y: float64 Y of skimage.color.rgb2ycbcr on uint8 RGB: 16 + 65.481 r + 128.553 g + 24.966 b, r,g,b = u8 * (1/255).
 * ssim   : structural_similarity with the 11x11 Gaussian window (window11 = HOST pointer to the 11 weights of
 *          cv2.getGaussianKernel(11, 1.5)), valid region, float64; dt 0 = n_img x uint8 [H,W,n_planes] (mean over the
 *          planes, metrics.py:46-50), dt 1 = n_img x float64 [H,W] (n_planes = 1).  out_per_img: fp64 [n_img]. */
int vmgb_sse_u8(const void* a, const void* b, long long n, long long* out_sum, void* stream);
long long vmgb_metrics_scratch_doubles(int n_img, int H, int W, int n_planes);
int vmgb_sse_f32(const float* a, const float* b, long long n, int clamp01, double* scratch, double* out_sum, void* stream);
int vmgb_sse_f64(const double* a, const double* b, long long n, double* scratch, double* out_sum, void* stream);
int vmgb_rgb_to_y(const void* rgb_u8, double* y, long long pixels, void* stream);
int vmgb_ssim(const void* a, const void* b, int dt, int n_img, int H, int W, int n_planes, const double* window11, double* scratch,
              double* out_per_img, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VMGB_H_ */
