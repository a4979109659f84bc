/*
 * nsr_b200.h -- C ABI of the B200-native NeRF per-ray renderer (libnsr_b200.so).
 *
 * The reference (gyhandy/Neural-Sim-NeRF) is pure Python/PyTorch and has no FFI; these
 * entry points are what a binding for its render hot path would call.  Each one names
 * the reference function it replaces:
 *
 *   RN = optimization/utils/run_nerf_noscale.py     RH = optimization/utils/run_nerf_helpers.py
 *
 * Conventions
 *   - every pointer is a DEVICE pointer (CUDA, current device) unless the name ends in _host;
 *   - tensors are dense row-major fp32 unless stated otherwise;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is
 *     enqueued on it and nothing synchronises the host;
 *   - return value: 0 on success, a negative NSR_E_* code otherwise; nsr_last_error() gives the
 *     message for the calling thread.  Nothing is written to the outputs on a parameter error;
 *   - the caller owns every buffer (inputs, outputs, packed weights, workspace).
 */
#ifndef NSR_B200_H_
#define NSR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NSR_OK 0
#define NSR_E_INVALID (-1)     /* bad argument (null pointer, unsupported size ...) */
#define NSR_E_UNSUPPORTED (-2) /* valid in the reference but not implemented here (e.g. S > 256) */
#define NSR_E_CUDA (-3)        /* a CUDA runtime call failed; see nsr_last_error() */
#define NSR_E_DEVICE (-4)      /* not an sm_100 device */

/* render flags (bit mask) */
#define NSR_FLAG_LINDISP 1u    /* RN:443  sample linearly in inverse depth */
#define NSR_FLAG_WHITE_BKGD 2u /* RN:384  composite onto white */
#define NSR_FLAG_PTS_INPUT 4u  /* nsr_mlp_forward: `z_or_pts` holds explicit points [n,S,3] (RN:26 run_network) */
#define NSR_FLAG_FAST_FP16 8u  /* MLP products as single fp16 x fp16 MMAs (fp32 accumulate) instead of the default
                                  error-compensated hi/lo split (3 MMAs): ~3x the MLP throughput, but misses the
                                  1e-3 parity bar on rays that graze sharp surfaces ("precision") */

#define NSR_FLAG_MIXED_F8 16u  /* MLP products of the first three layers as the fp16 hi/lo split, of the later layers as one fp16
                                  MMA plus the two residual products in e4m3 (kind::f8f6f4, double rate): 2.25 instead of 3
                                  tensor-core passes per product; inside the 1e-3 bar on the fitted scene (3.7e-4) but not on
                                  every network (2.6e-3 on 3x-scaled random ones): opt-in ("precision").
                                  Forward only; NSR_FLAG_FAST_FP16 wins if both are set. */

#define NSR_FLAG_EMBEDDED_INPUT 64u /* nsr_mlp_forward: `z_or_pts` holds already-embedded inputs [n*S, 90] = 63 xyz + 27 view-dir channels
                                      (what NeRF.forward takes, RH:99-122); `rays` is not read and may be NULL */
#define NSR_FLAG_DENSE 32u     /* evaluate EVERY sample point with the default fp16 hi/lo arithmetic.  Without it the render entry points
                                  use the two-tier evaluation ("active set", below) wherever its outputs are bit-identical. */

/* network geometry this library is specialised for (RN:261-278, CFG): D=8, W=256, skips=[4],
 * multires=10 (63 ch), multires_views=4 (27 ch), use_viewdirs=True. */
#define NSR_NET_NUM_TENSORS 12 /* pts_linears.0-7, views_linears.0, feature_linear, alpha_linear, rgb_linear */

int nsr_version(void);
const char* nsr_last_error(void);

/* Number of bytes of one packed network (weights re-laid-out as fp16 hi/lo UMMA operand chunks in the
 * order the kernel streams them, followed by the fp32 biases and the alpha / rgb heads). */
size_t nsr_packed_net_bytes(void);

/*
 * Pack one NeRF MLP (RH:70-97 `class NeRF`, state_dict order below) for the tensor-core kernel.
 *   weights[i], biases[i]: device fp32, PyTorch nn.Linear layout [out, in] / [out]
 *     i = 0..7  pts_linears.i   (256x63, 256x256 x4, 256x319, 256x256 x2)
 *     i = 8     views_linears.0 (128x283)
 *     i = 9     feature_linear  (256x256)
 *     i = 10    alpha_linear    (1x256)
 *     i = 11    rgb_linear      (3x128)
 *   packed_out: device buffer of nsr_packed_net_bytes() bytes, 128-byte aligned.
 * Must be called again after the module's parameters change (the Python layer keys its cache on
 * the tensors' version counters).
 */
int nsr_pack_net(const float* const* weights, const float* const* biases, void* packed_out, void* stream);

/*
 * Positional encoding + MLP for S points on each of n rays -> raw [n,S,4] = (rgb_raw[3], sigma_raw).
 * Replaces RN:26-40 run_network + RH:18-66 Embedder + RH:99-122 NeRF.forward.
 *   rays      [n,11]  o(3) d(3) near far viewdir(3)                        (RN:106-112)
 *   z_or_pts  [n,S]   sample depths, points are o + d*z (RN:463), or, with NSR_FLAG_PTS_INPUT,
 *             [n,S,3] explicit points (then only rays[:,8:11] is read)
 */
int nsr_mlp_forward(const float* rays, const float* z_or_pts, int64_t n_rays, int n_samples, const void* packed_net,
                    uint32_t flags, float* raw_out, void* stream);

/*
 * Alpha compositing.  Replaces RN:343-387 raw2outputs (raw_noise_std = 0; pass pre-noised raw otherwise).
 *   raw [n,S,4], z_vals [n,S], rays_d [n,ld_rays_d] (first 3 columns are the direction; ld = 3 or 11)
 *   outputs (each may be NULL): rgb_map [n,3], disp_map [n], acc_map [n], weights [n,S], depth_map [n]
 * disp_map is NaN where acc_map == 0, like the reference (RN:381).
 */
int nsr_raw2outputs(const float* raw, const float* z_vals, const float* rays_d, int ld_rays_d, int64_t n_rays,
                    int n_samples, uint32_t flags, float* rgb_map, float* disp_map, float* acc_map, float* weights,
                    float* depth_map, void* stream);

/*
 * Inverse-CDF resampling.  Replaces RH:199-243 sample_pdf.
 *   bins [n,B], weights [n,B-1]; u = NULL -> deterministic linspace(0,1,N) (det=True, RH:207-209),
 *   else u [n,N] uniform draws supplied by the caller (det=False, RH:211).  samples_out [n,N].
 */
int nsr_sample_pdf(const float* bins, const float* weights, int64_t n_rays, int n_bins, int n_new, const float* u,
                   float* samples_out, void* stream);

/*
 * Hierarchical step between the coarse and the fine pass.  Replaces RN:473-477 and RN:495:
 *   z_mid = .5(z[1:]+z[:-1]); z_samples = sample_pdf(z_mid, weights[:,1:-1], N_importance);
 *   z_fine = sort(cat[z_coarse, z_samples]); z_std = std(z_samples, unbiased=False)
 *   z_coarse [n,S], weights [n,S], u = NULL or [n,N_importance]
 *   outputs: z_fine [n,S+N_importance], z_samples (may be NULL) [n,N_importance], z_std (may be NULL) [n]
 */
int nsr_resample_merge(const float* z_coarse, const float* weights, int64_t n_rays, int n_samples, int n_importance,
                       const float* u, float* z_fine, float* z_samples, float* z_std, void* stream);

/* Bytes of scratch nsr_render_rays_forward needs for n rays.  Layout (each block padded to 256 bytes), left valid
 * after the call: z0 [n,S] | weights0 [n,S] | raw0 [n,S,4] | z_fine [n,S+Ni] | raw_fine [n,S+Ni,4] | active set of the coarse pass |
 * active set of the fine pass -- the coarse pass's depths and raw outputs are what a backward through rgb0 needs.
 * nsr_render_workspace_layout writes the 7 block offsets and the total (8 values, same order) to offsets_out and returns 8. */
size_t nsr_render_workspace_bytes(int64_t n_rays, int n_samples, int n_importance);
int nsr_render_workspace_layout(int64_t n_rays, int n_samples, int n_importance, size_t* offsets_out, int capacity);

/*
 * The whole per-ray renderer, forward.  Replaces RN:390-501 render_rays (perturb = 0 unless
 * t_rand / u are given, raw_noise_std = 0):
 *   coarse z (RN:439-445) [+ stratified jitter with caller-supplied t_rand [n,S], RN:447-461]
 *   -> encode + coarse MLP -> raw2outputs -> sample_pdf -> sort/merge -> encode + fine MLP -> raw2outputs.
 *   packed_fine = NULL uses the coarse network for the fine pass (RN:481).  n_importance = 0 stops
 *   after the coarse pass (rgb0/disp0/acc0/z_std are then not written).
 * Outputs (NULL = not wanted): rgb_map [n,3], disp_map [n], acc_map [n], rgb0 [n,3], disp0 [n],
 *   acc0 [n], z_std [n], raw [n,S+Ni,4] (retraw), z_vals_out [n,S+Ni] (the fine depths; the backward
 *   pass needs them), weights_out [n,S+Ni].
 */
int nsr_render_rays_forward(const float* rays, int64_t n_rays, const void* packed_coarse, const void* packed_fine,
                            int n_samples, int n_importance, uint32_t flags, const float* t_rand, const float* u,
                            float* rgb_map, float* disp_map, float* acc_map, float* rgb0, float* disp0, float* acc0,
                            float* z_std, float* raw, float* z_vals_out, float* weights_out, void* workspace,
                            size_t workspace_bytes, void* stream);

/*
 * The same with one more output: relu_mask (NULL, or nsr_relu_mask_bytes(n_rays, n_samples + n_importance) bytes, 16-byte aligned)
 * receives one bit per ReLU output of the LAST network pass (2176 bits = 272 B per sample point; nothing else of the
 * activations is kept).  Handing it to nsr_render_rays_backward_ex lets the backward pass skip its forward recompute
 * (10 of its 22 GEMM steps): the pose-gradient route of RN:168-181 for callers that can spare 272 B per point.
 * dump_out (NULL, or nsr_mlp_dump_bytes() bytes, 128-byte aligned; needs relu_mask and the default precision): the last pass also
 * writes every layer's input activations as fp16 (the operand half of the weight-gradient dump), so that a backward call with
 * BOTH relu_mask and dump can produce dL/dW, dL/db without recomputing the forward pass either (what nsr_train_step does).
 *
 * Two-tier evaluation and the ACTIVE SET.  RN:356 computes alpha = 1 - exp(-relu(sigma) dist): a sample point with sigma <= 0 has
 * alpha == 0 and weight == 0 EXACTLY, so neither its colour nor the value of its sigma reaches any output of raw2outputs (RN:343-387)
 * or any gradient.  Unless NSR_FLAG_DENSE (or an opt-in precision flag, or dump_out) is given, each network pass therefore runs as
 *   tier 1  every point, ONE fp16 MMA per product, pts_linears.0-7 + alpha head only -> sigma~; points with sigma~ <= -tau
 *           (default 4.0; |sigma~ - sigma| is ~1e-2 there and <= 0.53 anywhere on the fitted test scene) are certified empty;
 *   tier 2  the remaining points (the "active set", a compacted index list) with the default error-compensated arithmetic: bit for
 *           bit what the dense pass computes for them, since every row of an MMA tile is independent of the others;
 *   verify  tier 2 records max |sigma~ - sigma| over the active points; above verify_max (default 1.5; measured on the fitted scene: 0.74) a third launch re-evaluates
 *           EVERY point densely (it exits immediately otherwise).  A coarse pass whose active fraction exceeds force_fraction
 *           (default 0.30), or that failed its verification, makes the fine pass skip tier 1 and run densely.
 * All decisions are taken on the device (no host synchronisation).  Every map output (rgb/disp/acc/weights/z_std) is bit-identical
 * to the dense evaluation; only `raw` differs: certified-empty points hold (0, 0, 0, sigma~).  Hence `raw` and `relu_mask` (whose
 * tile order follows the active list) are only produced by the two-tier route when the caller passes `active_set`, a buffer of
 * nsr_active_set_bytes(n_rays, n_total_samples) bytes (256-byte aligned) describing the LAST pass: 16 u32 of control words
 * [0] number of active points, [1] float bits of max |sigma~ - sigma|, [2] "dense: tier 1 skipped", [3] "dense: verification
 * failed", padded to 256 B, then int32 point indices (ray * T + sample).  nsr_render_rays_backward_ex takes the same buffer and
 * back-propagates the active points only (dL/draw is exactly 0 at every other point).  active_set = NULL with raw / relu_mask
 * requested: dense evaluation, as before.
 */
size_t nsr_relu_mask_bytes(int64_t n_rays, int n_total_samples);
size_t nsr_active_set_bytes(int64_t n_rays, int n_total_samples);
int nsr_render_rays_forward_ex(const float* rays, int64_t n_rays, const void* packed_coarse, const void* packed_fine,
                               int n_samples, int n_importance, uint32_t flags, const float* t_rand, const float* u,
                               float* rgb_map, float* disp_map, float* acc_map, float* rgb0, float* disp0, float* acc0,
                               float* z_std, float* raw, float* z_vals_out, float* weights_out, void* relu_mask,
                               void* dump_out, void* active_set, void* workspace, size_t workspace_bytes, void* stream);

/* One network pass of the two-tier evaluation by itself (what nsr_render_rays_forward runs per pass), on depth input:
 * stages is a bit mask -- 1: tier 1 (zeroes the control words, writes (0,0,0,sigma~) to raw_out for every point and fills the active
 * list), 2: tier 2 (default arithmetic on the active points, raw_out and optionally relu_mask in active-list tile order),
 * 4: the conditional dense re-evaluation.  7 = the whole pass.  For profiling and tests; raw_out [n,S,4], active_set as above. */
int nsr_mlp_two_tier(const float* rays, const float* z_vals, int64_t n_rays, int n_samples, const void* packed_net, float* raw_out,
                     void* active_set, void* relu_mask, int stages, void* stream);

/* Knobs of the two-tier evaluation (process-wide; change them only while no work is being enqueued).  enabled = 0 makes every
 * pass dense.  Requires tau > verify_max >= 0.  nsr_get_two_tier: any pointer may be NULL. */
int nsr_set_two_tier(int enabled, float tau, float verify_max, float force_fraction);
int nsr_get_two_tier(int* enabled, float* tau, float* verify_max, float* force_fraction);
/* Tier 1 as clusters of two CTAs (tcgen05 ... cta_group::2: one instruction stream drives the tensor cores of both SMs of a TPC, each
 * CTA holding half of every weight chunk).  Process-wide, default off (environment NSR_TIER1_PAIR=1 turns it on at load); returns the
 * previous setting.  Results are bit-identical either way; DESIGN.md 3.1a has the measurements. */
int nsr_set_tier1_pair(int enabled);

/* Coarse-pass refinement.  sample_pdf (RH:199-243) normalises the coarse weights over the ray, so on a ray that barely touches the
 * object (acc0 ~ 1e-3, one sample with sigma ~ 0.02) the tensor-core arithmetic's absolute error on sigma (~1e-4) is a per-cent
 * error of the pdf, the fine samples move, and at a silhouette the pixel leaves the 1e-3 bar (4 of 160 000 rays of the test image).
 * nsr_render_rays_forward therefore re-evaluates the density of those few coarse points with fp64 accumulation on the CUDA cores before the coarse
 * compositing: on every ray whose coarse optical depth is below 2.303 (acc0 < 0.9), every sample with sigma > -0.01; raw[p].sigma
 * is overwritten in place.  On by default whenever N_importance > 0 and the precision is the default one; nsr_set_coarse_refine
 * returns the previous setting.  nsr_coarse_refine is the stage by itself on raw [n,S,4] / z_vals [n,S] (workspace:
 * nsr_coarse_refine_workspace_bytes(n_rays) bytes, 256-byte aligned; its first u32 holds the number of points found, at most
 * 2 n_rays + 1024 are re-evaluated). */
int nsr_set_coarse_refine(int enabled);
int nsr_set_coarse_refine_limit(float acc_limit);   /* rays with coarse acc0 below this (default 0.9) are refined */
int nsr_set_coarse_refine_sigma(float sigma_hi);    /* on every other ray too: the samples with -0.01 < sigma < sigma_hi */
size_t nsr_coarse_refine_workspace_bytes(int64_t n_rays);
int nsr_coarse_refine(const float* rays, const float* z_vals, int64_t n_rays, int n_samples, const void* packed_net, float* raw, void* workspace,
                      size_t workspace_bytes, void* stream);

/* Bytes of scratch nsr_render_rays_backward needs for n rays of T = n_samples + n_importance depths. */
size_t nsr_render_backward_workspace_bytes(int64_t n_rays, int n_total_samples);

/*
 * Backward of render_rays for the pose path.  Replaces the autograd tape behind
 *   dLdray = torch.autograd.grad(rgb_p, batch_rays, grad_outputs=patch_grad_E)            (RN:177-178)
 * at the render_rays boundary: given dL/drgb_map [n,3] it returns dL/d(ray_batch) [n,11] (columns 0:3 origin,
 * 3:6 direction, 8:11 unit view direction; 6:8 near/far are constants of the graph and get 0).  Only the last
 * network pass carries gradient (z_samples is detached, RN:475): `packed_net` is the network that produced
 * `raw` (the fine one, or the coarse one when N_importance == 0 / network_fine is None), `z_vals` [n,T] and
 * `raw` [n,T,4] are the z_vals_out / raw outputs of nsr_render_rays_forward.  Activations are recomputed, not
 * stored.  Flags: NSR_FLAG_WHITE_BKGD as in the forward call.
 *
 * Parameter gradients (the loss.backward() of RN:691-707): pass dW[12] / dB[12] (device fp32 tensors in
 * the nsr_pack_net order and shapes, e.g. dW[5] is [256,319]) and a `dump` scratch buffer of nsr_mlp_dump_bytes() bytes;
 * dL/dW and dL/db of THIS pass are ADDED to them (zero them first; call once per pass that carries gradient -- the fine
 * pass through rgb_map, the coarse pass through rgb0).  dW = dB = dump = NULL skips all of that.
 */
size_t nsr_mlp_dump_bytes(int64_t n_rays, int n_total_samples);
int nsr_render_rays_backward(const float* rays, const float* z_vals, const float* raw, int64_t n_rays, int n_total_samples,
                             const void* packed_net, uint32_t flags, const float* d_rgb_map, float* d_rays, void* dump,
                             float* const* dW, float* const* dB, void* workspace, size_t workspace_bytes, void* stream);

/* The same, reading the ReLU sign bits nsr_render_rays_forward_ex saved for this pass (relu_mask != NULL: no recompute).  With
 * relu_mask AND dW / dB, `dump` must be the buffer that forward call filled through dump_out (activations); this call adds the
 * gradients to it.  relu_mask = NULL: identical to nsr_render_rays_backward (everything recomputed).
 * active_set (NULL, or the buffer that forward call filled): back-propagate the active points only -- same dL/d(rays), since
 * dL/draw == 0 exactly at every certified-empty point; needs dW = dB = dump = NULL. */
int nsr_render_rays_backward_ex(const float* rays, const float* z_vals, const float* raw, int64_t n_rays, int n_total_samples,
                                const void* packed_net, uint32_t flags, const float* d_rgb_map, float* d_rays, void* dump,
                                float* const* dW, float* const* dB, const void* relu_mask, const void* active_set, void* workspace,
                                size_t workspace_bytes, void* stream);

/* The MLP stage of the backward pass by itself (what nsr_render_rays_backward_ex runs between the compositing backward and the per-ray
 * reduction): dL/draw [n,T,4] -> d_pts [n,T,8] = (dL/dpoint[3], 0, dL/dviewdir[3], 0) per sample.  relu_mask / active_set as above
 * (with an active set the other points' rows are zeroed).  For profiling and tests. */
int nsr_mlp_backward(const float* rays, const float* z_vals, int64_t n_rays, int n_total_samples, const void* packed_net, const float* d_raw,
                     float* d_pts, const void* relu_mask, const void* active_set, void* stream);

/*
 * Ray generation + packing.  Replaces RH:156-165 get_rays and RN:91-112 (use_viewdirs, ndc=False):
 *   K_host[9], c2w_host[12] are HOST row-major 3x3 / 3x4; rays_out [H*W,11] device.
 */
int nsr_make_rays(int H, int W, const float* K_host, const float* c2w_host, float near_, float far_, float* rays_out,
                  void* stream);

/* The same from a c2w that lives on the DEVICE (rows ld_c2w >= 4 floats apart: a [3,4] or the top of a [4,4] matrix), e.g. the
 * pose sampler's output (LL:63-72): the pose never visits the host. */
/* RN:91-112 for rays generated by the caller (render(rays=...), RN:163-170): rays_o, rays_d [n,3] device -> rays_out [n,11] =
 * (o, d, near, far, d / |d|), the layout every other entry point takes. */
int nsr_pack_rays(const float* rays_o, const float* rays_d, int64_t n_rays, float near_, float far_, float* rays_out, void* stream);
int nsr_make_rays_dev(int H, int W, const float* K_host, const float* c2w_dev, int ld_c2w, float near_, float far_,
                      float* rays_out, void* stream);

/* RH:14 to8b: out[i] = uint8(255 * clip(x[i], 0, 1)) (fp32 product, truncation), n_values floats -> n_values bytes.  Replaces
 * the host-side conversion in front of imageio.imwrite (RN:200-206, RN:245-250): an image leaves the GPU as H*W*3 bytes. */
int nsr_to8b(const float* x, int64_t n_values, uint8_t* out, void* stream);

/*
 * Pull dL/d(ray_batch) [n,11] back through get_rays (RH:156-165) and the view-direction normalisation (RN:97) to the camera
 * pose: d_c2w [12] (device, row-major 3x4; = or += with accumulate != 0).  Replaces the get_rays part of the tape behind
 *   dLdpsi = torch.autograd.grad(batch_rays, categorical_prob, grad_outputs=dLdray)                     (RN:179-181)
 * -- the caller chains d_c2w through the 8-float pose sampler.  rays / d_rays: [n,11] as produced by nsr_make_rays and
 * nsr_render_rays_backward.  pixel_idx = NULL: the rays are the whole H*W image in row-major order; else int32 [n] pixel index
 * (row*W + column) of each ray.  workspace: nsr_c2w_grad_workspace_bytes() bytes, 8-byte aligned.  Deterministic (no atomics).
 */
size_t nsr_c2w_grad_workspace_bytes(void);
int nsr_rays_grad_to_c2w(int H, int W, const float* K_host, const float* rays, const float* d_rays, const int32_t* pixel_idx,
                         int64_t n_rays, float* d_c2w, int accumulate, void* workspace, void* stream);

/*
 * One image, forward: ray generation (host OR device c2w; exactly one non-NULL) -> nsr_render_rays_forward over all H*W rays ->
 * to8b.  Replaces one iteration of render_path's loop (RN:229-250: render(H, W, K, c2w=...) + to8b).  Outputs (NULL = not
 * wanted): rgb8 [H,W,3] uint8, rgb_map [H*W,3], disp_map, acc_map, rgb0, disp0, acc0, z_std [H*W].  The packed rays [H*W,11]
 * are left at the head of the workspace (256-byte aligned, nsr_render_image_workspace_bytes()).
 */
size_t nsr_render_image_workspace_bytes(int H, int W, int n_samples, int n_importance);
int nsr_render_image_forward(int H, int W, const float* K_host, const float* c2w_host, const float* c2w_dev, int ld_c2w,
                             float near_, float far_, const void* packed_coarse, const void* packed_fine, int n_samples,
                             int n_importance, uint32_t flags, uint8_t* rgb8, float* rgb_map, float* disp_map, float* acc_map,
                             float* rgb0, float* disp0, float* acc0, float* z_std, void* workspace, size_t workspace_bytes,
                             void* stream);

/*
 * One image, forward + backward to the pose: what one iteration of render_path_grad's loop (RN:141-194) computes for ALL rays of the
 * image at once -- rays from the device-resident c2w (RN:148), render (RN:168-170, saving one bit per ReLU), the backward of RN:177-178
 * from d_rgb_map [H*W,3] (= grad_E permuted to HWC, RN:154-155) without recompute, and the get_rays part of RN:179-181 in closed form.
 * Outputs: rgb_map [H*W,3] (may be NULL), d_c2w [12] device (= or += with accumulate != 0).  No autograd tape, one call, no host sync.
 */
size_t nsr_render_image_grad_workspace_bytes(int H, int W, int n_samples, int n_importance);
int nsr_render_image_grad(int H, int W, const float* K_host, const float* c2w_dev, int ld_c2w, float near_, float far_,
                          const void* packed_coarse, const void* packed_fine, int n_samples, int n_importance, uint32_t flags,
                          const float* d_rgb_map, float* rgb_map, float* d_c2w, int accumulate, void* workspace,
                          size_t workspace_bytes, void* stream);

/*
 * Counter-based random numbers (Philox4x32-10): out[i] ~ U[0,1), a pure function of (seed, stream_id, i).  What the training
 * kwargs draw with torch.rand for the stratified jitter (RN:455, t_rand [n,S]) and the inverse-CDF samples (RH:211, u [n,Ni]).
 */
int nsr_random_uniform(uint64_t seed, uint32_t stream_id, float* out, int64_t count, void* stream);

/* raw[p][3] += std * N(0,1) for p < n_points (RN:365-366 raw_noise_std regularisation), N(0,1) from (seed, stream_id, p). */
int nsr_add_sigma_noise(uint64_t seed, uint32_t stream_id, float* raw, int64_t n_points, float std, void* stream);

/*
 * One optimisation step of the NeRF training loop, RN:691-707, entirely on the device:
 *   render(rays) with the training kwargs (perturb: stratified depths + random inverse-CDF draws; raw_noise_std) ->
 *   loss = img2mse(rgb, target) + img2mse(rgb0, target) -> backward to the parameters of both networks (the fine one through
 *   rgb, the coarse one through rgb0; z_samples detached, RN:475) -> Adam (torch.optim.Adam, RN:287: no weight decay) ->
 *   re-pack the operand blobs.  The learning-rate schedule (RN:710-715) stays with the caller: pass this step's `lr`.
 *   nsr_train_net: HOST arrays of 24 DEVICE pointers each -- 12 weights in nsr_pack_net order, then the 12 biases -- for the
 *   parameters and Adam's exp_avg / exp_avg_sq, plus the network's packed blob (updated in place).  fine = NULL: the coarse
 *   network also evaluates the fine pass (RN:481).  step: 1-based Adam step count.  losses_out: device [2] = img_loss, img_loss0.
 *   rgb_out (may be NULL): [n,3].  Random tensors are functions of (seed, element index): streams 0 t_rand, 1 u, 2/3 sigma noise.
 */
typedef struct nsr_train_net {
  float* const* params;
  float* const* exp_avg;
  float* const* exp_avg_sq;
  void* packed;
} nsr_train_net;
size_t nsr_train_workspace_bytes(int64_t n_rays, int n_samples, int n_importance);
int nsr_train_step(const float* rays, const float* target, int64_t n_rays, const nsr_train_net* coarse, const nsr_train_net* fine,
                   int n_samples, int n_importance, uint32_t flags, int perturb, float raw_noise_std, uint64_t seed, float lr,
                   float beta1, float beta2, float eps, int64_t step, float* losses_out, float* rgb_out, void* workspace,
                   size_t workspace_bytes, void* stream);

/* Introspection (host only, no GPU): the order in which the MLP kernels consume the weight chunks of forward GEMM step `step`
 * (0..7 pts_linears, 8 feature_linear, 9 views_linears): slot i -> (accumulator half, K chunk).  Returns the number of slots
 * (halves x K chunks), or a negative error code.  Two-half steps run (h0, K early) (h1, K early) (h0, K late) (h1, K late), see
 * DESIGN.md "Pipeline". */
int nsr_chunk_issue_order(int step, int* half_out, int* kc_out, int capacity);

/* Number of kernels this library has launched since load (all threads); used by bench.py's gpu_launches. */
uint64_t nsr_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* NSR_B200_H_ */
