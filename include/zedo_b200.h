/* zedo_b200.h -- C ABI of the B200-native ZeDO per-pose optimisation loop.
 *
 * Drop-in boundary for the hot path of ipl-uw/ZeDO-Release.  The
 * reference is pure Python over PyTorch, so its "FFI" is a ctypes binding: every entry
 * point below replaces one reference function (cited as file:line relative to the
 * reference root) and is what the reference-side stub in INTEGRATION.md binds.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch / C++ types in any signature.
 *   - All array pointers are DEVICE pointers to contiguous float32 data in the reference's
 *     own layouts unless a parameter is marked "host".
 *   - All work is enqueued on the caller's stream (cudaStream_t passed as void*; NULL =
 *     legacy default stream); no hidden synchronisation except where noted.
 *   - No device allocation after set-up: zedo_plan_create sizes the workspaces for max_batch and
 *     zedo_plan_reserve the per-step bias tables for the longest schedule; only a call that exceeds what
 *     was reserved grows a buffer (stream-ordered, with one stream synchronisation).  Index lists
 *     (keylist, joint_subset) travel in the kernel parameters.
 *   - Return value: 0 = OK, <0 = ZEDO_E_* argument error, >0 = cudaError_t.  Never throws.
 *   - A plan is bound to one device and one stream at a time and is not thread-safe
 *     (one plan per process/GPU, matching one-process-per-GPU sharding).
 *   - There is no CPU fallback: every compute entry point fails with a CUDA error when no
 *     sm_100 device is present.
 */
#ifndef ZEDO_B200_H
#define ZEDO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZEDO_B200_ABI_VERSION 2

/* argument errors */
#define ZEDO_E_INVALID   (-1)  /* NULL pointer / bad enum */
#define ZEDO_E_SHAPE     (-2)  /* unsupported shape (J, hidden, batch > plan capacity ...) */
#define ZEDO_E_MISSING   (-3)  /* a state_dict tensor the network needs was not supplied */
#define ZEDO_E_NOMEM     (-4)  /* host allocation failed */
#define ZEDO_E_STATE     (-5)  /* call order violated (e.g. step before set_schedule) */

/* GEMM arithmetic of the score network (all accumulate in float32) */
#define ZEDO_GEMM_SPLIT3 0  /* tcgen05, fp16 hi/lo 3-product split (2^-22 product error)        */
#define ZEDO_GEMM_FP16   1  /* tcgen05, single-pass fp16 inputs: fast mode             */
#define ZEDO_GEMM_FP32   2  /* CUDA-core float32 FFMA: validation kernel                */
#define ZEDO_GEMM_SPLIT2 3  /* tcgen05, fp16 activations x (hi+lo) weights: 2 MMA passes  */
#define ZEDO_GEMM_FP8LO  4  /* tcgen05, fp16 main product + the two low-order products in e4m3 (kind::f8f6f4):
                               2 fp16-pass equivalents, 2^-15 product error (split3: 2^-22, split2: 2^-12); holds every parity
                               bound of split3 and is what the host layer uses when no mode is named.
                               The e4m3 weight images sit under one power-of-two scale per matrix that keeps four
                               significant bits down to max / 2^10; a plan with a 1024x1024 weight whose max / median |w|
                               exceeds 1024 (too heavy-tailed even for that) runs this mode as SPLIT3; ZEDO_FP8LO_FORCE=1 overrides. */

/* network kinds */
#define ZEDO_NET_SCORE_FC_ADV 0  /* ScoreModelFC_Adv          (model.py:97-298)          */
#define ZEDO_NET_CONTROL      1  /* Control_ScoreModelFC_Adv  (control_model.py:97-382)  */

/* predictors (sampling.py:180-255) and SDEs (sde_lib.py:112-261) for zedo_sde_step */
#define ZEDO_PRED_EULER_MARUYAMA    0
#define ZEDO_PRED_REVERSE_DIFFUSION 1

typedef struct zedo_plan zedo_plan;

typedef struct zedo_net_desc {
  int32_t kind;      /* ZEDO_NET_*                                   */
  int32_t n_joints;  /* J; 3*J <= 64                                 */
  int32_t hidden;    /* hidden_dim; must be 1024 (GroupNorm(32, hidden) = groups of 32 channels) */
  int32_t embed;     /* embed_dim, multiple of 16 (reference: 512)    */
  int32_t n_blocks;  /* residual blocks (reference: 2)               */
  float   gn_eps;    /* GroupNorm eps (1e-5)                         */
} zedo_net_desc;

/* ---- plan: packed weights + workspaces ---------------------------------------------------
 * Replaces: ScoreModelFC_Adv.__init__/load_state_dict/.to(device) (run/opt_main.py:69-137).
 * names[i] is a state_dict key ("pre_dense.weight", "b1_gnorm2.bias", ...; a leading
 * "module." is ignored, run/opt_main.py:130-132) and tensors[i] the float32 data in the
 * reference layout ([out,in] for Linear weights); host or device pointers are both accepted.
 * max_batch = largest number of poses a single call will carry.  Synchronises the device. */
int zedo_plan_create(zedo_plan** out, const zedo_net_desc* desc, int32_t n_tensors,
                     const char* const* names, const float* const* tensors,
                     const int64_t* numels, int64_t max_batch, int32_t device);
int zedo_plan_destroy(zedo_plan* plan);
int64_t zedo_plan_capacity(const zedo_plan* plan);
/* Size the per-step bias tables for schedules of up to max_steps steps (config.ZeDO.OIL_iterations,
 * run/opt_main.py:197) and, for gemm_mode == ZEDO_GEMM_FP32, the float32 validation workspaces, so that no
 * later call allocates.  Synchronises `stream`. */
int zedo_plan_reserve(zedo_plan* plan, int32_t max_steps, int32_t gemm_mode, void* stream);

/* ---- process-wide tuning options ---------------------------------------------------------------
 * Defaults are the product path; none selects a non-CUDA path.  Initial values may be given through the
 * environment variable named beside each option (read once, at first use). */
#define ZEDO_OPT_GEOM_KERNEL     0  /* geometry kernel: 0 = by batch size, 1 = warp per pose, 2 = 128-pose CTAs, 3 = (zedo_oil_loop) rays precomputed once per loop (ZEDO_GEOM) */
#define ZEDO_OPT_PDL             1  /* programmatic dependent launch between the kernels of a step (ZEDO_PDL), default 1 */
#define ZEDO_OPT_SMALL_TILES     2  /* batches of at most this many 128-row tiles use 64-channel tiles (ZEDO_SMALL_TILES), 18;
                                       read by zedo_plan_create */
#define ZEDO_OPT_CTA_PAIRS       3  /* cta_group::2 kernel for the 1024x1024 layers (ZEDO_TC2), default 1; read by plan_create */
#define ZEDO_OPT_FP8LO_FORCE     4  /* e4m3 low-order products even for heavy-tailed weights (ZEDO_FP8LO_FORCE), default 0 */
#define ZEDO_OPT_EXPERIMENT      5  /* timing experiments; only in builds with -DZEDO_EXPERIMENTS=1 (ZEDO_DBG) */
#define ZEDO_OPT_LEAN_EW         6  /* epilogue warps of the K = 64 first layer (no residual / addend operands): 16 (default)
                                      or 8 (ZEDO_LEAN_EW) */
#define ZEDO_OPT_GRAPH           7  /* zedo_oil_loop as ONE cudaGraphLaunch: the call is captured the first time it is seen and
                                      replayed while it repeats verbatim (same buffers, sizes, schedule); default 0 --
                                      capturing ~7 launches per step costs about a third of a small-batch loop, so it pays
                                      for callers that replay a loop on persistent buffers; calls on the legacy default
                                      stream (which cannot be captured) are launched directly (ZEDO_GRAPH) */
#define ZEDO_OPT_TMA_2SM         8  /* CTA-pair kernel: operand stages by tensor-map TMA whose completion is credited to the leader
                                      CTA's barrier (cp.async.bulk.tensor ... cta_group::2) instead of linear bulk copies plus a
                                      relay hop from the peer CTA; same bytes, same results; default 1 (ZEDO_TMA_2SM) */
#define ZEDO_OPT_COUNT           9
int zedo_set_option(int32_t option, int32_t value);
int zedo_get_option(int32_t option, int32_t* value);

/* ---- score network forward -----------------------------------------------------------------
 * Replaces: ScoreModelFC_Adv.forward(batch, t, condition, mask) (model.py:215-298) in eval
 * mode for a batch-uniform time label t999 (= 999*t, utils.py:762); condition/mask are never
 * read by the reference forward.  x, out: [B, J, 3]. */
int zedo_score_forward(zedo_plan* plan, const float* x, float t999, float* out, int64_t B,
                       int32_t gemm_mode, void* stream);

/* ---- per-step geometry -----------------------------------------------------------------------
 * Replaces: gradient_field_gen(key2d, key3d, K, t=T|None, conf=conf|None, returnT=True)
 * (simple_zeroshot_opt.py:46-125), noise_type=None.
 * uv [B,J,2], x [B,J,3], K [B,3,3], conf [B,J] or NULL (clamped IN PLACE to [1e-4,1] like
 * the reference :64-66 when clamp_conf_inplace != 0), T [B,3] in/out: read when solve_T == 0,
 * written (least-squares solve, sign flip) when solve_T != 0.  g [B,J,3] out (may be NULL);
 * when x_out != NULL it receives x + g (may alias x). */
int zedo_grad_field(const float* uv, const float* x, const float* K, float* conf, float* T,
                    int32_t solve_T, int32_t clamp_conf_inplace, float* g, float* x_out,
                    int64_t B, int32_t J, void* stream);

/* ---- one predictor update with the network in the loop -------------------------------------
 * Replaces: pc_sampler(...) (sampling.py:450-527) = NoneCorrector + Predictor.update_fn
 * (sampling.py:180-205) over RSDE.sde / RSDE.discretize (sde_lib.py:93-107) and get_score_fn
 * (utils.py:751-777) for the sub-VP SDE.  x [B,J,3] in; x_next, x_mean [B,J,3] out (either
 * may be NULL or alias x).  z: injected randn_like(x) or NULL (= zeros); ignored when
 * probability_flow != 0.  t is the continuous time (0.01 .. 0.1 in the shipped configs). */
int zedo_sde_step(zedo_plan* plan, const float* x, float t, const float* z, int32_t predictor,
                  int32_t probability_flow, float beta_min, float beta_max, int32_t n_scales,
                  float* x_next, float* x_mean, int64_t B, int32_t gemm_mode, void* stream);

/* ---- noise-bearing predictors / correctors with caller-injected noise ---------------------------------
 * Replaces: AncestralSamplingPredictor (sampling.py:208-244), LangevinCorrector (:258-287) and
 * AnnealedLangevinDynamics (:290-324) over get_score_fn (utils.py:751-795), for a batch-uniform time.
 * Two calls per update so that a sharded run can make the Langevin batch means GLOBAL in between:
 *
 *   zedo_score_stats   network forward with time label `label` (999 t for VP / sub-VP, the marginal std for a
 *                      continuous VE SDE); the output stays in the plan.  When `stats` != NULL it receives
 *                      { sum over rows of |score_row|_2, sum over rows of |z_row|_2, rows } (device double[3],
 *                      overwritten, summed in a fixed order): sampling.py:281-282 before the `.mean()`.
 *                      A multi-GPU caller all-reduces (SUM) `stats` over the ranks here.
 *   zedo_noise_update  the elementwise update from the network output of the preceding zedo_score_stats /
 *                      zedo_score_forward call on this plan (ZEDO_E_STATE if B differs), one float32 rounding per
 *                      tensor op of the reference.  x, z [B,J,3]; x_next / x_mean may be NULL or alias x.
 *
 * std_div: score = (-net) / std_div (VP, sub-VP: the marginal std) or score = net when std_div == 0 (VE). */
#define ZEDO_UPD_ANCESTRAL_VP 0  /* p0 = discrete_betas[timestep]                                  (:233-241) */
#define ZEDO_UPD_ANCESTRAL_VE 1  /* p0 = sigma, p1 = adjacent sigma                                (:220-231) */
#define ZEDO_UPD_LANGEVIN     2  /* p0 = snr, p1 = alpha; step from the batch means in `stats`      (:277-285) */
#define ZEDO_UPD_ALD          3  /* p0 = snr, p1 = alpha, p2 = marginal std; stats unused           (:314-321) */
int zedo_score_stats(zedo_plan* plan, const float* x, float label, const float* z, float std_div, double* stats,
                     int64_t B, int32_t gemm_mode, void* stream);
int zedo_noise_update(zedo_plan* plan, int32_t kind, const float* x, const float* z, float std_div, float p0,
                      float p1, float p2, const double* stats, float* x_next, float* x_mean, int64_t B,
                      void* stream);

/* ---- the whole OIL loop ---------------------------------------------------------------------
 * Replaces: the `for i in range(sample_num)` body of run/opt_main.py:202-220 (and
 * run/inference.py:211-229): steps x {gradient_field_gen -> x += g -> pc_sampler}, state
 * resident on the device for all steps.  x [B,J,3] in/out (rotated hypothesis R x0),
 * T [B,3] in/out, t_sched host float[steps] (= torch.linspace(sde.T, eps, steps)).
 * Steps i < phase_switch keep T; later steps re-solve it (phase_switch = steps/5 in
 * opt_main.py:203, 950 in opt_main_infant.py:310).  dump (nullable) [n_dump,B,J,3] receives
 * the pose after step dump_steps[k] (host int array, strictly ascending). */
int zedo_oil_loop(zedo_plan* plan, float* x, float* T, const float* uv, const float* K,
                  float* conf, const float* t_sched, int32_t steps, int32_t phase_switch,
                  float beta_min, float beta_max, int32_t n_scales, float* dump,
                  const int32_t* dump_steps, int32_t n_dump, int64_t B, int32_t gemm_mode,
                  void* stream);

/* ---- IPO: rotation / scale fit ---------------------------------------------------------------
 * Replaces: T0 (run/opt_main.py:177-179) + RotOpt (simple_zeroshot_opt.py:8-31) +
 * quaternion_to_matrix (utils.py:59-88) + the 500-iteration Adam loop (run/opt_main.py:180-195)
 * with the analytic gradient of the mean L1 reprojection loss.
 * x0 [B,J,3] hypothesis, uv [B,J,2], K [B,3,3], keylist host int[nkey]; axes_mask bit0 = x,
 * bit1 = y, bit2 = z trainable (config.ZeDO.RotAxes); B_global = batch size of the loss mean
 * (global batch when poses are sharded).  Outputs R [B,9], T [B,3] = T0*clamp(scale) and
 * x_rot [B,J,3] = R x0 (may be NULL).  qs (nullable) [B,5] receives (w,x,y,z,scale). */
int zedo_ipo_fit(const float* x0, const float* uv, const float* K, const int32_t* keylist,
                 int32_t nkey, int32_t axes_mask, float ipo_T, float minT, float maxT,
                 int32_t iters, int64_t B_global, float lr, float* R, float* T, float* x_rot,
                 float* qs, int64_t B, int32_t J, void* stream);

/* Infant driver variant (run/opt_main_infant.py:255-300): the pelvis pixel is (uv[pelvis_a] + uv[pelvis_b]) / 2
 * (joint 0 for Mini-RGBD, joints 0 and 3 for SyRIP) and, when ray_init != 0, x_rot = R . (back-projected 2D rays
 * scaled so the pelvis ray has length |T|, pelvis-subtracted) instead of R . x0. */
int zedo_ipo_fit_ex(const float* x0, const float* uv, const float* K, const int32_t* keylist,
                    int32_t nkey, int32_t axes_mask, int32_t pelvis_a, int32_t pelvis_b,
                    int32_t ray_init, float ipo_T, float minT, float maxT, int32_t iters,
                    int64_t B_global, float lr, float* R, float* T, float* x_rot, float* qs, int64_t B,
                    int32_t J, void* stream);

/* RotOpt.forward / its backward for drivers that keep autograd + torch.optim.Adam
 * (simple_zeroshot_opt.py:20-31).  q [B,4], scale [B], xk [B,nk,3], T0 [B,3], K [B,9];
 * uv_out [B,nk,2].  Backward: d_uv [B,nk,2] -> d_q [B,4], d_scale [B]. */
int zedo_rotopt_forward(const float* q, const float* scale, const float* xk, const float* T0,
                        const float* K, float minT, float maxT, float* uv_out, int64_t B,
                        int32_t nk, void* stream);
int zedo_rotopt_backward(const float* q, const float* scale, const float* xk, const float* T0,
                         const float* K, float minT, float maxT, const float* d_uv, float* d_q,
                         float* d_scale, int64_t B, int32_t nk, void* stream);

/* ---- evaluation --------------------------------------------------------------------------------
 * Replaces: eval_multi (lib/dataset/h36m.py:365-442, pw3d.py:286-345) + align_to_gt /
 * procrustes (lib/utils/transforms.py:42-148).  pred [N,S,J,3] float32, gt [N,J,3] float64
 * root-relative metres; joint_subset host int[n_sub] or NULL (all J).  Outputs per pose:
 * err_min [N] float64 = amin over hypotheses of the mean per-joint error (after Procrustes
 * when protocol2 != 0), argmin [N] int32 (first minimum wins, numpy semantics),
 * err_all [N,S] float64 (nullable) and aligned [N,S,J,3] float64 (nullable) = the pose that was
 * scored (align_to_gt's output Z in protocol 2).  Computed in float64 like the reference's numpy path. */
int zedo_eval_multi(const float* pred, const double* gt, int32_t protocol2, int64_t N, int32_t S,
                    int32_t J, const int32_t* joint_subset, int32_t n_sub, double* err_min,
                    int32_t* argmin, double* err_all, double* aligned, void* stream);

/* PCK / AUC of MPI-INF-3DHP.  Replaces: compute_PCK / compute_AUC (lib/algorithms/advanced/utils.py:814-849,
 * called from lib/dataset/mpii3dHP.py:480-481) on the selected hypothesis select[n] (NULL = hypothesis 0).
 * counts: device uint64[31], counts[k] = number of (pose, joint) with error_mm < 5 k (thresholds
 * linspace(0, 150, 31)); PCK = 100 counts[30] / total, AUC = mean_k 100 counts[k] / total. */
int zedo_pck_counts(const float* pred, const double* gt, const int32_t* select, int64_t N, int32_t S,
                    int32_t J, const int32_t* joint_subset, int32_t n_sub, uint64_t* counts, void* stream);

/* Diversity of the hypotheses.  Replaces the "std" report of lib/dataset/mpii3dHP.py:487-490:
 * multi_preds_cam - multi_preds_cam[:, :, [0], :], root dropped, .std(axis=1) over the S hypotheses.
 * out_std: device float64 [N, J-1, 3] (population std, ddof = 0); the reference prints its mean over
 * poses and joints per coordinate. */
int zedo_hypothesis_std(const float* pred, int64_t N, int32_t S, int32_t J, double* out_std, void* stream);

/* ---- cluster-pose generation -------------------------------------------------------------------------
 * Produces what the drivers load from clusters/{h36m,3dhp}_cluster{S}.npy (run/opt_main.py:58-65): S centres of
 * Lloyd's k-means over training poses.  The reference ships the files, not their generator (run/opt_main_infant.py:25,34
 * only imports scipy.cluster.vq / sklearn KMeans).  x [N, D] float32 (D = 3 J flattened poses), centers [S, D] float32
 * in/out (initial centres in, fitted centres out), assign [N] int32 out (label under the final centres), dist [N]
 * float64 out (nullable; squared distance to the assigned centre).  `iters` Lloyd iterations; distances and sums in
 * float64 with a fixed summation order (bit-reproducible); ties go to the lowest centre index; an empty cluster keeps
 * its centre. */
int zedo_kmeans_fit(const float* x, int64_t N, int32_t D, int32_t S, int32_t iters, float* centers,
                    int32_t* assign, double* dist, void* stream);

/* ---- misc ---------------------------------------------------------------------------------------- */
const char* zedo_strerror(int code);
int zedo_abi_version(void);
/* number of kernels this library has launched in the calling process (bench.py gpu_launches) */
int64_t zedo_launch_count(void);
/* Live kernel timing for bench.py's roofline: while enabled, every `stride`-th launch of each
 * kernel kind inside zedo_oil_loop / zedo_score_forward is bracketed by CUDA events on the
 * launching stream (at most 256 samples per kind).  kinds: 0 = first layer (K=64), 1 = hidden
 * layer (K=1024, the dominant kernel), 2 = post_dense, 3 = geometry, 4 = SDE update.
 * zedo_plan_profile_read synchronises the recorded events and returns the mean duration. */
int zedo_plan_profile(zedo_plan* plan, int32_t enable, int32_t stride);
int zedo_plan_profile_read(zedo_plan* plan, int32_t kind, float* mean_ms, int32_t* n_samples);
/* host-side helpers exposed for the CPU test-suite (no GPU needed):
 * sub-VP scalars in the reference's float32 op order (sde_lib.py:187-198). */
int zedo_subvp_scalars(float t, float beta_min, float beta_max, float* beta_t, float* diffusion,
                       float* std);
/* byte offset of element (row, col) of a [rows, cols] fp16 operand inside the blocked,
 * core-matrix-interleaved layout the tcgen05 kernels read (DESIGN.md "data layout"); tile_rows = 128
 * for activations, 256/64 for weights; hl = 0 (hi) / 1 (lo). */
int64_t zedo_blocked_offset(int64_t row, int64_t col, int64_t cols, int32_t tile_rows, int32_t hl);

#ifdef __cplusplus
}
#endif
#endif /* ZEDO_B200_H */
