// Declarations shared between the kernel translation units and the C-ABI glue (api.cu).
#pragma once
#include "common.cuh"

namespace zedo {

struct LayerArgs {
  const __half* A;
  const __half* W;
  const float* cbias;
  const float* gamma;
  const float* beta;
  const __half* addend;
  const __half* resid;
  __half* out;
  float* out_f32;
  int ld_out;
  int m_tiles, n_tiles, num_kb;
  float descale;
  float gn_eps;
  int dbg;  // timing experiments only (env ZEDO_DBG): 1 = no bulk copies after the first fill, 2 = no MMAs
  int a_fmt;  // block format of A (common.cuh): 0 = [hi16 | lo16], 1 = [hi16 | hi8 | lo8 | lo16]
  int o_fmt;  // block format of out / resid / addend
  // CTA-pair kernel, optional: tensor maps (device pointers to CUtensorMap, 64-byte aligned) that describe the A buffer
  // and the W tiles as [bytes / 128][128] byte matrices with a 32 KiB box; with them both CTAs' stage copies complete on the
  // LEADER's barrier (cp.async.bulk.tensor ... cta_group::2) and the peer's relay hop disappears.  NULL = linear bulk copies.
  const void* tmapA;
  const void* tmapW;
  int o_flags;  // format-1 output: bit 0 = some consumer reads the e4m3 images (hi8, lo8), bit 1 = some consumer reads
                // lo16 (residual / addend epilogue, post_dense); images nobody reads are neither formed nor stored
};
// a short host-side index list (IPO key joints, evaluated joint subset) travels by value in the kernel parameters:
// no device allocation or copy per call (n = 0: "no list")
struct IntList {
  int n = 0;
  int v[32] = {};
};
constexpr int EPI_GN_SILU = 0, EPI_LINEAR_ACT = 1, EPI_LINEAR_F32 = 2;
int launch_layer_tc(const LayerArgs& a, int bn, int nprod, int epi, int num_sms, cudaStream_t st);
int launch_layer_tc2(const LayerArgs& a, int nprod, int epi, int num_sms, cudaStream_t st);
// eps_prev/prev/dump: optional fused tail of the previous OIL step (predictor update with that step's eps)
int launch_grad_field(const float* uv, const float* x, const float* K, float* conf, float* T, int solve_T,
                      int clamp_inplace, float* g, float* x_out, __half* xa, int64_t B, int J, cudaStream_t st,
                      const float* eps_prev = nullptr, const SdeCoef* prev = nullptr, float* dump = nullptr);
// the OIL loop's geometry on rays precomputed once per loop (geom.cu): selected by batch size / ZEDO_OPT_GEOM_KERNEL = 3
bool oil_rays_selected(int64_t B);
size_t oil_rays_slots(int64_t rows, int J);  // entries of rays_a / rays_b for `rows` poses
int launch_oil_rays(const float* uv, const float* K, float* conf, float4* rays_a, float2* rays_b, double* pose_c,
                    int64_t B, int J, cudaStream_t st);
int launch_oil_geom(const float4* rays_a, const float2* rays_b, const double* pose_c, float* x, float* T, int solve_T,
                    __half* xa, int64_t B, int J, cudaStream_t st, const float* eps_prev, const SdeCoef* prev,
                    float* dump);
int launch_pack_x(const float* x, __half* xa, int64_t B, int D, cudaStream_t st);
int launch_sde_update(const float* x, const float* eps, int ld_eps, const float* z, const SdeCoef& c, int predictor,
                      int probability_flow, float* x_next, float* x_mean, int64_t B, int D, cudaStream_t st);
int launch_row_norm_stats(const float* eps, int ld_eps, const float* z, float std_div, float* norms, double* stats,
                          int64_t B, int D, cudaStream_t st);
int launch_noise_update(int kind, const float* x, const float* eps, int ld_eps, const float* z, float std_div,
                        float p0, float p1, float p2, const double* stats, float* x_next, float* x_mean, int64_t B,
                        int D, cudaStream_t st);
int launch_kmeans(const float* x, int64_t N, int D, int S, int iters, float* centers, int* assign, double* dist,
                  cudaStream_t st);
int launch_sgemm_tn(const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc, int M,
                    int N, int K, cudaStream_t st, int accumulate = 0);
int launch_timestep_embedding(const float* tin, const float* freqs, float* emb, int n_steps, int half, int fourier,
                              cudaStream_t st);
int launch_silu_inplace(float* v, int64_t n, cudaStream_t st);
int launch_gn_silu_rows(const float* in, const float* cbias, const float* addend, const float* gamma,
                        const float* beta, const float* resid, float* out, int64_t M, int C, float eps,
                        cudaStream_t st);
int launch_ipo_fit(const float* x0, const float* uv, const float* K, const IntList& keylist, int axes_mask,
                   int pelvis_a, int pelvis_b, int ray_init, float ipo_T, float minT, float maxT, int iters, int64_t B_global, float lr, float* R, float* T,
                   float* x_rot, float* qs, int64_t B, int J, cudaStream_t st);
int launch_rotopt_forward(const float* q, const float* scale, const float* xk, const float* T0, const float* K,
                          float minT, float maxT, float* uv_out, int64_t B, int nk, cudaStream_t st);
int launch_rotopt_backward(const float* q, const float* scale, const float* xk, const float* T0, const float* K,
                           float minT, float maxT, const float* d_uv, float* d_q, float* d_scale, int64_t B, int nk,
                           cudaStream_t st);
int launch_hypothesis_std(const float* pred, int64_t N, int S, int J, double* out, cudaStream_t st);
int launch_pck_counts(const float* pred, const double* gt, const int* select, int64_t N, int S, int J,
                      const IntList& subset, unsigned long long* counts, cudaStream_t st);
int launch_eval_multi(const float* pred, const double* gt, int protocol2, int64_t N, int S, int J,
                      const IntList& subset, double* err_min, int* argmin, double* err_all,
                      double* aligned, cudaStream_t st);


}  // namespace zedo
