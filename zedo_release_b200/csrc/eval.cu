// K5: multi-hypothesis evaluation -- MPJPE / PA-MPJPE and the per-pose argmin over hypotheses.
//
// Restates eval_multi (lib/dataset/h36m.py:394-417, pw3d.py:303-338) and align_to_gt / procrustes
// (lib/utils/transforms.py:42-148: scaling=True, reflection='best', i.e. NO determinant fix, so
// reflections are allowed).  One thread per pose, the S hypotheses are a serial loop so "first
// minimum wins" exactly like np.argmin.  Float64 throughout: the reference's
// numpy path promotes to float64 (gt comes from a float64 pickle), and the selection indices
// have to be bit-exact.
#include "kernels.cuh"

namespace zedo {

// One-sided Jacobi SVD of a 3x3 matrix M (row-major): on exit the columns of M are U_i * s_i and
// V accumulates the right rotations, M_in = U diag(s) V^T.
__device__ void svd3x3_onesided(double* M, double* V) {
  V[0] = V[4] = V[8] = 1.0;
  V[1] = V[2] = V[3] = V[5] = V[6] = V[7] = 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0;
#pragma unroll
    for (int pair = 0; pair < 3; ++pair) {
      const int p = pair == 2 ? 1 : 0;
      const int q = pair == 0 ? 1 : 2;
      const double alpha = M[p] * M[p] + M[3 + p] * M[3 + p] + M[6 + p] * M[6 + p];
      const double beta = M[q] * M[q] + M[3 + q] * M[3 + q] + M[6 + q] * M[6 + q];
      const double gamma = M[p] * M[q] + M[3 + p] * M[3 + q] + M[6 + p] * M[6 + q];
      off = fmax(off, fabs(gamma) / sqrt(fmax(alpha * beta, 1e-300)));
      if (fabs(gamma) <= 1e-300) continue;
      const double zeta = (beta - alpha) / (2.0 * gamma);
      const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
      const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const double mp = M[3 * r + p], mq = M[3 * r + q];
        M[3 * r + p] = c * mp - s * mq;
        M[3 * r + q] = s * mp + c * mq;
        const double vp = V[3 * r + p], vq = V[3 * r + q];
        V[3 * r + p] = c * vp - s * vq;
        V[3 * r + q] = s * vp + c * vq;
      }
    }
    if (off < 1e-15) break;
  }
}

// One THREAD per pose: the S hypotheses are a serial loop in that thread ("first minimum wins" exactly like np.argmin)
// and every lane of a warp runs its OWN 3x3 SVD on different data.  (The first version gave a pose to a warp with one
// joint per lane; all 32 lanes then repeated the same float64 Jacobi SVD.)  pred / gt rows are re-read per pass from
// L1: a warp's 32 poses touch 32 x 3J floats, which the first pass leaves resident.
__global__ void __launch_bounds__(128)
eval_multi_kernel(const float* __restrict__ pred, const double* __restrict__ gt, int protocol2, int64_t N, int S,
                  int J, const IntList subset, double* __restrict__ err_min,
                  int* __restrict__ argmin, double* __restrict__ err_all, double* __restrict__ aligned) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  // joints that count in the mean: all J, or the listed subset (e.g. the 12 SyRIP joints)
  uint32_t counted = J >= 32 ? 0xffffffffu : ((1u << J) - 1u);
  int n_counted = J;
  if (subset.n > 0) {
    counted = 0;
    for (int i = 0; i < subset.n; ++i) counted |= 1u << subset.v[i];
    n_counted = subset.n;
  }
  const double* g = gt + n * J * 3;
  // gt statistics for Procrustes (transforms.py:74-85)
  const double invJ = 1.0 / J;
  double am0 = 0, am1 = 0, am2 = 0, a_norm = 0;
  if (protocol2) {
    for (int j = 0; j < J; ++j) {
      am0 += g[3 * j];
      am1 += g[3 * j + 1];
      am2 += g[3 * j + 2];
    }
    am0 *= invJ;
    am1 *= invJ;
    am2 *= invJ;
    for (int j = 0; j < J; ++j) {
      const double a0 = g[3 * j] - am0, a1 = g[3 * j + 1] - am1, a2 = g[3 * j + 2] - am2;
      a_norm += a0 * a0 + a1 * a1 + a2 * a2;
    }
    a_norm = sqrt(a_norm);
  }

  double best = 0.0;
  int best_idx = 0;
  for (int s = 0; s < S; ++s) {
    const float* pp = pred + (n * S + s) * J * 3;
    double k = 1.0, bm0 = 0, bm1 = 0, bm2 = 0, b_norm = 1.0;
    double Rm[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (protocol2) {
      for (int j = 0; j < J; ++j) {
        bm0 += (double)pp[3 * j];
        bm1 += (double)pp[3 * j + 1];
        bm2 += (double)pp[3 * j + 2];
      }
      bm0 *= invJ;
      bm1 *= invJ;
      bm2 *= invJ;
      double bn = 0;
      for (int j = 0; j < J; ++j) {
        const double b0 = pp[3 * j] - bm0, b1 = pp[3 * j + 1] - bm1, b2 = pp[3 * j + 2] - bm2;
        bn += b0 * b0 + b1 * b1 + b2 * b2;
      }
      b_norm = sqrt(bn);
      double M[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, V[9];  // M = A0^T B0 of the normalised, centred point sets
      for (int j = 0; j < J; ++j) {
        const double an0 = (g[3 * j] - am0) / a_norm, an1 = (g[3 * j + 1] - am1) / a_norm,
                     an2 = (g[3 * j + 2] - am2) / a_norm;
        const double b0 = (pp[3 * j] - bm0) / b_norm, b1 = (pp[3 * j + 1] - bm1) / b_norm,
                     b2 = (pp[3 * j + 2] - bm2) / b_norm;
        M[0] += an0 * b0; M[1] += an0 * b1; M[2] += an0 * b2;
        M[3] += an1 * b0; M[4] += an1 * b1; M[5] += an1 * b2;
        M[6] += an2 * b0; M[7] += an2 * b1; M[8] += an2 * b2;
      }
      svd3x3_onesided(M, V);
      // R = V U^T = sum_i v_i u_i^T with u_i = M[:, i] / s_i ; trace(S) = sum_i s_i
#pragma unroll
      for (int i = 0; i < 9; ++i) Rm[i] = 0.0;
      double tr = 0.0;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double si = sqrt(M[i] * M[i] + M[3 + i] * M[3 + i] + M[6 + i] * M[6 + i]);
        tr += si;
        const double inv = si > 0.0 ? 1.0 / si : 0.0;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) Rm[3 * r + c] += V[3 * r + i] * (M[3 * c + i] * inv);
      }
      k = a_norm * tr;  // Z = A_norm * trace * (B0 R) + A_bar   (transforms.py:113)
    }
    double esum = 0.0;
    for (int j = 0; j < J; ++j) {
      double p0 = pp[3 * j], p1 = pp[3 * j + 1], p2 = pp[3 * j + 2];
      if (protocol2) {
        const double b0 = (p0 - bm0) / b_norm, b1 = (p1 - bm1) / b_norm, b2 = (p2 - bm2) / b_norm;
        p0 = k * (b0 * Rm[0] + b1 * Rm[3] + b2 * Rm[6]) + am0;
        p1 = k * (b0 * Rm[1] + b1 * Rm[4] + b2 * Rm[7]) + am1;
        p2 = k * (b0 * Rm[2] + b1 * Rm[5] + b2 * Rm[8]) + am2;
      }
      if (aligned != nullptr) {  // the pose eval_multi scores (align_to_gt output in protocol 2)
        double* ap = aligned + ((n * S + s) * J + j) * 3;
        ap[0] = p0;
        ap[1] = p1;
        ap[2] = p2;
      }
      if ((counted >> j) & 1u) {
        const double d0 = p0 - g[3 * j], d1 = p1 - g[3 * j + 1], d2 = p2 - g[3 * j + 2];
        esum += sqrt(d0 * d0 + d1 * d1 + d2 * d2);
      }
    }
    const double err = esum / n_counted;
    if (err_all != nullptr) err_all[n * S + s] = err;
    if (s == 0 || err < best) {
      best = err;
      best_idx = s;
    }
  }
  err_min[n] = best;
  argmin[n] = best_idx;
}

// PCK / AUC of MPI-INF-3DHP (utils.py:814-849, used by mpii3dHP.py:480-481): per-joint error (mm) of the
// selected hypothesis against 31 thresholds linspace(0, 150, 31); counts[k] += #(error_mm < 5 k).
__global__ void __launch_bounds__(128)
pck_counts_kernel(const float* __restrict__ pred, const double* __restrict__ gt, const int* __restrict__ select,
                  int64_t N, int S, int J, const IntList subset,
                  unsigned long long* __restrict__ counts) {
  __shared__ unsigned int local[31];
  if (threadIdx.x < 31) local[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n < N && lane < J) {
    bool counted = true;
    if (subset.n > 0) {
      counted = false;
      for (int i = 0; i < subset.n; ++i) counted |= (subset.v[i] == lane);
    }
    if (counted) {
      const int s = select != nullptr ? select[n] : 0;
      const float* pp = pred + ((n * S + s) * J + lane) * 3;
      const double* gp = gt + (n * J + lane) * 3;
      const double d0 = (double)pp[0] - gp[0], d1 = (double)pp[1] - gp[1], d2 = (double)pp[2] - gp[2];
      const double e_mm = sqrt(d0 * d0 + d1 * d1 + d2 * d2) * 1000.0;
      for (int k = 0; k < 31; ++k)
        if (e_mm < 5.0 * k) atomicAdd(&local[k], 1u);
    }
  }
  __syncthreads();
  if (threadIdx.x < 31 && local[threadIdx.x] != 0) atomicAdd(&counts[threadIdx.x], (unsigned long long)local[threadIdx.x]);
}

int launch_pck_counts(const float* pred, const double* gt, const int* select, int64_t N, int S, int J,
                      const IntList& subset, unsigned long long* counts, cudaStream_t st) {
  if (N == 0) return 0;
  const int warps = 4;
  pck_counts_kernel<<<(unsigned)((N + warps - 1) / warps), warps * 32, 0, st>>>(pred, gt, select, N, S, J, subset,
                                                                             counts);
  ZEDO_LAUNCH_CHECK();
  return 0;
}

// Diversity of the hypotheses (mpii3dHP.py:487-490): root-relative joints 1..J-1, population standard deviation
// over the S hypotheses of every coordinate.  One thread per (pose, joint, coordinate), two passes, float64.
__global__ void hypothesis_std_kernel(const float* __restrict__ pred, int64_t N, int S, int J,
                                      double* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int per_pose = (J - 1) * 3;
  if (t >= N * per_pose) return;
  const int64_t n = t / per_pose;
  const int e = (int)(t - n * per_pose), j = 1 + e / 3, c = e % 3;
  const float* base = pred + n * S * J * 3;
  double mean = 0.0;
  for (int s = 0; s < S; ++s) mean += (double)base[(s * J + j) * 3 + c] - (double)base[s * J * 3 + c];
  mean /= S;
  double var = 0.0;
  for (int s = 0; s < S; ++s) {
    const double d = (double)base[(s * J + j) * 3 + c] - (double)base[s * J * 3 + c] - mean;
    var += d * d;
  }
  out[t] = sqrt(var / S);
}

int launch_hypothesis_std(const float* pred, int64_t N, int S, int J, double* out, cudaStream_t st) {
  const int64_t total = N * (J - 1) * 3;
  if (total <= 0) return 0;
  hypothesis_std_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(pred, N, S, J, out);
  ZEDO_LAUNCH_CHECK();
  return 0;
}

int launch_eval_multi(const float* pred, const double* gt, int protocol2, int64_t N, int S, int J,
                      const IntList& subset, double* err_min, int* argmin, double* err_all,
                      double* aligned, cudaStream_t st) {
  if (N == 0) return 0;
  eval_multi_kernel<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(pred, gt, protocol2, N, S, J, subset, err_min, argmin,
                                                              err_all, aligned);
  ZEDO_LAUNCH_CHECK();
  return 0;
}

}  // namespace zedo
