// CUDA-core float32 kernels:
//   * the per-step bias tables of the time-embedding path (batch-invariant, model.py:253-259 and
//     every `*_t` Linear at model.py:265,273,281) -- computed once per time schedule;
//   * the ZEDO_GEMM_FP32 validation mode of the score network (plain FFMA GEMM + GroupNorm/SiLU),
//     used on the device to cross-check the tcgen05 path element by element.
#include "kernels.cuh"

namespace zedo {

// C[M,N] (+)= A[M,K] (row-major, lda) * W[N,K]^T (row-major, ldw) + bias[N] (nullable)
// 64x64 tile, BK = 16, 256 threads, 4x4 outputs per thread.
constexpr int TS = 64, TK = 16;

__global__ void __launch_bounds__(256)
sgemm_tn_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
                const float* __restrict__ bias, float* C, int ldc, int M, int N, int K, int accumulate) {
  __shared__ float As[TK][TS + 1];
  __shared__ float Ws[TK][TS + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * TS, n0 = blockIdx.x * TS;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += TK) {
    for (int i = threadIdx.x; i < TS * TK; i += 256) {
      const int r = i / TK, c = i % TK;
      const int gm = m0 + r, gn = n0 + r, gk = k0 + c;
      As[c][r] = (gm < M && gk < K) ? A[(int64_t)gm * lda + gk] : 0.f;
      Ws[c][r] = (gn < N && gk < K) ? W[(int64_t)gn * ldw + gk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = Ws[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn < N) {
        const float prev = accumulate ? C[(int64_t)gm * ldc + gn] : 0.f;
        C[(int64_t)gm * ldc + gn] = prev + acc[i][j] + (bias ? bias[gn] : 0.f);
      }
    }
  }
}

// embedding of the time labels, one row of E = 2*half per step: [sin(arg), cos(arg)] with
//   positional (model.py:81-95):  arg = t999 * f_k
//   fourier    (model.py:27-36):  arg = ((log t999 * W_k) * 2) * pi, one float32 rounding per torch op (tin = log t999)
__global__ void timestep_embedding_kernel(const float* __restrict__ tin, const float* __restrict__ freqs,
                                          float* __restrict__ emb, int n_steps, int half, int fourier) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_steps * half) return;
  const int s = i / half, k = i % half;
  const float prod = __fmul_rn(tin[s], freqs[k]);
  const float arg = fourier ? __fmul_rn(__fmul_rn(prod, 2.0f), 3.14159274101257324f) : prod;
  emb[(int64_t)s * 2 * half + k] = sinf(arg);
  emb[(int64_t)s * 2 * half + half + k] = cosf(arg);
}

__global__ void silu_inplace_kernel(float* __restrict__ v, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float y = v[i];
    v[i] = y / (1.f + expf(-y));
  }
}

// one warp per row: v = in + cbias (+ addend) -> GroupNorm(groups of 32) -> SiLU (+ resid) -> out
__global__ void gn_silu_rows_kernel(const float* __restrict__ in, const float* __restrict__ cbias,
                                    const float* __restrict__ addend, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* resid, float* out, int64_t M,
                                    int C, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  for (int g = 0; g < C / 32; ++g) {
    const int col = g * 32 + lane;
    float v = in[row * C + col] + cbias[col];
    if (addend != nullptr) v += addend[row * C + col];
    const float mean = warp_sum(v) * (1.f / 32.f);
    const float d = v - mean;
    const float var = warp_sum(d * d) * (1.f / 32.f);
    float y = d * (1.f / sqrtf(var + eps)) * gamma[col] + beta[col];
    y = y / (1.f + expf(-y));
    if (resid != nullptr) y += resid[row * C + col];
    out[row * C + col] = y;
  }
}

int launch_sgemm_tn(const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc, int M,
                    int N, int K, cudaStream_t st, int accumulate) {
  if (M == 0 || N == 0) return 0;
  dim3 grid((N + TS - 1) / TS, (M + TS - 1) / TS);
  sgemm_tn_kernel<<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, C, ldc, M, N, K, accumulate);
  ZEDO_LAUNCH_CHECK();
  return 0;
}

int launch_timestep_embedding(const float* tin, const float* freqs, float* emb, int n_steps, int half, int fourier,
                              cudaStream_t st) {
  const int n = n_steps * half;
  timestep_embedding_kernel<<<(n + 255) / 256, 256, 0, st>>>(tin, freqs, emb, n_steps, half, fourier);
  ZEDO_LAUNCH_CHECK();
  return 0;
}

int launch_silu_inplace(float* v, int64_t n, cudaStream_t st) {
  silu_inplace_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(v, n);
  ZEDO_LAUNCH_CHECK();
  return 0;
}

int launch_gn_silu_rows(const float* in, const float* cbias, const float* addend, const float* gamma,
                        const float* beta, const float* resid, float* out, int64_t M, int C, float eps,
                        cudaStream_t st) {
  if (M == 0) return 0;
  const int warps = 8;
  gn_silu_rows_kernel<<<(unsigned)((M + warps - 1) / warps), warps * 32, 0, st>>>(in, cbias, addend, gamma, beta,
                                                                                resid, out, M, C, eps);
  ZEDO_LAUNCH_CHECK();
  return 0;
}

}  // namespace zedo
