// Cluster-pose generation: Lloyd's k-means over training poses, the producer of the `clusters/*_cluster{S}.npy`
// files the drivers load as hypothesis initialisations (reference run/opt_main.py:58-65; the reference ships the
// files, not the generator -- run/opt_main_infant.py:25,34 only imports scipy.cluster.vq / sklearn KMeans).
//   assign : one thread per pose, centres staged in shared memory, squared distance in float64, first minimum wins
//   update : one CTA per centre, fixed summation order (thread-strided partial sums, then a shared-memory tree),
//            so a fit is bit-reproducible; an empty cluster keeps its centre
#include "kernels.cuh"

namespace zedo {

__global__ void __launch_bounds__(256)
kmeans_assign_kernel(const float* __restrict__ x, const float* __restrict__ centers, int64_t N, int D, int S,
                     int* __restrict__ assign, double* __restrict__ dist) {
  extern __shared__ float sc[];  // [S, D]
  for (int i = threadIdx.x; i < S * D; i += blockDim.x) sc[i] = centers[i];
  __syncthreads();
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float* xp = x + n * D;
  double best = 0.0;
  int best_c = 0;
  for (int c = 0; c < S; ++c) {
    double d = 0.0;
    for (int k = 0; k < D; ++k) {
      const double t = (double)xp[k] - (double)sc[c * D + k];
      d += t * t;
    }
    if (c == 0 || d < best) {
      best = d;
      best_c = c;
    }
  }
  assign[n] = best_c;
  if (dist != nullptr) dist[n] = best;
}

__global__ void __launch_bounds__(256)
kmeans_update_kernel(const float* __restrict__ x, const int* __restrict__ assign, int64_t N, int D,
                     float* __restrict__ centers) {
  __shared__ double sh[256];
  __shared__ int shc[256];
  const int c = blockIdx.x;
  int cnt = 0;
  for (int64_t n = threadIdx.x; n < N; n += blockDim.x) cnt += assign[n] == c;
  shc[threadIdx.x] = cnt;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) shc[threadIdx.x] += shc[threadIdx.x + o];
    __syncthreads();
  }
  const int total = shc[0];
  if (total == 0) return;  // empty cluster: keep the centre
  for (int k = 0; k < D; ++k) {
    double s = 0.0;
    for (int64_t n = threadIdx.x; n < N; n += blockDim.x)
      if (assign[n] == c) s += (double)x[n * D + k];
    __syncthreads();
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) centers[c * D + k] = (float)(sh[0] / total);
  }
}

int launch_kmeans(const float* x, int64_t N, int D, int S, int iters, float* centers, int* assign, double* dist,
                  cudaStream_t st) {
  if (N == 0 || S == 0) return 0;
  const size_t smem = (size_t)S * D * sizeof(float);
  if (smem > 200 * 1024) return ZEDO_E_SHAPE;
  if (smem > 48 * 1024) ZEDO_CUDA_TRY(ensure_max_smem((const void*)kmeans_assign_kernel, (int)smem));
  const unsigned grid = (unsigned)((N + 255) / 256);
  for (int it = 0; it < iters; ++it) {
    kmeans_assign_kernel<<<grid, 256, smem, st>>>(x, centers, N, D, S, assign, nullptr);
    ZEDO_LAUNCH_CHECK();
    kmeans_update_kernel<<<S, 256, 0, st>>>(x, assign, N, D, centers);
    ZEDO_LAUNCH_CHECK();
  }
  kmeans_assign_kernel<<<grid, 256, smem, st>>>(x, centers, N, D, S, assign, dist);  // labels of the final centres
  ZEDO_LAUNCH_CHECK();
  return 0;
}

}  // namespace zedo
