// K4: the IPO rotation/scale fit -- all Adam iterations of one pose inside one thread.
//
// Restates run/opt_main.py:175-195 + RotOpt (simple_zeroshot_opt.py:8-31) + quaternion_to_matrix
// (utils.py:59-88) with the analytic gradient of mean |uv - uv*|.
// The reference spends ~60 tiny kernels + autograd per iteration (30k launches per hypothesis);
// here the quaternion, scale and Adam moments never leave registers, the key joints live in
// shared memory (struct-of-arrays, conflict-free), and the only HBM traffic is one read of
// x0/uv/K and one write of R/T/x_rot per pose.
#include "kernels.cuh"

namespace zedo {

constexpr int kIpoThreads = 64;

struct Quat {
  float w, x, y, z;
};

__device__ __forceinline__ void quat_to_R(const Quat& q, float* R) {
  // utils.py:71-87 (q is not normalised: two_s = 2 / |q|^2)
  const float two_s = 2.0f / (q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  R[0] = 1 - two_s * (q.y * q.y + q.z * q.z);
  R[1] = two_s * (q.x * q.y - q.z * q.w);
  R[2] = two_s * (q.x * q.z + q.y * q.w);
  R[3] = two_s * (q.x * q.y + q.z * q.w);
  R[4] = 1 - two_s * (q.x * q.x + q.z * q.z);
  R[5] = two_s * (q.y * q.z - q.x * q.w);
  R[6] = two_s * (q.x * q.z - q.y * q.w);
  R[7] = two_s * (q.y * q.z + q.x * q.w);
  R[8] = 1 - two_s * (q.x * q.x + q.y * q.y);
}

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// forward of RotOpt for one key joint: p = R X + T0*clamp(s); P = K p; uv = P.xy / P.z
__device__ __forceinline__ void project_one(const float* R, const float* Km, const float* Tc, const float* X,
                                            float* P, float& u, float& v) {
  const float p0 = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + Tc[0];
  const float p1 = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + Tc[1];
  const float p2 = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + Tc[2];
  P[0] = Km[0] * p0 + Km[1] * p1 + Km[2] * p2;
  P[1] = Km[3] * p0 + Km[4] * p1 + Km[5] * p2;
  P[2] = Km[6] * p0 + Km[7] * p1 + Km[8] * p2;
  u = P[0] / P[2];
  v = P[1] / P[2];
}

// gradient of sum_k <d_uv_k, uv_k> w.r.t. (q, scale); d_uv supplied through a functor
template <class DUV>
__device__ __forceinline__ void rotopt_grad(const Quat& q, float scale, const float* Km, const float* T0, float minT,
                                            float maxT, int nk, const float* sx, int stride, DUV duv, float* dq,
                                            float& dscale) {
  float R[9];
  quat_to_R(q, R);
  const float sc = clampf(scale, minT, maxT);
  const float Tc[3] = {T0[0] * sc, T0[1] * sc, T0[2] * sc};
  float G[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  float gs = 0.f;
  for (int k = 0; k < nk; ++k) {
    const float X[3] = {sx[(k * 5 + 0) * stride], sx[(k * 5 + 1) * stride], sx[(k * 5 + 2) * stride]};
    float P[3], u, v;
    project_one(R, Km, Tc, X, P, u, v);
    float du, dv;
    duv(k, u, v, du, dv);
    const float dP0 = du / P[2], dP1 = dv / P[2];
    const float dP2 = -(du * P[0] + dv * P[1]) / (P[2] * P[2]);
    const float d0 = Km[0] * dP0 + Km[3] * dP1 + Km[6] * dP2;  // K^T dP
    const float d1 = Km[1] * dP0 + Km[4] * dP1 + Km[7] * dP2;
    const float d2 = Km[2] * dP0 + Km[5] * dP1 + Km[8] * dP2;
    G[0] += d0 * X[0]; G[1] += d0 * X[1]; G[2] += d0 * X[2];
    G[3] += d1 * X[0]; G[4] += d1 * X[1]; G[5] += d1 * X[2];
    G[6] += d2 * X[0]; G[7] += d2 * X[1]; G[8] += d2 * X[2];
    gs += d0 * T0[0] + d1 * T0[1] + d2 * T0[2];
  }
  dscale = (scale >= minT && scale <= maxT) ? gs : 0.f;
  const float w = q.w, x = q.x, y = q.y, z = q.z;
  const float n = w * w + x * x + y * y + z * z;
  const float s2 = 2.0f / n;
  const float A[9] = {-(y * y + z * z), x * y - z * w, x * z + y * w, x * y + z * w, -(x * x + z * z),
                      y * z - x * w,    x * z - y * w, y * z + x * w, -(x * x + y * y)};
  float GA = 0.f;
#pragma unroll
  for (int i = 0; i < 9; ++i) GA += G[i] * A[i];
  const float c = -4.0f / (n * n);
  dq[0] = c * w * GA + s2 * (-z * G[1] + y * G[2] + z * G[3] - x * G[5] - y * G[6] + x * G[7]);
  dq[1] = c * x * GA + s2 * (y * G[1] + z * G[2] + y * G[3] - 2 * x * G[4] - w * G[5] + z * G[6] + w * G[7] - 2 * x * G[8]);
  dq[2] = c * y * GA + s2 * (-2 * y * G[0] + x * G[1] + w * G[2] + x * G[3] + z * G[5] - w * G[6] + z * G[7] - 2 * y * G[8]);
  dq[3] = c * z * GA + s2 * (-2 * z * G[0] - w * G[1] + x * G[2] + w * G[3] - 2 * z * G[4] + y * G[5] + x * G[6] + y * G[7]);
}

__device__ __forceinline__ void adam_step(float& p, float g, float& m, float& v, float step_size, float bc2_sqrt,
                                          float b1, float b2, float eps) {
  // torch.optim.Adam (single tensor): lerp on exp_avg, mul+addcmul on exp_avg_sq, addcdiv on the param
  m = m + (g - m) * (1.0f - b1);
  v = v * b2 + (1.0f - b2) * g * g;
  const float denom = sqrtf(v) / bc2_sqrt + eps;
  p = p - step_size * (m / denom);
}

__global__ void __launch_bounds__(kIpoThreads)
ipo_fit_kernel(const float* __restrict__ x0, const float* __restrict__ uv, const float* __restrict__ Kmat,
               const IntList keylist, int axes_mask, int pelvis_a, int pelvis_b, int ray_init,
               float ipo_T, float minT, float maxT, int iters,
               float lam, float lr, float* __restrict__ Rout, float* __restrict__ Tout, float* __restrict__ x_rot,
               float* __restrict__ qs, int64_t B, int J) {
  extern __shared__ float sm[];
  const int nk = keylist.n;
  float* step_size = sm;            // [iters]
  float* bc2_sqrt = sm + iters;     // [iters]
  float* sx = sm + 2 * iters;       // [nk][5][kIpoThreads]: X.x, X.y, X.z, u*, v*
  const int tid = threadIdx.x;
  for (int i = tid; i < iters; i += kIpoThreads) {
    const double bc1 = 1.0 - pow(0.9, (double)(i + 1));
    const double bc2 = 1.0 - pow(0.999, (double)(i + 1));
    step_size[i] = (float)((double)lr / bc1);
    bc2_sqrt[i] = (float)sqrt(bc2);
  }
  const int64_t pose = (int64_t)blockIdx.x * kIpoThreads + tid;
  const bool live = pose < B;
  const int64_t pc = live ? pose : 0;
  for (int k = 0; k < nk; ++k) {
    const int j = keylist.v[k];
    sx[(k * 5 + 0) * kIpoThreads + tid] = x0[(pc * J + j) * 3 + 0];
    sx[(k * 5 + 1) * kIpoThreads + tid] = x0[(pc * J + j) * 3 + 1];
    sx[(k * 5 + 2) * kIpoThreads + tid] = x0[(pc * J + j) * 3 + 2];
    sx[(k * 5 + 3) * kIpoThreads + tid] = uv[(pc * J + j) * 2 + 0];
    sx[(k * 5 + 4) * kIpoThreads + tid] = uv[(pc * J + j) * 2 + 1];
  }
  float Km[9], Ki[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) Km[i] = Kmat[pc * 9 + i];
  inv3x3(Km, Ki);
  // T0 = IPO_T * normalize(K^-1 [u_pelvis, v_pelvis, 1])   (run/opt_main.py:177-179).  The pelvis is joint 0, or
  // the mean of joints 0 and 3 for SyRIP (run/opt_main_infant.py:259-262): (uv[a] + uv[b]) / 2 with a == b allowed.
  const float up = (uv[(pc * J + pelvis_a) * 2 + 0] + uv[(pc * J + pelvis_b) * 2 + 0]) / 2;
  const float vp = (uv[(pc * J + pelvis_a) * 2 + 1] + uv[(pc * J + pelvis_b) * 2 + 1]) / 2;
  float T0[3] = {Ki[0] * up + Ki[1] * vp + Ki[2], Ki[3] * up + Ki[4] * vp + Ki[5], Ki[6] * up + Ki[7] * vp + Ki[8]};
  const float tn = sqrtf(T0[0] * T0[0] + T0[1] * T0[1] + T0[2] * T0[2]);
  T0[0] = T0[0] / tn * ipo_T;
  T0[1] = T0[1] / tn * ipo_T;
  T0[2] = T0[2] / tn * ipo_T;
  __syncthreads();

  Quat q = {1.f, 0.f, 0.f, 0.f};
  float scale = 1.f;
  float mq[4] = {0, 0, 0, 0}, vq[4] = {0, 0, 0, 0}, ms = 0.f, vs = 0.f;
  const float* mysx = sx + tid;
  for (int it = 0; it < iters; ++it) {
    float dq[4], dscale;
    rotopt_grad(q, scale, Km, T0, minT, maxT, nk, mysx, kIpoThreads,
                [&](int k, float u, float v, float& du, float& dv) {
                  const float eu = u - mysx[(k * 5 + 3) * kIpoThreads];
                  const float ev = v - mysx[(k * 5 + 4) * kIpoThreads];
                  du = eu > 0.f ? lam : (eu < 0.f ? -lam : 0.f);  // d|e|/de = sign(e), sign(0) = 0
                  dv = ev > 0.f ? lam : (ev < 0.f ? -lam : 0.f);
                },
                dq, dscale);
    const float ss = step_size[it], bs = bc2_sqrt[it];
    adam_step(q.w, dq[0], mq[0], vq[0], ss, bs, 0.9f, 0.999f, 1e-8f);
    if (axes_mask & 1) adam_step(q.x, dq[1], mq[1], vq[1], ss, bs, 0.9f, 0.999f, 1e-8f);
    if (axes_mask & 2) adam_step(q.y, dq[2], mq[2], vq[2], ss, bs, 0.9f, 0.999f, 1e-8f);
    if (axes_mask & 4) adam_step(q.z, dq[3], mq[3], vq[3], ss, bs, 0.9f, 0.999f, 1e-8f);
    adam_step(scale, dscale, ms, vs, ss, bs, 0.9f, 0.999f, 1e-8f);
  }
  if (!live) return;
  float R[9];
  quat_to_R(q, R);
  const float sc = clampf(scale, minT, maxT);
#pragma unroll
  for (int i = 0; i < 9; ++i) Rout[pose * 9 + i] = R[i];
  Tout[pose * 3 + 0] = T0[0] * sc;
  Tout[pose * 3 + 1] = T0[1] * sc;
  Tout[pose * 3 + 2] = T0[2] * sc;
  if (qs != nullptr) {
    qs[pose * 5 + 0] = q.w;
    qs[pose * 5 + 1] = q.x;
    qs[pose * 5 + 2] = q.y;
    qs[pose * 5 + 3] = q.z;
    qs[pose * 5 + 4] = scale;
  }
  if (x_rot != nullptr && ray_init) {
    // infant driver (run/opt_main_infant.py:281-292,300): the hypothesis is replaced by the back-projected 2D
    // rays, scaled so the pelvis ray has length |T|, pelvis-subtracted, then rotated by the fitted R
    auto ray = [&](int j, float* r) {
      const float u = uv[(pose * J + j) * 2 + 0], v = uv[(pose * J + j) * 2 + 1];
      r[0] = Ki[0] * u + Ki[1] * v + Ki[2];
      r[1] = Ki[3] * u + Ki[4] * v + Ki[5];
      r[2] = Ki[6] * u + Ki[7] * v + Ki[8];
    };
    float ra[3], rb[3];
    ray(pelvis_a, ra);
    ray(pelvis_b, rb);
    const float root[3] = {(ra[0] + rb[0]) / 2, (ra[1] + rb[1]) / 2, (ra[2] + rb[2]) / 2};
    const float rn = sqrtf(root[0] * root[0] + root[1] * root[1] + root[2] * root[2]);
    const float Tn = sqrtf((T0[0] * sc) * (T0[0] * sc) + (T0[1] * sc) * (T0[1] * sc) + (T0[2] * sc) * (T0[2] * sc));
    for (int j = 0; j < J; ++j) {
      float rj[3];
      ray(j, rj);
      const float a = rj[0] / rn * Tn - root[0] / rn * Tn, b = rj[1] / rn * Tn - root[1] / rn * Tn,
                  c = rj[2] / rn * Tn - root[2] / rn * Tn;
      x_rot[(pose * J + j) * 3 + 0] = R[0] * a + R[1] * b + R[2] * c;
      x_rot[(pose * J + j) * 3 + 1] = R[3] * a + R[4] * b + R[5] * c;
      x_rot[(pose * J + j) * 3 + 2] = R[6] * a + R[7] * b + R[8] * c;
    }
  } else if (x_rot != nullptr) {
    for (int j = 0; j < J; ++j) {  // denoise_x = rot_mat.bmm(x^T)^T   (run/opt_main.py:201)
      const float a = x0[(pose * J + j) * 3 + 0], b = x0[(pose * J + j) * 3 + 1], c = x0[(pose * J + j) * 3 + 2];
      x_rot[(pose * J + j) * 3 + 0] = R[0] * a + R[1] * b + R[2] * c;
      x_rot[(pose * J + j) * 3 + 1] = R[3] * a + R[4] * b + R[5] * c;
      x_rot[(pose * J + j) * 3 + 2] = R[6] * a + R[7] * b + R[8] * c;
    }
  }
}

// RotOpt.forward for autograd-driven callers: uv_out [B, nk, 2]
__global__ void rotopt_forward_kernel(const float* __restrict__ q, const float* __restrict__ scale,
                                      const float* __restrict__ xk, const float* __restrict__ T0,
                                      const float* __restrict__ Kmat, float minT, float maxT,
                                      float* __restrict__ uv_out, int64_t B, int nk) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * nk) return;
  const int64_t b = i / nk;
  const Quat qq = {q[b * 4], q[b * 4 + 1], q[b * 4 + 2], q[b * 4 + 3]};
  float R[9], Km[9], P[3], u, v;
  quat_to_R(qq, R);
#pragma unroll
  for (int k = 0; k < 9; ++k) Km[k] = Kmat[b * 9 + k];
  const float sc = clampf(scale[b], minT, maxT);
  const float Tc[3] = {T0[b * 3] * sc, T0[b * 3 + 1] * sc, T0[b * 3 + 2] * sc};
  const float X[3] = {xk[i * 3], xk[i * 3 + 1], xk[i * 3 + 2]};
  project_one(R, Km, Tc, X, P, u, v);
  uv_out[i * 2] = u;
  uv_out[i * 2 + 1] = v;
}

// backward of RotOpt.forward: one thread per pose, key joints read straight from global memory
__global__ void rotopt_backward_kernel(const float* __restrict__ q, const float* __restrict__ scale,
                                       const float* __restrict__ xk, const float* __restrict__ T0,
                                       const float* __restrict__ Kmat, float minT, float maxT,
                                       const float* __restrict__ d_uv, float* __restrict__ d_q,
                                       float* __restrict__ d_scale, int64_t B, int nk) {
  extern __shared__ float sm[];  // [nk][5][blockDim]
  const int tid = threadIdx.x;
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + tid;
  const int64_t bc = b < B ? b : 0;
  for (int k = 0; k < nk; ++k) {
    sm[(k * 5 + 0) * blockDim.x + tid] = xk[(bc * nk + k) * 3 + 0];
    sm[(k * 5 + 1) * blockDim.x + tid] = xk[(bc * nk + k) * 3 + 1];
    sm[(k * 5 + 2) * blockDim.x + tid] = xk[(bc * nk + k) * 3 + 2];
    sm[(k * 5 + 3) * blockDim.x + tid] = d_uv[(bc * nk + k) * 2 + 0];
    sm[(k * 5 + 4) * blockDim.x + tid] = d_uv[(bc * nk + k) * 2 + 1];
  }
  if (b >= B) return;
  const Quat qq = {q[b * 4], q[b * 4 + 1], q[b * 4 + 2], q[b * 4 + 3]};
  float Km[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) Km[k] = Kmat[b * 9 + k];
  const float T[3] = {T0[b * 3], T0[b * 3 + 1], T0[b * 3 + 2]};
  const float* mysm = sm + tid;
  const int stride = blockDim.x;
  float dq[4], ds;
  rotopt_grad(qq, scale[b], Km, T, minT, maxT, nk, mysm, stride,
              [&](int k, float, float, float& du, float& dv) {
                du = mysm[(k * 5 + 3) * stride];
                dv = mysm[(k * 5 + 4) * stride];
              },
              dq, ds);
  d_q[b * 4 + 0] = dq[0];
  d_q[b * 4 + 1] = dq[1];
  d_q[b * 4 + 2] = dq[2];
  d_q[b * 4 + 3] = dq[3];
  d_scale[b] = ds;
}

int launch_ipo_fit(const float* x0, const float* uv, const float* K, const IntList& keylist, int axes_mask,
                   int pelvis_a, int pelvis_b, int ray_init, float ipo_T, float minT, float maxT, int iters, int64_t B_global, float lr, float* R, float* T,
                   float* x_rot, float* qs, int64_t B, int J, cudaStream_t st) {
  if (B == 0) return 0;
  const int nk = keylist.n;
  const size_t smem = (size_t)(2 * iters + nk * 5 * kIpoThreads) * sizeof(float);
  if (smem > 200 * 1024) return ZEDO_E_SHAPE;
  if (smem > 48 * 1024) ZEDO_CUDA_TRY(ensure_max_smem((const void*)ipo_fit_kernel, (int)smem));
  const float lam = (float)(1.0 / ((double)B_global * nk * 2));
  ipo_fit_kernel<<<(unsigned)((B + kIpoThreads - 1) / kIpoThreads), kIpoThreads, smem, st>>>(
      x0, uv, K, keylist, axes_mask, pelvis_a, pelvis_b, ray_init, ipo_T, minT, maxT, iters, lam, lr, R, T,
      x_rot, qs, B, J);
  ZEDO_LAUNCH_CHECK();
  return 0;
}

int launch_rotopt_forward(const float* q, const float* scale, const float* xk, const float* T0, const float* K,
                          float minT, float maxT, float* uv_out, int64_t B, int nk, cudaStream_t st) {
  if (B == 0) return 0;
  const int64_t n = B * nk;
  rotopt_forward_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(q, scale, xk, T0, K, minT, maxT, uv_out, B, nk);
  ZEDO_LAUNCH_CHECK();
  return 0;
}

int launch_rotopt_backward(const float* q, const float* scale, const float* xk, const float* T0, const float* K,
                           float minT, float maxT, const float* d_uv, float* d_q, float* d_scale, int64_t B, int nk,
                           cudaStream_t st) {
  if (B == 0) return 0;
  const int threads = 64;
  const size_t smem = (size_t)nk * 5 * threads * sizeof(float);
  if (smem > 48 * 1024) return ZEDO_E_SHAPE;
  rotopt_backward_kernel<<<(unsigned)((B + threads - 1) / threads), threads, smem, st>>>(q, scale, xk, T0, K, minT,
                                                                                       maxT, d_uv, d_q, d_scale, B, nk);
  ZEDO_LAUNCH_CHECK();
  return 0;
}

}  // namespace zedo
