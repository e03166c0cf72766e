// K3 / K2: per-step reprojection fit and the fused SDE predictor update.
//
// The geometry kernels restate gradient_field_gen (reference simple_zeroshot_opt.py:46-125):
//   rays r_j = K^-1 [u_j, v_j, 1], r_j /= r_j.z                                     (:61-71)
//   optional least-squares translation, rows weighted conf^2, sign flip on T_z < 0   (:73-93)
//   r^_j = r_j/|r_j|; p_j = X_j + T; g_j = (p_j . r^_j) r^_j - p_j                   (:33-36,99,109)
// Per pose they move x (r/w), uv, conf, K, T = 672 bytes at J = 17 and (optionally) emit the first GEMM's fp16
// hi/lo operand in the blocked interleaved layout (common.cuh).  Two general kernels, one arithmetic (a third form,
// for the OIL loop at large batches, evaluates everything that does not depend on the pose once per loop: see
// oil_rays_kernel / oil_geom_kernel below):
//   * grad_field_warp_kernel: one warp per pose, lane j owns joint j (J <= 32) -- lowest latency, small batches;
//   * grad_field_block_kernel: 128 poses per CTA staged in shared memory by coalesced loads, four threads per
//     pose (joints q, q+4, ...) -- ~3.5x fewer instructions per pose, used once the batch fills the GPU.
// Both call the same per-joint functions below, written with explicit-rounding intrinsics (one float32 rounding
// per reference tensor op, no compiler-chosen FMA contraction) and reduce the normal equations in the same order
// (stride-4 partial sums, then (p0+p1)+(p2+p3)), so a pose's result is bit-identical in either kernel, at any
// position in the batch and however the batch is sharded.
#include <cstdlib>
#include <cstring>

#include "kernels.cuh"

namespace zedo {

constexpr int kGeomWarps = 8;                   // warp kernel: poses per CTA
constexpr int kGeomPoses = 128;                 // block kernel: poses per CTA
constexpr int kGeomTpp = 4;                     // block kernel: threads per pose
constexpr int kGeomThreads = kGeomPoses * kGeomTpp;
constexpr int64_t kGeomBlockMinPoses = 32768;   // >= 256 CTAs of 128 poses

// ---- the arithmetic, shared by both kernels ---------------------------------------------------------------
// Euler-Maruyama probability-flow update of one coordinate, one rounding per tensor op of sampling.py:185-191 /
// sde_lib.py:93-100 / utils.py:762-776: score = -eps/std; drift = (-0.5 beta) x - g^2 score; x + drift*dt
__device__ __forceinline__ float em_pf_update(float xv, float e, float neg_half_beta, float g2, float std, float dt) {
  const float score = __fdiv_rn(-e, std);
  const float drift = __fsub_rn(__fmul_rn(neg_half_beta, xv), __fmul_rn(g2, score));
  return __fadd_rn(xv, __fmul_rn(drift, dt));
}

// The same update with the division (-e) / std evaluated through the correctly rounded reciprocal r = RN(1 / std), which
// is uniform over the launch: q0 = RN(-e r), rem = -e - q0 std (exact in one FMA), q = RN(q0 + rem r).  By Markstein's
// theorem q is the correctly rounded quotient whenever nothing under- or overflows on the way, so inside the guarded
// exponent range the result has the bits of __fdiv_rn (checked on 8e7 random pairs); outside it (zeros,
// denormals, huge values) the IEEE division itself runs.  6 instructions instead of the ~12 of the generic sequence.
__device__ __noinline__ float fdiv_rn_rare(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float em_pf_update_rcp(float xv, float e, float neg_half_beta, float g2, float std, float rstd,
                                                  bool std_ok, float dt) {
  const float ne = -e, ae = fabsf(e);
  float score;
  if (std_ok && ae >= 0x1p-60f && ae <= 0x1p60f) {
    const float q0 = __fmul_rn(ne, rstd);
    const float rem = __fmaf_rn(-q0, std, ne);
    score = __fmaf_rn(rem, rstd, q0);
  } else {
    score = fdiv_rn_rare(ne, std);  // out of line: keeps the 16 unrolled call sites of the staging loop small
  }
  const float drift = __fsub_rn(__fmul_rn(neg_half_beta, xv), __fmul_rn(g2, score));
  return __fadd_rn(xv, __fmul_rn(drift, dt));
}

// general 3x3 inverse by the adjugate (skew allowed), every product and sum rounded once
__device__ __forceinline__ float det2_rn(float a, float b, float c, float d) {
  return __fsub_rn(__fmul_rn(a, d), __fmul_rn(b, c));
}
__device__ __forceinline__ void inv3x3_rn(const float* m, float* o) {
  const float A = det2_rn(m[4], m[5], m[7], m[8]), B = -det2_rn(m[3], m[5], m[6], m[8]);
  const float C = det2_rn(m[3], m[4], m[6], m[7]);
  const float det = __fadd_rn(__fadd_rn(__fmul_rn(m[0], A), __fmul_rn(m[1], B)), __fmul_rn(m[2], C));
  const float r = __fdiv_rn(1.0f, det);
  o[0] = __fmul_rn(A, r);
  o[1] = __fmul_rn(-det2_rn(m[1], m[2], m[7], m[8]), r);
  o[2] = __fmul_rn(det2_rn(m[1], m[2], m[4], m[5]), r);
  o[3] = __fmul_rn(B, r);
  o[4] = __fmul_rn(det2_rn(m[0], m[2], m[6], m[8]), r);
  o[5] = __fmul_rn(-det2_rn(m[0], m[2], m[3], m[5]), r);
  o[6] = __fmul_rn(C, r);
  o[7] = __fmul_rn(-det2_rn(m[0], m[1], m[6], m[7]), r);
  o[8] = __fmul_rn(det2_rn(m[0], m[1], m[3], m[4]), r);
}

// K^-1 [u, v, 1] (a 3-term dot product accumulated left to right), then ray / ray.z  (:61-71)
struct Ray {
  float x, y, z;
};
__device__ __forceinline__ Ray back_project(const float* Ki, float u, float v) {
  Ray r;
  r.x = __fadd_rn(__fmaf_rn(Ki[1], v, __fmul_rn(Ki[0], u)), Ki[2]);
  r.y = __fadd_rn(__fmaf_rn(Ki[4], v, __fmul_rn(Ki[3], u)), Ki[5]);
  const float z = __fadd_rn(__fmaf_rn(Ki[7], v, __fmul_rn(Ki[6], u)), Ki[8]);
  r.x = __fdiv_rn(r.x, z);
  r.y = __fdiv_rn(r.y, z);
  r.z = (z != 0.f && fabsf(z) <= 3.402823466e+38f) ? 1.f : __int_as_float(0x7fc00000);  // == z / z
  return r;
}

// One joint's contribution to A^T A and A^T b (:73-90).  The rows of A and b are formed and weighted in float32
// exactly as the reference forms them; the products of A^T A / A^T b are float32 too (its bmm inputs), but the
// sums run in float64: the system is ill-conditioned along the depth axis (cond ~ 1e3) and float64 sums + solve
// return the exact solution of the reference's system, so the kernel differs from the reference only by the
// reference's own float32 rounding.
struct NormalEq {
  double S, Sxz, Syz, Szz, b0, b1, b2;  // M = [[S,0,Sxz],[0,S,Syz],[Sxz,Syz,Szz]], rhs = (b0, b1, b2)
};
__device__ __forceinline__ NormalEq joint_normal_eq(const Ray& r, float X0, float X1, float X2, float w) {
  const float bx = __fmul_rn(__fsub_rn(X0, __fmul_rn(X2, r.x)), w);
  const float by = __fmul_rn(__fsub_rn(X1, __fmul_rn(X2, r.y)), w);
  const float ax = __fmul_rn(r.x, w), ay = __fmul_rn(r.y, w), am = -w;
  NormalEq t;
  t.S = (double)__fmul_rn(am, am);
  t.Sxz = (double)__fmul_rn(am, ax);
  t.Syz = (double)__fmul_rn(am, ay);
  t.Szz = __dadd_rn((double)__fmul_rn(ax, ax), (double)__fmul_rn(ay, ay));
  t.b0 = (double)__fmul_rn(am, bx);
  t.b1 = (double)__fmul_rn(am, by);
  t.b2 = __dadd_rn((double)__fmul_rn(ax, bx), (double)__fmul_rn(ay, by));
  return t;
}

// eliminate the two S rows (Schur complement on the depth axis); T[T_z < 0] *= -1  (:91-93)
__device__ __forceinline__ void solve_translation(const NormalEq& e, float& T0, float& T1, float& T2) {
  const double iS = __ddiv_rn(1.0, e.S);
  const double den = __dsub_rn(e.Szz, __dmul_rn(__dadd_rn(__dmul_rn(e.Sxz, e.Sxz), __dmul_rn(e.Syz, e.Syz)), iS));
  const double num = __dsub_rn(e.b2, __dmul_rn(__dadd_rn(__dmul_rn(e.Sxz, e.b0), __dmul_rn(e.Syz, e.b1)), iS));
  const double tz = __ddiv_rn(num, den);
  T0 = (float)__dmul_rn(__dsub_rn(e.b0, __dmul_rn(e.Sxz, tz)), iS);
  T1 = (float)__dmul_rn(__dsub_rn(e.b1, __dmul_rn(e.Syz, tz)), iS);
  T2 = (float)tz;
  if (T2 < 0.f) {
    T0 = -T0;
    T1 = -T1;
    T2 = -T2;
  }
}

// unit ray r^ = r / |r|  (:33-36)
__device__ __forceinline__ void unit_ray(const Ray& r, float& hx, float& hy, float& hz) {
  const float nrm =
      __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(r.x, r.x), __fmul_rn(r.y, r.y)), __fmul_rn(r.z, r.z)));
  hx = __fdiv_rn(r.x, nrm);
  hy = __fdiv_rn(r.y, nrm);
  hz = __fdiv_rn(r.z, nrm);
}

// p = X + T, g = (p . r^) r^ - p  (:99,109); returns g, and X + g in n
__device__ __forceinline__ void project_on_unit_ray(float hx, float hy, float hz, float X0, float X1, float X2,
                                                    float T0, float T1, float T2, float* g, float* n) {
  const float p0 = __fadd_rn(X0, T0), p1 = __fadd_rn(X1, T1), p2 = __fadd_rn(X2, T2);
  const float d = __fadd_rn(__fadd_rn(__fmul_rn(p0, hx), __fmul_rn(p1, hy)), __fmul_rn(p2, hz));
  g[0] = __fsub_rn(__fmul_rn(d, hx), p0);
  g[1] = __fsub_rn(__fmul_rn(d, hy), p1);
  g[2] = __fsub_rn(__fmul_rn(d, hz), p2);
  n[0] = __fadd_rn(X0, g[0]);
  n[1] = __fadd_rn(X1, g[1]);
  n[2] = __fadd_rn(X2, g[2]);
}

__device__ __forceinline__ void project_on_ray(const Ray& r, float X0, float X1, float X2, float T0, float T1,
                                               float T2, float* g, float* n) {
  float hx, hy, hz;
  unit_ray(r, hx, hy, hz);
  project_on_unit_ray(hx, hy, hz, X0, X1, X2, T0, T1, T2, g, n);
}

__device__ __forceinline__ float clamp_conf(float c) {
  if (c > 1.f) c = 1.f;        // conf[conf > 1] = 1        (:65)
  if (c < 1e-4f) c = 1e-4f;    // conf[conf < 1e-4] = 1e-4  (:66)
  return c;
}

__device__ __forceinline__ double quad_sum_f64(double v) {  // (p0 + p1) + (p2 + p3) in every lane of a quad
  v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}

// ---- one warp per pose ----------------------------------------------------------------------------------------
// lane j holds joint j's addend (zero beyond J).  Lanes 0..3 accumulate the stride-4 partial sums left to right
// -- the order in which a thread of the block kernel walks its joints -- then the quad sum; result in every lane.
__device__ __forceinline__ double warp_sum_quad_order(double v, int J) {
  double acc = v;
  for (int o = kGeomTpp; o < J; o += kGeomTpp) acc = __dadd_rn(acc, __shfl_down_sync(0xffffffffu, v, o));
  acc = quad_sum_f64(acc);
  return __shfl_sync(0xffffffffu, acc, 0);
}

__global__ void __launch_bounds__(kGeomWarps * 32)
grad_field_warp_kernel(const float* __restrict__ uv, const float* x, const float* __restrict__ Kmat, float* conf,
                       float* T, int solve_T, int clamp_inplace, float* g_out, float* x_out,
                       __half* __restrict__ xa, int64_t B, int J, const float* __restrict__ eps_prev,
                       float neg_half_beta, float gsq, float std, float dt, float* __restrict__ dump) {
  __shared__ float stage[kGeomWarps][kXaCols];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int64_t pose = (int64_t)blockIdx.x * kGeomWarps + wib;
  griddep_wait();  // PDL: x / eps / T come from the previous kernels
  if (pose >= B) return;
  const bool active = lane < J;

  // intrinsics: lanes 0..8 load one entry each, broadcast, invert
  float kv = lane < 9 ? Kmat[pose * 9 + lane] : 0.f;
  float Km[9], Ki[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) Km[i] = __shfl_sync(0xffffffffu, kv, i);
  inv3x3_rn(Km, Ki);

  float u = 0.f, v = 0.f, X0 = 0.f, X1 = 0.f, X2 = 0.f, c = 1.f;
  if (active) {
    const float2 p2 = *reinterpret_cast<const float2*>(uv + (pose * J + lane) * 2);
    u = p2.x;
    v = p2.y;
    const float* xp = x + (pose * J + lane) * 3;
    X0 = xp[0];
    X1 = xp[1];
    X2 = xp[2];
    if (eps_prev != nullptr) {
      // fused tail of the previous OIL step: the predictor update with that step's network output
      const float* ep = eps_prev + pose * 64 + lane * 3;
      X0 = em_pf_update(X0, ep[0], neg_half_beta, gsq, std, dt);
      X1 = em_pf_update(X1, ep[1], neg_half_beta, gsq, std, dt);
      X2 = em_pf_update(X2, ep[2], neg_half_beta, gsq, std, dt);
      if (dump != nullptr) {
        float* dp = dump + (pose * J + lane) * 3;
        dp[0] = X0;
        dp[1] = X1;
        dp[2] = X2;
      }
    }
    if (conf != nullptr) {
      c = clamp_conf(conf[pose * J + lane]);
      if (clamp_inplace) conf[pose * J + lane] = c;
    }
  }
  const Ray ray = back_project(Ki, u, v);

  float T0, T1, T2;
  if (solve_T) {
    // row weight conf*conf on A and on b (:85-88); an idle lane contributes exact zeros
    NormalEq e = joint_normal_eq(ray, X0, X1, X2, active ? __fmul_rn(c, c) : 0.f);
    e.S = warp_sum_quad_order(e.S, J);
    e.Sxz = warp_sum_quad_order(e.Sxz, J);
    e.Syz = warp_sum_quad_order(e.Syz, J);
    e.Szz = warp_sum_quad_order(e.Szz, J);
    e.b0 = warp_sum_quad_order(e.b0, J);
    e.b1 = warp_sum_quad_order(e.b1, J);
    e.b2 = warp_sum_quad_order(e.b2, J);
    solve_translation(e, T0, T1, T2);
    if (lane == 0) {
      T[pose * 3 + 0] = T0;
      T[pose * 3 + 1] = T1;
      T[pose * 3 + 2] = T2;
    }
  } else {
    float tv = lane < 3 ? T[pose * 3 + lane] : 0.f;
    T0 = __shfl_sync(0xffffffffu, tv, 0);
    T1 = __shfl_sync(0xffffffffu, tv, 1);
    T2 = __shfl_sync(0xffffffffu, tv, 2);
  }

  float g[3], n[3];
  project_on_ray(ray, X0, X1, X2, T0, T1, T2, g, n);
  if (active) {
    const int64_t o = (pose * J + lane) * 3;
    if (g_out != nullptr) {
      g_out[o] = g[0];
      g_out[o + 1] = g[1];
      g_out[o + 2] = g[2];
    }
    if (x_out != nullptr) {
      x_out[o] = n[0];
      x_out[o + 1] = n[1];
      x_out[o + 2] = n[2];
    }
  }
  if (xa != nullptr) {
    // emit row `pose` of the first GEMM's A operand: 64 halves (3J padded with zeros), hi and lo
    float* st = stage[wib];
    st[lane] = 0.f;
    st[lane + 32] = 0.f;
    __syncwarp();
    if (active) {
      st[lane * 3 + 0] = n[0];
      st[lane * 3 + 1] = n[1];
      st[lane * 3 + 2] = n[2];
    }
    __syncwarp();
    if (lane < 8) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        __half h0, l0, h1, l1;
        split_hi_lo(st[lane * 8 + 2 * e], h0, l0);
        split_hi_lo(st[lane * 8 + 2 * e + 1], h1, l1);
        hi[e] = pack_half2(h0, h1);
        lo[e] = pack_half2(l0, l1);
      }
      const int64_t oh = blocked_half_offset(pose, lane * 8, kXaCols, kActTileRows, 0);
      const int64_t ol = blocked_half_offset(pose, lane * 8, kXaCols, kActTileRows, 1);
      *reinterpret_cast<uint4*>(xa + oh) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(xa + ol) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

// ---- 128 poses per CTA, four threads per pose -------------------------------------------------------------
// elements i = tid, tid + 512, ... of an [n, len] row-major block: f(i, row, col) without a division per element
template <class F>
__device__ __forceinline__ void for_block_elems(int n, int len, F f) {
  int row = (int)threadIdx.x / len, col = (int)threadIdx.x - row * len;
  const int drow = kGeomThreads / len, dcol = kGeomThreads - drow * len;
#pragma unroll 4  // several independent global loads in flight per thread
  for (int i = threadIdx.x; i < n * len; i += kGeomThreads) {
    f(i, row, col);
    row += drow;
    col += dcol;
    if (col >= len) {
      col -= len;
      ++row;
    }
  }
}

__host__ __device__ inline int geom_smem_floats(int J) {
  const int Jp = J | 1, Dp = (3 * J) | 1;  // odd strides keep the per-pose reads spread over the banks
  return kGeomPoses * (2 * Jp + Dp + Jp + 9 + 3);
}

__global__ void __launch_bounds__(kGeomThreads)
grad_field_block_kernel(const float* __restrict__ uv, const float* x, const float* __restrict__ Kmat, float* conf,
                        float* T, int solve_T, int clamp_inplace, float* g_out, float* x_out,
                        __half* __restrict__ xa, int64_t B, int J, const float* __restrict__ eps_prev,
                        float neg_half_beta, float gsq, float std, float dt, float* __restrict__ dump) {
  extern __shared__ __align__(16) float geom_smem[];
  const int D = 3 * J, Jp = J | 1, Dp = D | 1;
  float2* us = reinterpret_cast<float2*>(geom_smem);  // [P][Jp] (u, v)
  float* xs = geom_smem + kGeomPoses * 2 * Jp;         // [P][Dp] pose (updated in place)
  float* cs = xs + kGeomPoses * Dp;                    // [P][Jp] clamped confidence
  float* ks = cs + kGeomPoses * Jp;                    // [P][9]
  float* ts = ks + kGeomPoses * 9;                     // [P][3]
  const int64_t p0 = (int64_t)blockIdx.x * kGeomPoses;
  const int n = (int)((B - p0) < kGeomPoses ? (B - p0) : kGeomPoses);
  const int tid = threadIdx.x;

  griddep_wait();  // PDL: x / eps / T come from the previous kernels
  const float2* uv2 = reinterpret_cast<const float2*>(uv) + p0 * J;
  for_block_elems(n, J, [&](int i, int r, int c) { us[r * Jp + c] = uv2[i]; });
  for (int i = tid; i < n * 9; i += kGeomThreads) ks[i] = Kmat[p0 * 9 + i];
  if (conf != nullptr) {
    float* cg = conf + p0 * J;
    for_block_elems(n, J, [&](int i, int r, int c) {
      const float cv = clamp_conf(cg[i]);
      if (clamp_inplace) cg[i] = cv;
      cs[r * Jp + c] = cv;
    });
  }
  {
    const float* xg = x + p0 * D;
    float* dg = dump != nullptr ? dump + p0 * D : nullptr;
    const float* eg = eps_prev != nullptr ? eps_prev + p0 * 64 : nullptr;
    for_block_elems(n, D, [&](int i, int r, int c) {
      float xv = xg[i];
      if (eg != nullptr) {
        // fused tail of the previous OIL step: the predictor update with that step's network output
        xv = em_pf_update(xv, eg[r * 64 + c], neg_half_beta, gsq, std, dt);
        if (dg != nullptr) dg[i] = xv;
      }
      xs[r * Dp + c] = xv;
    });
  }
  if (!solve_T)
    for (int i = tid; i < n * 3; i += kGeomThreads) ts[i] = T[p0 * 3 + i];
  __syncthreads();

  if (((tid & ~31) >> 2) < n) {
    // per-pose phase (warps with a live pose): thread (pose, q) owns joints q, q + 4, ...; the idle quads of a
    // ragged last warp recompute pose n - 1 (the warp stays converged for the shuffles) and discard the result
    const int pl = tid >> 2, q = tid & 3;
    const bool live = pl < n;
    const int pp = live ? pl : n - 1;
    float Ki[9];
    inv3x3_rn(ks + pp * 9, Ki);
    float* xp = xs + pp * Dp;
    const float2* up = us + pp * Jp;
    float T0, T1, T2;
    if (solve_T) {
      NormalEq e{0, 0, 0, 0, 0, 0, 0};
#pragma unroll 4
      for (int j = q; j < J; j += kGeomTpp) {
        const float2 p2 = up[j];
        const float c = conf != nullptr ? cs[pp * Jp + j] : 1.f;
        const NormalEq t = joint_normal_eq(back_project(Ki, p2.x, p2.y), xp[3 * j], xp[3 * j + 1], xp[3 * j + 2],
                                           __fmul_rn(c, c));
        e.S = __dadd_rn(e.S, t.S);
        e.Sxz = __dadd_rn(e.Sxz, t.Sxz);
        e.Syz = __dadd_rn(e.Syz, t.Syz);
        e.Szz = __dadd_rn(e.Szz, t.Szz);
        e.b0 = __dadd_rn(e.b0, t.b0);
        e.b1 = __dadd_rn(e.b1, t.b1);
        e.b2 = __dadd_rn(e.b2, t.b2);
      }
      e.S = quad_sum_f64(e.S);
      e.Sxz = quad_sum_f64(e.Sxz);
      e.Syz = quad_sum_f64(e.Syz);
      e.Szz = quad_sum_f64(e.Szz);
      e.b0 = quad_sum_f64(e.b0);
      e.b1 = quad_sum_f64(e.b1);
      e.b2 = quad_sum_f64(e.b2);
      solve_translation(e, T0, T1, T2);
      if (live && q == 0) {
        ts[pl * 3 + 0] = T0;
        ts[pl * 3 + 1] = T1;
        ts[pl * 3 + 2] = T2;
      }
    } else {
      T0 = ts[pp * 3 + 0];
      T1 = ts[pp * 3 + 1];
      T2 = ts[pp * 3 + 2];
    }
    if (live) {
      float* gp = g_out != nullptr ? g_out + (p0 + pl) * D : nullptr;
#pragma unroll 4
      for (int j = q; j < J; j += kGeomTpp) {
        const float2 p2 = up[j];
        float g[3], nx[3];
        project_on_ray(back_project(Ki, p2.x, p2.y), xp[3 * j], xp[3 * j + 1], xp[3 * j + 2], T0, T1, T2, g, nx);
        if (gp != nullptr) {
          gp[3 * j] = g[0];
          gp[3 * j + 1] = g[1];
          gp[3 * j + 2] = g[2];
        }
        xp[3 * j] = nx[0];
        xp[3 * j + 1] = nx[1];
        xp[3 * j + 2] = nx[2];
      }
    }
  }
  __syncthreads();

  if (x_out != nullptr) {
    float* og = x_out + p0 * D;
    for_block_elems(n, D, [&](int i, int r, int c) { og[i] = xs[r * Dp + c]; });
  }
  if (solve_T)
    for (int i = tid; i < n * 3; i += kGeomThreads) T[p0 * 3 + i] = ts[i];
  if (xa != nullptr) {
    // rows p0 .. p0 + n of the first GEMM's A operand: 64 halves per row (3J padded with zeros), hi and lo;
    // consecutive threads write consecutive 16-byte chunks of the blocked layout
    for (int item = tid; item < kGeomPoses * (kXaCols / 8); item += kGeomThreads) {
      const int r = item & (kGeomPoses - 1), ch = item / kGeomPoses;
      if (r >= n) continue;
      const float* xp = xs + r * Dp;
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c0 = ch * 8 + 2 * e;
        const float a = c0 < D ? xp[c0] : 0.f;
        const float b = c0 + 1 < D ? xp[c0 + 1] : 0.f;
        __half h0, l0, h1, l1;
        split_hi_lo(a, h0, l0);
        split_hi_lo(b, h1, l1);
        hi[e] = pack_half2(h0, h1);
        lo[e] = pack_half2(l0, l1);
      }
      *reinterpret_cast<uint4*>(xa + blocked_half_offset(p0 + r, ch * 8, kXaCols, kActTileRows, 0)) =
          make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(xa + blocked_half_offset(p0 + r, ch * 8, kXaCols, kActTileRows, 1)) =
          make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

// ---- the OIL loop's geometry step on precomputed rays ----------------------------------------------------------
// uv, K and (after the first call's in-place clamp) conf do not change during zedo_oil_loop, so everything of
// gradient_field_gen that depends on them alone is evaluated ONCE per loop by oil_rays_kernel -- with the very functions
// above, in the summation order of the kernels above, hence bit-identical -- and the per-step kernel is left with the
// part that depends on the pose: b = A^T (w b_rows) of the normal equations, the back-substitution and the projection.
// Per step and pose that removes 7 IEEE divisions and a square root per joint, the K inverse and four of the seven
// float64 reductions (~3x fewer instructions); the price is 24 B per joint slot of ray data instead of 8 B of (u, v).
//
// Layout (one entry per lane of the step kernel's per-pose phase, so a warp's loads are 512 / 256 contiguous bytes):
//   slot(pose, it, q) = ((pose / 8) * n_it + it) * 32 + (pose % 8) * 4 + q      joint j = q + 4 it, n_it = ceil(J / 4)
//   rays_a[slot] = (r^_x, r^_y, r^_z, w = conf^2)      rays_b[slot] = (r_x, r_y)      zeros for j >= J
//   pose_c[pose] = (1/S, den, Sxz, Syz)  float64: the pose-independent part of solve_translation
constexpr int kRayPoses = 64;                  // poses per CTA of the step kernel
constexpr int kRayThreads = kRayPoses * kGeomTpp;

__host__ __device__ inline int ray_iters(int J) { return (J + kGeomTpp - 1) / kGeomTpp; }

__global__ void __launch_bounds__(kRayThreads)
oil_rays_kernel(const float* __restrict__ uv, const float* __restrict__ Kmat, float* conf, float4* __restrict__ rays_a,
                float2* __restrict__ rays_b, double* __restrict__ pose_c, int64_t B, int J) {
  const int64_t p0 = (int64_t)blockIdx.x * kRayPoses;
  const int tid = threadIdx.x, pl = tid >> 2, q = tid & 3;
  const int64_t pose = p0 + pl;  // B_pad8 rows are covered by the grid; rows >= B get zero entries
  const int n_it = ray_iters(J);
  const bool live = pose < B;
  const int64_t pp = live ? pose : B - 1;  // idle quads recompute the last pose (converged shuffles), results dropped
  float Km[9], Ki[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) Km[i] = Kmat[pp * 9 + i];
  inv3x3_rn(Km, Ki);
  NormalEq e{0, 0, 0, 0, 0, 0, 0};
  const int64_t slot0 = ((pose >> 3) * n_it) * 32 + (pose & 7) * 4 + q;
  for (int it = 0; it < n_it; ++it) {
    const int j = q + kGeomTpp * it;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    float2 b = make_float2(0.f, 0.f);
    if (j < J) {
      const float2 p2 = *reinterpret_cast<const float2*>(uv + (pp * J + j) * 2);
      float c = 1.f;
      if (conf != nullptr) {
        c = clamp_conf(conf[pp * J + j]);
        if (live) conf[pp * J + j] = c;  // the reference clamps in place at its first call (:65-66)
      }
      const Ray r = back_project(Ki, p2.x, p2.y);
      const float w = __fmul_rn(c, c);
      unit_ray(r, a.x, a.y, a.z);
      a.w = w;
      b = make_float2(r.x, r.y);
      // the X-independent sums of joint_normal_eq, same products, same order
      const float ax = __fmul_rn(r.x, w), ay = __fmul_rn(r.y, w), am = -w;
      e.S = __dadd_rn(e.S, (double)__fmul_rn(am, am));
      e.Sxz = __dadd_rn(e.Sxz, (double)__fmul_rn(am, ax));
      e.Syz = __dadd_rn(e.Syz, (double)__fmul_rn(am, ay));
      e.Szz = __dadd_rn(e.Szz, __dadd_rn((double)__fmul_rn(ax, ax), (double)__fmul_rn(ay, ay)));
    }
    if (live) {
      rays_a[slot0 + (int64_t)it * 32] = a;
      rays_b[slot0 + (int64_t)it * 32] = b;
    } else {
      rays_a[slot0 + (int64_t)it * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
      rays_b[slot0 + (int64_t)it * 32] = make_float2(0.f, 0.f);
    }
  }
  e.S = quad_sum_f64(e.S);
  e.Sxz = quad_sum_f64(e.Sxz);
  e.Syz = quad_sum_f64(e.Syz);
  e.Szz = quad_sum_f64(e.Szz);
  if (live && q == 0) {
    const double iS = __ddiv_rn(1.0, e.S);
    const double den = __dsub_rn(e.Szz, __dmul_rn(__dadd_rn(__dmul_rn(e.Sxz, e.Sxz), __dmul_rn(e.Syz, e.Syz)), iS));
    double* pc = pose_c + pose * 4;
    pc[0] = iS;
    pc[1] = den;
    pc[2] = e.Sxz;
    pc[3] = e.Syz;
  }
}

// One OIL step's geometry for kRayPoses poses per CTA: fused predictor update of the previous step (as in the kernels
// above), then -- four threads per pose, thread q owns joints q, q + 4, ... -- the translation solve (SOLVE) and the
// projection on the precomputed unit rays; x is updated in place and the first GEMM's operand emitted.
// JCT = the number of joints as a compile-time constant (17, 12; 0 = read it from the argument): row lengths, strides and
// trip counts fold, which removes a fifth of the kernel's (issue-bound) instructions.
template <bool SOLVE, int JCT>
__global__ void __launch_bounds__(kRayThreads, 4)
oil_geom_kernel(const float4* __restrict__ rays_a, const float2* __restrict__ rays_b,
                const double* __restrict__ pose_c, float* x, float* T, __half* __restrict__ xa, int64_t B, int J_rt,
                const float* __restrict__ eps_prev, float neg_half_beta, float gsq, float std, float dt,
                float* __restrict__ dump) {
  extern __shared__ __align__(16) float geom_smem[];
  constexpr int NIT = JCT ? (JCT + kGeomTpp - 1) / kGeomTpp : 0;
  const int J = JCT ? JCT : J_rt;
  const int D = 3 * J, Dp = D | 1;
  float* xs = geom_smem;               // [P][Dp]
  float* ts = xs + kRayPoses * Dp;     // [P][3]
  const int64_t p0 = (int64_t)blockIdx.x * kRayPoses;
  const int n = (int)((B - p0) < kRayPoses ? (B - p0) : kRayPoses);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int kWarps = kRayThreads / 32;
  const int n_it = NIT ? NIT : ray_iters(J);

  griddep_wait();  // PDL: x / eps / T come from the previous kernels
  // stage the poses, one row per warp iteration (coalesced), applying the previous step's predictor update.  All
  // loads of a warp's rows are issued before the first one is consumed: the phase is bound by global-load latency
  // (r02 ncu: 43 % of the stall samples sat on the first use of eps), not by bandwidth or issue slots.
  {
    constexpr int kRows = kRayPoses / kWarps;  // rows per warp
    const int c1 = lane + 32;                  // D <= 64: at most two columns per lane
    const float rstd = __frcp_rn(std);
    const bool std_ok = fabsf(std) >= 0x1p-60f && fabsf(std) <= 0x1p60f;
    float xv[kRows][2], ev[kRows][2];
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
      const int r = warp + k * kWarps;
      const float* xg = x + (p0 + r) * D;
      const float* eg = eps_prev + (p0 + r) * 64;
      const bool ok = r < n;
      xv[k][0] = ok && lane < D ? xg[lane] : 0.f;
      xv[k][1] = ok && c1 < D ? xg[c1] : 0.f;
      ev[k][0] = ok && eps_prev != nullptr && lane < D ? eg[lane] : 0.f;
      ev[k][1] = ok && eps_prev != nullptr && c1 < D ? eg[c1] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
      const int r = warp + k * kWarps;
      if (r >= n) break;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = lane + 32 * h;
        if (c < D) {
          float v = xv[k][h];
          if (eps_prev != nullptr) {
            v = em_pf_update_rcp(v, ev[k][h], neg_half_beta, gsq, std, rstd, std_ok, dt);
            if (dump != nullptr) dump[(p0 + r) * D + c] = v;
          }
          xs[r * Dp + c] = v;
        }
      }
    }
  }
  if (!SOLVE)
    for (int i = tid; i < n * 3; i += kRayThreads) ts[i] = T[p0 * 3 + i];
  __syncthreads();

  if (((tid & ~31) >> 2) < n) {
    const int pl = tid >> 2, q = tid & 3;
    const bool live = pl < n;
    const int pp = live ? pl : n - 1;
    const int64_t slot0 = (((p0 + pp) >> 3) * n_it) * 32 + (pp & 7) * 4 + q;
    const float4* ra = rays_a + slot0;
    float* xp = xs + pp * Dp;
    constexpr int kMaxIt = NIT ? NIT : 6;  // 3 J <= 64 -> J <= 21
    float4 hw[kMaxIt];
#pragma unroll
    for (int it = 0; it < kMaxIt; ++it)
      hw[it] = it < n_it ? ra[it * 32] : make_float4(0.f, 0.f, 0.f, 0.f);
    float T0, T1, T2;
    if (SOLVE) {
      const float2* rb = rays_b + slot0;
      float2 rxy[kMaxIt];
#pragma unroll
      for (int it = 0; it < kMaxIt; ++it) rxy[it] = it < n_it ? rb[it * 32] : make_float2(0.f, 0.f);
      const double* pc = pose_c + (p0 + pp) * 4;
      const double iS = pc[0], den = pc[1], Sxz = pc[2], Syz = pc[3];
      double b0 = 0.0, b1 = 0.0, b2 = 0.0;
#pragma unroll
      for (int it = 0; it < kMaxIt; ++it) {
        const int j = q + kGeomTpp * it;
        if (it < n_it && j < J) {
          // the X-dependent part of joint_normal_eq, same products, same order
          const float X0 = xp[3 * j], X1 = xp[3 * j + 1], X2 = xp[3 * j + 2];
          const float w = hw[it].w, rx = rxy[it].x, ry = rxy[it].y;
          const float bx = __fmul_rn(__fsub_rn(X0, __fmul_rn(X2, rx)), w);
          const float by = __fmul_rn(__fsub_rn(X1, __fmul_rn(X2, ry)), w);
          const float ax = __fmul_rn(rx, w), ay = __fmul_rn(ry, w), am = -w;
          b0 = __dadd_rn(b0, (double)__fmul_rn(am, bx));
          b1 = __dadd_rn(b1, (double)__fmul_rn(am, by));
          b2 = __dadd_rn(b2, __dadd_rn((double)__fmul_rn(ax, bx), (double)__fmul_rn(ay, by)));
        }
      }
      b0 = quad_sum_f64(b0);
      b1 = quad_sum_f64(b1);
      b2 = quad_sum_f64(b2);
      // solve_translation with the pose-independent factors precomputed
      const double num = __dsub_rn(b2, __dmul_rn(__dadd_rn(__dmul_rn(Sxz, b0), __dmul_rn(Syz, b1)), iS));
      const double tz = __ddiv_rn(num, den);
      T0 = (float)__dmul_rn(__dsub_rn(b0, __dmul_rn(Sxz, tz)), iS);
      T1 = (float)__dmul_rn(__dsub_rn(b1, __dmul_rn(Syz, tz)), iS);
      T2 = (float)tz;
      if (T2 < 0.f) {
        T0 = -T0;
        T1 = -T1;
        T2 = -T2;
      }
      if (live && q == 0) {
        ts[pl * 3 + 0] = T0;
        ts[pl * 3 + 1] = T1;
        ts[pl * 3 + 2] = T2;
      }
    } else {
      T0 = ts[pp * 3 + 0];
      T1 = ts[pp * 3 + 1];
      T2 = ts[pp * 3 + 2];
    }
    if (live) {
#pragma unroll
      for (int it = 0; it < kMaxIt; ++it) {
        const int j = q + kGeomTpp * it;
        if (it < n_it && j < J) {
          float g[3], nx[3];
          project_on_unit_ray(hw[it].x, hw[it].y, hw[it].z, xp[3 * j], xp[3 * j + 1], xp[3 * j + 2], T0, T1, T2, g,
                              nx);
          xp[3 * j] = nx[0];
          xp[3 * j + 1] = nx[1];
          xp[3 * j + 2] = nx[2];
        }
      }
    }
  }
  __syncthreads();

#pragma unroll 2
  for (int r = warp; r < n; r += kWarps) {
    float* og = x + (p0 + r) * D;
    for (int c = lane; c < D; c += 32) og[c] = xs[r * Dp + c];
  }
  if (SOLVE)
    for (int i = tid; i < n * 3; i += kRayThreads) T[p0 * 3 + i] = ts[i];
  if (xa != nullptr) {
    // rows p0 .. p0 + n of the first GEMM's A operand: 64 halves per row (3J padded with zeros), hi and lo;
    // consecutive threads write consecutive 16-byte chunks of the blocked layout
    for (int item = tid; item < kRayPoses * (kXaCols / 8); item += kRayThreads) {
      const int r = item & (kRayPoses - 1), ch = item / kRayPoses;
      if (r >= n) continue;
      const float* xp = xs + r * Dp;
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c0 = ch * 8 + 2 * e;
        const float a = c0 < D ? xp[c0] : 0.f;
        const float b = c0 + 1 < D ? xp[c0 + 1] : 0.f;
        __half h0, l0, h1, l1;
        split_hi_lo(a, h0, l0);
        split_hi_lo(b, h1, l1);
        hi[e] = pack_half2(h0, h1);
        lo[e] = pack_half2(l0, l1);
      }
      *reinterpret_cast<uint4*>(xa + blocked_half_offset(p0 + r, ch * 8, kXaCols, kActTileRows, 0)) =
          make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(xa + blocked_half_offset(p0 + r, ch * 8, kXaCols, kActTileRows, 1)) =
          make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

// x [B, D] float32 -> blocked hi/lo A operand of the first GEMM (one k-block of 64 columns)
__global__ void pack_x_kernel(const float* __restrict__ x, __half* __restrict__ xa, int64_t B, int D) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t row = t >> 3;
  const int chunk = (int)(t & 7);
  griddep_wait();
  if (row >= B) return;
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c0 = chunk * 8 + 2 * e;
    const float a = c0 < D ? x[row * D + c0] : 0.f;
    const float b = c0 + 1 < D ? x[row * D + c0 + 1] : 0.f;
    __half h0, l0, h1, l1;
    split_hi_lo(a, h0, l0);
    split_hi_lo(b, h1, l1);
    hi[e] = pack_half2(h0, h1);
    lo[e] = pack_half2(l0, l1);
  }
  *reinterpret_cast<uint4*>(xa + blocked_half_offset(row, chunk * 8, kXaCols, kActTileRows, 0)) =
      make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(xa + blocked_half_offset(row, chunk * 8, kXaCols, kActTileRows, 1)) =
      make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// K2: the predictor update, elementwise over [B, D]; eps is the network output with row
// stride ld_eps.  Float32 op order of sampling.py:185-191 / sde_lib.py:93-107 / utils.py:762-776:
//   score = -eps/std
//   Euler-Maruyama:     drift = (-0.5 beta) x - g^2 score ; x_mean = x + drift*dt ; x = x_mean + (g sqrt(-dt)) z
//   reverse diffusion:  f = (-0.5 beta) x * dt' ; rev_f = f - G^2 score ; x_mean = x - rev_f ; x = x_mean + G z
__global__ void sde_update_kernel(const float* __restrict__ x, const float* __restrict__ eps, int ld_eps,
                                  const float* __restrict__ z, float neg_half_beta, float g2, float std,
                                  float dt, float noise_scale, int predictor, float* x_next, float* x_mean,
                                  int64_t B, int D) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  griddep_wait();
  if (idx >= B * D) return;
  const int64_t row = idx / D;
  const int e = (int)(idx - row * D);
  const float xv = x[idx];
  const float score = -eps[row * ld_eps + e] / std;
  float xm;
  if (predictor == ZEDO_PRED_EULER_MARUYAMA) {
    xm = em_pf_update(xv, eps[row * ld_eps + e], neg_half_beta, g2, std, dt);
  } else {
    const float f = neg_half_beta * xv * dt;
    const float rev_f = f - g2 * score;
    xm = xv - rev_f;
  }
  if (x_mean != nullptr) x_mean[idx] = xm;
  if (x_next != nullptr) {
    float xn = xm;
    if (z != nullptr && noise_scale != 0.f) xn = xm + noise_scale * z[idx];
    x_next[idx] = xn;
  }
}

// ---- host launchers ---------------------------------------------------------------------------------

int launch_grad_field(const float* uv, const float* x, const float* K, float* conf, float* T, int solve_T,
                      int clamp_inplace, float* g, float* x_out, __half* xa, int64_t B, int J, cudaStream_t st,
                      const float* eps_prev, const SdeCoef* prev, float* dump) {
  if (B == 0) return 0;
  float nhb = 0.f, g2 = 0.f, sd = 1.f, dt = 0.f;
  if (eps_prev != nullptr && prev != nullptr) {
    nhb = -0.5f * prev->beta_t;
    g2 = prev->diffusion * prev->diffusion;
    sd = prev->std;
    dt = prev->dt;
  } else {
    eps_prev = nullptr;
  }
  // ZEDO_OPT_GEOM_KERNEL forces one kernel (tests); default: 128-pose CTAs once the batch fills the GPU
  const int forced = option_get(ZEDO_OPT_GEOM_KERNEL);
  if (forced >= 2 || (forced == 0 && B >= kGeomBlockMinPoses)) {
    const size_t smem = (size_t)geom_smem_floats(J) * sizeof(float);
    if (smem > 48 * 1024) ZEDO_CUDA_TRY(ensure_max_smem((const void*)grad_field_block_kernel, (int)smem));
    ZEDO_CUDA_TRY(launch_pdl(grad_field_block_kernel, dim3((unsigned)((B + kGeomPoses - 1) / kGeomPoses)),
                             dim3(kGeomThreads), smem, st, uv, x, K, conf, T, solve_T, clamp_inplace, g, x_out, xa, B,
                             J, eps_prev, nhb, g2, sd, dt, dump));
  } else {
    ZEDO_CUDA_TRY(launch_pdl(grad_field_warp_kernel, dim3((unsigned)((B + kGeomWarps - 1) / kGeomWarps)),
                             dim3(kGeomWarps * 32), 0, st, uv, x, K, conf, T, solve_T, clamp_inplace, g, x_out, xa, B,
                             J, eps_prev, nhb, g2, sd, dt, dump));
  }
  ZEDO_LAUNCH_CHECK();
  return 0;
}

// ---- OIL loop on precomputed rays ----
bool oil_rays_selected(int64_t B) {
  const int forced = option_get(ZEDO_OPT_GEOM_KERNEL);
  return forced == 3 || (forced == 0 && B >= kGeomBlockMinPoses);
}

size_t oil_rays_slots(int64_t rows, int J) { return (size_t)((rows + 7) / 8) * ray_iters(J) * 32; }

int launch_oil_rays(const float* uv, const float* K, float* conf, float4* rays_a, float2* rays_b, double* pose_c,
                    int64_t B, int J, cudaStream_t st) {
  if (B == 0) return 0;
  ZEDO_CUDA_TRY(launch_pdl(oil_rays_kernel, dim3((unsigned)((B + kRayPoses - 1) / kRayPoses)), dim3(kRayThreads), 0,
                           st, uv, K, conf, rays_a, rays_b, pose_c, B, J));
  ZEDO_LAUNCH_CHECK();
  return 0;
}

template <bool SOLVE, int JCT>
static cudaError_t launch_oil_geom_t(const float4* rays_a, const float2* rays_b, const double* pose_c, float* x,
                                     float* T, __half* xa, int64_t B, int J, cudaStream_t st, const float* eps_prev,
                                     float nhb, float g2, float sd, float dt, float* dump) {
  const size_t smem = (size_t)kRayPoses * (((3 * J) | 1) + 3) * sizeof(float);
  return launch_pdl(oil_geom_kernel<SOLVE, JCT>, dim3((unsigned)((B + kRayPoses - 1) / kRayPoses)), dim3(kRayThreads),
                    smem, st, rays_a, rays_b, pose_c, x, T, xa, B, J, eps_prev, nhb, g2, sd, dt, dump);
}

int launch_oil_geom(const float4* rays_a, const float2* rays_b, const double* pose_c, float* x, float* T, int solve_T,
                    __half* xa, int64_t B, int J, cudaStream_t st, const float* eps_prev, const SdeCoef* prev,
                    float* dump) {
  if (B == 0) return 0;
  if (J < 1 || 3 * J > kXaCols) return ZEDO_E_SHAPE;  // the staging phase holds a row in two columns per lane
  float nhb = 0.f, g2 = 0.f, sd = 1.f, dt = 0.f;
  if (eps_prev != nullptr && prev != nullptr) {
    nhb = -0.5f * prev->beta_t;
    g2 = prev->diffusion * prev->diffusion;
    sd = prev->std;
    dt = prev->dt;
  } else {
    eps_prev = nullptr;
  }
  cudaError_t e;
#define ZEDO_OIL_GEOM(S, N) \
  e = launch_oil_geom_t<S, N>(rays_a, rays_b, pose_c, x, T, xa, B, J, st, eps_prev, nhb, g2, sd, dt, dump)
  if (solve_T) {
    if (J == 17) ZEDO_OIL_GEOM(true, 17);
    else if (J == 12) ZEDO_OIL_GEOM(true, 12);
    else ZEDO_OIL_GEOM(true, 0);
  } else {
    if (J == 17) ZEDO_OIL_GEOM(false, 17);
    else if (J == 12) ZEDO_OIL_GEOM(false, 12);
    else ZEDO_OIL_GEOM(false, 0);
  }
#undef ZEDO_OIL_GEOM
  ZEDO_CUDA_TRY(e);
  ZEDO_LAUNCH_CHECK();
  return 0;
}

int launch_pack_x(const float* x, __half* xa, int64_t B, int D, cudaStream_t st) {
  if (B == 0) return 0;
  const int64_t threads = B * 8;
  ZEDO_CUDA_TRY(launch_pdl(pack_x_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, st, x, xa, B, D));
  ZEDO_LAUNCH_CHECK();
  return 0;
}

int launch_sde_update(const float* x, const float* eps, int ld_eps, const float* z, const SdeCoef& c,
                      int predictor, int probability_flow, float* x_next, float* x_mean, int64_t B, int D,
                      cudaStream_t st) {
  if (B == 0) return 0;
  float g2, dt, noise_scale;
  if (predictor == ZEDO_PRED_EULER_MARUYAMA) {
    g2 = c.diffusion * c.diffusion;
    dt = c.dt;
    // diffusion[:, None, None] * np.sqrt(-dt): float32 tensor times a python double scalar
    noise_scale = probability_flow ? 0.f : c.diffusion * (float)sqrt(-(double)c.dt);
  } else {
    const float dtp = -c.dt;  // 1/N
    const float G = c.diffusion * sqrtf(dtp);
    g2 = G * G;
    dt = dtp;
    noise_scale = probability_flow ? 0.f : G;
  }
  const int64_t n = B * D;
  ZEDO_CUDA_TRY(launch_pdl(sde_update_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, x, eps, ld_eps, z,
                           -0.5f * c.beta_t, g2, c.std, dt, noise_scale, predictor, x_next, x_mean, B, D));
  ZEDO_LAUNCH_CHECK();
  return 0;
}

}  // namespace zedo

// ---- noise-bearing predictor / corrector updates with caller-injected noise (sampling.py:208-324) -------------
namespace zedo {

// score of one element under get_score_fn's conventions (utils.py:751-795): VP / sub-VP: (-eps) / std; VE: eps
__device__ __forceinline__ float score_of(float e, float std_div) { return std_div > 0.f ? __fdiv_rn(-e, std_div) : e; }

// per row: |score_row|_2 and |z_row|_2 (torch.norm(v.reshape(B, -1), dim=-1), sampling.py:281-282); one warp per row
__global__ void __launch_bounds__(256)
row_norms_kernel(const float* __restrict__ eps, int ld_eps, const float* __restrict__ z, float std_div,
                 float* __restrict__ norms, int64_t B, int D) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  griddep_wait();
  if (row >= B) return;
  float sg = 0.f, sz = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float g = score_of(eps[row * ld_eps + c], std_div);
    const float zz = z[row * D + c];
    sg = fmaf(g, g, sg);
    sz = fmaf(zz, zz, sz);
  }
  sg = warp_sum(sg);
  sz = warp_sum(sz);
  if (lane == 0) {
    norms[2 * row] = sqrtf(sg);
    norms[2 * row + 1] = sqrtf(sz);
  }
}

// stats = (sum_rows |score_row|, sum_rows |z_row|, rows): one CTA, fixed summation order -> the same bits on every run
__global__ void __launch_bounds__(1024) norm_stats_kernel(const float* __restrict__ norms, int64_t B, double* stats) {
  __shared__ double sh[2][1024];
  double a = 0.0, b = 0.0;
  for (int64_t r = threadIdx.x; r < B; r += 1024) {
    a += (double)norms[2 * r];
    b += (double)norms[2 * r + 1];
  }
  sh[0][threadIdx.x] = a;
  sh[1][threadIdx.x] = b;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    stats[0] = sh[0][0];
    stats[1] = sh[1][0];
    stats[2] = (double)B;
  }
}

// One rounding per tensor op of the reference, t uniform over the batch:
//   ancestral VP (:233-241)  x_mean = (x + beta score) / sqrt(1 - beta);           x' = x_mean + sqrt(beta) z        p0 = beta
//   ancestral VE (:220-231)  x_mean = x + score (s^2 - a^2);  x' = x_mean + sqrt(a^2 (s^2 - a^2) / s^2) z   p0 = s, p1 = a
//   Langevin     (:277-285)  step = (snr |z|_mean / |grad|_mean)^2 * 2 * alpha;  x_mean = x + step grad;
//                            x' = x_mean + sqrt(step * 2) z                                    p0 = snr, p1 = alpha
//   ALD          (:314-321)  step = (snr std_m)^2 * 2 * alpha; same update                    p0 = snr, p1 = alpha, p2 = std_m
__global__ void noise_update_kernel(int kind, const float* __restrict__ x, const float* __restrict__ eps, int ld_eps,
                                    const float* __restrict__ z, float std_div, float p0, float p1, float p2,
                                    const double* __restrict__ stats, float* x_next, float* x_mean, int64_t B, int D) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  griddep_wait();
  if (idx >= B * D) return;
  const int64_t row = idx / D;
  const int e = (int)(idx - row * D);
  const float xv = x[idx], zv = z != nullptr ? z[idx] : 0.f;
  const float score = score_of(eps[row * ld_eps + e], std_div);
  float xm, xn;
  if (kind == ZEDO_UPD_ANCESTRAL_VP) {
    xm = __fdiv_rn(__fadd_rn(xv, __fmul_rn(p0, score)), __fsqrt_rn(__fsub_rn(1.f, p0)));
    xn = __fadd_rn(xm, __fmul_rn(__fsqrt_rn(p0), zv));
  } else if (kind == ZEDO_UPD_ANCESTRAL_VE) {
    const float d2 = __fsub_rn(__fmul_rn(p0, p0), __fmul_rn(p1, p1));
    xm = __fadd_rn(xv, __fmul_rn(score, d2));
    const float sd = __fsqrt_rn(__fdiv_rn(__fmul_rn(__fmul_rn(p1, p1), d2), __fmul_rn(p0, p0)));
    xn = __fadd_rn(xm, __fmul_rn(sd, zv));
  } else {
    float base;
    if (kind == ZEDO_UPD_LANGEVIN) {
      const float gn = (float)(stats[0] / stats[2]), nn = (float)(stats[1] / stats[2]);
      base = __fdiv_rn(__fmul_rn(p0, nn), gn);
    } else {
      base = __fmul_rn(p0, p2);
    }
    const float step = __fmul_rn(__fmul_rn(__fmul_rn(base, base), 2.f), p1);
    xm = __fadd_rn(xv, __fmul_rn(step, score));
    xn = __fadd_rn(xm, __fmul_rn(__fsqrt_rn(__fmul_rn(step, 2.f)), zv));
  }
  if (x_mean != nullptr) x_mean[idx] = xm;
  if (x_next != nullptr) x_next[idx] = xn;
}

int launch_row_norm_stats(const float* eps, int ld_eps, const float* z, float std_div, float* norms, double* stats,
                          int64_t B, int D, cudaStream_t st) {
  if (B == 0) return 0;
  row_norms_kernel<<<(unsigned)((B + 7) / 8), 256, 0, st>>>(eps, ld_eps, z, std_div, norms, B, D);
  ZEDO_LAUNCH_CHECK();
  norm_stats_kernel<<<1, 1024, 0, st>>>(norms, B, stats);
  ZEDO_LAUNCH_CHECK();
  return 0;
}

int launch_noise_update(int kind, const float* x, const float* eps, int ld_eps, const float* z, float std_div,
                        float p0, float p1, float p2, const double* stats, float* x_next, float* x_mean, int64_t B,
                        int D, cudaStream_t st) {
  if (B == 0) return 0;
  const int64_t n = B * D;
  noise_update_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(kind, x, eps, ld_eps, z, std_div, p0, p1, p2, stats,
                                                                x_next, x_mean, B, D);
  ZEDO_LAUNCH_CHECK();
  return 0;
}

}  // namespace zedo
