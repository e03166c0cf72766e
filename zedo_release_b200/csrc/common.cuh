// Shared helpers for the ZeDO B200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/zedo_b200.h"

namespace zedo {

// ---- launch accounting (zedo_launch_count) -----------------------------------------------------
void count_launch(int n = 1);

#define ZEDO_CUDA_TRY(expr)                        \
  do {                                             \
    cudaError_t _e = (expr);                       \
    if (_e != cudaSuccess) return (int)_e;         \
  } while (0)

#define ZEDO_LAUNCH_CHECK()                        \
  do {                                             \
    ::zedo::count_launch();                        \
    cudaError_t _e = cudaGetLastError();           \
    if (_e != cudaSuccess) return (int)_e;         \
  } while (0)

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------
// Kernels of the OIL loop are launched with cudaLaunchAttributeProgrammaticStreamSerialization: the next
// kernel may be scheduled while the previous one drains, runs its prologue (barrier init, TMEM allocation)
// and then blocks in griddep_wait() until the previous grid has completed and flushed.  Every kernel
// launched that way calls griddep_wait() before it touches global memory written by a predecessor.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
bool pdl_enabled();  // option ZEDO_OPT_PDL (default on)

// ---- process-wide tuning options (zedo_set_option; initial values may come from ZEDO_* environment variables,
// read ONCE at first use -- never on a launch path) -------------------------------------------------------
int option_get(int opt);
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (device, kernel) instead of on every launch
cudaError_t ensure_max_smem(const void* func, int bytes);

// Experiment code paths of the layer kernels (stale-stage / no-MMA timing runs) exist only in builds
// with -DZEDO_EXPERIMENTS=1; the shipped library carries none of them.
#ifndef ZEDO_EXPERIMENTS
#define ZEDO_EXPERIMENTS 0
#endif

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// ---- blocked, core-matrix-interleaved fp16 operand layout ------------------------------------------
// A [rows, cols] fp16 operand (cols padded to a multiple of 64) is stored as tiles of
// TILE_ROWS x 64 halves; tile (rt, kb) has a "hi" image followed by a "lo" image, each
// TILE_ROWS*128 bytes.  Inside an image the eight 16-byte column chunks (8 halves each) are the
// outer index and the rows the inner one: byte offset = chunk*(TILE_ROWS*16) + r*16.  That is the
// canonical K-major SWIZZLE_NONE ("interleaved") UMMA layout -- 8x16B core matrices, SBO = 128 B
// between core matrices along M/N, LBO = TILE_ROWS*16 B between the two K chunks of one MMA -- so
//   * one linear bulk copy (cp.async.bulk) of the tile image is directly consumable by tcgen05.mma;
//   * an epilogue warp (lane = row) writes 32 consecutive rows of one chunk = 512 contiguous bytes
//     per store instruction (4 L1 wavefronts instead of 32 for a row-major or 128B-swizzled image).
#ifndef ZEDO_BLOCK_K
#define ZEDO_BLOCK_K 64
#endif
constexpr int kBlockK = ZEDO_BLOCK_K;  // columns (K) per block: 64, or 32 in the -DZEDO_BLOCK_K=32 layout study (r02c)
constexpr int kXaCols = 64;        // width of the first layer's operand x (3J <= 64) and K padding of every weight
static_assert(kBlockK == 64 || kBlockK == 32, "block width");
constexpr int kActTileRows = 128;  // BLOCK_M

__host__ __device__ inline int64_t blocked_half_offset(int64_t row, int64_t col, int64_t cols_padded,
                                                       int tile_rows, int hl) {
  const int64_t num_kb = cols_padded / kBlockK;
  const int64_t rt = row / tile_rows, r = row % tile_rows;
  const int64_t kb = col / kBlockK, c = col % kBlockK;
  const int64_t image = (int64_t)tile_rows * kBlockK;  // halves per hi or lo image
  return ((rt * num_kb + kb) * 2 + hl) * image + (c >> 3) * ((int64_t)tile_rows * 8) + r * 8 + (c & 7);
}

// Activation block formats.  A block is the 128 x 64 slice (row tile, k-block) of an activation matrix:
//   format 0  [hi16 | lo16]                2 x 16 KiB   v = hi16 + lo16 (fp16 pair, ~22 bits)
//   format 1  [hi16 | hi8 | lo8 | lo16]    16 + 8 + 8 + 16 KiB, used by the ZEDO_GEMM_FP8LO mode:
//             hi8 = e4m3(hi16), lo8 = e4m3(lo16 * 2^11) are the operands of the two low-order products, which only
//             need ~4 bits (they sit 2^-11 below the main product) and run at twice the fp16 MMA rate;
//             the first 32 KiB are what a GEMM stage loads (one bulk copy), lo16 is kept for the consumers that add
//             the activation itself (residual / addend epilogues, post_dense).
// An 8-bit image uses the same interleaved layout with 16 elements per 16-byte chunk:
// byte offset = (c / 16) * (rows * 16) + r * 16 + c % 16.
constexpr int kImgHalves = kActTileRows * kBlockK;  // one fp16 image of a block (16 KiB)
__host__ __device__ constexpr int act_block_halves(int fmt) { return fmt ? 3 * kImgHalves : 2 * kImgHalves; }
__host__ __device__ constexpr int act_lo16_off(int fmt) { return fmt ? 2 * kImgHalves : kImgHalves; }  // in halves
constexpr int kHi8ByteOff = kImgHalves * 2;             // byte offset of the hi8 image inside a format-1 block
constexpr int kLo8ByteOff = kImgHalves * 2 + kImgHalves;  // ... of the lo8 image
constexpr float kLo8Scale = 2048.f;                     // lo8 = e4m3(lo * 2^11); the matching W image carries 2^-11

__host__ __device__ inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

// ---- hi/lo split --------------------------------------------------------------------------------
__device__ __forceinline__ void split_hi_lo(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

__device__ __forceinline__ uint32_t pack_half2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- small 3x3 helpers (row-major) ----------------------------------------------------------------
__device__ __forceinline__ void inv3x3(const float* m, float* o) {
  const float a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
  const float A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
  const float det = a * A + b * B + c * C;
  const float r = 1.0f / det;
  o[0] = A * r;
  o[1] = -(b * i - c * h) * r;
  o[2] = (b * f - c * e) * r;
  o[3] = B * r;
  o[4] = (a * i - c * g) * r;
  o[5] = -(a * f - c * d) * r;
  o[6] = C * r;
  o[7] = -(a * h - b * g) * r;
  o[8] = (a * e - b * d) * r;
}

// sub-VP scalars of one step, float32 op order of sde_lib.py:187-198 / utils.py:762-776
struct SdeCoef {
  float beta_t;     // beta(t)
  float diffusion;  // g(t)
  float std;        // marginal std (1 - exp(2 lmc))
  float dt;         // -1/N (Euler-Maruyama) ; +1/N for the reverse-diffusion discretisation
};
SdeCoef subvp_coef(float t, float beta_min, float beta_max, int n_scales);

}  // namespace zedo
