// K1: one fused layer of the score network on the 5th-generation tensor cores.
//
//   out = EPILOGUE( A[M,K] . W[N,K]^T * descale + cbias[N] (+ addend) )            (model.py:264-290)
//
// A and W are fp16 "hi/lo" pairs (v = hi + lo carries ~22 mantissa bits) in the blocked
// core-matrix-interleaved layout of common.cuh, so every pipeline stage is filled by two linear
// bulk async copies (cp.async.bulk -> SASS UBLKCP) that complete on an mbarrier; no tensor map.
// The product is formed by three tcgen05.mma passes into ONE float32 TMEM accumulator:
//   D += A_hi.W_hi ; D += A_lo.W_hi ; D += A_hi.W_lo           (ZEDO_GEMM_SPLIT3, parity mode)
// or two passes A_hi.(W_hi + W_lo) (ZEDO_GEMM_SPLIT2: exact weights, fp16-rounded activations)
// or a single pass A_hi.W_hi (ZEDO_GEMM_FP16, fast mode).
//
// CTA = 128 rows (poses) x BN columns, persistent over tiles, 10 warps:
//   warp 0   : producer  -- one lane issues the bulk copies of a stage
//   warp 1   : MMA issuer -- one lane issues tcgen05.mma / tcgen05.commit; owns TMEM alloc
//   warps 2-9: epilogue  -- thread = one row (TMEM lane) x half of the columns; tcgen05.ld 32 columns at a time;
//              GroupNorm(32 contiguous channels) is therefore thread-local: bias table add,
//              mean/var, affine, SiLU, residual, hi/lo split, blocked store -- no shuffles.
// TMEM holds two accumulator stages (2 x BN columns) so the epilogue of tile i overlaps the
// MMAs of tile i+1.
#include <cstdlib>

#include "tc_common.cuh"

namespace zedo {

template <int BN, int NPROD>
struct TileCfg {
  // NPROD = 3: A_hi.W_hi + A_lo.W_hi + A_hi.W_lo | 2: A_hi.W_hi + A_hi.W_lo | 1: A_hi.W_hi
  static constexpr int kAImage = kActTileRows * kBlockK * 2;  // bytes of one hi or lo A image (16 KiB)
  static constexpr int kBImage = BN * kBlockK * 2;
  // NPROD = 4 (ZEDO_GEMM_FP8LO): stages hold the [hi16 | hi8 | lo8] images of both operands, see mlp_tc2.cu
  static constexpr int kABytes = NPROD == 4 ? 2 * kAImage : kAImage * (NPROD == 3 ? 2 : 1);
  static constexpr int kBBytes = NPROD == 4 ? 2 * kBImage : kBImage * (NPROD >= 2 ? 2 : 1);
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kMaxStages = (225 * 1024) / kStageBytes;
  static constexpr int kStages = kMaxStages > 8 ? 8 : kMaxStages;
  static constexpr int kTmemCols = 2 * BN;  // two accumulator stages (power of two >= 32)
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN, int NPROD, int EPI, int EW, bool RES = true>
__global__ void __launch_bounds__(tc_threads(EW), 1) layer_tc_kernel(const LayerArgs args) {
  using Cfg = TileCfg<BN, NPROD>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + S * Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + S;
  uint64_t* tmem_full = bars + 2 * S;
  uint64_t* tmem_empty = bars + 2 * S + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], EW * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: everything above overlapped with the tail of the previous kernel; from here on its output is read
  griddep_wait();
  griddep_launch_dependents();  // persistent single-wave grid: the next kernel may queue up behind our CTAs

  const int num_tiles = args.m_tiles * args.n_tiles;
  const int num_kb = args.num_kb;

  if (warp == 0) {
    // ===================== producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / args.n_tiles, nt = tile - mt * args.n_tiles;
        const int64_t a_blk = act_block_halves(args.a_fmt);
        const __half* a_src = args.A + ((int64_t)mt * num_kb) * a_blk;
        const __half* w_src = args.W + ((int64_t)nt * num_kb) * 2 * (BN * kBlockK);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (ZEDO_EXPERIMENTS && (args.dbg & 1) && (tile != (int)blockIdx.x || kb >= S)) {
            mbar_arrive(&full[stage]);  // experiment: MMA on stale tiles, no L2->SMEM traffic
          } else {
            mbar_arrive_expect_tx(&full[stage], Cfg::kStageBytes);
            if (NPROD == 3 && args.a_fmt == 1) {
              // format-1 block: hi16 and lo16 are not adjacent (the 8-bit images sit between them)
              bulk_g2s(sA + stage * Cfg::kABytes, a_src + (int64_t)kb * a_blk, Cfg::kAImage, &full[stage]);
              bulk_g2s(sA + stage * Cfg::kABytes + Cfg::kAImage, a_src + (int64_t)kb * a_blk + act_lo16_off(1),
                       Cfg::kAImage, &full[stage]);
            } else {
              bulk_g2s(sA + stage * Cfg::kABytes, a_src + (int64_t)kb * a_blk, Cfg::kABytes, &full[stage]);
            }
            bulk_g2s(sB + stage * Cfg::kBBytes, w_src + (int64_t)kb * 2 * (BN * kBlockK), Cfg::kBBytes,
                     &full[stage]);
          }
          if (++stage == S) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(128, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * Cfg::kABytes);
          const uint32_t b_addr = smem_u32(sB + stage * Cfg::kBBytes);
          const uint64_t a_hi = make_kmajor_desc(a_addr, kActTileRows);
          const uint64_t b_hi = make_kmajor_desc(b_addr, BN);
          // one MMA consumes K = 16 halves = two 16-byte chunks; chunks are tile_rows*16 bytes apart
          constexpr uint32_t kAStep = (2 * kActTileRows * 16) >> 4, kBStep = (2 * BN * 16) >> 4;
          if (ZEDO_EXPERIMENTS && (args.dbg & 2)) {  // experiment: feed only, no tensor work
            umma_commit(&empty[stage]);
            if (++stage == S) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            umma_f16(d_tmem, a_hi + kAStep * k, b_hi + kBStep * k, idesc, (kb | k) != 0);
          if (NPROD == 4) {
            // the two low-order products in e4m3, same order as the CTA-pair kernel (bit-identical results)
            const uint64_t a_hi8 = make_kmajor_desc(a_addr + Cfg::kAImage, kActTileRows);
            const uint64_t a_lo8 = make_kmajor_desc(a_addr + Cfg::kAImage + Cfg::kAImage / 2, kActTileRows);
            const uint64_t b_hi8 = make_kmajor_desc(b_addr + Cfg::kBImage, BN);
            const uint64_t b_lo8 = make_kmajor_desc(b_addr + Cfg::kBImage + Cfg::kBImage / 2, BN);
#pragma unroll
            for (int k = 0; k < kBlockK / 32; ++k) umma_f8(d_tmem, a_lo8 + kAStep * k, b_hi8 + kBStep * k, idesc, 1);
#pragma unroll
            for (int k = 0; k < kBlockK / 32; ++k) umma_f8(d_tmem, a_hi8 + kAStep * k, b_lo8 + kBStep * k, idesc, 1);
          }
          if (NPROD == 3) {
            const uint64_t a_lo = make_kmajor_desc(a_addr + Cfg::kAImage, kActTileRows);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) umma_f16(d_tmem, a_lo + kAStep * k, b_hi + kBStep * k, idesc, 1);
          }
          if (NPROD == 2 || NPROD == 3) {
            const uint64_t b_lo = make_kmajor_desc(b_addr + Cfg::kBImage, BN);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) umma_f16(d_tmem, a_hi + kAStep * k, b_lo + kBStep * k, idesc, 1);
          }
          umma_commit(&empty[stage]);  // frees the smem stage when these MMAs retire
          if (++stage == S) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[as]);  // accumulator ready for the epilogue
      }
    }
  } else {
    // ===================== epilogue: thread = one row x half of the tile's columns =====================
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int chalf = (warp - 2) >> 2;  // which slice of the BN columns
    const int r = q * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int mt = tile / args.n_tiles, nt = tile - mt * args.n_tiles;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      epilogue_tile<BN, EPI, EW, RES>(args, tmem_base + (uint32_t)(as * BN), q, chalf, r, mt, nt);
      tc_fence_before();
      mbar_arrive(&tmem_empty[as]);
    }
  }

  __syncwarp();  // reconverge the single-lane roles before the block-wide barrier
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---- host launcher ------------------------------------------------------------------------------------

template <int BN, int NPROD, int EPI, int EW, bool RES = true>
static int launch_ew(const LayerArgs& a, int num_sms, cudaStream_t st) {
  using Cfg = TileCfg<BN, NPROD>;
  auto kern = layer_tc_kernel<BN, NPROD, EPI, EW, RES>;
  ZEDO_CUDA_TRY(ensure_max_smem((const void*)kern, Cfg::kSmemBytes));
  const int tiles = a.m_tiles * a.n_tiles;
  if (tiles == 0) return 0;
  const int grid = tiles < num_sms ? tiles : num_sms;
  ZEDO_CUDA_TRY(launch_pdl(kern, dim3(grid), dim3(tc_threads(EW)), Cfg::kSmemBytes, st, a));
  ZEDO_LAUNCH_CHECK();
  return 0;
}

template <int BN, int NPROD, int EPI>
static int launch_one(const LayerArgs& a, int num_sms, cudaStream_t st) {
  return launch_ew<BN, NPROD, EPI, 8>(a, num_sms, st);
}

template <int BN, int EPI>
static int launch_nprod(const LayerArgs& a, int nprod, int num_sms, cudaStream_t st) {
  switch (nprod) {
    case 3: return launch_one<BN, 3, EPI>(a, num_sms, st);
    case 2: return launch_one<BN, 2, EPI>(a, num_sms, st);
    case 1: return launch_one<BN, 1, EPI>(a, num_sms, st);
    default: return ZEDO_E_INVALID;
  }
}

// bn: 256 (hidden layers) or 64 (post_dense); nprod: 3 / 2 / 1 MMA passes, 4 = fp8lo (64-channel tiles only); epi: EPI_*
int launch_layer_tc(const LayerArgs& a_in, int bn, int nprod, int epi, int num_sms, cudaStream_t st) {
  LayerArgs a = a_in;
  a.dbg = ZEDO_EXPERIMENTS ? option_get(ZEDO_OPT_EXPERIMENT) : 0;
  if (nprod == 4) {  // small-batch form of the fp8lo layers: format-1 A blocks, [hi16 | hi8 | lo8] weight tiles
    if (bn != 64 || a.a_fmt != 1) return ZEDO_E_INVALID;
    if (epi == EPI_GN_SILU) return launch_ew<64, 4, EPI_GN_SILU, 8>(a, num_sms, st);
    if (epi == EPI_LINEAR_ACT) return launch_ew<64, 4, EPI_LINEAR_ACT, 8>(a, num_sms, st);
    return ZEDO_E_INVALID;
  }
  // the K = 64 first layer is all epilogue: without residual / addend registers 16 epilogue warps fit (4 per scheduler)
  // (r02, same box: 0.566 -> 0.502 ms inside the loop; the lean epilogue with 8 warps is no faster than the general one)
  if (bn == 256 && epi == EPI_GN_SILU && nprod == 3 && a.resid == nullptr && a.addend == nullptr &&
      option_get(ZEDO_OPT_LEAN_EW) == 16)
    return launch_ew<256, 3, EPI_GN_SILU, 16, false>(a, num_sms, st);
  if (bn == 256 && epi == EPI_GN_SILU) return launch_nprod<256, EPI_GN_SILU>(a, nprod, num_sms, st);
  if (bn == 256 && epi == EPI_LINEAR_ACT) return launch_nprod<256, EPI_LINEAR_ACT>(a, nprod, num_sms, st);
  if (bn == 64 && epi == EPI_LINEAR_F32) return launch_nprod<64, EPI_LINEAR_F32>(a, nprod, num_sms, st);
  // narrow tiles for small batches: 4x more CTAs, each with a 4x shorter MMA chain (latency mode)
  if (bn == 64 && epi == EPI_GN_SILU) return launch_nprod<64, EPI_GN_SILU>(a, nprod, num_sms, st);
  if (bn == 64 && epi == EPI_LINEAR_ACT) return launch_nprod<64, EPI_LINEAR_ACT>(a, nprod, num_sms, st);
  return ZEDO_E_INVALID;
}

}  // namespace zedo
