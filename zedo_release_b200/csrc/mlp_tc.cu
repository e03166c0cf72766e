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

#include "kernels.cuh"

namespace zedo {

// ---- PTX wrappers --------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}

// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem] . B[smem]^T, fp16 inputs, float32 accumulate (SASS: UTCHMMA)
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// arrive on an mbarrier once every previously issued MMA of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 consecutive float32 columns of this thread's TMEM lane (SASS: LDTM)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, SWIZZLE_NONE ("interleaved") shared-memory matrix descriptor (cute::UMMA::SmemDescriptor
// bit layout): [0,14) start address >> 4 | [16,30) LBO >> 4 = byte stride between the two 16-byte K
// chunks of one MMA (= tile_rows * 16) | [32,46) SBO >> 4 = 128 B between 8-row core matrices |
// [46,48) version = 1 | [61,64) layout type = 0 (no swizzle).
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr, uint32_t tile_rows) {
  uint64_t d = (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((tile_rows * 16u) >> 4) << 16;
  d |= (uint64_t)(128 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = F16, both K-major
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

template <int BN, int NPROD>
struct TileCfg {
  // NPROD = 3: A_hi.W_hi + A_lo.W_hi + A_hi.W_lo | 2: A_hi.W_hi + A_hi.W_lo | 1: A_hi.W_hi
  static constexpr int kAImage = kActTileRows * kBlockK * 2;  // bytes of one hi or lo A image (16 KiB)
  static constexpr int kBImage = BN * kBlockK * 2;
  static constexpr int kABytes = kAImage * (NPROD == 3 ? 2 : 1);
  static constexpr int kBBytes = kBImage * (NPROD >= 2 ? 2 : 1);
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kMaxStages = (225 * 1024) / kStageBytes;
  static constexpr int kStages = kMaxStages > 8 ? 8 : kMaxStages;
  static constexpr int kTmemCols = 2 * BN;  // two accumulator stages (power of two >= 32)
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

constexpr int kEpiWarps = 8;                       // two warps per TMEM lane quarter, each owns half the columns
constexpr int kTcThreads = 64 + kEpiWarps * 32;    // producer warp + MMA warp + epilogue warps

// hi/lo split of two floats with packed conversions (F2FP.PACK_AB instead of two F2F)
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ void add_hi_lo(float* v, const uint4& h4, const uint4& l4) {
  const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
    const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lw[e]));
    v[2 * e] += hf.x + lf.x;
    v[2 * e + 1] += hf.y + lf.y;
  }
}

template <int BN, int NPROD, int EPI>
__global__ void __launch_bounds__(kTcThreads, 1) layer_tc_kernel(const LayerArgs args) {
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
      mbar_init(&tmem_empty[i], kEpiWarps * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = args.m_tiles * args.n_tiles;
  const int num_kb = args.num_kb;

  if (warp == 0) {
    // ===================== producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / args.n_tiles, nt = tile - mt * args.n_tiles;
        const __half* a_src = args.A + ((int64_t)mt * num_kb) * 2 * (kActTileRows * kBlockK);
        const __half* w_src = args.W + ((int64_t)nt * num_kb) * 2 * (BN * kBlockK);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if ((args.dbg & 1) && (tile != (int)blockIdx.x || kb >= S)) {
            mbar_arrive(&full[stage]);  // experiment: MMA on stale tiles, no L2->SMEM traffic
          } else {
            mbar_arrive_expect_tx(&full[stage], Cfg::kStageBytes);
            bulk_g2s(sA + stage * Cfg::kABytes, a_src + (int64_t)kb * 2 * (kActTileRows * kBlockK), Cfg::kABytes,
                     &full[stage]);
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
          if (args.dbg & 2) {  // experiment: feed only, no tensor work
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
          if (NPROD == 3) {
            const uint64_t a_lo = make_kmajor_desc(a_addr + Cfg::kAImage, kActTileRows);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) umma_f16(d_tmem, a_lo + kAStep * k, b_hi + kBStep * k, idesc, 1);
          }
          if (NPROD >= 2) {
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
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int chalf = (warp - 2) >> 2;   // which half of the BN columns
    const int r = q * 32 + lane;
    constexpr int kGroupsPerWarp = BN / 64;
    const int n_total = args.n_tiles * BN;
    const int nkb_out = n_total / kBlockK;
    constexpr int64_t kLoOff = kActTileRows * kBlockK;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int mt = tile / args.n_tiles, nt = tile - mt * args.n_tiles;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
#pragma unroll 1
      for (int gg = 0; gg < kGroupsPerWarp; ++gg) {
        const int g = chalf * kGroupsPerWarp + gg;
        const int col0 = nt * BN + g * 32;
        // position of this thread's 32 columns inside the blocked [M_pad, N_pad] activation layout
        const int kbo = col0 / kBlockK;
        const int hsel = (col0 / 32) & 1;
        // chunk c of this row lives at c * (128 rows * 8 halves) + r * 8 inside the (mt, kbo) hi image
        const int64_t row_off =
            (((int64_t)mt * nkb_out + kbo) * 2) * (kActTileRows * kBlockK) + (int64_t)r * 8;
        constexpr int kChunkStride = kActTileRows * 8;
        // residual / addend loads are issued before the TMEM read so their latency overlaps it
        uint4 rh[4], rl[4];
        if (EPI != EPI_LINEAR_F32 && args.resid != nullptr) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int pc = (hsel * 4 + j) * kChunkStride;
            rh[j] = *reinterpret_cast<const uint4*>(args.resid + row_off + pc);
            rl[j] = *reinterpret_cast<const uint4*>(args.resid + row_off + kLoOff + pc);
          }
        }
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + g * 32), v);
        const float4* cb = reinterpret_cast<const float4*>(args.cbias + col0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b4 = __ldg(cb + i);
          v[4 * i + 0] = fmaf(v[4 * i + 0], args.descale, b4.x);
          v[4 * i + 1] = fmaf(v[4 * i + 1], args.descale, b4.y);
          v[4 * i + 2] = fmaf(v[4 * i + 2], args.descale, b4.z);
          v[4 * i + 3] = fmaf(v[4 * i + 3], args.descale, b4.w);
        }
        if (EPI == EPI_LINEAR_F32) {
          float4* dst = reinterpret_cast<float4*>(args.out_f32 + ((int64_t)mt * kActTileRows + r) * args.ld_out + col0);
#pragma unroll
          for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          continue;
        }
        if (args.addend != nullptr) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int pc = (hsel * 4 + j) * kChunkStride;
            add_hi_lo(v + 8 * j, *reinterpret_cast<const uint4*>(args.addend + row_off + pc),
                      *reinterpret_cast<const uint4*>(args.addend + row_off + kLoOff + pc));
          }
        }
        if (EPI == EPI_GN_SILU) {
          float sum = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) sum += v[i];
          const float mean = sum * (1.f / 32.f);
          float sq = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v[i] -= mean;
            sq = fmaf(v[i], v[i], sq);
          }
          const float rstd = 1.f / sqrtf(sq * (1.f / 32.f) + args.gn_eps);
          const float4* gp = reinterpret_cast<const float4*>(args.gamma + col0);
          const float4* bp = reinterpret_cast<const float4*>(args.beta + col0);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 g4 = __ldg(gp + i), b4 = __ldg(bp + i);
            const float ga[4] = {g4.x, g4.y, g4.z, g4.w}, be[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float y = fmaf(v[4 * i + e] * rstd, ga[e], be[e]);
              v[4 * i + e] = __fdividef(y, 1.f + __expf(-y));
            }
          }
        }
        if (args.resid != nullptr) {
#pragma unroll
          for (int j = 0; j < 4; ++j) add_hi_lo(v + 8 * j, rh[j], rl[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) split_pair(v[8 * j + 2 * e], v[8 * j + 2 * e + 1], hi[e], lo[e]);
          const int pc = (hsel * 4 + j) * kChunkStride;
          *reinterpret_cast<uint4*>(args.out + row_off + pc) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(args.out + row_off + kLoOff + pc) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
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

template <int BN, int NPROD, int EPI>
static int launch_one(const LayerArgs& a, int num_sms, cudaStream_t st) {
  using Cfg = TileCfg<BN, NPROD>;
  static bool configured = false;
  auto kern = layer_tc_kernel<BN, NPROD, EPI>;
  if (!configured) {
    ZEDO_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const int tiles = a.m_tiles * a.n_tiles;
  if (tiles == 0) return 0;
  const int grid = tiles < num_sms ? tiles : num_sms;
  kern<<<grid, kTcThreads, Cfg::kSmemBytes, st>>>(a);
  ZEDO_LAUNCH_CHECK();
  return 0;
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

// bn: 256 (hidden layers) or 64 (post_dense); nprod: 3 / 2 / 1 MMA passes; epi: EPI_*
int launch_layer_tc(const LayerArgs& a_in, int bn, int nprod, int epi, int num_sms, cudaStream_t st) {
  static const int dbg = getenv("ZEDO_DBG") ? atoi(getenv("ZEDO_DBG")) : 0;
  LayerArgs a = a_in;
  a.dbg = dbg;
  if (bn == 256 && epi == EPI_GN_SILU) return launch_nprod<256, EPI_GN_SILU>(a, nprod, num_sms, st);
  if (bn == 256 && epi == EPI_LINEAR_ACT) return launch_nprod<256, EPI_LINEAR_ACT>(a, nprod, num_sms, st);
  if (bn == 64 && epi == EPI_LINEAR_F32) return launch_nprod<64, EPI_LINEAR_F32>(a, nprod, num_sms, st);
  return ZEDO_E_INVALID;
}

}  // namespace zedo
