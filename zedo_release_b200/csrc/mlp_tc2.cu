// K1, CTA-pair version for the 1024 -> 1024 layers: tcgen05.mma.cta_group::2, M = 256 per pair.
//
// A cluster of two CTAs (one TPC = two SMs) computes 256 poses x 256 channels.  Each CTA stages
// only ITS 128 rows of A (hi+lo, 32 KiB) and ITS half (128 channels) of the W tile (hi+lo, 32 KiB)
// per 64-wide k-block: 64 KiB per stage instead of 96, so three stages fit (the one-CTA kernel is
// limited to two and stalls on load latency) and the L2 -> SMEM traffic per flop drops by a third.
// The leader CTA's MMA lane issues every tcgen05.mma for the pair; accumulator rows 0-127 land in the
// leader's TMEM, rows 128-255 in the peer's, so each CTA's epilogue warps drain their own TMEM exactly
// as in mlp_tc.cu.
//
// Stage copies and cross-CTA signalling.  Default (r02c, LayerArgs::tmapA / tmapW set): the operand buffers are described
// to the TMA unit as byte matrices [bytes / 128][128] with a 32 KiB box, and both CTAs issue
// cp.async.bulk.tensor.2d ... cta_group::2 copies into their OWN shared memory whose completion bytes are credited to the
// LEADER's full[s]; the leader expects 2 x 64 KiB per stage and its MMA lane waits once.  Fall-back / A-B form
// (ZEDO_OPT_TMA_2SM = 0, or no cuTensorMapEncodeTiled in the driver): linear cp.async.bulk copies, which can only complete
// on a barrier of the destination CTA, so the peer relays (this hand-off cost 11 % of the layer, profiles/r02c_tma2sm.md).
//   full[s]       leader: both CTAs' tensor-map copies landed | linear form: own bulk copies landed (per CTA)
//   peer_full[s]  linear form, leader only: the peer's relay lane observed its full[s] and arrived remotely
//   empty[s]      tcgen05.commit multicast (mask 0b11): the MMAs that read stage s in BOTH CTAs retired
//   tmem_full[a]  tcgen05.commit multicast: accumulator a complete in both TMEMs
//   tmem_empty[a] leader only, count 2*EW: one arrive per epilogue warp of both CTAs (peer: remote)
#include <cstdlib>

#include "tc_common.cuh"

namespace zedo {

// Stage-event trace of CTA pair 0 (experiments build, experiment bit 16): clock64 of this SM at the hand-offs of the
// operand pipeline, for stage iterations [kTraceSkip, kTraceSkip + kTraceLen) of the launch.  Events: 0 producer starts
// waiting for the stage to be free, 1 it is free, 2 copies issued, 3 MMA / relay lane sees the stage full, 4 leader sees
// the peer's stage full, 5 MMAs + commit issued (peer: relay arrive sent).  Read back with zedo_debug_stage_trace.
#if ZEDO_EXPERIMENTS
constexpr int kTraceLen = 512, kTraceSkip = 96, kTraceEvents = 6;
__device__ unsigned long long g_stage_trace[2][kTraceEvents][kTraceLen];
#define ZEDO_TRACE(ev, it)                                                                                  \
  do {                                                                                                      \
    if ((args.dbg & 16) && blockIdx.x < 2 && (it) >= kTraceSkip && (it) < kTraceSkip + kTraceLen)            \
      g_stage_trace[blockIdx.x][ev][(it) - kTraceSkip] = (unsigned long long)clock64();                      \
  } while (0)
#else
#define ZEDO_TRACE(ev, it) do { } while (0)
#endif

__device__ __forceinline__ void umma_f16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// arrive (once all prior MMAs of this thread retire) on the barrier at this offset in both CTAs
// the same for 8-bit operands (e4m3 x e4m3, K = 32 per instruction, twice the fp16 rate); kind::f8f6f4 shares the
// instruction-descriptor layout of kind::f16 and its format code 0 is E4M3, so make_idesc_f16 serves both
__device__ __forceinline__ void umma_f8_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}

template <int NPROD>
struct PairCfg {
  static constexpr int kBN = 256;                              // channels per pair tile
  static constexpr int kHalfN = 128;                           // W rows staged by each CTA
  static constexpr int kAImage = kActTileRows * kBlockK * 2;   // 16 KiB
  static constexpr int kBImage = kHalfN * kBlockK * 2;         // 16 KiB
  // NPROD = 4 (ZEDO_GEMM_FP8LO): A_hi16.W_hi16 in fp16 + A_lo8.W_hi8 + A_hi8.W_lo8 in e4m3; a stage holds the
  // [hi16 | hi8 | lo8] images of both operands (format-1 blocks, common.cuh) = 32 KiB each
  static constexpr int kABytes = NPROD == 4 ? 2 * kAImage : kAImage * (NPROD == 3 ? 2 : 1);
  static constexpr int kBBytes = NPROD == 4 ? 2 * kBImage : kBImage * (NPROD >= 2 ? 2 : 1);
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kMaxStages = (225 * 1024) / kStageBytes;
  static constexpr int kStages = kMaxStages > 6 ? 6 : kMaxStages;
  static constexpr int kTmemCols = 2 * kBN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

template <int NPROD, int EPI, int EW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(tc_threads(EW), 1)
layer_tc2_kernel(const LayerArgs args) {
  using Cfg = PairCfg<NPROD>;
  constexpr int S = Cfg::kStages;
  constexpr int BN = Cfg::kBN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + S * Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes);
  uint64_t* full = bars;
  uint64_t* peer_full = bars + S;
  uint64_t* empty = bars + 2 * S;
  uint64_t* tmem_full = bars + 3 * S;
  uint64_t* tmem_empty = bars + 3 * S + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * S + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&peer_full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * EW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync_all();  // both CTAs' barriers exist before anyone signals across the pair
  if (warp == 1) tmem_alloc_2cta(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();                // PDL: the prologue above overlapped with the previous kernel's tail
  griddep_launch_dependents();

  const int num_kb = args.num_kb;
  // tensor-map copies completing on the leader's barrier need 32 KiB operand stages (NPROD 3 / 4 with 64-column blocks)
  const bool tma2 = Cfg::kABytes == 32768 && Cfg::kBBytes == 32768 && args.tmapA != nullptr && args.tmapW != nullptr &&
                    !(ZEDO_EXPERIMENTS && (args.dbg & 5));
  const int num_pairs = (args.m_tiles / 2) * args.n_tiles;  // m_tiles is even (plan pads to 256 rows)
  const int pair0 = blockIdx.x >> 1, pair_stride = gridDim.x >> 1;

  if (warp == 0) {
    // ===================== producer (both CTAs): own A rows + own half of W =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int sit = 0;  // stage iteration of this launch (trace index)
      for (int pt = pair0; pt < num_pairs; pt += pair_stride) {
        const int mp = pt / args.n_tiles, nt = pt - mp * args.n_tiles;
        const int mt = 2 * mp + (int)rank;
        // an A block is 32 KiB (format 0) or 48 KiB (format 1); either way the stage takes its first kABytes
        const int64_t a_blk = act_block_halves(args.a_fmt);
        const __half* a_src = args.A + ((int64_t)mt * num_kb) * a_blk;
        const __half* w_src = args.W + ((int64_t)(2 * nt + (int)rank) * num_kb) * 2 * (Cfg::kHalfN * kBlockK);
        for (int kb = 0; kb < num_kb; ++kb, ++sit) {
          ZEDO_TRACE(0, sit);
          mbar_wait(&empty[stage], phase ^ 1);
          ZEDO_TRACE(1, sit);
          if (ZEDO_EXPERIMENTS && (args.dbg & 1) && (pt != pair0 || kb >= S)) {
            mbar_arrive(&full[stage]);  // experiment: MMA on stale tiles, no L2 -> SMEM traffic
            if (++stage == S) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          if (ZEDO_EXPERIMENTS && (args.dbg & 4)) {
            // experiment: 3/4 of the stage bytes cross L2 -> SMEM (what a stage without the hi8 images would cost);
            // results are garbage by design
            mbar_arrive_expect_tx(&full[stage], Cfg::kStageBytes / 4 * 3);
            bulk_g2s(sA + stage * Cfg::kABytes, a_src + (int64_t)kb * a_blk, Cfg::kABytes / 4 * 3, &full[stage]);
            bulk_g2s(sB + stage * Cfg::kBBytes, w_src + (int64_t)kb * 2 * (Cfg::kHalfN * kBlockK), Cfg::kBBytes / 4 * 3,
                     &full[stage]);
            if (++stage == S) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          if (tma2) {
            // both CTAs' copies are credited to the LEADER's full[stage]: one wait for the MMA lane, no relay hop.
            // A box = 256 rows of 128 bytes = 32 KiB; a block of A is a_blk * 2 bytes, a W half tile 32 KiB.
            const uint32_t lead_full = mapa_u32(smem_u32(&full[stage]), 0);
            if (leader) mbar_arrive_expect_tx(&full[stage], 2 * Cfg::kStageBytes);
            tma2d_g2s_2cta(sA + stage * Cfg::kABytes, args.tmapA, 0,
                           (int)((((int64_t)mt * num_kb + kb) * a_blk * 2) >> 7), lead_full);
            tma2d_g2s_2cta(sB + stage * Cfg::kBBytes, args.tmapW, 0,
                           (int)((((int64_t)(2 * nt + (int)rank) * num_kb + kb) * Cfg::kBBytes) >> 7), lead_full);
          } else {
            mbar_arrive_expect_tx(&full[stage], Cfg::kStageBytes);
            bulk_g2s(sA + stage * Cfg::kABytes, a_src + (int64_t)kb * a_blk, Cfg::kABytes, &full[stage]);
            bulk_g2s(sB + stage * Cfg::kBBytes, w_src + (int64_t)kb * 2 * (Cfg::kHalfN * kBlockK), Cfg::kBBytes,
                     &full[stage]);
          }
          ZEDO_TRACE(2, sit);
          if (++stage == S) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int sit = 0;
      if (leader) {
        // ===================== MMA issuer (leader CTA only) =====================
        constexpr uint32_t idesc = make_idesc_f16(256, BN);
        int it = 0;
        for (int pt = pair0; pt < num_pairs; pt += pair_stride, ++it) {
          const int as = it & 1;
          const uint32_t aphase = (it >> 1) & 1;
          mbar_wait(&tmem_empty[as], aphase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
          for (int kb = 0; kb < num_kb; ++kb, ++sit) {
            if (ZEDO_EXPERIMENTS && (args.dbg & 64)) mbar_wait_spin(&full[stage], phase); else mbar_wait(&full[stage], phase);
            ZEDO_TRACE(3, sit);
            if (!tma2) {
              if (ZEDO_EXPERIMENTS && (args.dbg & 64)) mbar_wait_spin(&peer_full[stage], phase); else mbar_wait(&peer_full[stage], phase);
            }
            ZEDO_TRACE(4, sit);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(sA + stage * Cfg::kABytes);
            const uint32_t b_addr = smem_u32(sB + stage * Cfg::kBBytes);
            const uint64_t a_hi = make_kmajor_desc(a_addr, kActTileRows);
            const uint64_t b_hi = make_kmajor_desc(b_addr, Cfg::kHalfN);
            constexpr uint32_t kAStep = (2 * kActTileRows * 16) >> 4, kBStep = (2 * Cfg::kHalfN * 16) >> 4;
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              umma_f16_2cta(d_tmem, a_hi + kAStep * k, b_hi + kBStep * k, idesc, (kb | k) != 0);
            if (NPROD == 4) {
              // 64 e4m3 columns = four 16-byte chunks = two K = 32 instructions per product; same chunk strides
              const uint64_t a_hi8 = make_kmajor_desc(a_addr + Cfg::kAImage, kActTileRows);
              const uint64_t a_lo8 = make_kmajor_desc(a_addr + Cfg::kAImage + Cfg::kAImage / 2, kActTileRows);
              const uint64_t b_hi8 = make_kmajor_desc(b_addr + Cfg::kBImage, Cfg::kHalfN);
              const uint64_t b_lo8 = make_kmajor_desc(b_addr + Cfg::kBImage + Cfg::kBImage / 2, Cfg::kHalfN);
#pragma unroll
              for (int k = 0; k < kBlockK / 32; ++k)
                umma_f8_2cta(d_tmem, a_lo8 + kAStep * k, b_hi8 + kBStep * k, idesc, 1);
#pragma unroll
              for (int k = 0; k < kBlockK / 32; ++k)
                umma_f8_2cta(d_tmem, a_hi8 + kAStep * k, b_lo8 + kBStep * k, idesc, 1);
            }
            if (NPROD == 3) {
              const uint64_t a_lo = make_kmajor_desc(a_addr + Cfg::kAImage, kActTileRows);
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k)
                umma_f16_2cta(d_tmem, a_lo + kAStep * k, b_hi + kBStep * k, idesc, 1);
            }
            if (NPROD == 2 || NPROD == 3) {
              const uint64_t b_lo = make_kmajor_desc(b_addr + Cfg::kBImage, Cfg::kHalfN);
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k)
                umma_f16_2cta(d_tmem, a_hi + kAStep * k, b_lo + kBStep * k, idesc, 1);
            }
            umma_commit_2cta(&empty[stage]);
            ZEDO_TRACE(5, sit);
            if (++stage == S) {
              stage = 0;
              phase ^= 1;
            }
          }
          umma_commit_2cta(&tmem_full[as]);
        }
      } else {
        // ===================== relay (peer CTA): tell the leader this CTA's stage has landed =====================
        // (not needed with tensor-map copies: they complete on the leader's barrier themselves)
        for (int pt = pair0; pt < num_pairs && !tma2; pt += pair_stride) {
          for (int kb = 0; kb < num_kb; ++kb, ++sit) {
            if (ZEDO_EXPERIMENTS && (args.dbg & 64)) mbar_wait_spin(&full[stage], phase); else mbar_wait(&full[stage], phase);
            ZEDO_TRACE(3, sit);
            mbar_arrive_cluster(&peer_full[stage], 0);
            ZEDO_TRACE(5, sit);
            if (++stage == S) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else {
    // ===================== epilogue (both CTAs): own 128 rows =====================
    const int q = warp & 3;
    const int chalf = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    int it = 0;
    for (int pt = pair0; pt < num_pairs; pt += pair_stride, ++it) {
      const int mp = pt / args.n_tiles, nt = pt - mp * args.n_tiles;
      const int mt = 2 * mp + (int)rank;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      if (!(ZEDO_EXPERIMENTS && (args.dbg & 8)))  // experiment bit 8: no epilogue work at all (feed + MMA only)
        epilogue_tile<BN, EPI, EW>(args, tmem_base + (uint32_t)(as * BN), q, chalf, r, mt, nt);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(&tmem_empty[as], 0);  // the leader's MMA lane owns accumulator reuse
    }
  }

  __syncwarp();
  tc_fence_before();
  cluster_sync_all();  // nobody leaves while the partner may still read its SMEM / signal its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, Cfg::kTmemCols);
  }
}

template <int NPROD, int EPI, int EW>
static int launch_pair_ew(const LayerArgs& a, int num_sms, cudaStream_t st) {
  using Cfg = PairCfg<NPROD>;
  auto kern = layer_tc2_kernel<NPROD, EPI, EW>;
  ZEDO_CUDA_TRY(ensure_max_smem((const void*)kern, Cfg::kSmemBytes));
  if (a.m_tiles % 2 != 0) return ZEDO_E_SHAPE;
  const int pairs = (a.m_tiles / 2) * a.n_tiles;
  if (pairs == 0) return 0;
  int max_pairs = num_sms / 2;
  // experiment (bits 8..15 of the experiment option): run on that many CTA pairs only -- does the layer follow the number
  // of SMs that ingest (a per-SM bound) or stay put (the chip-wide L2 output cap)?
  if (ZEDO_EXPERIMENTS && (a.dbg >> 8) > 0 && (a.dbg >> 8) < max_pairs) max_pairs = a.dbg >> 8;
  const int grid = 2 * (pairs < max_pairs ? pairs : max_pairs);
  ZEDO_CUDA_TRY(launch_pdl(kern, dim3(grid), dim3(tc_threads(EW)), Cfg::kSmemBytes, st, a));
  ZEDO_LAUNCH_CHECK();
  return 0;
}

template <int NPROD, int EPI>
static int launch_pair(const LayerArgs& a, int num_sms, cudaStream_t st) {
  return launch_pair_ew<NPROD, EPI, 8>(a, num_sms, st);  // 12 / 16 epilogue warps measured slower (r01, r02)
}

template <int EPI>
static int launch_pair_nprod(const LayerArgs& a, int nprod, int num_sms, cudaStream_t st) {
  switch (nprod) {
    case 3: return launch_pair<3, EPI>(a, num_sms, st);
    case 2: return launch_pair<2, EPI>(a, num_sms, st);
    case 1: return launch_pair<1, EPI>(a, num_sms, st);
    case 4: return a.a_fmt == 1 ? launch_pair<4, EPI>(a, num_sms, st) : ZEDO_E_INVALID;
    default: return ZEDO_E_INVALID;
  }
}

// hidden layers (N multiple of 256, weights packed with 128-row tiles); a.m_tiles must be even
int launch_layer_tc2(const LayerArgs& a_in, int nprod, int epi, int num_sms, cudaStream_t st) {
  LayerArgs a = a_in;
  a.dbg = ZEDO_EXPERIMENTS ? option_get(ZEDO_OPT_EXPERIMENT) : 0;
  if (epi == EPI_GN_SILU) return launch_pair_nprod<EPI_GN_SILU>(a, nprod, num_sms, st);
  if (epi == EPI_LINEAR_ACT) return launch_pair_nprod<EPI_LINEAR_ACT>(a, nprod, num_sms, st);
  return ZEDO_E_INVALID;
}

}  // namespace zedo

#if ZEDO_EXPERIMENTS
// experiments build only: the stage-event trace of the last traced launch, [2 CTAs][6 events][512 iterations] clocks
extern "C" int zedo_debug_stage_trace(unsigned long long* host, int n) {
  if (host == nullptr || n != 2 * zedo::kTraceEvents * zedo::kTraceLen) return -1;
  return (int)cudaMemcpyFromSymbol(host, zedo::g_stage_trace, sizeof(unsigned long long) * (size_t)n);
}
#endif
