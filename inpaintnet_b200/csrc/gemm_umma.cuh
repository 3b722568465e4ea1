// tcgen05 (UMMA) GEMM core: D[128 x G*BNG tile] accumulated in TMEM from TMA-fed SWIZZLE_128B
// shared-memory stages, warp specialised (warp0 = TMA producer, warp1 = MMA issuer + TMEM owner,
// warps 2..5 = epilogue, one TMEM lane == one output row per thread).  bf16 x bf16 -> fp32.
//
//   D[m, g, n] = sum_seg sum_k A_seg[m, k] * B_seg[g*gate_stride + n, k]
//
// Operands may be K-major (row-major [rows, K]) or MN-major (row-major [K, rows]) independently,
// so forward (NT), dgrad (NN) and wgrad (TN) GEMMs all run without materialised transposes.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace ipn {

constexpr int UMMA_BM = 128;
constexpr int UMMA_BK = 64;  // bf16 elements per stage along K (= 128 bytes, one swizzle row)
constexpr int UMMA_THREADS = 320;  // warp0 TMA, warp1 MMA, warps 2..9 epilogue (2 warps per TMEM lane quadrant)

struct UmmaSeg {
  alignas(64) CUtensorMap tmA;
  alignas(64) CUtensorMap tmB;
  int K;
  int a_c0, a_c1;  // base coordinates (inner, outer) added to the tile coordinates
  int b_c0, b_c1;
};

template <class Epi>
struct UmmaProblem {
  UmmaSeg seg[2];
  int nseg;
  int M, N;         // N = columns per gate
  int gate_stride;  // row (K-major B) / column (MN-major B) distance between gates
  typename Epi::Params epi;
};

template <class Epi>
struct UmmaBatch {
  UmmaProblem<Epi> p[2];
  int split_k;
};

template <int G_, int BNG_, bool TA_, bool TB_>
struct UmmaCfg {
  static constexpr int G = G_;
  static constexpr int BNG = BNG_;
  static constexpr bool TA = TA_;
  static constexpr bool TB = TB_;
  static constexpr int BN = G_ * BNG_;
  static constexpr int A_BYTES = UMMA_BM * UMMA_BK * 2;
  static constexpr int B_BYTES = BN * UMMA_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // Two CTAs per SM: one CTA's epilogue (global-memory latency bound) overlaps the other's mainloop and
  // the memory-level parallelism per SM doubles.  ~110 KB of stages per CTA, at least 2 stages.
  static constexpr int MAX_SMEM = 110 * 1024;
  static constexpr int STAGES_RAW = MAX_SMEM / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 4 ? 4 : (STAGES_RAW < 2 ? 2 : STAGES_RAW);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr uint32_t TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : BN <= 256 ? 256 : 512;
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N must be a multiple of 16 in [16,256] for M=128");
  static_assert(!TB_ || (BNG_ % 64 == 0), "MN-major B needs 64-wide blocks");
  static_assert(BNG_ % 16 == 0, "epilogue works on 16-column chunks");
};

template <class Cfg, class Epi>
__global__ void __launch_bounds__(UMMA_THREADS, 2) umma_gemm_kernel(const __grid_constant__ UmmaBatch<Epi> batch) {
  static_assert(Epi::G == Cfg::G, "epilogue / tile gate count mismatch");
  constexpr int G = Cfg::G, BNG = Cfg::BNG, BN = Cfg::BN, STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int prob = blockIdx.z / batch.split_k;
  const int ksplit = blockIdx.z - prob * batch.split_k;
  const UmmaProblem<Epi>& P = batch.p[prob];
  const int n0 = blockIdx.x * BNG;
  const int m0 = blockIdx.y * UMMA_BM;

  // k-chunk range handled by this CTA (over the concatenation of all segments)
  int chunks_seg0 = (P.seg[0].K + UMMA_BK - 1) / UMMA_BK;
  int chunks_total = chunks_seg0 + (P.nseg > 1 ? (P.seg[1].K + UMMA_BK - 1) / UMMA_BK : 0);
  const int per = (chunks_total + batch.split_k - 1) / batch.split_k;
  const int kc_begin = ksplit * per;
  const int kc_end = min(chunks_total, kc_begin + per);
  const int nchunks = kc_end - kc_begin;

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        ptx::mbar_init(&full_bar[s], 1);
        ptx::mbar_init(&empty_bar[s], 1);
      }
      ptx::mbar_init(tmem_full_bar, 1);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    ptx::tmem_relinquish();
  } else if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&P.seg[0].tmA);
    ptx::prefetch_tmap(&P.seg[0].tmB);
    if (P.nseg > 1) {
      ptx::prefetch_tmap(&P.seg[1].tmA);
      ptx::prefetch_tmap(&P.seg[1].tmB);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && nchunks > 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kc = kc_begin; kc < kc_end; ++kc) {
        const int si = (kc >= chunks_seg0) ? 1 : 0;
        const UmmaSeg& S = P.seg[si];
        const int k0 = (si ? kc - chunks_seg0 : kc) * UMMA_BK;
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
        uint8_t* sb = sa + Cfg::A_BYTES;
        ptx::mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
        if (!Cfg::TA) {
          ptx::tma_load_2d(sa, &S.tmA, &full_bar[stage], S.a_c0 + k0, S.a_c1 + m0);
        } else {
#pragma unroll
          for (int i = 0; i < UMMA_BM / 64; ++i)
            ptx::tma_load_2d(sa + i * 8192, &S.tmA, &full_bar[stage], S.a_c0 + m0 + 64 * i, S.a_c1 + k0);
        }
        if (!Cfg::TB) {
#pragma unroll
          for (int g = 0; g < G; ++g)
            ptx::tma_load_2d(sb + g * BNG * 128, &S.tmB, &full_bar[stage], S.b_c0 + k0,
                             S.b_c1 + g * P.gate_stride + n0);
        } else {
#pragma unroll
          for (int g = 0; g < G; ++g)
#pragma unroll
            for (int i = 0; i < BNG / 64; ++i)
              ptx::tma_load_2d(sb + (g * (BNG / 64) + i) * 8192, &S.tmB, &full_bar[stage],
                               S.b_c0 + g * P.gate_stride + n0 + 64 * i, S.b_c1 + k0);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && nchunks > 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(UMMA_BM, BN, Cfg::TA ? 1 : 0, Cfg::TB ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int kc = 0; kc < nchunks; ++kc) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES);
        const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
        for (int kk = 0; kk < UMMA_BK / 16; ++kk) {
          const uint64_t da = Cfg::TA ? ptx::make_smem_desc(sa + kk * 2048, 8192, 1024)
                                      : ptx::make_smem_desc(sa + kk * 32, 16, 1024);
          const uint64_t db = Cfg::TB ? ptx::make_smem_desc(sb + kk * 2048, 8192, 1024)
                                      : ptx::make_smem_desc(sb + kk * 32, 16, 1024);
          ptx::umma_bf16(tmem_base, da, db, idesc, (kc > 0 || kk > 0) ? 1u : 0u);
        }
        ptx::umma_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs retire
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      ptx::umma_commit(tmem_full_bar);
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int q = warp & 3;          // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;  // the two warps of a quadrant take alternate 16-column chunks
    const int row = m0 + q * 32 + lane;
    if (nchunks > 0) {
      ptx::mbar_wait(tmem_full_bar, 0);
      ptx::tc_fence_after();
    }
#pragma unroll 1
    for (int c = half; c < BNG / 16; c += 2) {
      const int col0 = n0 + c * 16;
      if (col0 >= P.N) break;  // warp uniform
      float acc[G][16];
      if (nchunks > 0) {
#pragma unroll
        for (int g = 0; g < G; ++g)
          ptx::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * BNG + c * 16), acc[g]);
        ptx::tmem_ld_wait();
      } else {
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[g][i] = 0.f;
      }
      if (row < P.M) Epi::template apply<16>(P.epi, row, col0, min(16, P.N - col0), acc);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

}  // namespace ipn
