// tcgen05 (UMMA) GEMM core, "lanes = output columns" formulation.
//
//   D[row, g, col] = sum_seg sum_k X_seg[row, k] * W_seg[g*gate_stride + col, k]
//
// The WEIGHT-side operand W feeds the M dimension of the MMA (128 TMEM lanes = 128 output columns),
// the ACTIVATION-side operand X feeds the N dimension (BR output rows = BR TMEM columns per gate), so in
// the epilogue a warp's 32 lanes own 32 consecutive output columns of the same row: all global traffic
// of the epilogue is coalesced without shared-memory staging (see epilogues.cuh).  Gate g of a recurrent
// cell is a separate accumulator (TMEM columns [g*BR, (g+1)*BR)) fed by its own W sub-tile, so one thread
// holds r, z, n (or i, f, g, o) of the same hidden unit.
//
// warp 0 = TMA producer (cp.async.bulk.tensor.2d, SWIZZLE_128B, mbarrier expect_tx), warp 1 = MMA issuer
// (single thread, tcgen05.mma kind::f16, fp32 accumulate in TMEM) + TMEM owner, warps 2..9 = epilogue
// (tcgen05.ld 32x32b.x16).  Both operands may be K-major or MN-major independently, so forward (NT), dgrad
// (NN) and wgrad (TN) GEMMs run without materialised transposes.  Up to two K segments and two
// independent problems per launch; split-K across CTAs for the weight gradients.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace ipn {

constexpr int UMMA_BC = 128;       // output columns per tile (TMEM lanes, MMA M)
constexpr int UMMA_BK = 64;        // bf16 elements per stage along K (= 128 bytes, one swizzle row)
constexpr int UMMA_THREADS = 320;  // default: warp0 TMA, warp1 MMA, warps 2..9 epilogue (2 warps per TMEM lane quadrant)

struct UmmaSeg {
  alignas(64) CUtensorMap tmW;
  alignas(64) CUtensorMap tmX;
  int K;
  int w_c0, w_c1;  // base coordinates (inner, outer) added to the tile coordinates
  int x_c0, x_c1;
};

template <class Epi>
struct UmmaProblem {
  UmmaSeg seg[2];
  int nseg;
  int M, N;         // M = output rows (X side), N = output columns per gate (W side)
  int gate_stride;  // row (K-major W) / column (MN-major W) distance between gates
  typename Epi::Params epi;
};

template <class Epi>
struct UmmaBatch {
  UmmaProblem<Epi> p[2];
  int split_k;
};

// G gates, BR output rows per tile, TW / TX: W / X operand is MN-major (row-major [K, cols|rows])
template <int G_, int BR_, bool TW_, bool TX_, int MAX_SMEM_KB = 110, int EPI_WARPS_ = 8>
struct UmmaCfg {
  static constexpr int EPI_WARPS = EPI_WARPS_;            // multiple of 4 (warps per TMEM lane quadrant x 4)
  static constexpr int THREADS = 64 + 32 * EPI_WARPS_;
  static constexpr int PARTS = EPI_WARPS_ / 4;            // warps sharing a lane quadrant take every PARTS-th chunk
  static constexpr int G = G_;
  static constexpr int BR = BR_;
  static constexpr bool TW = TW_;
  static constexpr bool TX = TX_;
  static constexpr int W_BYTES = G_ * UMMA_BC * UMMA_BK * 2;
  static constexpr int X_BYTES = BR_ * UMMA_BK * 2;
  static constexpr int STAGE_BYTES = W_BYTES + X_BYTES;
  static constexpr int STAGES_RAW = (MAX_SMEM_KB * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : (STAGES_RAW < 2 ? 2 : STAGES_RAW);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 512 /*barriers*/;
  static constexpr int ACC_COLS = G_ * BR_;
  static constexpr uint32_t TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
  static constexpr int CTAS_PER_SM = (2 * SMEM_BYTES <= 226 * 1024 && 2 * TMEM_COLS <= 512) ? 2 : 1;
  static_assert(BR_ % 16 == 0 && BR_ >= 16 && BR_ <= 256, "UMMA N must be a multiple of 16 in [16,256] for M=128");
  static_assert(ACC_COLS <= 512, "accumulators exceed TMEM");
  static_assert(!TX_ || (BR_ % 64 == 0), "MN-major X needs 64-wide blocks");
  static_assert(SMEM_BYTES <= 227 * 1024, "stage configuration exceeds shared memory");
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// optional per-CTA phase timestamps (diagnostics): Epi::dbg_buf(params) returns a device buffer or nullptr
template <class Epi>
__device__ __forceinline__ void dbg_stamp(const typename Epi::Params& p, int slot) {
  unsigned long long* buf = Epi::dbg_buf(p);
  if (buf != nullptr) {
    const long long cta = (long long)blockIdx.x + (long long)gridDim.x * (blockIdx.y + (long long)gridDim.y * blockIdx.z);
    buf[cta * 8 + slot] = gtimer();
  }
}

template <class Cfg, class Epi>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::CTAS_PER_SM)
umma_gemm_kernel(const __grid_constant__ UmmaBatch<Epi> batch) {
  static_assert(Epi::G == Cfg::G, "epilogue / tile gate count mismatch");
  constexpr int G = Cfg::G, BR = Cfg::BR, STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* st_full = tmem_full_bar + 1;   // epilogue-input staging ring (8 chunk buffers max)
  uint64_t* st_empty = st_full + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(st_empty + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int prob = blockIdx.z / batch.split_k;
  const int ksplit = blockIdx.z - prob * batch.split_k;
  const UmmaProblem<Epi>& P = batch.p[prob];
  const int c0 = blockIdx.x * UMMA_BC;  // first output column of this tile
  const int r0 = blockIdx.y * BR;       // first output row

  // k-chunk range handled by this CTA (over the concatenation of all segments)
  const int chunks_seg0 = (P.seg[0].K + UMMA_BK - 1) / UMMA_BK;
  const int chunks_total = chunks_seg0 + (P.nseg > 1 ? (P.seg[1].K + UMMA_BK - 1) / UMMA_BK : 0);
  const int per = (chunks_total + batch.split_k - 1) / batch.split_k;
  const int kc_begin = ksplit * per;
  const int kc_end = min(chunks_total, kc_begin + per);
  const int nchunks = kc_end - kc_begin;

  if (threadIdx.x == 0) dbg_stamp<Epi>(P.epi, 0);
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        ptx::mbar_init(&full_bar[s], 1);
        ptx::mbar_init(&empty_bar[s], 1);
      }
      ptx::mbar_init(tmem_full_bar, 1);
      for (int s = 0; s < 8; ++s) {
        ptx::mbar_init(&st_full[s], 1);
        ptx::mbar_init(&st_empty[s], 4);  // the 4 lane-quadrant warps that consume a chunk
      }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    ptx::tmem_relinquish();
  } else if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&P.seg[0].tmW);
    ptx::prefetch_tmap(&P.seg[0].tmX);
    if (P.nseg > 1) {
      ptx::prefetch_tmap(&P.seg[1].tmW);
      ptx::prefetch_tmap(&P.seg[1].tmX);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) dbg_stamp<Epi>(P.epi, 1);

  // Epilogue-input staging: once the accumulators are complete the mainloop stages are dead, so warp 0
  // refills them with the epilogue's input rows (bulk copies of up to 128 columns per row) chunk by chunk
  // (16 rows), and the epilogue warps read their operands from shared memory instead of stalling on
  // global loads.  A ring of nbuf chunk buffers with full/empty mbarriers handles chunks that do not all fit.
  bool stage_on = false;
  int st_off[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  int st_na = 0, st_chunk_bytes = 0, st_nbuf = 1;
  StageArr st_arr[Epi::NARR > 0 ? 8 : 1];
  if constexpr (Epi::NARR > 0) {
    if (Epi::stage_on(P.epi) && nchunks > 0) {
      st_na = Epi::stage_arrays(P.epi, st_arr);
#pragma unroll
      for (int a = 0; a < Epi::NARR; ++a) st_off[a + 1] = st_off[a] + (a < st_na ? 16 * 128 * st_arr[a].eb : 0);
      st_chunk_bytes = (st_off[Epi::NARR] + 127) & ~127;
      st_nbuf = min(8, (STAGES * Cfg::STAGE_BYTES) / st_chunk_bytes);
      stage_on = st_nbuf >= 1;
    }
  }

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
        uint8_t* sw = smem + stage * Cfg::STAGE_BYTES;
        uint8_t* sx = sw + Cfg::W_BYTES;
        ptx::mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
#pragma unroll
        for (int g = 0; g < G; ++g) {
          if (!Cfg::TW) {
            ptx::tma_load_2d(sw + g * 16384, &S.tmW, &full_bar[stage], S.w_c0 + k0, S.w_c1 + g * P.gate_stride + c0);
          } else {
#pragma unroll
            for (int i = 0; i < UMMA_BC / 64; ++i)
              ptx::tma_load_2d(sw + g * 16384 + i * 8192, &S.tmW, &full_bar[stage],
                               S.w_c0 + g * P.gate_stride + c0 + 64 * i, S.w_c1 + k0);
          }
        }
        if (!Cfg::TX) {
          ptx::tma_load_2d(sx, &S.tmX, &full_bar[stage], S.x_c0 + k0, S.x_c1 + r0);
        } else {
#pragma unroll
          for (int i = 0; i < BR / 64; ++i)
            ptx::tma_load_2d(sx + i * 8192, &S.tmX, &full_bar[stage], S.x_c0 + r0 + 64 * i, S.x_c1 + k0);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      dbg_stamp<Epi>(P.epi, 2);  // all TMA loads issued
    }
    if constexpr (Epi::NARR > 0) {
      if (stage_on) {
        __syncwarp();
        ptx::mbar_wait(tmem_full_bar, 0);  // every MMA has finished reading the stage buffers
        const int ncols = min(UMMA_BC, P.N - c0);
        int row_bytes = 0;
#pragma unroll
        for (int a = 0; a < Epi::NARR; ++a)
          if (a < st_na) row_bytes += ncols * st_arr[a].eb;
        for (int c = 0; c < BR / 16; ++c) {
          const int row0 = r0 + c * 16;
          if (row0 >= P.M) break;
          const int nvr = min(16, P.M - row0);
          const int b = c % st_nbuf;
          if (c >= st_nbuf) ptx::mbar_wait(&st_empty[b], ((c / st_nbuf) - 1) & 1);
          if (lane == 0) ptx::mbar_arrive_expect_tx(&st_full[b], (uint32_t)(nvr * row_bytes));
          __syncwarp();
          uint8_t* dst = smem + b * st_chunk_bytes;
          if (lane < nvr) {
#pragma unroll
            for (int a = 0; a < Epi::NARR; ++a) {
              if (a < st_na) {
                const StageArr& A = st_arr[a];
                const long long row = row0 + lane;
                const long long src_row = A.gather != nullptr ? (long long)A.gather[row + A.gather_add] : row + A.row_add;
                ptx::bulk_copy_g2s(dst + st_off[a] + lane * 128 * A.eb,
                                   A.ptr + src_row * A.row_stride_bytes + (long long)c0 * A.eb, (uint32_t)(ncols * A.eb),
                                   &st_full[b]);
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the loop (warp-uniform descriptor arithmetic); one elected lane issues.
    if (nchunks > 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(UMMA_BC, BR, Cfg::TW ? 1 : 0, Cfg::TX ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t smem_u = ptx::smem_u32(smem);
      const uint64_t dw_base = Cfg::TW ? ptx::make_smem_desc(smem_u, 8192, 1024) : ptx::make_smem_desc(smem_u, 16, 1024);
      const uint64_t dx_base = Cfg::TX ? ptx::make_smem_desc(smem_u + Cfg::W_BYTES, 8192, 1024)
                                       : ptx::make_smem_desc(smem_u + Cfg::W_BYTES, 16, 1024);
      constexpr uint32_t W_KK = (Cfg::TW ? 2048 : 32) >> 4, X_KK = (Cfg::TX ? 2048 : 32) >> 4;
      for (int kc = 0; kc < nchunks; ++kc) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint64_t so = (uint64_t)((stage * Cfg::STAGE_BYTES) >> 4);
        if (ptx::elect_one()) {
#pragma unroll
          for (int kk = 0; kk < UMMA_BK / 16; ++kk) {
            const uint64_t dx = dx_base + so + (uint64_t)(kk * X_KK);
#pragma unroll
            for (int g = 0; g < G; ++g) {
              const uint64_t dw = dw_base + so + (uint64_t)(g * (16384 >> 4) + kk * W_KK);
              ptx::umma_bf16(tmem_base + (uint32_t)(g * BR), dw, dx, idesc, (kc > 0 || kk > 0) ? 1u : 0u);
            }
          }
          ptx::umma_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs retire
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (ptx::elect_one()) ptx::umma_commit(tmem_full_bar);
      __syncwarp();
      if (lane == 0) dbg_stamp<Epi>(P.epi, 3);  // all MMAs issued
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int q = warp & 3;            // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;  // the PARTS warps of a quadrant take every PARTS-th 16-row chunk
    const int col = c0 + q * 32 + lane;
    const bool col_ok = col < P.N;
    typename Epi::Col cc;
    if (col_ok) Epi::col_init(P.epi, col, cc);  // overlaps the mainloop
    if (nchunks > 0) {
      ptx::mbar_wait(tmem_full_bar, 0);
      ptx::tc_fence_after();
    }
    if (threadIdx.x == 64) dbg_stamp<Epi>(P.epi, 4);  // accumulators complete
#pragma unroll 1
    for (int c = half; c < BR / 16; c += Cfg::PARTS) {
      const int row0 = r0 + c * 16;
      if (row0 >= P.M) break;  // warp uniform
      float acc[G][16];
      if (nchunks > 0) {
#pragma unroll
        for (int g = 0; g < G; ++g)
          ptx::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * BR + c * 16), acc[g]);
        ptx::tmem_ld_wait();
      } else {
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[g][i] = 0.f;
      }
      const int nv = min(16, P.M - row0);
      if constexpr (Epi::NARR > 0) {
        if (stage_on) {
          const int b = c % st_nbuf;
          ptx::mbar_wait(&st_full[b], (c / st_nbuf) & 1);
          const char* chunk = reinterpret_cast<const char*>(smem) + b * st_chunk_bytes;
          typename Epi::template Pre<8> pre[2];
          if (col_ok) {
#pragma unroll
            for (int h = 0; h < 2; ++h) Epi::template preload_smem<8>(P.epi, cc, chunk, q * 32 + lane, h * 8, pre[h]);
          }
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&st_empty[b]);  // this warp is done with the chunk buffer
          if (col_ok) {
            float a8[G][8];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
              for (int g = 0; g < G; ++g)
#pragma unroll
                for (int i = 0; i < 8; ++i) a8[g][i] = acc[g][h * 8 + i];
              const int nvh = nv - h * 8;
              if (nvh > 0) Epi::template applyT<8>(P.epi, cc, col, row0 + h * 8, min(8, nvh), a8, pre[h]);
            }
          }
          continue;
        }
      }
      if (col_ok) {
        // two phases of 8 rows (bounds the registers held by an epilogue's load phase)
        float a8[G][8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int g = 0; g < G; ++g)
#pragma unroll
            for (int i = 0; i < 8; ++i) a8[g][i] = acc[g][h * 8 + i];
          const int nvh = nv - h * 8;
          if (nvh > 0) Epi::template applyT<8>(P.epi, cc, col, row0 + h * 8, min(8, nvh), a8);
        }
      }
    }
  }
  if (threadIdx.x == 64) dbg_stamp<Epi>(P.epi, 5);  // first epilogue warp done
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) dbg_stamp<Epi>(P.epi, 6);  // all warps done
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// =============================================================================================
// Persistent variant (ipn_gemm): one CTA per SM loops over output tiles.  The TMA ring never drains
// between tiles, accumulators are double-buffered in TMEM (tile i's epilogue overlaps tile i+1's
// mainloop), and there is no wave quantisation.  Tiles are ordered column-tile fastest so that the CTAs
// running at the same time share the X row tile (streamed from HBM once) while W stays L2-resident.
// =============================================================================================
template <int BR_, bool TW_, bool TX_, bool PAIR_ = false>
struct UmmaPCfg {
#ifndef IPN_PGEMM_EPI_WARPS
#define IPN_PGEMM_EPI_WARPS 16
#endif
  static constexpr int EPI_WARPS = IPN_PGEMM_EPI_WARPS;   // 4 per TMEM lane quadrant: the epilogue of K <= 1024 products is latency bound
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  static constexpr int PARTS = EPI_WARPS / 4;
  static constexpr int G = 1;
  static constexpr int BR = BR_;             // output rows of the tile (MMA N)
  static constexpr bool TW = TW_;
  static constexpr bool TX = TX_;
  static constexpr bool PAIR = PAIR_;        // cta_group::2: the pair computes 256 output columns x BR rows,
  static constexpr int XR = PAIR_ ? BR_ / 2 : BR_;   // each CTA stages its own W tile and HALF of the X tile
  static constexpr int W_BYTES = UMMA_BC * UMMA_BK * 2;
  static constexpr int X_BYTES = XR * UMMA_BK * 2;
  static constexpr int STAGE_BYTES = W_BYTES + X_BYTES;
  static constexpr int STAGES_RAW = (200 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 512;
  static constexpr uint32_t TMEM_COLS = 2 * BR_ <= 32 ? 32 : 2 * BR_ <= 64 ? 64 : 2 * BR_ <= 128 ? 128 : 2 * BR_ <= 256 ? 256 : 512;
  static_assert(BR_ % 16 == 0 && BR_ >= 16 && BR_ <= 256, "UMMA N must be a multiple of 16 in [16,256]");
  static_assert(!TX_ || (XR % 64 == 0), "MN-major X needs 64-wide blocks");
  static_assert(!PAIR_ || BR_ % 32 == 0, "pair tiles split the X rows in two");
};

// ntx counts column tiles of 128 (PAIR: column-tile PAIRS of 256); grid.x CTAs (PAIR: clusters of 2)
template <class Cfg, class Epi>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
umma_gemm_persist_kernel(const __grid_constant__ UmmaBatch<Epi> batch, int ntx, int nty, int ntiles, int row_fastest) {
  static_assert(Epi::G == 1, "the persistent GEMM serves single-accumulator epilogues");
  constexpr int BR = Cfg::BR, STAGES = Cfg::STAGES;
  constexpr bool PAIR = Cfg::PAIR;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? ptx::cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;        // tile stream this CTA (pair) follows
  const int nworkers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
      for (int b = 0; b < 2; ++b) {
        ptx::mbar_init(&tmem_full[b], 1);
        ptx::mbar_init(&tmem_empty[b], PAIR ? 2 * Cfg::EPI_WARPS : Cfg::EPI_WARPS);
      }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    if (PAIR) { ptx::tmem_alloc_pair<Cfg::TMEM_COLS>(tmem_slot); ptx::tmem_relinquish_pair(); }
    else { ptx::tmem_alloc<Cfg::TMEM_COLS>(tmem_slot); ptx::tmem_relinquish(); }
  } else if (warp == 0 && lane == 0) {
    for (int pi = 0; pi < 2; ++pi) {
      if (batch.p[pi].nseg <= 0) continue;
      ptx::prefetch_tmap(&batch.p[pi].seg[0].tmW);
      ptx::prefetch_tmap(&batch.p[pi].seg[0].tmX);
      if (batch.p[pi].nseg > 1) { ptx::prefetch_tmap(&batch.p[pi].seg[1].tmW); ptx::prefetch_tmap(&batch.p[pi].seg[1].tmX); }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (PAIR) ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile -> (problem, k-split, column tile, row tile) and its k-chunk range.  Both CTAs of a pair derive the
  // SAME k range (validity is decided per pair) so that their pipelines stay in lock step.
  struct Tile { int prob, c0, r0, kc_begin, kc_end, chunks_seg0; };
  auto decode = [&](int t) {
    Tile o;
    // column tile fastest (default): concurrent CTAs share the X row tile, W stays in L2; row tile fastest when
    // the streamed (large) operand sits on the W side (roles swapped, EpiBlockedP)
    int tx, ty, z;
    if (row_fastest) { ty = t % nty; const int rest = t / nty; tx = rest % ntx; z = rest / ntx; }
    else { tx = t % ntx; const int rest = t / ntx; ty = rest % nty; z = rest / nty; }
    o.prob = z / batch.split_k;
    const int ksplit = z - o.prob * batch.split_k;
    const UmmaProblem<Epi>& P = batch.p[o.prob];
    const int pc0 = tx * (PAIR ? 2 * UMMA_BC : UMMA_BC);
    o.c0 = pc0 + (int)rank * UMMA_BC;
    o.r0 = ty * BR;
    o.chunks_seg0 = (P.seg[0].K + UMMA_BK - 1) / UMMA_BK;
    const int total = o.chunks_seg0 + (P.nseg > 1 ? (P.seg[1].K + UMMA_BK - 1) / UMMA_BK : 0);
    const int per = (total + batch.split_k - 1) / batch.split_k;
    o.kc_begin = ksplit * per;
    o.kc_end = min(total, o.kc_begin + per);
    if (pc0 >= P.N || o.r0 >= P.M) o.kc_end = o.kc_begin;   // tile outside this problem: nothing to do
    return o;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = worker; t < ntiles; t += nworkers) {
        const Tile ti = decode(t);
        const UmmaProblem<Epi>& P = batch.p[ti.prob];
        const int xr0 = ti.r0 + (PAIR ? (int)rank * Cfg::XR : 0);   // this CTA's part of the X tile
        for (int kc = ti.kc_begin; kc < ti.kc_end; ++kc) {
          const int si = (kc >= ti.chunks_seg0) ? 1 : 0;
          const UmmaSeg& S = P.seg[si];
          const int k0 = (si ? kc - ti.chunks_seg0 : kc) * UMMA_BK;
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sw = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sx = sw + Cfg::W_BYTES;
          if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES * (PAIR ? 2 : 1));
          auto ld = [&](void* dst, const CUtensorMap* m, int c0, int c1) {
            if (PAIR) ptx::tma_load_2d_pair(dst, m, &full_bar[stage], c0, c1);
            else ptx::tma_load_2d(dst, m, &full_bar[stage], c0, c1);
          };
          if (!Cfg::TW) {
            ld(sw, &S.tmW, S.w_c0 + k0, S.w_c1 + ti.c0);
          } else {
#pragma unroll
            for (int i = 0; i < UMMA_BC / 64; ++i) ld(sw + i * 8192, &S.tmW, S.w_c0 + ti.c0 + 64 * i, S.w_c1 + k0);
          }
          if (!Cfg::TX) {
            ld(sx, &S.tmX, S.x_c0 + k0, S.x_c1 + xr0);
          } else {
#pragma unroll
            for (int i = 0; i < Cfg::XR / 64; ++i) ld(sx + i * 8192, &S.tmX, S.x_c0 + xr0 + 64 * i, S.x_c1 + k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp loops, one elected lane issues; leader CTA of a pair) ==========
    if (leader) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(PAIR ? 2 * UMMA_BC : UMMA_BC, BR, Cfg::TW ? 1 : 0, Cfg::TX ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t smem_u = ptx::smem_u32(smem);
      const uint64_t dw_base = Cfg::TW ? ptx::make_smem_desc(smem_u, 8192, 1024) : ptx::make_smem_desc(smem_u, 16, 1024);
      const uint64_t dx_base = Cfg::TX ? ptx::make_smem_desc(smem_u + Cfg::W_BYTES, 8192, 1024)
                                       : ptx::make_smem_desc(smem_u + Cfg::W_BYTES, 16, 1024);
      constexpr uint32_t W_KK = (Cfg::TW ? 2048 : 32) >> 4, X_KK = (Cfg::TX ? 2048 : 32) >> 4;
      int it = 0;
      for (int t = worker; t < ntiles; t += nworkers, ++it) {
        const Tile ti = decode(t);
        const int b = it & 1, n = it >> 1;
        ptx::mbar_wait(&tmem_empty[b], (n & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t dcol = tmem_base + (uint32_t)(b * BR);
        const int nchunks = ti.kc_end - ti.kc_begin;
        for (int kc = 0; kc < nchunks; ++kc) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint64_t so = (uint64_t)((stage * Cfg::STAGE_BYTES) >> 4);
          if (ptx::elect_one()) {
#pragma unroll
            for (int kk = 0; kk < UMMA_BK / 16; ++kk) {
              const uint64_t dw = dw_base + so + (uint64_t)(kk * W_KK), dx = dx_base + so + (uint64_t)(kk * X_KK);
              if (PAIR) ptx::umma_bf16_pair(dcol, dw, dx, idesc, (kc > 0 || kk > 0) ? 1u : 0u);
              else ptx::umma_bf16(dcol, dw, dx, idesc, (kc > 0 || kk > 0) ? 1u : 0u);
            }
            if (PAIR) ptx::umma_commit_pair(&empty_bar[stage]);
            else ptx::umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (ptx::elect_one()) {
          if (nchunks > 0) {
            if (PAIR) ptx::umma_commit_pair(&tmem_full[b]);
            else ptx::umma_commit(&tmem_full[b]);
          } else {
            ptx::mbar_arrive(&tmem_full[b]);
            if (PAIR) ptx::mbar_arrive_remote(&tmem_full[b], 1);
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;
    const int part = (warp - 2) >> 2;
    int it = 0;
    for (int t = worker; t < ntiles; t += nworkers, ++it) {
      const Tile ti = decode(t);
      const UmmaProblem<Epi>& P = batch.p[ti.prob];
      const int b = it & 1, n = it >> 1;
      const int col = ti.c0 + q * 32 + lane;
      const bool tile_ok = ti.c0 < P.N && ti.r0 < P.M;
      const bool col_ok = col < P.N && tile_ok;
      typename Epi::Col cc;
      if (col_ok) Epi::col_init(P.epi, col, cc);
      const int nchunks = ti.kc_end - ti.kc_begin;
      ptx::mbar_wait(&tmem_full[b], n & 1);
      ptx::tc_fence_after();
      if (tile_ok) {
#pragma unroll 1
        for (int c = part; c < BR / 16; c += Cfg::PARTS) {
          const int row0 = ti.r0 + c * 16;
          if (row0 >= P.M) break;  // warp uniform
          float acc[1][16];
          if (nchunks > 0) {
            ptx::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * BR + c * 16), acc[0]);
            ptx::tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[0][i] = 0.f;
          }
          const int nv = min(16, P.M - row0);
          if (col_ok) {
            float a8[1][8];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
              for (int i = 0; i < 8; ++i) a8[0][i] = acc[0][h * 8 + i];
              const int nvh = nv - h * 8;
              if (nvh > 0) Epi::template applyT<8>(P.epi, cc, col, row0 + h * 8, min(8, nvh), a8);
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) ptx::mbar_arrive(&tmem_empty[b]);
        else ptx::mbar_arrive_remote(&tmem_empty[b], 0);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (PAIR) ptx::cluster_sync_all();
  if (warp == 1) {
    ptx::tc_fence_after();
    if (PAIR) ptx::tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem_base);
    else ptx::tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

}  // namespace ipn
