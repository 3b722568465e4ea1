// Persistent LSTM layer kernels (forward and backward) for the 384-tick AnticipationRNN stacks
// (AnticipationRNN/anticipation_rnn_gauss_reg_model.py:14-39: torch.nn.LSTM, one layer per call, zero initial state).
//
// One launch runs ALL timesteps of a layer.  A thread-block CLUSTER of NC = H/64 CTAs owns a tile of 128 batch rows
// for the whole sequence; CTA j of the cluster owns the hidden units [64j, 64j+64):
//   * its slice of W_hh -- the (i, f, g, o) rows of those 64 units, 4 x 64 x H bf16 = 128 KB at H = 256 -- is loaded
//     ONCE by TMA and stays resident in shared memory for all timesteps;
//   * forward, per step: acc[128 rows, 4 x 64 gate columns] = h_{t-1}[128, H] . W_slice^T on tcgen05 (M = 128, N = 256,
//     K = H, accumulator in TMEM), gate math in the epilogue with the cell state kept in REGISTERS across steps, and the
//     new h_t slice (128 x 64 bf16) written straight into the A-operand tile of EVERY CTA of the cluster through
//     distributed shared memory (st.shared::cluster), signalled with cluster-scope mbarrier arrives -- one exchange per
//     step, no global-memory round trip on the serial chain;
//   * backward, per step: the CTA differentiates its own 64 units (dG = [di, df, dg, do], written to the A tile and
//     TMA-stored to dP for the hoisted weight-gradient GEMMs), multiplies dG[128, 4 x 64] by its W_hh rows (K-split of
//     dh_{t-1} = dG . W_hh: a partial sum over its 256 gate rows for ALL H columns) and reduce-scatters the partial
//     sums to the owners of the columns through distributed shared memory (st.async with the bytes accounted on the
//     receiver's mbarrier; bf16 partials, fp32 sum).
// Forward exchange: the 16 KB h_t slice goes to every peer as ONE bulk DSMEM copy (cp.async.bulk shared::cta ->
// shared::cluster, complete_tx on the peer's barrier) issued by the store warp once the 16 epilogue warps have written
// the local copy; measured (tests/dev/lstm_persist_time.py) remote st.shared::cluster from all threads ran at ~7 B/clk/SM.
// Per-element tensors that only these kernels touch use the blocked layout of gru_persist.cuh (coalesced 16-byte
// vectors): the input projection P (4 arrays, produced directly by ipn_lstm_inproj_blocked) and the saved state
// (5 arrays: i, f, g, o, c_t).  h (row-major, read by the hoisted GEMMs) leaves through TMA stores from the A tile.
#include "gru_persist.cuh"
#include <stdlib.h>

namespace ipn {

constexpr int LP_ROWS = 128;
constexpr int LP_THREADS = 640;               // producer, MMA, store, (idle), 16 epilogue warps
constexpr int LP_KB = LP_ROWS * 128;          // A k-block: 128 rows x 64 bf16 (SWIZZLE_128B)
constexpr int LP_WKB = 4 * 64 * 128;          // W k-block: (4 gates x 64 units) rows x 64 k
constexpr int LP_ARR = 5;                     // saved per step: i, f, g, o, c

struct LstmPersistFwd {
  alignas(64) CUtensorMap tmW;   // W_hh [4H, H] as {k: H, unit: H (stride H), gate: 4 (stride H*H)}, box {64, 64, 4}
  alignas(64) CUtensorMap tmH;   // hseq [(T+1)*B, H], box 64 x 128 (initial-state load, per-step stores)
  alignas(64) CUtensorMap tmY;   // y [T*B, ld_y], box 64 x 128 (stores)
  const uint4* Pblk;             // blocked [T*B, 4, H]
  uint4* gates;                  // blocked [T*B, 5, H], nullable
  float* cseq;                   // fp32 [(T+1)*B, H]: slot s_begin read, slot s_end written
  int T, B;
  int s_begin, s_end;
  int has_y, y_col0, y_reverse_time;
  unsigned long long* timing;    // per-CTA cycle counters (ipn_dbg_set_timing_buffer), normally null
};

struct LstmPersistBwd {
  alignas(64) CUtensorMap tmW;   // W_hh [4H, H] as {64 n, 4H k (stride H), H/64 n-blocks (stride 64)}, box {64, 64, H/64} (MN-major B)
  alignas(64) CUtensorMap tmDP;  // dP [T*B, 4H], box 64 x 128 (stores)
  alignas(64) CUtensorMap tmDY;  // dY [T*B, ld_dy], box 64 x 128 (loads)
  const uint4* gates;            // blocked [T*B, 5, H]
  int T, B;
  int y_col0, y_reverse_time;
  unsigned long long* timing;
};

__device__ __forceinline__ void st_cluster_u4(uint32_t caddr, const uint4& u) { ptx::st_cluster_v4(caddr, u.x, u.y, u.z, u.w); }

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int NC>
static inline int lpf_smem_bytes() { return NC * LP_WKB + NC * LP_KB + 16 * 8 + 16; }   // 9 barriers + the TMEM slot

template <int NC, bool SAVE>
__global__ void __launch_bounds__(LP_THREADS, 1) lstm_persist_fwd_kernel(const __grid_constant__ LstmPersistFwd p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int H = NC * 64;
  constexpr int VPR = H / 8;                 // 16-byte vectors per row of one array
  const int T = p.T, B = p.B;
  const int NS = p.s_end - p.s_begin;
  const uint32_t rank = ptx::cluster_ctarank();
  const int rbase = (blockIdx.x / NC) * LP_ROWS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  uint8_t* sW = smem;
  uint8_t* sA = sW + NC * LP_WKB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + NC * LP_KB);
  uint64_t* w_full = bars;          // W slice landed (once)
  uint64_t* a0_full = bars + 1;     // initial h tile landed (once)
  uint64_t* a_ready = bars + 2;     // [NC] k-block kb of h_t written by the 16 epilogue warps of CTA kb
  uint64_t* mma_all = bars + 6;     // the MMAs of this step are complete in ALL CTAs of the cluster (multicast commits)
  uint64_t* st_free = bars + 7;     // the TMA store of the previous h_t has read this CTA's k-block
  uint64_t* tmem_full = bars + 8;   // THIS CTA's MMAs of the step are complete (local commit: no cluster round trip)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  if ((ptx::smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("inpaintnet_b200: lstm_persist_fwd: shared memory base not 1024-byte aligned\n");
    __trap();
  }
  if (warp == 1) {
    if (lane == 0) {
      ptx::mbar_init(w_full, 1);
      ptx::mbar_init(a0_full, 1);
      // own k-block: the 16 epilogue warps arrive; a peer's k-block: one local arrive.expect_tx + the bytes of the
      // peer's bulk copy (complete_tx)
      for (int k = 0; k < NC; ++k) ptx::mbar_init(&a_ready[k], k == (int)rank ? 16 : 1);
      ptx::mbar_init(mma_all, NC);
      ptx::mbar_init(st_free, 1);
      ptx::mbar_init(tmem_full, 1);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc<256>(tmem_slot);
    ptx::tmem_relinquish();
  } else if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.tmW);
    ptx::prefetch_tmap(&p.tmH);
    if (p.has_y) ptx::prefetch_tmap(&p.tmY);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();   // every CTA's barriers are initialised before any remote arrive / multicast commit
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    ptx::setmaxnreg_dec<56>();
    if (warp == 0) {
      // ===================== producer: resident W slice, initial state, L2 prefetch of the input projection ==========
      if (lane == 0) {
        ptx::mbar_arrive_expect_tx(w_full, NC * LP_WKB);
        for (int kb = 0; kb < NC; ++kb) ptx::tma_load_3d(sW + kb * LP_WKB, &p.tmW, w_full, kb * 64, (int)rank * 64, 0);
        ptx::mbar_arrive_expect_tx(a0_full, NC * LP_KB);
        for (int kb = 0; kb < NC; ++kb) ptx::tma_load_2d(sA + kb * LP_KB, &p.tmH, a0_full, kb * 64, p.s_begin * B + rbase);
      }
    } else if (warp == 1) {
      // ===================== MMA issuer: 4 x NC instructions per step (M 128, N 256, K 16 each) =====================
      constexpr uint32_t idesc = ptx::make_idesc_bf16(128, 256, 0, 0);
      const uint64_t descA0 = ptx::make_smem_desc(ptx::smem_u32(sA), 16, 1024);
      const uint64_t descW0 = ptx::make_smem_desc(ptx::smem_u32(sW), 16, 1024);
      ptx::mbar_wait(w_full, 0);
      const bool tm = p.timing != nullptr;
      long long w_own = 0, w_peer = 0;
      const long long t_begin = clock64();
      for (int t = 0; t < NS; ++t) {
        if (t == 0) {
          ptx::mbar_wait(a0_full, 0);
        } else {
          // own k-block first: its arrival also says that this CTA's epilogue has read the accumulator of step t-1
          const long long c0 = tm ? clock64() : 0;
          ptx::mbar_wait_cluster(&a_ready[rank], (t - 1) & 1);
          const long long c1 = tm ? clock64() : 0;
          for (int kb = 0; kb < NC; ++kb)
            if (kb != (int)rank) ptx::mbar_wait_cluster(&a_ready[kb], (t - 1) & 1);
          if (tm) { w_own += c1 - c0; w_peer += clock64() - c1; }
        }
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          // arm the peers' k-blocks of h_t BEFORE the commit below: a peer sends only after mma_all of this step
          if (t + 1 < NS) {
#pragma unroll
            for (int kb = 0; kb < NC; ++kb)
              if (kb != (int)rank) ptx::mbar_arrive_expect_tx(&a_ready[kb], LP_KB);
          }
#pragma unroll
          for (int kb = 0; kb < NC; ++kb)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              ptx::umma_bf16(tmem_base, descA0 + (uint64_t)((kb * LP_KB) >> 4) + (uint64_t)(kk * 2),
                             descW0 + (uint64_t)((kb * LP_WKB) >> 4) + (uint64_t)(kk * 2), idesc, (kb > 0 || kk > 0) ? 1u : 0u);
          ptx::umma_commit(tmem_full);                                           // own epilogue may start
          ptx::umma_commit_multicast(mma_all, (uint16_t)((1u << NC) - 1u));     // every A tile of the cluster may be overwritten
        }
        __syncwarp();
      }
      if (tm && lane == 0) {
        unsigned long long* o = p.timing + (long long)blockIdx.x * 16;
        o[0] = (unsigned long long)(clock64() - t_begin); o[1] = w_own; o[2] = w_peer;
      }
    } else if (warp == 2) {
      // ===================== store warp: own k-block of h_t -> hseq slot (+ y) =====================
      if (lane == 0) {
        const uint32_t src = ptx::smem_u32(sA) + rank * LP_KB;
        uint32_t dst[NC], dbar[NC];
#pragma unroll
        for (int pr = 0; pr < NC; ++pr) {
          dst[pr] = ptx::mapa(src, (uint32_t)pr);
          dbar[pr] = ptx::mapa(ptx::smem_u32(&a_ready[rank]), (uint32_t)pr);
        }
        for (int t = 0; t < NS; ++t) {
          const int s = p.s_begin + t;
          ptx::mbar_wait(&a_ready[rank], t & 1);
          // the exchange: this CTA's 16 KB slice of h_t goes to k-block `rank` of every peer's A tile as ONE bulk
          // copy per peer through distributed shared memory (completes on the peer's a_ready[rank]); the peers' A
          // tiles are free once the MMAs of the whole cluster are done (long since: the epilogue ran in between)
          ptx::mbar_wait(mma_all, t & 1);
          if (t + 1 < NS) {
#pragma unroll
            for (int pr = 0; pr < NC; ++pr)
              if (pr != (int)rank) ptx::bulk_copy_s2s_cluster(dst[pr], src, LP_KB, dbar[pr]);
          }
          ptx::tma_store_2d(&p.tmH, sA + rank * LP_KB, (int)rank * 64, (s + 1) * B + rbase);
          if (p.has_y) ptx::tma_store_2d(&p.tmY, sA + rank * LP_KB, p.y_col0 + (int)rank * 64, (p.y_reverse_time ? T - 1 - s : s) * B + rbase);
          ptx::bulk_commit();
          ptx::bulk_wait_read0();
          ptx::mbar_arrive(st_free);
        }
        ptx::bulk_wait0();
      }
    }
  } else {
    // ===================== epilogue warps 4..19: warp = (TMEM lane quadrant, 16-unit sub-chunk) =====================
    ptx::setmaxnreg_inc<104>();
    const int q = warp & 3;
    const int sub = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const int u0 = (int)rank * 64 + sub * 16;      // first hidden unit of this thread
    const uint32_t sA_u = ptx::smem_u32(sA) + rank * LP_KB;   // this CTA's k-block inside an A tile
    const uint32_t sw = (uint32_t)(row & 7);
    const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(sub * 16);
    // the input-projection tiles are pulled from HBM into L2 four steps ahead (one gate array per sub-chunk warp group);
    // issued from the epilogue, which is in lockstep with the chain by construction
    auto prefetch = [&](int t) {
      if (t >= NS || q != 0 || lane != 0) return;
      const long long rtp = ((long long)(p.s_begin + t) * B + rbase) >> 7;
      ptx::bulk_prefetch_l2(p.Pblk + ((rtp * 4 + sub) * VPR + rank * 8) * 128, 8 * 128 * 16);
    };
    for (int t = 0; t < 4; ++t) prefetch(t);
    const bool tm = p.timing != nullptr && threadIdx.x == 128;
    long long w_mma = 0, w_st = 0, w_work = 0, w_sig = 0;
    const long long te0 = clock64();
    float c[2][8];
    {
      const float4* cp = reinterpret_cast<const float4*>(p.cseq + ((long long)p.s_begin * B + rbase + row) * H + u0);
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const float4 a = cp[2 * v], b = cp[2 * v + 1];
        c[v][0] = a.x; c[v][1] = a.y; c[v][2] = a.z; c[v][3] = a.w;
        c[v][4] = b.x; c[v][5] = b.y; c[v][6] = b.z; c[v][7] = b.w;
      }
    }
    for (int t = 0; t < NS; ++t) {
      const int s = p.s_begin + t;
      const long long rt = ((long long)s * B + rbase) >> 7;
      prefetch(t + 4);
      const uint4* pb = p.Pblk + ((rt * 4) * VPR + (u0 >> 3)) * 128 + row;
      uint4 pv[4][2];
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int v = 0; v < 2; ++v) pv[g][v] = ldg_stream(pb + (g * VPR + v) * 128);
      const long long c0 = tm ? clock64() : 0;
      ptx::mbar_wait(tmem_full, t & 1);   // own MMAs only: the epilogue writes nothing outside this CTA
      ptx::tc_fence_after();
      const long long c1 = tm ? clock64() : 0;
      if (t > 0) ptx::mbar_wait(st_free, (t - 1) & 1);
      const long long c2 = tm ? clock64() : 0;
      uint4* gp = SAVE ? p.gates + ((rt * LP_ARR) * VPR + (u0 >> 3)) * 128 + row : nullptr;
      uint4 gsave[2][LP_ARR];
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        float acc[4][8];
#pragma unroll
        for (int g = 0; g < 4; ++g) ptx::tmem_ld8(tacc + (uint32_t)(g * 64 + v * 8), acc[g]);
        float pi[8], pf[8], pg[8], po[8], gi[8], gf[8], gg[8], go[8], hh[8];
        unpack8(pv[0][v], pi); unpack8(pv[1][v], pf); unpack8(pv[2][v], pg); unpack8(pv[3][v], po);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float i_ = fmaf(0.5f, tanh_fast(fmaf(0.5f, acc[0][k], pi[k])), 0.5f);
          const float f_ = fmaf(0.5f, tanh_fast(fmaf(0.5f, acc[1][k], pf[k])), 0.5f);
          const float g_ = tanh_fast(acc[2][k] + pg[k]);
          const float o_ = fmaf(0.5f, tanh_fast(fmaf(0.5f, acc[3][k], po[k])), 0.5f);
          const float cn = fmaf(f_, c[v][k], i_ * g_);
          c[v][k] = cn;
          gi[k] = i_; gf[k] = f_; gg[k] = g_; go[k] = o_;
          hh[k] = o_ * tanh_fast(cn);
        }
        st_shared_v4(sA_u + (uint32_t)row * 128u + ((((uint32_t)(sub * 2 + v)) ^ sw) << 4), pack8(hh));
        if (SAVE) {
          gsave[v][0] = pack8(gi); gsave[v][1] = pack8(gf); gsave[v][2] = pack8(gg); gsave[v][3] = pack8(go);
          gsave[v][4] = pack8(c[v]);
        }
      }
      const long long c3 = tm ? clock64() : 0;
      ptx::tc_fence_before();
      ptx::fence_proxy_async();   // own h_t slice (shared::cta) -> visible to the bulk copies / the TMA store / the MMAs
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&a_ready[rank]);
      // the saved state goes out AFTER the signal: nothing on the serial chain waits for these global stores
      if (SAVE) {
#pragma unroll
        for (int v = 0; v < 2; ++v)
#pragma unroll
          for (int a = 0; a < LP_ARR; ++a) stg_stream(gp + (a * VPR + v) * 128, gsave[v][a]);
      }
      if (tm) { w_mma += c1 - c0; w_st += c2 - c1; w_work += c3 - c2; w_sig += clock64() - c3; }
    }
    if (tm) {
      unsigned long long* o = p.timing + (long long)blockIdx.x * 16;
      o[4] = (unsigned long long)(clock64() - te0); o[5] = w_mma; o[6] = w_st; o[7] = w_work; o[8] = w_sig;
    }
    {   // cell state after the last processed step
      float4* cp = reinterpret_cast<float4*>(p.cseq + ((long long)p.s_end * B + rbase + row) * H + u0);
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        cp[2 * v] = make_float4(c[v][0], c[v][1], c[v][2], c[v][3]);
        cp[2 * v + 1] = make_float4(c[v][4], c[v][5], c[v][6], c[v][7]);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();   // peers' shared memory and barriers stay alive until every CTA is done
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<256>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
template <int NC>
static inline int lpb_smem_bytes() { return 4 * NC * 8192 + 4 * LP_KB + 2 * LP_KB + 16 * 8 + 16; }

template <int NC>
__global__ void __launch_bounds__(LP_THREADS, 1) lstm_persist_bwd_kernel(const __grid_constant__ LstmPersistBwd p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int H = NC * 64;
  constexpr int VPR = H / 8;
  constexpr int WG = NC * 8192;               // bytes of one gate's W block: NC n-blocks x 64 k rows x 128 B
  const int T = p.T, B = p.B;
  const uint32_t rank = ptx::cluster_ctarank();
  const int rbase = (blockIdx.x / NC) * LP_ROWS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nM = T - 1;                       // GEMM phases (the state before step 0 is the zero initial state)

  uint8_t* sW = smem;                         // 4 gate blocks (MN-major B: [n-block][k][64 n])
  uint8_t* sA = sW + 4 * WG;                  // 4 k-blocks (one per gate) of dG; between GEMM phases: partial sums received from the peers
  uint8_t* sDY = sA + 4 * LP_KB;              // 2 stages of the dY tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDY + 2 * LP_KB);
  uint64_t* w_full = bars;            // 1
  uint64_t* dy_full = bars + 1;       // [2]
  uint64_t* dy_empty = bars + 3;      // [2] count 16
  uint64_t* dg_ready = bars + 5;      // dG of this step is in the A tile (16 epilogue warps)
  uint64_t* dp_read = bars + 6;       // the TMA stores of dP have read the A tile
  uint64_t* mma_all = bars + 7;       // GEMM phase complete in ALL CTAs (and their dP stores have read their A tiles)
  uint64_t* recv_full = bars + 8;     // partial sums of the (NC-1) peers have landed: (NC-1) x 4 warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  if ((ptx::smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("inpaintnet_b200: lstm_persist_bwd: shared memory base not 1024-byte aligned\n");
    __trap();
  }
  if (warp == 1) {
    if (lane == 0) {
      ptx::mbar_init(w_full, 1);
      for (int s = 0; s < 2; ++s) { ptx::mbar_init(&dy_full[s], 1); ptx::mbar_init(&dy_empty[s], 16); }
      ptx::mbar_init(dg_ready, 16);
      ptx::mbar_init(dp_read, 1);
      ptx::mbar_init(mma_all, NC);
      ptx::mbar_init(recv_full, 1);   // one local arrive.expect_tx per GEMM phase + the bytes of the peers' st.async
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc<256>(tmem_slot);
    ptx::tmem_relinquish();
  } else if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.tmW);
    ptx::prefetch_tmap(&p.tmDP);
    ptx::prefetch_tmap(&p.tmDY);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    ptx::setmaxnreg_dec<56>();
    if (warp == 0) {
      // ===================== producer: resident W rows, dY tiles, L2 prefetch of the saved state =====================
      if (lane == 0) {
        ptx::mbar_arrive_expect_tx(w_full, 4 * WG);
        for (int g = 0; g < 4; ++g) ptx::tma_load_3d(sW + g * WG, &p.tmW, w_full, 0, g * H + (int)rank * 64, 0);
        auto prefetch = [&](int it) {
          if (it >= T) return;
          const int s = T - 1 - it;
          const long long rt = ((long long)s * B + rbase) >> 7;
#pragma unroll
          for (int a = 0; a < 4; ++a) ptx::bulk_prefetch_l2(p.gates + ((rt * LP_ARR + a) * VPR + rank * 8) * 128, 8 * 128 * 16);
          if (s > 0) {
            const long long rp = ((long long)(s - 1) * B + rbase) >> 7;
            ptx::bulk_prefetch_l2(p.gates + ((rp * LP_ARR + 4) * VPR + rank * 8) * 128, 8 * 128 * 16);
          }
        };
        for (int it = 0; it < 3; ++it) prefetch(it);
        for (int it = 0; it < T; ++it) {
          const int s = T - 1 - it;
          const int st = it & 1;
          ptx::mbar_wait(&dy_empty[st], ((it >> 1) & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(&dy_full[st], LP_KB);
          ptx::tma_load_2d(sDY + st * LP_KB, &p.tmDY, &dy_full[st], p.y_col0 + (int)rank * 64,
                           (p.y_reverse_time ? T - 1 - s : s) * B + rbase);
          prefetch(it + 3);
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer: partial dh_{s-1}[128, H] = dG_s[128, 4 x 64] . W_hh[(g, own units), :] ==========
      const uint32_t idesc = ptx::make_idesc_bf16(128, H, 0, 1);
      const uint64_t descA0 = ptx::make_smem_desc(ptx::smem_u32(sA), 16, 1024);
      const uint64_t descW0 = ptx::make_smem_desc(ptx::smem_u32(sW), 8192, 1024);
      ptx::mbar_wait(w_full, 0);
      for (int it = 0; it < nM; ++it) {
        ptx::mbar_wait(dg_ready, it & 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              ptx::umma_bf16(tmem_base, descA0 + (uint64_t)((g * LP_KB) >> 4) + (uint64_t)(kk * 2),
                             descW0 + (uint64_t)((g * WG) >> 4) + (uint64_t)(kk * 128), idesc, (g > 0 || kk > 0) ? 1u : 0u);
        }
        __syncwarp();
        // the peers write their partial sums into this CTA's A tile once mma_all fires: the dP stores must have read it
        ptx::mbar_wait(dp_read, it & 1);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(recv_full, (NC - 1) * LP_KB);   // armed before any peer can send (they wait for mma_all)
          ptx::umma_commit_multicast(mma_all, (uint16_t)((1u << NC) - 1u));
        }
        __syncwarp();
      }
    } else if (warp == 2) {
      // ===================== store warp: dG k-blocks -> dP (gate column blocks of this CTA's units) =====================
      if (lane == 0) {
        for (int it = 0; it < T; ++it) {
          const int s = T - 1 - it;
          ptx::mbar_wait(dg_ready, it & 1);
#pragma unroll
          for (int g = 0; g < 4; ++g) ptx::tma_store_2d(&p.tmDP, sA + g * LP_KB, g * H + (int)rank * 64, s * B + rbase);
          ptx::bulk_commit();
          ptx::bulk_wait_read0();
          ptx::mbar_arrive(dp_read);
        }
        ptx::bulk_wait0();
      }
    }
  } else {
    // ===================== epilogue warps 4..19 =====================
    ptx::setmaxnreg_inc<104>();
    const int q = warp & 3;
    const int sub = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const int u0 = (int)rank * 64 + sub * 16;
    const uint32_t sA_u = ptx::smem_u32(sA), sDY_u = ptx::smem_u32(sDY);
    const uint32_t sw = (uint32_t)(row & 7);
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    // between GEMM phases every thread sends, to each peer, the partial sums of ITS row for the 16 columns that the
    // peer's thread (row, sub) owns: peer (rank + pr) % NC receives them in slot NC - 1 - pr of its A tile
    uint32_t send_base[NC], send_bar[NC];
#pragma unroll
    for (int pr = 1; pr < NC; ++pr) {
      const uint32_t peer = (rank + (uint32_t)pr) % (uint32_t)NC;
      send_base[pr] = ptx::mapa(sA_u + (uint32_t)(NC - 1 - pr) * LP_KB, peer);
      send_bar[pr] = ptx::mapa(ptx::smem_u32(recv_full), peer);
    }
    const bool tm = p.timing != nullptr && threadIdx.x == 128;
    long long w_dy = 0, w_recv = 0, w_e = 0, w_mma = 0, w_s = 0;
    const long long te0 = clock64();
    float dc[2][8], cc[2][8];   // carried dL/dc and the cell state c_s of the step being differentiated
#pragma unroll
    for (int v = 0; v < 2; ++v)
#pragma unroll
      for (int k = 0; k < 8; ++k) dc[v][k] = 0.f;
    {
      const long long rt = ((long long)(T - 1) * B + rbase) >> 7;
      const uint4* gp = p.gates + ((rt * LP_ARR + 4) * VPR + (u0 >> 3)) * 128 + row;
#pragma unroll
      for (int v = 0; v < 2; ++v) unpack8(ldg_stream(gp + v * 128), cc[v]);
    }
    for (int it = 0; it < T; ++it) {
      const int s = T - 1 - it;
      const long long rt = ((long long)s * B + rbase) >> 7;
      const uint4* gp = p.gates + ((rt * LP_ARR) * VPR + (u0 >> 3)) * 128 + row;
      uint4 gv[2][5];
#pragma unroll
      for (int v = 0; v < 2; ++v) {
#pragma unroll
        for (int a = 0; a < 4; ++a) gv[v][a] = ldg_stream(gp + (a * VPR + v) * 128);
        if (s > 0) {
          const long long rp = ((long long)(s - 1) * B + rbase) >> 7;
          gv[v][4] = ldg_stream(p.gates + ((rp * LP_ARR + 4) * VPR + (u0 >> 3) + v) * 128 + row);
        } else {
          gv[v][4] = make_uint4(0, 0, 0, 0);   // zero initial cell state
        }
      }
      const int st = it & 1;
      const long long c0 = tm ? clock64() : 0;
      ptx::mbar_wait(&dy_full[st], (it >> 1) & 1);
      const long long c1 = tm ? clock64() : 0;
      float dh[2][8];
#pragma unroll
      for (int v = 0; v < 2; ++v)
        unpack8(ld_shared_v4(sDY_u + st * LP_KB + (uint32_t)row * 128u + ((((uint32_t)(sub * 2 + v)) ^ sw) << 4)), dh[v]);
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&dy_empty[st]);
      if (it > 0) {
        // recurrent part: own partial sum (accumulator) + the partial sums the peers sent (bf16, swizzled rows)
        const long long r0 = tm ? clock64() : 0;
        ptx::mbar_wait_cluster(recv_full, (it - 1) & 1);
        ptx::tc_fence_after();
        if (tm) w_recv += clock64() - r0;
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          float own[8];
          ptx::tmem_ld8(tlane + (uint32_t)(u0 + v * 8), own);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 8; ++k) dh[v][k] += own[k];
#pragma unroll
          for (int sl = 0; sl < NC - 1; ++sl) {
            float part[8];
            unpack8(ld_shared_v4(sA_u + sl * LP_KB + (uint32_t)row * 128u + ((((uint32_t)(sub * 2 + v)) ^ sw) << 4)), part);
#pragma unroll
            for (int k = 0; k < 8; ++k) dh[v][k] += part[k];
          }
        }
        // the received partial sums alias the A tile: every epilogue thread has read them before dG overwrites it
        ptx::named_bar_sync(1, 512);
      }
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        float gi[8], gf[8], gg[8], go[8], cp[8], di[8], df[8], dgg[8], dox[8];
        unpack8(gv[v][0], gi); unpack8(gv[v][1], gf); unpack8(gv[v][2], gg); unpack8(gv[v][3], go); unpack8(gv[v][4], cp);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float tc = tanh_fast(cc[v][k]);
          const float d_o = dh[v][k] * tc;
          const float dcc = fmaf(dh[v][k] * go[k], 1.f - tc * tc, dc[v][k]);
          dox[k] = d_o * go[k] * (1.f - go[k]);
          di[k] = dcc * gg[k] * gi[k] * (1.f - gi[k]);
          dgg[k] = dcc * gi[k] * (1.f - gg[k] * gg[k]);
          df[k] = dcc * cp[k] * gf[k] * (1.f - gf[k]);
          dc[v][k] = dcc * gf[k];
          cc[v][k] = cp[k];
        }
        const uint32_t so = (uint32_t)row * 128u + ((((uint32_t)(sub * 2 + v)) ^ sw) << 4);
        st_shared_v4(sA_u + so, pack8(di));
        st_shared_v4(sA_u + LP_KB + so, pack8(df));
        st_shared_v4(sA_u + 2 * LP_KB + so, pack8(dgg));
        st_shared_v4(sA_u + 3 * LP_KB + so, pack8(dox));
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async();   // dG tile -> visible to the MMAs and the dP TMA stores
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(dg_ready);
      const long long c2 = tm ? clock64() : 0;
      if (tm) { w_dy += c1 - c0; w_e += c2 - c1; }
      if (it < nM) {
        // ---- after the GEMM phase: send the partial sums of the peers' columns (reduce-scatter through DSMEM)
        ptx::mbar_wait(mma_all, it & 1);
        ptx::tc_fence_after();
        const long long c3 = tm ? clock64() : 0;
#pragma unroll
        for (int pr = 1; pr < NC; ++pr) {
          const uint32_t peer = (rank + (uint32_t)pr) % (uint32_t)NC;
          float part[16];
          ptx::tmem_ld16(tlane + (uint32_t)(peer * 64 + sub * 16), part);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            const uint4 u = pack8(part + 8 * v);
            ptx::st_async_v4(send_base[pr] + (uint32_t)row * 128u + ((((uint32_t)(sub * 2 + v)) ^ sw) << 4), u.x, u.y, u.z, u.w, send_bar[pr]);
          }
        }
        ptx::tc_fence_before();
        if (tm) { w_mma += c3 - c2; w_s += clock64() - c3; }
      }
    }
    if (tm) {
      unsigned long long* o = p.timing + (long long)blockIdx.x * 16;
      o[4] = (unsigned long long)(clock64() - te0); o[5] = w_dy; o[6] = w_recv; o[7] = w_e; o[8] = w_mma; o[9] = w_s;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<256>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static bool lp_al16(const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; }

bool lstm_persist_shape_ok(int core, int act_dt, int B, int H) {
  static const int on = getenv("IPN_LSTM_PERSIST") ? atoi(getenv("IPN_LSTM_PERSIST")) : 1;
  return on && core == IPN_CORE_UMMA && act_dt == IPN_BF16 && (H == 128 || H == 256) && B > 0 && B % LP_ROWS == 0;
}

template <class K, class P>
static int lp_launch(K kern, const P& p, int ntiles, int NC, int smem, bool* configured, cudaStream_t stream) {
  if (!*configured) {
    IPN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    *configured = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(ntiles * NC), 1, 1);
  cfg.blockDim = dim3(LP_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NC;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  IPN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  IPN_LAUNCH_CHECK();
  return IPN_OK;
}

int lstm_persist_fwd(const IpnLstmLayer* L, int s_begin, int s_end, cudaStream_t stream) {
  const int T = L->T, H = L->H, B = L->B;
  IPN_REQUIRE(lstm_persist_shape_ok(L->core, L->act_dt, B, H), IPN_ERR_ARG, "lstm_persist_fwd: shape not eligible (B=%d H=%d)", B, H);
  IPN_REQUIRE(L->table == nullptr && L->P_blocked && lp_al16(L->P) && lp_al16(L->hseq) && lp_al16(L->cseq) &&
              (L->gates == nullptr || lp_al16(L->gates)), IPN_ERR_ARG, "lstm_persist_fwd: needs a blocked P, 16-byte aligned buffers and no token table");
  IPN_REQUIRE(L->y == nullptr || (lp_al16(L->y) && L->ld_y % 8 == 0 && L->y_col0 % 8 == 0), IPN_ERR_ALIGN, "lstm_persist_fwd: y alignment");
  LstmPersistFwd p;
  memset(&p, 0, sizeof(p));
  IPN_PROPAGATE(get_tensor_map_3d(&p.tmW, L->w_hh, (unsigned long long)H, (unsigned long long)H, 4ULL, H, (long long)H * H, 64, 4));
  IPN_PROPAGATE(get_tensor_map(&p.tmH, L->hseq, (unsigned long long)H, (unsigned long long)(T + 1) * B, H, LP_ROWS));
  p.has_y = L->y != nullptr;
  if (p.has_y) IPN_PROPAGATE(get_tensor_map(&p.tmY, L->y, (unsigned long long)L->ld_y, (unsigned long long)T * B, L->ld_y, LP_ROWS));
  p.Pblk = reinterpret_cast<const uint4*>(L->P);
  p.gates = reinterpret_cast<uint4*>(L->gates);
  p.cseq = L->cseq;
  p.T = T; p.B = B; p.s_begin = s_begin; p.s_end = s_end;
  p.y_col0 = L->y_col0; p.y_reverse_time = L->y_reverse_time;
  p.timing = g_dbg_timing;
  const bool save = L->gates != nullptr;
  const double rows = (double)(s_end - s_begin) * B;
  ProfScope prof("lstm_layer_fwd_persist", 2.0 * rows * 4.0 * H * H, rows * H * 2.0 * (4 + 1 + (save ? LP_ARR : 0) + (p.has_y ? 1 : 0)), stream);
  static bool cfgd[4] = {false, false, false, false};
  const int ntiles = B / LP_ROWS;
  if (H == 256) {
    if (save) return lp_launch(lstm_persist_fwd_kernel<4, true>, p, ntiles, 4, lpf_smem_bytes<4>(), &cfgd[0], stream);
    return lp_launch(lstm_persist_fwd_kernel<4, false>, p, ntiles, 4, lpf_smem_bytes<4>(), &cfgd[1], stream);
  }
  if (save) return lp_launch(lstm_persist_fwd_kernel<2, true>, p, ntiles, 2, lpf_smem_bytes<2>(), &cfgd[2], stream);
  return lp_launch(lstm_persist_fwd_kernel<2, false>, p, ntiles, 2, lpf_smem_bytes<2>(), &cfgd[3], stream);
}

int lstm_persist_bwd(const IpnLstmLayerBwd* L, cudaStream_t stream) {
  const int T = L->T, H = L->H, B = L->B;
  IPN_REQUIRE(lstm_persist_shape_ok(L->core, L->act_dt, B, H), IPN_ERR_ARG, "lstm_persist_bwd: shape not eligible (B=%d H=%d)", B, H);
  IPN_REQUIRE(L->dY != nullptr && lp_al16(L->dY) && L->ld_dy % 8 == 0 && L->y_col0 % 8 == 0 && lp_al16(L->dP) && lp_al16(L->gates),
              IPN_ERR_ALIGN, "lstm_persist_bwd: needs dY, 16-byte aligned buffers and ld_dy %% 8 == 0");
  LstmPersistBwd p;
  memset(&p, 0, sizeof(p));
  IPN_PROPAGATE(get_tensor_map_3d(&p.tmW, L->w_hh, 64ULL, 4ULL * H, (unsigned long long)(H / 64), H, 64, 64, (unsigned)(H / 64)));
  IPN_PROPAGATE(get_tensor_map(&p.tmDP, L->dP, 4ULL * H, (unsigned long long)T * B, 4LL * H, LP_ROWS));
  IPN_PROPAGATE(get_tensor_map(&p.tmDY, L->dY, (unsigned long long)L->ld_dy, (unsigned long long)T * B, L->ld_dy, LP_ROWS));
  p.gates = reinterpret_cast<const uint4*>(L->gates);
  p.T = T; p.B = B;
  p.y_col0 = L->y_col0; p.y_reverse_time = L->y_reverse_time;
  p.timing = g_dbg_timing;
  const double rows = (double)T * B;
  ProfScope prof("lstm_layer_bwd_persist", 2.0 * rows * 4.0 * H * H, rows * H * 2.0 * (LP_ARR + 1 + 4), stream);
  static bool cfgd[2] = {false, false};
  const int ntiles = B / LP_ROWS;
  if (H == 256) return lp_launch(lstm_persist_bwd_kernel<4>, p, ntiles, 4, lpb_smem_bytes<4>(), &cfgd[0], stream);
  return lp_launch(lstm_persist_bwd_kernel<2>, p, ntiles, 2, lpb_smem_bytes<2>(), &cfgd[1], stream);
}

}  // namespace ipn

extern "C" int ipn_lstm_persist_eligible(int core, int act_dt, int B, int H) {
  return ipn::lstm_persist_shape_ok(core, act_dt, B, H) ? 1 : 0;
}
extern "C" int ipn_lstm_gates_cols(int H, int persistent) { return (persistent ? ipn::LP_ARR : 4) * H; }
