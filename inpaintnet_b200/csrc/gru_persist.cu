// Persistent GRU layer, forward (see gru_persist.cuh for the design).
#include "gru_persist.cuh"
#include <stdlib.h>
#include <type_traits>

namespace ipn {

// ---------------------------------------------------------------------------------------------
// input-projection prep: blocked bf16 pre-activation input of every gate, everything that does not depend
// on h folded in:   r, z: 0.5 * (P + table[tok] + pvec + b_hh)   (sigmoid(x) = 0.5 tanh(0.5 x) + 0.5)
//                   n   :        P + table[tok] + pvec            (b_hn stays inside r * (.))
// one CTA = 128 rows x 64 units of one gate; coalesced reads, swizzled smem transpose, coalesced writes
// ---------------------------------------------------------------------------------------------
struct PrepP {
  const __nv_bfloat16* P;
  long long ldP;
  int P_bcast;
  const float* table;
  long long ld_table;
  const int* tok;
  const float* pvec;
  const float* b_hh;
  uint4* out;
  int H, Bt;
  int ntw, tt_min, row0;   // window: blockIdx.x = (tt - tt_min) * ntw + tile ; rows tt*Bt + row0 + tile*128 ...
};

__global__ void __launch_bounds__(256) gru_prep_p_kernel(PrepP p) {
  __shared__ uint4 tile[128 * 8];
  const int g = blockIdx.y, c = blockIdx.z;
  const int lt = blockIdx.x;                               // local tile index (also the output tile index)
  const long long Rbase = (long long)(p.tt_min + lt / p.ntw) * p.Bt + p.row0 + (lt % p.ntw) * 128;
  const int H = p.H;
  const int i = threadIdx.x;
  const int vec = i & 7;
  const int col = g * H + c * 64 + vec * 8;
  float cst[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    cst[k] = p.pvec != nullptr ? p.pvec[col + k] : 0.f;
    if (g < 2) cst[k] += p.b_hh[col + k];
  }
  const float scale = g < 2 ? 0.5f : 1.f;
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int row = pass * 32 + (i >> 3);
    const long long R = Rbase + row;
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = 0.f;
    if (p.P != nullptr) {
      const long long sr = p.P_bcast ? (R % p.Bt) : R;
      const uint4 u = *reinterpret_cast<const uint4*>(p.P + sr * p.ldP + col);
      unpack8(u, f);
    }
    if (p.table != nullptr) {
      const float4* tp = reinterpret_cast<const float4*>(p.table + (long long)p.tok[R] * p.ld_table + col);
      const float4 a = tp[0], b = tp[1];
      f[0] += a.x; f[1] += a.y; f[2] += a.z; f[3] += a.w;
      f[4] += b.x; f[5] += b.y; f[6] += b.z; f[7] += b.w;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = (f[k] + cst[k]) * scale;
    tile[row * 8 + (vec ^ (row & 7))] = pack8(f);
  }
  __syncthreads();
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int v = pass * 2 + (i >> 7), row = i & 127;
    const uint4 u = tile[row * 8 + (v ^ (row & 7))];
    p.out[(((long long)lt * 3 + g) * (H / 8) + c * 8 + v) * 128 + row] = u;
  }
}

// Token-table input projection (encoder layer 0: P is a row of a <= 128-row table): nothing is materialised.  The
// table is folded once per call (same arithmetic and rounding as gru_prep_p_kernel: bf16((table + b_hh) * 0.5) for
// r,z and bf16(table) for n) and the layer kernel's epilogue gathers its rows by token id (L2-resident, 3H * 2 B per
// row) -- for 98304 context measures that removes 14.5 GB of HBM writes and the same amount of reads per layer.
__global__ void gru_fold_table_kernel(const float* table, long long ld_table, int rows, const float* b_hh, int H,
                                      __nv_bfloat16* out) {   // b_hh == nullptr: it is folded into the other P term
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * 3 * H) return;
  const int r = i / (3 * H), c = i - r * 3 * H;
  float f = 0.f;
  f += table[(long long)r * ld_table + c];
  const float cst = (c < 2 * H && b_hh != nullptr) ? b_hh[c] : 0.f;
  out[i] = __float2bfloat16_rn((f + cst) * (c < 2 * H ? 0.5f : 1.f));
}

int gru_fold_table(const float* table, long long ld_table, int rows, const float* b_hh, int H, void* out, cudaStream_t stream) {
  ProfScope prof("gru_fold_table", 0.0, (double)rows * 3 * H * 6, stream);
  gru_fold_table_kernel<<<(rows * 3 * H + 255) / 256, 256, 0, stream>>>(table, ld_table, rows, b_hh, H,
                                                                         reinterpret_cast<__nv_bfloat16*>(out));
  IPN_LAUNCH_CHECK();
  return IPN_OK;
}

// y[R, col0 + u] = keep[R, col0 + u] ? y * scale : 0   (inter-layer dropout applied after the layer kernel)
__global__ void gru_mask_y_kernel(__nv_bfloat16* y, long long ld_y, const unsigned char* mask, long long ld_mask,
                                  int col0, int H, long long rows, float scale, int nrows, int Bt, int tt_min, int row0) {
  // rows = NS * nrows window rows: local row l -> time tt_min + l / nrows, batch row row0 + l % nrows
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int vpr = H / 8;
  if (idx >= rows * vpr) return;
  const long long l = idx / vpr;
  const long long R = (long long)(tt_min + l / nrows) * Bt + row0 + l % nrows;
  const int u = (int)(idx - l * vpr) * 8;
  uint4* yp = reinterpret_cast<uint4*>(y + R * ld_y + col0 + u);
  const uint2 m = *reinterpret_cast<const uint2*>(mask + R * ld_mask + col0 + u);
  float f[8];
  unpack8(*yp, f);
  const unsigned char* mb = reinterpret_cast<const unsigned char*>(&m);
#pragma unroll
  for (int k = 0; k < 8; ++k) f[k] = mb[k] ? f[k] * scale : 0.f;
  *yp = pack8(f);
}

// ---------------------------------------------------------------------------------------------
// the persistent forward kernel
// ---------------------------------------------------------------------------------------------
constexpr int GPF_W_BYTES = 3 * GP_CH * 128;  // one (chunk, k-block) of W_hh: 3 gates x 64 rows x 64 bf16 = 24 KB
constexpr int GPF_W_RING = 3 * GPF_W_BYTES;   // shared memory of the W ring (3 stages; 6 half-size stages per CTA of a pair)
constexpr int GPF_NBAR = 56;

static inline int gpf_smem_bytes(int H) {
  return (H / 64) * GP_KB_BYTES + GPF_W_RING + GP_KB_BYTES + H * 4 + GPF_NBAR * 8 + 16;
}

// PAIR: two CTAs (adjacent row tiles of one direction) form a cta_group::2 pair: one tcgen05.mma covers both
// tiles (M = 256), each CTA stages only HALF of every W_hh tile (half the L2 traffic and shared-memory fill,
// twice the pipeline depth) and the operand fetch per MMA drops from 10 KB to 7 KB per SM.
// CS = 2 (on by default since round 2, IPN_GPF_CS=0 disables; validated on B200: tests/dev/persist_fwd.py + the GPU suite): COLUMN SPLIT.
// The two CTAs of a cluster own the SAME 128-row tile and half of the hidden units each (chunks
// [rank*KB/2, (rank+1)*KB/2): half of the W_hh stream, of the MMAs and of the epilogue per step), so the serial
// per-step latency halves and a 64-CTA encoder layer occupies 128 SMs.  Every CTA still needs the full h_{t-1}
// as its A operand: each stores its half of h_t to the hseq slot (TMA store, as before) and signals the PEER's
// h_stored[chunk] barrier once the store has completed; the A loaders reload all k-blocks from L2 as before.
template <bool SAVE, bool PAIR, int CS>
__global__ void __launch_bounds__(GP_THREADS, 1) gru_persist_fwd_kernel(const __grid_constant__ GruPersistFwd p) {
  static_assert(CS == 1 || CS == 2 || (CS == 4 && !PAIR), "column split: 2 or 4 CTAs per tile, or 2 CTA pairs per two tiles");
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int WST_BYTES = PAIR ? GPF_W_BYTES / 2 : GPF_W_BYTES;  // per-CTA bytes of one W stage
  constexpr int WSTAGES = GPF_W_RING / WST_BYTES;
  constexpr int UPC = PAIR ? 32 : 64;                              // hidden units of a chunk staged by this CTA
  const GruPersistFwdDir& D = p.d[blockIdx.y];
  const int H = p.H, KB = H >> 6, T = p.T;
  const int NS = p.s_end - p.s_begin;   // processing steps s_begin .. s_end-1 (loop index t = s - s_begin)
  const int Bt = p.Bt;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = (PAIR || CS > 1) ? ptx::cluster_ctarank() : 0u;
  const uint32_t rank = PAIR ? (crank & 1u) : 0u;   // pair: which half of the M = 256 tile / of every W stage
  const bool leader = rank == 0;                    // column split: every CTA (pair) runs its own complete pipeline
  const uint32_t lead_rank = crank & ~1u;           // cluster rank of this pair's leader CTA
  // PAIR + CS = 2: a cluster of 4 = two pairs; pair `crank >> 1` owns half of the hidden units of TWO adjacent row tiles
  const uint32_t colrank = PAIR ? (crank >> 1) : crank;
  const int tile_x = PAIR ? ((int)blockIdx.x / (2 * CS)) * 2 + (int)rank : (int)blockIdx.x / CS;   // row tile of this CTA
  const int rbase = p.row0 + tile_x * GP_ROWS;
  const uint16_t pair_mask = (uint16_t)(3u << lead_rank);
  const int c_lo = CS > 1 ? (int)colrank * (KB / CS) : 0, c_hi = c_lo + KB / CS;   // chunks (64 hidden units) of this CTA

  const bool trace_on = p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
  auto tr = [&](int role, int t, int ev) {
    if (trace_on && t >= 8 && t < 10 && ev < 32) p.trace[(role * 2 + (t - 8)) * 32 + ev] = (unsigned long long)clock64();
  };
  uint8_t* sA = smem;
  uint8_t* sW = sA + KB * GP_KB_BYTES;
  uint8_t* sStg = sW + GPF_W_RING;
  float* sBias = reinterpret_cast<float*>(sStg + GP_KB_BYTES);  // b_hn[H]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + H);
  uint64_t* w_full = bars;            // [6]  (leader CTA's copy is the one in use)
  uint64_t* w_empty = bars + 6;       // [6]
  uint64_t* a_full = bars + 12;       // [8]  TMA landed k-block kb of this step's h_{t-1} (leader's copy: both tiles)
  uint64_t* a_free = bars + 20;       // [8]  every MMA of this step has read k-block kb
  uint64_t* h_stored = bars + 28;     // [8]  chunk kb of h_t is in global memory
  uint64_t* tmem_full = bars + 36;    // [2]
  uint64_t* tmem_empty = bars + 38;   // [2]  (leader's copy collects the epilogue warps of both CTAs)
  uint64_t* stg_ready = bars + 40;    // staging tile written by the 8 epilogue warps
  uint64_t* stg_free = bars + 41;     // staging tile read out by the TMA store
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 42);
  // column split: the h_stored barriers of ODD steps.  A peer signals them remotely, so nothing in THIS CTA orders
  // its arrival for step t+1 after our wait for step t; with one barrier set per step parity the peer would have to
  // be two steps ahead to alias a phase, which the data dependence (it needs our step t+1 stores) rules out.
  uint64_t* h_stored_odd = bars + 44;  // [8]

  if ((ptx::smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("inpaintnet_b200: gru_persist_fwd: shared memory base not 1024-byte aligned\n");
    __trap();
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < 6; ++s) { ptx::mbar_init(&w_full[s], 1); ptx::mbar_init(&w_empty[s], 1); }
      for (int k = 0; k < 8; ++k) { ptx::mbar_init(&a_full[k], 1); ptx::mbar_init(&a_free[k], 1); ptx::mbar_init(&h_stored[k], 1); }
      if (CS > 1) for (int k = 0; k < 8; ++k) ptx::mbar_init(&h_stored_odd[k], 1);
      for (int b = 0; b < 2; ++b) { ptx::mbar_init(&tmem_full[b], 1); ptx::mbar_init(&tmem_empty[b], PAIR ? 32 : 16); }
      ptx::mbar_init(stg_ready, 16);
      ptx::mbar_init(stg_free, 1);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    if (PAIR) { ptx::tmem_alloc_pair<512>(tmem_slot); ptx::tmem_relinquish_pair(); }
    else { ptx::tmem_alloc<512>(tmem_slot); ptx::tmem_relinquish(); }
  } else if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&D.tmW);
    ptx::prefetch_tmap(&D.tmH);
    if (D.has_y) ptx::prefetch_tmap(&D.tmY);
  }
  for (int u = threadIdx.x; u < H; u += GP_THREADS) sBias[u] = D.b_hh[2 * H + u];
  ptx::tc_fence_before();
  __syncthreads();
  if (PAIR || CS > 1) ptx::cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / TMA signal
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp < 4) {
  ptx::setmaxnreg_dec<56>();
  if (warp == 0) {
    // ===================== W_hh producer: the whole matrix streams through the ring every step ==========
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // the epilogue's input-projection tiles are pulled from HBM into L2 well ahead of their use
      auto prefetch_p = [&](int t, int c) {
        if (t >= NS || (p.dbg & 16) || D.Pblk == nullptr) return;
        const int tt = D.reverse ? T - 1 - (p.s_begin + t) : (p.s_begin + t);
        const long long rt = (long long)tt * D.p_t_stride + D.p_t0 + tile_x;
        const int vpr = H >> 3;
#pragma unroll
        for (int g = 0; g < 3; ++g)
          ptx::bulk_prefetch_l2(D.Pblk + ((rt * 3 + g) * vpr + c * 8) * 128, 8 * 128 * 16);
      };
      prefetch_p(0, c_lo);
      for (int t = 0; t < NS; ++t)
        for (int c = c_lo; c < c_hi; ++c)
          for (int kb = 0; kb < KB; ++kb) {
            if (kb == 0) prefetch_p(c + 1 == c_hi ? t + 1 : t, c + 1 == c_hi ? c_lo : c + 1);
            ptx::mbar_wait(&w_empty[stage], phase ^ 1);
            const bool skip = (p.dbg & 8) && (t > 0 || c > c_lo);
            if (leader) {
              if (skip) ptx::mbar_arrive(&w_full[stage]);
              else ptx::mbar_arrive_expect_tx(&w_full[stage], GPF_W_BYTES);
            }
            if (!skip) {
              // one 3-D box: 64 k x UPC units x 3 gates -> rows (gate, unit) of the K-major B tile
              uint8_t* dst = sW + stage * WST_BYTES;
              if (PAIR) ptx::tma_load_3d_pair(dst, &D.tmW, &w_full[stage], kb * 64, c * 64 + (int)rank * UPC, 0);
              else ptx::tma_load_3d(dst, &D.tmW, &w_full[stage], kb * 64, c * 64, 0);
            }
            if (++stage == WSTAGES) { stage = 0; phase ^= 1; }
          }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only for a pair) =====================
    // The whole warp runs the loop so that the descriptor arithmetic stays warp-uniform; one elected lane issues.
    if (leader) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(PAIR ? 256 : 128, 3 * GP_CH, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int i = 0;
      const bool tm = p.timing != nullptr;
      long long w_te = 0, w_af = 0, w_wf = 0;
      const long long t_begin = clock64();
      const uint64_t descA0 = ptx::make_smem_desc(ptx::smem_u32(sA), 16, 1024);
      const uint64_t descW0 = ptx::make_smem_desc(ptx::smem_u32(sW), 16, 1024);
      for (int t = 0; t < NS; ++t)
        for (int c = c_lo; c < c_hi; ++c, ++i) {
          const int b = i & 1, n = i >> 1;
          wait_acc(&tmem_empty[b], (n & 1) ^ 1, tm, w_te);
          ptx::tc_fence_after();
          if (lane == 0) tr(0, t, (c - c_lo) * 3);
          const uint32_t dcol = tmem_base + (uint32_t)(b * 256);
          for (int kb = 0; kb < KB; ++kb) {
            if (c == c_lo) wait_acc(&a_full[kb], t & 1, tm, w_af);
            wait_acc(&w_full[stage], phase, tm, w_wf);
            ptx::tc_fence_after();
            if (lane == 0 && kb == KB - 1) tr(0, t, (c - c_lo) * 3 + 1);
            const uint64_t da0 = descA0 + (uint64_t)((kb * GP_KB_BYTES) >> 4);
            const uint64_t dw0 = descW0 + (uint64_t)((stage * WST_BYTES) >> 4);
            if (ptx::elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                if (PAIR) ptx::umma_bf16_pair(dcol, da0 + (uint64_t)(kk * 2), dw0 + (uint64_t)(kk * 2), idesc, (kb > 0 || kk > 0) ? 1u : 0u);
                else ptx::umma_bf16(dcol, da0 + (uint64_t)(kk * 2), dw0 + (uint64_t)(kk * 2), idesc, (kb > 0 || kk > 0) ? 1u : 0u);
              }
              if (PAIR) {
                ptx::umma_commit_pair(&w_empty[stage], pair_mask);
                if (c == c_hi - 1) ptx::umma_commit_pair(&a_free[kb], pair_mask);
              } else {
                ptx::umma_commit(&w_empty[stage]);
                if (c == c_hi - 1) ptx::umma_commit(&a_free[kb]);
              }
            }
            __syncwarp();
            if (++stage == WSTAGES) { stage = 0; phase ^= 1; }
          }
          if (ptx::elect_one()) {
            if (PAIR) ptx::umma_commit_pair(&tmem_full[b], pair_mask);
            else ptx::umma_commit(&tmem_full[b]);
          }
          __syncwarp();
          if (lane == 0) tr(0, t, (c - c_lo) * 3 + 2);
        }
      if (tm && lane == 0) {
        unsigned long long* o = p.timing + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * 16;
        o[0] = (unsigned long long)(clock64() - t_begin); o[1] = w_te; o[2] = w_af; o[3] = w_wf;
      }
    }
  } else if (warp == 2) {
    // ===================== store warp: staging tile -> hseq slot (+ y) =====================
    if (lane == 0) {
      int i = 0;
      const bool tm = p.timing != nullptr;
      long long w_sr = 0, w_st = 0;
      for (int t = 0; t < NS; ++t) {
        const int tt = D.reverse ? T - 1 - (p.s_begin + t) : (p.s_begin + t);
        const int out_slot = D.reverse ? tt : tt + 1;
        for (int c = c_lo; c < c_hi; ++c, ++i) {
          wait_acc(stg_ready, i & 1, tm, w_sr);
          tr(2, t, (c - c_lo) * 3);
          const long long ts0 = tm ? clock64() : 0;
          ptx::tma_store_2d(&D.tmH, sStg, c * 64, out_slot * Bt + rbase);
          if (D.has_y) ptx::tma_store_2d(&D.tmY, sStg, D.y_col0 + c * 64, tt * Bt + rbase);
          ptx::bulk_commit();
          ptx::bulk_wait_read0();
          ptx::mbar_arrive(stg_free);
          ptx::bulk_wait0();
          tr(2, t, (c - c_lo) * 3 + 1);
          if (CS > 1) {   // the peer reloads this chunk of h_t from L2 as a k-block of its next step's A operand
            uint64_t* hs = (t & 1) ? &h_stored_odd[c] : &h_stored[c];
            ptx::mbar_arrive(hs);
            ptx::fence_acq_rel_cluster();   // one release fence, then relaxed remote arrives (see ptx.cuh)
            if (PAIR) {
              ptx::mbar_arrive_remote_relaxed(hs, crank ^ 2u);   // same rows, the other half of the hidden units
            } else {
#pragma unroll
              for (uint32_t pr = 1; pr < (uint32_t)CS; ++pr) ptx::mbar_arrive_remote_relaxed(hs, (crank + pr) % (uint32_t)CS);
            }
          } else {
            ptx::mbar_arrive(&h_stored[c]);
          }
          tr(2, t, (c - c_lo) * 3 + 2);
          if (tm) w_st += clock64() - ts0;
        }
      }
      if (tm) {
        unsigned long long* o = p.timing + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * 16;
        o[4] = w_sr; o[5] = w_st;
      }
    }
  } else if (warp == 3) {
    // ===================== A loader: h_{t-1} tile, one 64-unit k-block at a time =====================
    if (lane == 0) {
      const bool tm = p.timing != nullptr;
      long long w_fr = 0, w_hs = 0;
      for (int t = 0; t < NS; ++t) {
        const int tt = D.reverse ? T - 1 - (p.s_begin + t) : (p.s_begin + t);
        const int in_slot = D.reverse ? tt + 1 : tt;
        for (int kb = 0; kb < KB; ++kb) {
          if (t > 0) {
            wait_acc(&a_free[kb], (t - 1) & 1, tm, w_fr);
            if (CS > 1) {   // step t-1's stores: barrier set (t-1) & 1, phase ((t-1) >> 1) & 1; peer chunks at cluster scope
              uint64_t* hs = ((t - 1) & 1) ? &h_stored_odd[kb] : &h_stored[kb];
              if ((kb < c_lo || kb >= c_hi) && !(p.dbg & 64)) ptx::mbar_wait_cluster(hs, ((t - 1) >> 1) & 1);
              else wait_acc(hs, ((t - 1) >> 1) & 1, tm, w_hs);
            } else {
              wait_acc(&h_stored[kb], (t - 1) & 1, tm, w_hs);
            }
            if (!(p.dbg & 32)) ptx::fence_proxy_async_all();
          }
          tr(3, t, kb);
          if (leader) ptx::mbar_arrive_expect_tx(&a_full[kb], PAIR ? 2 * GP_KB_BYTES : GP_KB_BYTES);
          if (PAIR) ptx::tma_load_2d_pair(sA + kb * GP_KB_BYTES, &D.tmH, &a_full[kb], kb * 64, in_slot * Bt + rbase);
          else ptx::tma_load_2d(sA + kb * GP_KB_BYTES, &D.tmH, &a_full[kb], kb * 64, in_slot * Bt + rbase);
        }
      }
      if (tm) {
        unsigned long long* o = p.timing + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * 16;
        o[6] = w_fr; o[7] = w_hs;
      }
    }
  }
  } else {
    // ===================== epilogue warps 4..19: warp = (TMEM lane quadrant, 16-unit sub-chunk) =====================
    ptx::setmaxnreg_inc<104>();
    const int q = warp & 3;
    const int sub = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const int vpr = H >> 3;            // 16-byte vectors per row of one array
    const uint32_t sA_u = ptx::smem_u32(sA), sStg_u = ptx::smem_u32(sStg);
    const uint32_t sw = (uint32_t)(row & 7);
    const long long astride = (long long)vpr * 128;
    const bool tm = p.timing != nullptr && threadIdx.x == 128;
    long long w_tf = 0, w_sf = 0;
    const long long te0 = clock64();
    const bool ldp = !(p.dbg & 1);
    // input projection, register double-buffered one chunk ahead (the tiles were prefetched into L2 earlier)
    int tok_t = -1, tokv = 0;   // token-table mode: this row's token of step tok_t
    auto load_p = [&](uint4 (&dst)[3][2], int t, int c) {
      const int tt = D.reverse ? T - 1 - (p.s_begin + t) : (p.s_begin + t);
      if (D.ftab != nullptr) {   // gather the folded table row of this thread's token (L2 / L1 hits)
        if (t != tok_t) {
          tokv = __ldg(D.tok + (long long)tt * Bt + rbase + row);
          tok_t = t;
        }
        const uint4* base = D.ftab + (long long)tokv * (3 * vpr) + c * 8 + sub * 2;
        // one 256-bit load per gate (the two 16-byte vectors are adjacent): a gather costs one LSU wavefront per lane
        // and instruction, so this halves its time (the folded table is 32-byte aligned: checked on the host)
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          if (ldp) ldg_nc_32B(base + g * vpr, dst[g][0], dst[g][1]);
          else dst[g][0] = dst[g][1] = make_uint4(0, 0, 0, 0);
        }
        if (D.Pblk != nullptr) {
          // two-term projection (tick GRU layer 0): table row of the previous token + the per-beat projection, the
          // latter a blocked tile that is the same for every step of the call (p_t_stride == 0)
          const long long rtb = (long long)tt * D.p_t_stride + D.p_t0 + tile_x;
          const uint4* pb = D.Pblk + (rtb * 3 * vpr + c * 8 + sub * 2) * 128 + row;
#pragma unroll
          for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int v = 0; v < 2; ++v) {
              float a[8], b[8];
              unpack8(dst[g][v], a);
              unpack8(ldg_stream(pb + (g * vpr + v) * 128), b);
#pragma unroll
              for (int k = 0; k < 8; ++k) a[k] += b[k];
              dst[g][v] = pack8(a);
            }
        }
        return;
      }
      const long long rt = (long long)tt * D.p_t_stride + D.p_t0 + tile_x;
      const uint4* base = D.Pblk + (rt * 3 * vpr + c * 8 + sub * 2) * 128 + row;
#pragma unroll
      for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int v = 0; v < 2; ++v) dst[g][v] = ldp ? ldg_stream(base + (g * vpr + v) * 128) : make_uint4(0, 0, 0, 0);
    };
    // one chunk of the flattened (t, c) sequence
    auto do_chunk = [&](int t, int c, int i) {
      const int tt = D.reverse ? T - 1 - (p.s_begin + t) : (p.s_begin + t);
      const long long rt = ((long long)tt * Bt + rbase) >> 7;  // 128-row tile index in time-ordered buffers
      const int b = i & 1, n = i >> 1;
      uint4 pv[3][2];
      load_p(pv, t, c);  // L2 hits (bulk-prefetched by the producer warp); latency hidden by the other 3 warps of the SMSP
      if (leader) ptx::mbar_wait(&a_full[c], t & 1);  // h_{t-1} k-block c visible to this thread (read below)
      if (threadIdx.x == 128) tr(1, t, (c - c_lo) * 4);
      wait_acc(&tmem_full[b], n & 1, tm, w_tf);
      ptx::tc_fence_after();
      if (threadIdx.x == 128) tr(1, t, (c - c_lo) * 4 + 1);
      if (!(p.dbg & 4)) {
        const uint32_t abase = sA_u + c * GP_KB_BYTES + row * 128;
        // accumulator column of (gate g, unit u of the chunk): pair: (u / 32) * 96 + g * 32 + u % 32 ; single: g * 64 + u
        constexpr int GSTR = PAIR ? 32 : 64;
        const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) +
                              (uint32_t)(b * 256 + (PAIR ? (sub >> 1) * 96 + (sub & 1) * 16 : sub * 16));
        uint4* gp = D.gates + ((rt * GP_GATE_ARRAYS) * vpr + c * 8 + sub * 2) * 128 + row;
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          float acc[3][8];
#pragma unroll
          for (int g = 0; g < 3; ++g) ptx::tmem_ld8(tacc + (uint32_t)(g * GSTR + v * 8), acc[g]);
          const uint4 hpv = ld_shared_v4(abase + (((uint32_t)(sub * 2 + v) ^ sw) << 4));
          float pr[8], pz[8], pn[8], hp[8], rr[8], zz[8], nn[8], hn[8], hh[8];
          unpack8(pv[0][v], pr);
          unpack8(pv[1][v], pz);
          unpack8(pv[2][v], pn);
          unpack8(hpv, hp);
          const float4 b0 = *reinterpret_cast<const float4*>(sBias + c * 64 + sub * 16 + v * 8);
          const float4 b1 = *reinterpret_cast<const float4*>(sBias + c * 64 + sub * 16 + v * 8 + 4);
          const float bn[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          ptx::tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float r = fmaf(0.5f, tanh_fast(fmaf(0.5f, acc[0][k], pr[k])), 0.5f);
            const float z = fmaf(0.5f, tanh_fast(fmaf(0.5f, acc[1][k], pz[k])), 0.5f);
            const float g = acc[2][k] + bn[k];
            const float nv = tanh_fast(fmaf(r, g, pn[k]));
            rr[k] = r; zz[k] = z; hn[k] = g; nn[k] = nv;
            hh[k] = fmaf(z, hp[k] - nv, nv);
          }
          if (v == 0) wait_acc(stg_free, (i & 1) ^ 1, tm, w_sf);  // previous chunk's tile has been read out
          st_shared_v4(sStg_u + row * 128 + (((uint32_t)(sub * 2 + v) ^ sw) << 4), pack8(hh));
          if (SAVE && !(p.dbg & 2)) {
            stg_stream(gp + v * 128, pack8(rr));
            stg_stream(gp + astride + v * 128, pack8(zz));
            stg_stream(gp + 2 * astride + v * 128, pack8(nn));
            stg_stream(gp + 3 * astride + v * 128, pack8(hn));
            stg_stream(gp + 4 * astride + v * 128, hpv);
          }
        }
      } else {
        wait_acc(stg_free, (i & 1) ^ 1, tm, w_sf);
      }
      if (threadIdx.x == 128) tr(1, t, (c - c_lo) * 4 + 2);
      ptx::tc_fence_before();
      ptx::fence_proxy_async();  // staging-tile writes -> visible to the TMA store
      __syncwarp();
      if (lane == 0) {
        if (leader) ptx::mbar_arrive(&tmem_empty[b]);
        else ptx::mbar_arrive_remote(&tmem_empty[b], lead_rank);
        ptx::mbar_arrive(stg_ready);
      }
    };
    {
      int i = 0, t = 0, c = c_lo;
      const int total = NS * (c_hi - c_lo);
      while (i < total) {
        do_chunk(t, c, i);
        ++i; if (++c == c_hi) { c = c_lo; ++t; }
      }
    }
    if (tm) {
      unsigned long long* o = p.timing + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * 16;
      o[8] = (unsigned long long)(clock64() - te0); o[9] = w_tf; o[10] = w_sf;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (PAIR || CS > 1) ptx::cluster_sync_all();  // the peer's shared memory / TMEM / barriers stay alive until both are done
  if (warp == 1) {
    ptx::tc_fence_after();
    if (PAIR) ptx::tmem_dealloc_pair<512>(tmem_base);
    else ptx::tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
bool persist_enabled() {
  static const int on = getenv("IPN_PERSIST") ? atoi(getenv("IPN_PERSIST")) : 1;
  return on != 0;
}

static bool al16(const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; }
static bool dir_gathers_table_plus_bcast(const IpnGruDir& D);

bool gru_persist_fwd_shape_ok(const IpnGruLayer* L) {
  if (!persist_enabled()) return false;
  if (L->core != IPN_CORE_UMMA || L->act_dt != IPN_BF16) return false;
  if (L->H % 64 != 0 || L->H < 64 || L->H > 512) return false;
  if (L->B_total % GP_ROWS != 0 || L->row0 % GP_ROWS != 0 || L->nrows % GP_ROWS != 0 || L->nrows <= 0) return false;
  if (L->row0 < 0 || L->row0 + L->nrows > L->B_total) return false;
  if (L->s_begin < 0 || L->s_begin >= L->s_end || L->s_end > L->T) return false;
  if (L->y != nullptr && (L->ld_y % 8 != 0 || !al16(L->y))) return false;
  if (L->mask != nullptr && (L->ld_mask % 8 != 0 || reinterpret_cast<uintptr_t>(L->mask) % 8 != 0)) return false;
  for (int d = 0; d < L->ndir; ++d) {
    const IpnGruDir& D = L->dir[d];
    if (D.y_col0 % 8 != 0) return false;
    if (D.P != nullptr && (D.ldP % 8 != 0 || !al16(D.P))) return false;
    if (D.table != nullptr && (D.ld_table % 4 != 0 || !al16(D.table))) return false;
    if (D.P_blocked && (D.P == nullptr || D.pvec != nullptr)) return false;
    // blocked P + token table: only as the two-term form (broadcast blocked tile + gathered table row)
    if (D.P_blocked && (D.table != nullptr || D.P_bcast) && !dir_gathers_table_plus_bcast(D)) return false;
    if (D.gates != nullptr && !al16(D.gates)) return false;
  }
  return true;
}

constexpr int GPF_FOLD_ROWS = 128;   // largest token table the gather path folds (vocabularies are 45-90 symbols)
// two-term input projection: P (blocked, one [B_total, 3, H] tile set reused at every step, biases folded in) + table[tok]
static bool dir_gathers_table_plus_bcast(const IpnGruDir& D) {
  return D.table != nullptr && D.tok != nullptr && D.P != nullptr && D.P_blocked && D.P_bcast && D.pvec == nullptr &&
         D.table_rows > 0 && D.table_rows <= GPF_FOLD_ROWS;
}
static bool dir_gathers_table(const IpnGruDir& D) {
  static const int on = getenv("IPN_GPF_GATHER") ? atoi(getenv("IPN_GPF_GATHER")) : 1;
  return on && D.table != nullptr && D.tok != nullptr && D.P == nullptr && D.pvec == nullptr && !D.P_blocked &&
         D.table_rows > 0 && D.table_rows <= GPF_FOLD_ROWS;
}

long long gru_persist_fwd_ws_bytes(const IpnGruLayer* L) {
  if (!gru_persist_fwd_shape_ok(L)) return 0;
  const long long per_dir = (long long)(L->s_end - L->s_begin) * L->nrows * 3 * L->H * 2;
  long long total = 0;   // per direction: the folded token table, or the blocked P of this call's window, or nothing
  for (int d = 0; d < L->ndir; ++d)
    total += (dir_gathers_table(L->dir[d]) || dir_gathers_table_plus_bcast(L->dir[d])) ? GPF_FOLD_ROWS * 3LL * L->H * 2
                                                                                          : (L->dir[d].P_blocked ? 0 : per_dir);
  return total > 16 ? total : 16;
}

int gru_persist_fwd(const IpnGruLayer* L, void* ws, long long ws_bytes, cudaStream_t stream) {
  const int T = L->T, H = L->H, Bt = L->B_total;
  IPN_REQUIRE(ws != nullptr && ws_bytes >= gru_persist_fwd_ws_bytes(L) && al16(ws), IPN_ERR_ARG,
              "gru_persist_fwd: workspace too small (%lld < %lld)", ws_bytes, gru_persist_fwd_ws_bytes(L));
  GruPersistFwd p;
  memset(&p, 0, sizeof(p));
  p.T = T; p.H = H; p.Bt = Bt;
  p.row0 = L->row0; p.s_begin = L->s_begin; p.s_end = L->s_end;
  const int NS = L->s_end - L->s_begin, ntw = L->nrows / GP_ROWS;   // steps and row tiles of this call's window
  static const int dbg = getenv("IPN_GPF_DBG") ? atoi(getenv("IPN_GPF_DBG")) : 0;
  p.dbg = dbg;
  p.timing = g_dbg_timing;   // (cleared below for the column-split grid: the buffer is sized for one CTA per tile)
  const long long per_dir = (long long)NS * L->nrows * 3 * H * 2;
  static const int pair_on = getenv("IPN_GPF_PAIR") ? atoi(getenv("IPN_GPF_PAIR")) : 1;
  // column split (see the kernel's header comment): on unless IPN_GPF_CS=0; only where all the
  // CTAs of the doubled grid are co-resident (otherwise the second wave waits and nothing is gained)
  static const int cs_on = getenv("IPN_GPF_CS") ? atoi(getenv("IPN_GPF_CS")) : 1;
  // 4 CTAs per row tile where that still fits one wave (a 4-CTA cluster grid places 132 CTAs): the 32-tile
  // uni-directional layers of a 4096-measure batch -- the beat GRU and every tick of the argmax decode
  const int cs = !cs_on ? 1 : ((H / 64) % 4 == 0 && cs_on != 2 && 4 * ntw * L->ndir <= 132) ? 4
                             : ((H / 64) % 2 == 0 && 2 * ntw * L->ndir <= 148) ? 2 : 1;
  // two CTA pairs per two row tiles (M = 256 MMAs on half of the hidden units each): every CTA stages HALF of each W_hh
  // tile, so the same ring holds twice as many tiles in flight -- the streaming rate of these kernels is (bytes in
  // flight) / (commit -> refill -> landed round trip), not L2 bandwidth.  IPN_GPF_PAIRCS=0 keeps the plain column split.
  static const int paircs_on = getenv("IPN_GPF_PAIRCS") ? atoi(getenv("IPN_GPF_PAIRCS")) : 1;
  const bool pair = pair_on && ntw % 2 == 0 && (cs == 1 || (cs == 2 && paircs_on && 2 * ntw * L->ndir <= 132));
  p.trace = g_dbg_timing != nullptr ? g_dbg_timing + 32768 : nullptr;   // the diagnostics buffer holds 65536 entries then
  if (cs > 1) p.timing = nullptr;
  char* wsp = reinterpret_cast<char*>(ws);
  bool save = false;
  for (int d = 0; d < L->ndir; ++d) {
    const IpnGruDir& D = L->dir[d];
    GruPersistFwdDir& o = p.d[d];
    IPN_PROPAGATE(get_tensor_map_3d(&o.tmW, D.w_hh, (unsigned long long)H, (unsigned long long)H, 3ULL, H, (long long)H * H,
                                    pair ? 32 : 64, 3));
    IPN_PROPAGATE(get_tensor_map(&o.tmH, D.hseq, (unsigned long long)H, (unsigned long long)(T + 1) * Bt, H, GP_ROWS));
    o.has_y = L->y != nullptr;
    if (o.has_y)
      IPN_PROPAGATE(get_tensor_map(&o.tmY, L->y, (unsigned long long)L->ld_y, (unsigned long long)T * Bt, L->ld_y, GP_ROWS));
    o.b_hh = D.b_hh;
    o.pvec = D.pvec;
    o.reverse = D.reverse;
    o.y_col0 = D.y_col0;
    o.gates = reinterpret_cast<uint4*>(D.gates);
    save = save || D.gates != nullptr;
    const int tt_min = D.reverse ? T - L->s_end : L->s_begin;   // earliest time index this call touches
    if (dir_gathers_table(D) || dir_gathers_table_plus_bcast(D)) {
      const bool plus = dir_gathers_table_plus_bcast(D);
      __nv_bfloat16* ft = reinterpret_cast<__nv_bfloat16*>(wsp);
      IPN_REQUIRE(reinterpret_cast<uintptr_t>(ft) % 32 == 0, IPN_ERR_ALIGN, "gru_persist_fwd: workspace must be 32-byte aligned");
      ProfScope prof("gru_fold_table", 0.0, (double)D.table_rows * 3 * H * 6, stream);
      gru_fold_table_kernel<<<(D.table_rows * 3 * H + 255) / 256, 256, 0, stream>>>(D.table, D.ld_table, D.table_rows,
                                                                                    plus ? nullptr : D.b_hh, H, ft);
      IPN_LAUNCH_CHECK();
      o.ftab = reinterpret_cast<const uint4*>(ft);
      o.tok = D.tok;
      o.Pblk = nullptr;
      if (plus) {   // the blocked per-row term: same tiles at every step
        o.Pblk = reinterpret_cast<const uint4*>(D.P);
        o.p_t_stride = 0;
        o.p_t0 = L->row0 / GP_ROWS;
      }
      wsp += GPF_FOLD_ROWS * 3LL * H * 2;
    } else if (D.P_blocked) {
      o.Pblk = reinterpret_cast<const uint4*>(D.P);   // global blocked layout over all T*Bt rows
      o.p_t_stride = Bt / GP_ROWS;
      o.p_t0 = L->row0 / GP_ROWS;
    } else {
      o.Pblk = reinterpret_cast<const uint4*>(wsp);   // local layout: only this call's window
      o.p_t_stride = ntw;
      o.p_t0 = -(long long)tt_min * ntw;
      PrepP q;
      q.P = reinterpret_cast<const __nv_bfloat16*>(D.P);
      q.ldP = D.ldP; q.P_bcast = D.P_bcast;
      q.table = D.table; q.ld_table = D.ld_table; q.tok = D.tok;
      q.pvec = D.pvec; q.b_hh = D.b_hh;
      q.out = reinterpret_cast<uint4*>(wsp);
      q.H = H; q.Bt = Bt;
      q.ntw = ntw; q.tt_min = tt_min; q.row0 = L->row0;
      dim3 grid((unsigned)((long long)NS * ntw), 3, H / 64);
      ProfScope prof("gru_prep_p", 0.0, (double)per_dir * (D.P != nullptr && !D.P_bcast ? 2.0 : 1.0), stream);
      gru_prep_p_kernel<<<grid, 256, 0, stream>>>(q);
      IPN_LAUNCH_CHECK();
      wsp += per_dir;
    }
  }
  for (int d = 0; d < L->ndir; ++d)
    IPN_REQUIRE((L->dir[d].gates != nullptr) == save, IPN_ERR_ARG, "gru_persist_fwd: gates must be given for all directions or none");
  const int smem = gpf_smem_bytes(H);
  auto launch = [&](auto kern, bool* configured) -> int {
    if (!*configured) {
      IPN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gpf_smem_bytes(512)));
      *configured = true;
    }
    // algorithmic work of the whole layer: recurrent GEMM flops; bytes = P read + h write + saved gates + y
    const double rows = (double)NS * L->nrows * L->ndir;
    ProfScope prof("gru_layer_fwd_persist", 2.0 * rows * 3.0 * H * H,
                   rows * H * 2.0 * (3 + 1 + (save ? GP_GATE_ARRAYS : 0) + (L->y ? 1 : 0)), stream);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(ntw * cs, L->ndir, 1);
    cfg.blockDim = dim3(GP_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pair ? 2 * cs : cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    IPN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
    IPN_LAUNCH_CHECK();
    return IPN_OK;
  };
  static bool cfgd[10] = {false, false, false, false, false, false, false, false, false, false};
  if (cs == 2 && pair && save) IPN_PROPAGATE(launch(gru_persist_fwd_kernel<true, true, 2>, &cfgd[8]));
  else if (cs == 2 && pair) IPN_PROPAGATE(launch(gru_persist_fwd_kernel<false, true, 2>, &cfgd[9]));
  else if (cs == 4 && save) IPN_PROPAGATE(launch(gru_persist_fwd_kernel<true, false, 4>, &cfgd[6]));
  else if (cs == 4) IPN_PROPAGATE(launch(gru_persist_fwd_kernel<false, false, 4>, &cfgd[7]));
  else if (cs == 2 && save) IPN_PROPAGATE(launch(gru_persist_fwd_kernel<true, false, 2>, &cfgd[4]));
  else if (cs == 2) IPN_PROPAGATE(launch(gru_persist_fwd_kernel<false, false, 2>, &cfgd[5]));
  else if (save && pair) IPN_PROPAGATE(launch(gru_persist_fwd_kernel<true, true, 1>, &cfgd[0]));
  else if (save) IPN_PROPAGATE(launch(gru_persist_fwd_kernel<true, false, 1>, &cfgd[1]));
  else if (pair) IPN_PROPAGATE(launch(gru_persist_fwd_kernel<false, true, 1>, &cfgd[2]));
  else IPN_PROPAGATE(launch(gru_persist_fwd_kernel<false, false, 1>, &cfgd[3]));
  // inter-layer dropout on the layer output
  if (L->y != nullptr && L->mask != nullptr) {
    for (int d = 0; d < L->ndir; ++d) {
      const long long rows = (long long)NS * L->nrows;
      const long long work = rows * (H / 8);
      const int tt_min = L->dir[d].reverse ? T - L->s_end : L->s_begin;
      ProfScope prof("gru_mask_y", 0.0, (double)rows * H * 5.0, stream);
      gru_mask_y_kernel<<<(unsigned)((work + 255) / 256), 256, 0, stream>>>(
          reinterpret_cast<__nv_bfloat16*>(L->y), L->ld_y, L->mask, L->ld_mask, L->dir[d].y_col0, H, rows, L->mask_scale,
          L->nrows, Bt, tt_min, L->row0);
      IPN_LAUNCH_CHECK();
    }
  }
  // final hidden state of every direction into its consumer's buffer
  for (int d = 0; d < L->ndir; ++d) {
    const IpnGruDir& D = L->dir[d];
    void* fo = D.final_out_dir != nullptr ? D.final_out_dir : L->final_out;
    if (fo == nullptr || L->s_end != T) continue;   // the final state exists once the last step has run
    const int fdt = D.final_out_dir != nullptr ? D.final_dir_dt : L->final_dt;
    const long long ldf = D.final_out_dir != nullptr ? D.ld_final_dir : L->ld_final;
    const int fes = fdt == IPN_BF16 ? 2 : 4;
    const char* src = reinterpret_cast<const char*>(D.hseq) + ((long long)(D.reverse ? 0 : T) * Bt + L->row0) * H * 2;
    char* dst = reinterpret_cast<char*>(fo) + ((long long)L->row0 * ldf + D.final_col0) * fes;
    IPN_PROPAGATE(ipn_convert_2d(src, IPN_BF16, H, dst, fdt, ldf, L->nrows, H, stream));
  }
  return IPN_OK;
}

}  // namespace ipn
