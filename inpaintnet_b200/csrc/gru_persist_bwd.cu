// Persistent GRU layer, backward (BPTT): one launch differentiates ALL timesteps of a layer.
//
// A CTA owns a tile of 128 batch rows of one direction (same tiling as the forward kernel).  Per step s
// (reverse processing order) it alternates two phases that are serial by data dependence:
//   E(s): pointwise gate derivatives of step s from dh_s (TMEM accumulator) + the saved gates (blocked
//         layout) + dY: writes dP_s / dGn_s (row-major, staging tile + TMA store: the hoisted weight-gradient
//         GEMMs consume them) and leaves the carry dh_s * z_s IN THE ACCUMULATOR (tcgen05.st);
//   M(s): dh_{s-1} = carry + [dP_r, dP_z, dGn]_s . W_hh  -- K = 3H streamed through shared memory (A k-blocks
//         re-read from L2, W_hh as an MN-major B operand), N = H in TMEM, accumulate on top of the carry.
#include "gru_persist.cuh"
#include <stdlib.h>

namespace ipn {

// dY prep: row-major dY[:, col0:col0+H] (* keep-mask * scale) -> blocked bf16 [rows, 1, H]
struct PrepDY {
  const __nv_bfloat16* dY;
  long long ld_dy;
  const unsigned char* mask;
  long long ld_mask;
  float scale;
  int col0;
  uint4* out;
  int H;
};
__global__ void __launch_bounds__(256) gru_prep_dy_kernel(PrepDY p) {
  __shared__ uint4 tile[128 * 8];
  const int rt = blockIdx.x, c = blockIdx.y;
  const int i = threadIdx.x;
  const int vec = i & 7;
  const int col = p.col0 + c * 64 + vec * 8;
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int row = pass * 32 + (i >> 3);
    const long long R = (long long)rt * 128 + row;
    uint4 u = *reinterpret_cast<const uint4*>(p.dY + R * p.ld_dy + col);
    if (p.mask != nullptr) {
      const uint2 m = *reinterpret_cast<const uint2*>(p.mask + R * p.ld_mask + col);
      const unsigned char* mb = reinterpret_cast<const unsigned char*>(&m);
      float f[8];
      unpack8(u, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = mb[k] ? f[k] * p.scale : 0.f;
      u = pack8(f);
    }
    tile[row * 8 + (vec ^ (row & 7))] = u;
  }
  __syncthreads();
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int v = pass * 2 + (i >> 7), row = i & 127;
    p.out[((long long)rt * (p.H / 8) + c * 8 + v) * 128 + row] = tile[row * 8 + (v ^ (row & 7))];
  }
}

constexpr int GPB_A_STAGES = 3;        // A k-blocks are re-read from L2 (latency ~1.5 kcycles): 2 stages starve the MMAs
constexpr int GPB_W_RING = 112 * 1024;   // 3 stages of a 32 KB column-split stage (2 starve the MMAs: the ring round trip is ~2.8 kcycles); all the shared memory there is
constexpr int GPB_NBAR = 48;
static inline int gpb_smem_bytes() { return GPB_A_STAGES * GP_KB_BYTES + GPB_W_RING + 4 * GP_KB_BYTES + GPB_NBAR * 8 + 16; }

// CS = 2 (on by default since round 2, IPN_GPB_CS=0 disables; validated on B200 with the GPU suite): COLUMN SPLIT, the
// mirror of the forward kernel's.  The two CTAs of a cluster own the same 128-row tile and one 256-column half of dh
// each (H = 512): half of the E phase (gate derivatives of their units) and half of every GEMM phase (N = 256 of
// the 512 dh columns).  The K = 3H operand [dP_r, dP_z, dGn] of the GEMM phase spans both halves: each CTA
// TMA-stores its chunks as before and signals the PEER's dg_stored barrier once the store has completed.
template <bool PAIR, int CS>
__global__ void __launch_bounds__(GP_THREADS, 1) gru_persist_bwd_kernel(const __grid_constant__ GruPersistBwd p) {
  static_assert(CS == 1 || CS == 2, "column split: 2 CTAs per tile, or 2 CTA pairs per two tiles");
  extern __shared__ __align__(1024) uint8_t smem[];
  const GruPersistBwdDir& D = p.d[blockIdx.y];
  const int H = p.H, KB = H >> 6, T = p.T, Bt = p.Bt;
  const int NH = H > 256 ? 2 : 1;        // accumulator halves (one MMA covers at most 256 columns)
  const int NPH = H / NH;                // columns per MMA
  const int WSTB = p.nbs * 8192;         // bytes of one W stage in this CTA
  const int WSTAGES = min(6, GPB_W_RING / WSTB);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = (PAIR || CS > 1) ? ptx::cluster_ctarank() : 0u;
  const uint32_t rank = PAIR ? (crank & 1u) : 0u;   // pair: which 128 rows of the M = 256 tile / which half of every W stage
  const bool leader = rank == 0;
  const uint32_t lead_rank = crank & ~1u;           // cluster rank of this pair's leader CTA
  const uint16_t pair_mask = (uint16_t)(3u << lead_rank);
  // PAIR + CS = 2 (as in the forward kernel): a cluster of 4 = two pairs, pair `crank >> 1` owns one 256-column half
  // of dh for TWO adjacent row tiles; each CTA stages half of every W stage, so the ring holds twice as many
  const uint32_t colrank = PAIR ? (crank >> 1) : crank;
  const int rbase = (PAIR ? ((int)blockIdx.x / (2 * CS)) * 2 + (int)rank : (int)blockIdx.x / CS) * GP_ROWS;
  const int h_lo = CS > 1 ? (int)colrank : 0, h_hi = CS > 1 ? h_lo + 1 : NH;            // accumulator halves of this CTA
  const int c_lo = CS > 1 ? (int)colrank * (KB / CS) : 0, c_hi = c_lo + KB / CS;        // 64-unit chunks of this CTA
  const int nM = D.dh0 != nullptr ? T : T - 1;   // number of GEMM phases (the last one only feeds dh0)

  uint8_t* sA = smem;
  uint8_t* sW = sA + GPB_A_STAGES * GP_KB_BYTES;
  uint8_t* sStg = sW + GPB_W_RING;               // 4 tiles: dP_r, dP_z, dP_n, dGn
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStg + 4 * GP_KB_BYTES);
  uint64_t* w_full = bars;             // [6]
  uint64_t* w_empty = bars + 6;        // [6]
  uint64_t* a_full = bars + 12;        // [4]
  uint64_t* a_empty = bars + 16;       // [4]
  uint64_t* dg_stored = bars + 20;     // [8] chunk c of dP/dGn of this step is in global memory
  uint64_t* tmem_full = bars + 28;     // GEMM phase complete: dh of the next step is in the accumulator
  uint64_t* tmem_free = bars + 29;     // epilogue finished reading dh and writing the carry
  uint64_t* stg_ready = bars + 30;
  uint64_t* stg_free = bars + 31;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);
  uint64_t* dg_stored_odd = bars + 34; // [8] column split: the barriers of odd steps (signalled remotely: see the forward kernel)

  if ((ptx::smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("inpaintnet_b200: gru_persist_bwd: shared memory base not 1024-byte aligned\n");
    __trap();
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < 6; ++s) { ptx::mbar_init(&w_full[s], 1); ptx::mbar_init(&w_empty[s], 1); }
      for (int s = 0; s < GPB_A_STAGES; ++s) { ptx::mbar_init(&a_full[s], 1); ptx::mbar_init(&a_empty[s], 1); }
      for (int k = 0; k < 8; ++k) ptx::mbar_init(&dg_stored[k], 1);
      if (CS > 1) for (int k = 0; k < 8; ++k) ptx::mbar_init(&dg_stored_odd[k], 1);
      ptx::mbar_init(tmem_full, 1);
      ptx::mbar_init(tmem_free, PAIR ? 32 : 16);
      ptx::mbar_init(stg_ready, 16);
      ptx::mbar_init(stg_free, 1);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    if (PAIR) { ptx::tmem_alloc_pair<512>(tmem_slot); ptx::tmem_relinquish_pair(); }
    else { ptx::tmem_alloc<512>(tmem_slot); ptx::tmem_relinquish(); }
  } else if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&D.tmW);
    ptx::prefetch_tmap(&D.tmDP);
    ptx::prefetch_tmap(&D.tmDG);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (PAIR || CS > 1) ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
  ptx::setmaxnreg_dec<56>();
  if (warp == 0) {
    // ===================== W_hh producer (same order as the MMA issuer: step, k-chunk, gate, half) ==========
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int vpr = H >> 3;
      // the epilogue's saved-gate / dY tiles are pulled into L2 one chunk ahead by this otherwise idle thread
      auto prefetch_e = [&](int it, int c) {
        if (it >= T || (p.dbg & 16)) return;
        const int s = T - 1 - it;
        const int tt = D.reverse ? T - 1 - s : s;
        const long long rt = ((long long)tt * Bt + rbase) >> 7;
#pragma unroll
        for (int a = 0; a < GP_GATE_ARRAYS; ++a)
          ptx::bulk_prefetch_l2(D.gates + ((rt * GP_GATE_ARRAYS + a) * vpr + c * 8) * 128, 8 * 128 * 16);
        if (D.dYblk != nullptr) ptx::bulk_prefetch_l2(D.dYblk + (rt * vpr + c * 8) * 128, 8 * 128 * 16);
      };
      for (int c = c_lo; c < c_hi; ++c) prefetch_e(0, c);
      for (int it = 0; it < nM; ++it)
        for (int kb = 0; kb < KB; ++kb) {
          if (kb >= c_lo && kb < c_hi) prefetch_e(it + 1, kb);
          for (int g = 0; g < 3; ++g)
            for (int h = h_lo; h < h_hi; ++h) {
              ptx::mbar_wait(&w_empty[stage], phase ^ 1);
              if (leader) ptx::mbar_arrive_expect_tx(&w_full[stage], (uint32_t)(WSTB * (PAIR ? 2 : 1)));
              const int nblk = h * (NPH >> 6) + (int)rank * p.nbs;
              if (PAIR) ptx::tma_load_3d_pair(sW + stage * WSTB, &D.tmW, &w_full[stage], 0, g * H + kb * 64, nblk);
              else ptx::tma_load_3d(sW + stage * WSTB, &D.tmW, &w_full[stage], 0, g * H + kb * 64, nblk);
              if (++stage == WSTAGES) { stage = 0; phase ^= 1; }
            }
        }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (leader) {
      const uint32_t idesc = ptx::make_idesc_bf16(PAIR ? 256 : 128, NPH, 0, 1);
      int ws = 0, as = 0;
      uint32_t wph = 0, aph = 0;
      const bool tm = p.timing != nullptr;
      long long w_tf = 0, w_af = 0, w_wf = 0;
      const long long t_begin = clock64();
      const uint64_t descA0 = ptx::make_smem_desc(ptx::smem_u32(sA), 16, 1024);
      const uint64_t descW0 = ptx::make_smem_desc(ptx::smem_u32(sW), 8192, 1024);
      for (int it = 0; it < nM; ++it) {
        wait_acc(tmem_free, it & 1, tm, w_tf);   // carry of this step is in the accumulator
        ptx::tc_fence_after();
        for (int j = 0; j < 3 * KB; ++j) {
          wait_acc(&a_full[as], aph, tm, w_af);
          const uint64_t da0 = descA0 + (uint64_t)((as * GP_KB_BYTES) >> 4);
          for (int h = h_lo; h < h_hi; ++h) {
            wait_acc(&w_full[ws], wph, tm, w_wf);
            ptx::tc_fence_after();
            const uint64_t dw0 = descW0 + (uint64_t)((ws * WSTB) >> 4);
            const uint32_t dcol = tmem_base + (uint32_t)(h * NPH);
            if (ptx::elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                if (PAIR) ptx::umma_bf16_pair(dcol, da0 + (uint64_t)(kk * 2), dw0 + (uint64_t)(kk * 128), idesc, 1u);
                else ptx::umma_bf16(dcol, da0 + (uint64_t)(kk * 2), dw0 + (uint64_t)(kk * 128), idesc, 1u);
              }
              if (PAIR) ptx::umma_commit_pair(&w_empty[ws], pair_mask);
              else ptx::umma_commit(&w_empty[ws]);
            }
            __syncwarp();
            if (++ws == WSTAGES) { ws = 0; wph ^= 1; }
          }
          if (ptx::elect_one()) {
            if (PAIR) ptx::umma_commit_pair(&a_empty[as], pair_mask);
            else ptx::umma_commit(&a_empty[as]);
          }
          __syncwarp();
          if (++as == GPB_A_STAGES) { as = 0; aph ^= 1; }
        }
        if (ptx::elect_one()) {
          if (PAIR) ptx::umma_commit_pair(tmem_full, pair_mask);
          else ptx::umma_commit(tmem_full);
        }
        __syncwarp();
      }
      if (tm && lane == 0) {
        unsigned long long* o = p.timing + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * 16;
        o[0] = (unsigned long long)(clock64() - t_begin); o[1] = w_tf; o[2] = w_af; o[3] = w_wf;
      }
    }
  } else if (warp == 2) {
    // ===================== store warp: 4 staging tiles -> dP (3 gate column blocks) and dGn =====================
    if (lane == 0) {
      int i = 0;
      for (int it = 0; it < T; ++it) {
        const int s = T - 1 - it;
        const int tt = D.reverse ? T - 1 - s : s;
        const int grow = tt * Bt + rbase;
        for (int c = c_lo; c < c_hi; ++c, ++i) {
          ptx::mbar_wait(stg_ready, i & 1);
#pragma unroll
          for (int g = 0; g < 3; ++g) ptx::tma_store_2d(&D.tmDP, sStg + g * GP_KB_BYTES, g * H + c * 64, grow);
          ptx::tma_store_2d(&D.tmDG, sStg + 3 * GP_KB_BYTES, c * 64, grow);
          ptx::bulk_commit();
          ptx::bulk_wait_read0();
          ptx::mbar_arrive(stg_free);
          ptx::bulk_wait0();
          if (CS > 1) {   // the peer streams these dP / dGn chunks as k-blocks of its GEMM phase
            uint64_t* ds = (it & 1) ? &dg_stored_odd[c] : &dg_stored[c];
            ptx::mbar_arrive(ds);
            ptx::fence_acq_rel_cluster();
            ptx::mbar_arrive_remote_relaxed(ds, PAIR ? (crank ^ 2u) : (crank ^ 1u));   // same rows, the other half of dh
          } else {
            ptx::mbar_arrive(&dg_stored[c]);
          }
        }
      }
    }
  } else {
    // ===================== A loader: [dP_r, dP_z, dGn] k-blocks of the step just differentiated =====================
    if (lane == 0) {
      int as = 0;
      uint32_t aph = 0;
      for (int it = 0; it < nM; ++it) {
        const int s = T - 1 - it;
        const int tt = D.reverse ? T - 1 - s : s;
        const int grow = tt * Bt + rbase;
        for (int kb = 0; kb < KB; ++kb) {
          if (CS > 1) {   // barrier set it & 1, phase (it >> 1) & 1; the peer's chunks at cluster scope
            uint64_t* ds = (it & 1) ? &dg_stored_odd[kb] : &dg_stored[kb];
            if (kb < c_lo || kb >= c_hi) ptx::mbar_wait_cluster(ds, (it >> 1) & 1);
            else ptx::mbar_wait(ds, (it >> 1) & 1);
          } else {
            ptx::mbar_wait(&dg_stored[kb], it & 1);
          }
          ptx::fence_proxy_async_all();
          for (int g = 0; g < 3; ++g) {
            ptx::mbar_wait(&a_empty[as], aph ^ 1);
            if (leader) ptx::mbar_arrive_expect_tx(&a_full[as], PAIR ? 2 * GP_KB_BYTES : GP_KB_BYTES);
            const CUtensorMap* tm = g < 2 ? &D.tmDP : &D.tmDG;
            const int col = g < 2 ? g * H + kb * 64 : kb * 64;
            if (PAIR) ptx::tma_load_2d_pair(sA + as * GP_KB_BYTES, tm, &a_full[as], col, grow);
            else ptx::tma_load_2d(sA + as * GP_KB_BYTES, tm, &a_full[as], col, grow);
            if (++as == GPB_A_STAGES) { as = 0; aph ^= 1; }
          }
        }
      }
    }
  }
  } else {
    // ===================== epilogue warps 4..19: warp = (TMEM lane quadrant, 16-unit sub-chunk) =====================
    ptx::setmaxnreg_inc<104>();
    const int q = warp & 3;
    const int sub = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const int vpr = H >> 3;
    const uint32_t sStg_u = ptx::smem_u32(sStg);
    const uint32_t sw = (uint32_t)(row & 7);
    const long long astride = (long long)vpr * 128;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    int i = 0;
    for (int it = 0; it < T; ++it) {
      const int s = T - 1 - it;
      const int tt = D.reverse ? T - 1 - s : s;
      const long long rt = ((long long)tt * Bt + rbase) >> 7;
      for (int c = c_lo; c < c_hi; ++c, ++i) {
        const uint4* gp = D.gates + ((rt * GP_GATE_ARRAYS) * vpr + c * 8 + sub * 2) * 128 + row;
        const uint4* yp = D.dYblk != nullptr ? D.dYblk + (rt * vpr + c * 8 + sub * 2) * 128 + row : nullptr;
        const int u0 = c * 64 + sub * 16;
        // all global loads of the chunk are issued up front (one memory round trip per chunk)
        uint4 gv[2][5], yv[2];
#pragma unroll
        for (int v = 0; v < 2; ++v) {
#pragma unroll
          for (int a = 0; a < 5; ++a) gv[v][a] = ldg_stream(gp + a * astride + v * 128);
          yv[v] = yp != nullptr ? ldg_stream(yp + v * 128) : make_uint4(0, 0, 0, 0);
        }
        if (c == c_lo && it > 0) {   // the first chunk's loads are in flight while the GEMM phase finishes
          ptx::mbar_wait(tmem_full, (it - 1) & 1);
          ptx::tc_fence_after();
        }
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const uint4 gr = gv[v][0], gz = gv[v][1], gn = gv[v][2], gh = gv[v][3], gq = gv[v][4];
          float dh[8];
          if (it > 0) ptx::tmem_ld8(tlane + (uint32_t)(u0 + v * 8), dh);
          else {
#pragma unroll
            for (int k = 0; k < 8; ++k) dh[k] = 0.f;
          }
          float r[8], z[8], n[8], hn[8], hp[8];
          unpack8(gr, r); unpack8(gz, z); unpack8(gn, n); unpack8(gh, hn); unpack8(gq, hp);
          if (it > 0) ptx::tmem_ld_wait();
          if (yp != nullptr) {
            float dy[8];
            unpack8(yv[v], dy);
#pragma unroll
            for (int k = 0; k < 8; ++k) dh[k] += dy[k];
          }
          if (it == 0 && D.dh_n != nullptr) {
            const float4* hp4 = reinterpret_cast<const float4*>(D.dh_n + (long long)(rbase + row) * D.ld_dhn + u0 + v * 8);
            const float4 a = hp4[0], b = hp4[1];
            dh[0] += a.x; dh[1] += a.y; dh[2] += a.z; dh[3] += a.w;
            dh[4] += b.x; dh[5] += b.y; dh[6] += b.z; dh[7] += b.w;
          }
          float dr[8], dz[8], dn[8], dg[8], cy[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float dnk = dh[k] * (1.f - z[k]) * (1.f - n[k] * n[k]);
            dn[k] = dnk;
            dz[k] = dh[k] * (hp[k] - n[k]) * z[k] * (1.f - z[k]);
            dr[k] = dnk * hn[k] * r[k] * (1.f - r[k]);
            dg[k] = dnk * r[k];
            cy[k] = dh[k] * z[k];
          }
          ptx::tmem_st8(tlane + (uint32_t)(u0 + v * 8), cy);   // carry stays in the accumulator
          if (v == 0) ptx::mbar_wait(stg_free, (i & 1) ^ 1);
          const uint32_t so = row * 128 + (((uint32_t)(sub * 2 + v) ^ sw) << 4);
          st_shared_v4(sStg_u + so, pack8(dr));
          st_shared_v4(sStg_u + GP_KB_BYTES + so, pack8(dz));
          st_shared_v4(sStg_u + 2 * GP_KB_BYTES + so, pack8(dn));
          st_shared_v4(sStg_u + 3 * GP_KB_BYTES + so, pack8(dg));
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(stg_ready);
      }
      // the accumulator now holds the carry of every column this warp owns: release it to the MMA issuer
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) ptx::mbar_arrive(tmem_free);
        else ptx::mbar_arrive_remote(tmem_free, lead_rank);
      }
    }
    if (D.dh0 != nullptr) {
      // gradient wrt the initial state: accumulator after the last GEMM phase
      ptx::mbar_wait(tmem_full, (T - 1) & 1);
      ptx::tc_fence_after();
      for (int c = c_lo; c < c_hi; ++c) {
        const int u0 = c * 64 + sub * 16;
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          float dh[8];
          ptx::tmem_ld8(tlane + (uint32_t)(u0 + v * 8), dh);
          ptx::tmem_ld_wait();
          const long long R = rbase + row;
          if (D.dh0_selu) {
            float h0[8];
            unpack8(*reinterpret_cast<const uint4*>(D.h0 + R * H + u0 + v * 8), h0);
#pragma unroll
            for (int k = 0; k < 8; ++k) dh[k] *= (h0[k] > 0.f ? 1.0507009873554804934193349852946f
                                                              : h0[k] + 1.0507009873554804934193349852946f * 1.6732632423543772848170429916717f);
          }
          if (D.dh0_dt == IPN_BF16) {
            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(D.dh0) + R * D.ld_dh0 + u0 + v * 8) = pack8(dh);
          } else {
            float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(D.dh0) + R * D.ld_dh0 + u0 + v * 8);
            o[0] = make_float4(dh[0], dh[1], dh[2], dh[3]);
            o[1] = make_float4(dh[4], dh[5], dh[6], dh[7]);
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (PAIR || CS > 1) ptx::cluster_sync_all();
  if (warp == 1) {
    ptx::tc_fence_after();
    if (PAIR) ptx::tmem_dealloc_pair<512>(tmem_base);
    else ptx::tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static bool al16b(const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; }

bool gru_persist_bwd_shape_ok(const IpnGruLayerBwd* L) {
  static const int on = getenv("IPN_PERSIST_BWD") ? atoi(getenv("IPN_PERSIST_BWD")) : 1;
  if (!on) return false;
  if (L->core != IPN_CORE_UMMA || L->act_dt != IPN_BF16) return false;
  if (L->H % 64 != 0 || L->H < 64 || L->H > 512) return false;
  if (L->B_total % GP_ROWS != 0 || L->row0 != 0 || L->nrows != L->B_total) return false;
  if (L->dY != nullptr && (L->ld_dy % 8 != 0 || !al16b(L->dY))) return false;
  if (L->mask != nullptr && (L->ld_mask % 8 != 0 || reinterpret_cast<uintptr_t>(L->mask) % 8 != 0)) return false;
  for (int d = 0; d < L->ndir; ++d) {
    const IpnGruBwdDir& D = L->dir[d];
    if (D.y_col0 % 8 != 0 || !al16b(D.gates) || !al16b(D.dP) || !al16b(D.dGn)) return false;
    if (D.dh_n != nullptr && (D.ld_dhn % 4 != 0 || !al16b(D.dh_n))) return false;
    if (D.dh0 != nullptr && (!al16b(D.dh0) || D.ld_dh0 % 8 != 0 || (D.dh0_dt != IPN_BF16 && D.dh0_dt != IPN_F32))) return false;
  }
  return true;
}

long long gru_persist_bwd_ws_bytes(const IpnGruLayerBwd* L) {
  if (!gru_persist_bwd_shape_ok(L)) return 0;
  const long long per_dir = (long long)L->T * L->B_total * L->H * 2;
  return L->dY != nullptr ? per_dir * L->ndir : 16;
}

int gru_persist_bwd(const IpnGruLayerBwd* L, void* ws, long long ws_bytes, cudaStream_t stream) {
  const int T = L->T, H = L->H, Bt = L->B_total;
  IPN_REQUIRE(ws != nullptr && ws_bytes >= gru_persist_bwd_ws_bytes(L) && al16b(ws), IPN_ERR_ARG,
              "gru_persist_bwd: workspace too small (%lld < %lld)", ws_bytes, gru_persist_bwd_ws_bytes(L));
  GruPersistBwd p;
  memset(&p, 0, sizeof(p));
  p.T = T; p.H = H; p.Bt = Bt;
  static const int dbg = getenv("IPN_GPB_DBG") ? atoi(getenv("IPN_GPB_DBG")) : 0;
  p.dbg = dbg;
  p.timing = g_dbg_timing;
  const int NH = H > 256 ? 2 : 1, NPH = H / NH;
  static const int pair_on = getenv("IPN_GPB_PAIR") ? atoi(getenv("IPN_GPB_PAIR")) : 1;
  // column split (kernel header comment): on unless IPN_GPB_CS=0
  static const int cs_on = getenv("IPN_GPB_CS") ? atoi(getenv("IPN_GPB_CS")) : 1;
  const bool cs = cs_on && NH == 2 && 2 * (Bt / GP_ROWS) * L->ndir <= 148;
  // CTA pairs inside the column split: opt-in (IPN_GPB_PAIRCS=1).  Unlike the forward kernel it measured SLOWER here
  // (layer backward 4.84 vs 4.35 ms per two train steps) once the W ring holds 3 full stages: the pair couples two
  // row tiles' E phases, and the streamed A operand already keeps the per-stage MMA work at 684 cycles.
  static const int paircs_on = getenv("IPN_GPB_PAIRCS") ? atoi(getenv("IPN_GPB_PAIRCS")) : 0;
  const bool pair = pair_on && (Bt / GP_ROWS) % 2 == 0 && NPH % 128 == 0 &&
                    (!cs || (paircs_on && 2 * (Bt / GP_ROWS) * L->ndir <= 132));
  if (cs) p.timing = nullptr;   // the diagnostics buffer is sized for one CTA per tile
  p.nbs = pair ? NPH / 128 : NPH / 64;
  const long long per_dir = (long long)T * Bt * H * 2;
  char* wsp = reinterpret_cast<char*>(ws);
  for (int d = 0; d < L->ndir; ++d) {
    const IpnGruBwdDir& D = L->dir[d];
    GruPersistBwdDir& o = p.d[d];
    // W_hh [3H, H] row-major viewed as {n inner 64, k 3H (stride H), n-block H/64 (stride 64)}
    IPN_PROPAGATE(get_tensor_map_3d(&o.tmW, D.w_hh, 64ULL, 3ULL * H, (unsigned long long)(H / 64), H, 64, 64, (unsigned)p.nbs));
    IPN_PROPAGATE(get_tensor_map(&o.tmDP, D.dP, 3ULL * H, (unsigned long long)T * Bt, 3LL * H, GP_ROWS));
    IPN_PROPAGATE(get_tensor_map(&o.tmDG, D.dGn, (unsigned long long)H, (unsigned long long)T * Bt, H, GP_ROWS));
    o.gates = reinterpret_cast<const uint4*>(D.gates);
    o.dh_n = D.dh_n; o.ld_dhn = D.ld_dhn;
    o.dh0 = D.dh0; o.ld_dh0 = D.ld_dh0; o.dh0_dt = D.dh0_dt; o.dh0_selu = D.dh0_selu;
    o.reverse = D.reverse;
    o.h0 = reinterpret_cast<const __nv_bfloat16*>(D.hseq) + (long long)(D.reverse ? T : 0) * Bt * H;
    if (L->dY != nullptr) {
      o.dYblk = reinterpret_cast<const uint4*>(wsp);
      PrepDY q;
      q.dY = reinterpret_cast<const __nv_bfloat16*>(L->dY); q.ld_dy = L->ld_dy;
      q.mask = L->mask; q.ld_mask = L->ld_mask; q.scale = L->mask_scale; q.col0 = D.y_col0;
      q.out = reinterpret_cast<uint4*>(wsp); q.H = H;
      dim3 grid((unsigned)((long long)T * Bt / 128), H / 64, 1);
      ProfScope prof("gru_prep_dy", 0.0, (double)per_dir * 2.0, stream);
      gru_prep_dy_kernel<<<grid, 256, 0, stream>>>(q);
      IPN_LAUNCH_CHECK();
      wsp += per_dir;
    }
  }
  auto launch = [&](auto kern, bool* configured) -> int {
    if (!*configured) {
      IPN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gpb_smem_bytes()));
      *configured = true;
    }
    const double rows = (double)T * Bt * L->ndir;
    ProfScope prof("gru_layer_bwd_persist", 2.0 * rows * 3.0 * H * H,
                   rows * H * 2.0 * (GP_GATE_ARRAYS + 4 + (L->dY ? 1 : 0)), stream);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(Bt / GP_ROWS * (cs ? 2 : 1), L->ndir, 1);
    cfg.blockDim = dim3(GP_THREADS, 1, 1);
    cfg.dynamicSmemBytes = gpb_smem_bytes();
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (pair && cs) ? 4 : (pair || cs) ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    IPN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
    IPN_LAUNCH_CHECK();
    return IPN_OK;
  };
  static bool cfgd[4] = {false, false, false, false};
  if (cs && pair) IPN_PROPAGATE(launch(gru_persist_bwd_kernel<true, 2>, &cfgd[3]));
  else if (cs) IPN_PROPAGATE(launch(gru_persist_bwd_kernel<false, 2>, &cfgd[2]));
  else if (pair) IPN_PROPAGATE(launch(gru_persist_bwd_kernel<true, 1>, &cfgd[0]));
  else IPN_PROPAGATE(launch(gru_persist_bwd_kernel<false, 1>, &cfgd[1]));
  return IPN_OK;
}

}  // namespace ipn
