// GRU layer forward / backward drivers: one fused kernel per timestep (recurrent GEMM + gate math
// in the epilogue), all directions of the layer in the same launch.
#include "launch.cuh"
#include <stdlib.h>

namespace ipn {

template <int W>
__global__ void gru_bwd_point_kernel(GruBwdPoint p0, GruBwdPoint p1, int nrows) {
  // thread = (column, chunk of W rows): consecutive threads -> consecutive columns (coalesced)
  const GruBwdPoint& p = blockIdx.z == 0 ? p0 : p1;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int rchunks = (nrows + W - 1) / W;
  if (idx >= (long long)rchunks * p.H) return;
  const int col = (int)(idx % p.H);
  const int row0 = (int)(idx / p.H) * W;
  float dh[W];
#pragma unroll
  for (int i = 0; i < W; ++i) dh[i] = 0.f;
  gru_bwd_pointwise<-1, W>(p, col, row0, min(W, nrows - row0), dh);
}

static inline const char* slot_ptr(const void* base, long long slot, long long B_total, int H, int dt) {
  return reinterpret_cast<const char*>(base) + slot * B_total * H * (dt == IPN_BF16 ? 2 : 4);
}

static int check_layer_common(int core, int act_dt, int T, int B_total, int H, int row0, int nrows, int ndir) {
  IPN_REQUIRE(core == IPN_CORE_SIMT || core == IPN_CORE_UMMA, IPN_ERR_ARG, "gru: unknown core %d", core);
  IPN_REQUIRE(act_dt == IPN_F32 || act_dt == IPN_BF16, IPN_ERR_ARG, "gru: bad act dtype");
  IPN_REQUIRE(core != IPN_CORE_UMMA || act_dt == IPN_BF16, IPN_ERR_ARG, "gru: the tcgen05 core needs bf16 activations");
  IPN_REQUIRE(T > 0 && B_total > 0 && H > 0, IPN_ERR_ARG, "gru: bad sizes T=%d B=%d H=%d", T, B_total, H);
  IPN_REQUIRE(row0 >= 0 && nrows > 0 && row0 + nrows <= B_total, IPN_ERR_ARG, "gru: bad row window");
  IPN_REQUIRE(ndir == 1 || ndir == 2, IPN_ERR_ARG, "gru: ndir must be 1 or 2");
  IPN_REQUIRE(core != IPN_CORE_UMMA || H % 8 == 0, IPN_ERR_ALIGN, "gru: tcgen05 core needs H %% 8 == 0 (H=%d)", H);
  return IPN_OK;
}

// gru_persist.cu / gru_persist_bwd.cu
bool persist_enabled();
bool gru_persist_fwd_shape_ok(const IpnGruLayer* L);
long long gru_persist_fwd_ws_bytes(const IpnGruLayer* L);
int gru_persist_fwd(const IpnGruLayer* L, void* ws, long long ws_bytes, cudaStream_t stream);
bool gru_persist_bwd_shape_ok(const IpnGruLayerBwd* L);
long long gru_persist_bwd_ws_bytes(const IpnGruLayerBwd* L);
int gru_persist_bwd(const IpnGruLayerBwd* L, void* ws, long long ws_bytes, cudaStream_t stream);

}  // namespace ipn

using namespace ipn;

extern "C" long long ipn_gru_layer_fwd_ws_bytes(const IpnGruLayer* L) {
  if (L == nullptr || L->T <= 0 || L->B_total <= 0 || L->H <= 0 || (L->ndir != 1 && L->ndir != 2)) return 0;
  return gru_persist_fwd_ws_bytes(L);
}
extern "C" long long ipn_gru_layer_bwd_ws_bytes(const IpnGruLayerBwd* L) {
  if (L == nullptr || L->T <= 0 || L->B_total <= 0 || L->H <= 0 || (L->ndir != 1 && L->ndir != 2)) return 0;
  return gru_persist_bwd_ws_bytes(L);
}
extern "C" int ipn_gru_gates_cols(int H) { return 5 * H; }
extern "C" int ipn_gru_persist_eligible(int core, int act_dt, int B_total, int H) {
  IpnGruLayerBwd L;
  memset(&L, 0, sizeof(L));
  L.core = core; L.act_dt = act_dt; L.T = 1; L.B_total = B_total; L.H = H; L.row0 = 0; L.nrows = B_total; L.ndir = 0;
  return (persist_enabled() && gru_persist_bwd_shape_ok(&L)) ? 1 : 0;
}

namespace ipn { unsigned long long* g_dbg_timing = nullptr; }
extern "C" void ipn_dbg_set_timing_buffer(void* dev_ptr) { g_dbg_timing = reinterpret_cast<unsigned long long*>(dev_ptr); }

extern "C" int ipn_gru_layer_fwd(const IpnGruLayer* L, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  IPN_REQUIRE(L != nullptr, IPN_ERR_ARG, "gru_layer_fwd: null descriptor");
  IPN_PROPAGATE(ensure_device());
  IPN_PROPAGATE(check_layer_common(L->core, L->act_dt, L->T, L->B_total, L->H, L->row0, L->nrows, L->ndir));
  IPN_REQUIRE(0 <= L->s_begin && L->s_begin < L->s_end && L->s_end <= L->T, IPN_ERR_ARG, "gru_layer_fwd: bad step range");
  const int T = L->T, H = L->H, dt = L->act_dt;
  const long long Bt = L->B_total;
  for (int d = 0; d < L->ndir; ++d) {
    const IpnGruDir& D = L->dir[d];
    IPN_REQUIRE(D.w_hh && D.b_hh && D.hseq, IPN_ERR_ARG, "gru_layer_fwd: null weight/state pointer (dir %d)", d);
    IPN_REQUIRE(D.P || D.table || D.pvec, IPN_ERR_ARG, "gru_layer_fwd: no input projection source (dir %d)", d);
    IPN_REQUIRE(!D.table || D.tok, IPN_ERR_ARG, "gru_layer_fwd: table without tokens");
  }
  if (L->ws != nullptr && gru_persist_fwd_shape_ok(L)) return gru_persist_fwd(L, L->ws, L->ws_bytes, stream);
  for (int d = 0; d < L->ndir; ++d)
    IPN_REQUIRE(!L->dir[d].P_blocked, IPN_ERR_ARG, "gru_layer_fwd: P_blocked needs the persistent kernel (workspace + eligible shape)");
  IPN_REQUIRE(!L->gates_blocked || (L->core == IPN_CORE_UMMA && Bt % 128 == 0 && H % 8 == 0), IPN_ERR_ARG,
              "gru_layer_fwd: gates_blocked needs the tcgen05 core and B_total %% 128 == 0");

  static const int dbg_epi = getenv("IPN_DBG_EPI") ? atoi(getenv("IPN_DBG_EPI")) : 0;
  static const int gru_br = getenv("IPN_GRU_BR") ? atoi(getenv("IPN_GRU_BR")) : 128;
  static const int stage_enable = getenv("IPN_STAGE") ? 2 : 0;   // IPN_STAGE=1 opts in to smem staging
  auto fill_epi = [&](GruFwdParams& e, const IpnGruDir& D, int s) {
    const int t = D.reverse ? T - 1 - s : s;
    const int in_slot = D.reverse ? t + 1 : t, out_slot = D.reverse ? t : t + 1;
    e.H = H; e.act_dt = dt; e.row0 = L->row0; e.trow = (long long)t * Bt;
    e.P = D.P; e.ldP = D.ldP; e.P_bcast = D.P_bcast; e.table = D.table; e.ld_table = D.ld_table; e.tok = D.tok; e.pvec = D.pvec;
    e.b_hh = D.b_hh;
    e.h_prev = slot_ptr(D.hseq, in_slot, Bt, H, dt);
    e.h_out = const_cast<char*>(slot_ptr(D.hseq, out_slot, Bt, H, dt));
    e.gates = D.gates;
    e.gates_blocked = L->gates_blocked;
    e.y = L->y; e.ld_y = L->ld_y; e.y_col0 = D.y_col0; e.mask = L->mask; e.ld_mask = L->ld_mask;
    e.mask_scale = L->mask_scale;
    const bool last = (s == T - 1);
    if (D.final_out_dir != nullptr) {
      e.final_out = last ? D.final_out_dir : nullptr;
      e.final_dt = D.final_dir_dt; e.ld_final = D.ld_final_dir;
    } else {
      e.final_out = last ? L->final_out : nullptr;
      e.final_dt = L->final_dt; e.ld_final = L->ld_final;
    }
    e.final_col0 = D.final_col0;
    e.dbg = dbg_epi;
    e.dbg_buf = g_dbg_timing;
    // shared-memory staging of the epilogue inputs needs 16-byte aligned row segments for every tile
    auto al16 = [](const void* ptr, long long stride_bytes) {
      return ptr == nullptr || (reinterpret_cast<uintptr_t>(ptr) % 16 == 0 && stride_bytes % 16 == 0);
    };
    e.stage = (stage_enable == 2 && L->core == IPN_CORE_UMMA && H % 16 == 0 && al16(D.P, D.ldP * 2) &&
               al16(D.table, D.ld_table * 4) && al16(D.hseq, H * 2) && (Bt * H * 2) % 16 == 0 &&
               al16(L->mask, L->ld_mask) && D.y_col0 % 16 == 0)
                  ? 1 : 0;
    return in_slot;
  };

  if (L->core == IPN_CORE_SIMT) {
    SimtBatch<EpiGruFwd> b;
    memset(&b, 0, sizeof(b));
    b.split_k = 1;
    for (int s = L->s_begin; s < L->s_end; ++s) {
      for (int d = 0; d < L->ndir; ++d) {
        const IpnGruDir& D = L->dir[d];
        SimtProblem<EpiGruFwd>& P = b.p[d];
        P.nseg = 1; P.M = L->nrows; P.N = H; P.gate_stride = H; P.in_dt = dt;
        const int in_slot = fill_epi(P.epi, D, s);
        HostOperand a{D.hseq, H, 0, (long long)(T + 1) * Bt, in_slot * Bt + L->row0, 0};
        HostOperand w{D.w_hh, H, 0, 3LL * H, 0, 0};
        fill_simt_seg(P.seg[0], a, w, H, dt);
      }
      IPN_PROPAGATE(launch_simt<EpiGruFwd>(b, L->ndir, L->nrows, H, stream, "gru_step_fwd_simt"));
    }
    return IPN_OK;
  }

  auto run_t = [&](auto cfg_tag, auto epi_tag) -> int {
    using Cfg = decltype(cfg_tag);
    // shared-memory staging of the epilogue inputs (EpiGruFwdT<.., true>) is implemented but measured slower
    // than direct loads in round 1 (256-byte bulk copies are TMA-issue bound; registers spill): opt-in only.
    using Epi = decltype(epi_tag);
    UmmaBatch<Epi> b;
    memset(&b, 0, sizeof(b));
    b.split_k = 1;
    for (int d = 0; d < L->ndir; ++d) {
      const IpnGruDir& D = L->dir[d];
      UmmaProblem<Epi>& P = b.p[d];
      P.nseg = 1; P.M = L->nrows; P.N = H; P.gate_stride = H;
      HostOperand a{D.hseq, H, 0, (long long)(T + 1) * Bt, 0, 0};
      HostOperand w{D.w_hh, H, 0, 3LL * H, 0, 0};
      IPN_PROPAGATE(fill_umma_seg(P.seg[0], a, w, H, Cfg::BR));
    }
    for (int s = L->s_begin; s < L->s_end; ++s) {
      for (int d = 0; d < L->ndir; ++d) {
        const int in_slot = fill_epi(b.p[d].epi, L->dir[d], s);
        b.p[d].seg[0].x_c1 = (int)(in_slot * Bt + L->row0);
      }
      IPN_PROPAGATE((launch_umma<Cfg, Epi>(b, L->ndir, L->nrows, H, stream, "gru_step_fwd_umma")));
    }
    return IPN_OK;
  };
  // 128 hidden units x BR rows x 3 gates per CTA
  auto run = [&](auto cfg_tag) -> int {
    if (stage_enable == 2) return run_t(cfg_tag, EpiGruFwdT<IPN_BF16, true>{});
    return run_t(cfg_tag, EpiGruFwdT<IPN_BF16, false>{});
  };
  if (gru_br == 64) return run(UmmaCfg<3, 64, false, false, 112>{});    // 192 TMEM columns, 2 CTAs/SM
  if (gru_br == 32) return run(UmmaCfg<3, 32, false, false, 112>{});    // 96 TMEM columns, 2 CTAs/SM
  if (gru_br == 1128) return run(UmmaCfg<3, 128, false, false, 200, 16>{});  // 16 epilogue warps
  return run(UmmaCfg<3, 128, false, false, 200>{});                     // 384 TMEM columns, 1 CTA/SM
}

extern "C" int ipn_gru_layer_bwd(const IpnGruLayerBwd* L, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  IPN_REQUIRE(L != nullptr, IPN_ERR_ARG, "gru_layer_bwd: null descriptor");
  IPN_PROPAGATE(ensure_device());
  IPN_PROPAGATE(check_layer_common(L->core, L->act_dt, L->T, L->B_total, L->H, L->row0, L->nrows, L->ndir));
  IPN_REQUIRE(L->dhz_ws != nullptr, IPN_ERR_ARG, "gru_layer_bwd: null workspace");
  const int T = L->T, H = L->H, dt = L->act_dt;
  const long long Bt = L->B_total;
  for (int d = 0; d < L->ndir; ++d) {
    const IpnGruBwdDir& D = L->dir[d];
    IPN_REQUIRE(D.w_hh && D.hseq && D.gates && D.dP && D.dGn, IPN_ERR_ARG, "gru_layer_bwd: null pointer (dir %d)", d);
  }
  IPN_REQUIRE(!L->gates_persist || (L->B_total % 128 == 0 && H % 8 == 0 && dt == IPN_BF16), IPN_ERR_ARG,
              "gru_layer_bwd: gates_persist set for a shape the persistent forward kernel cannot have produced");
  if (L->gates_persist && L->ws != nullptr && gru_persist_bwd_shape_ok(L)) return gru_persist_bwd(L, L->ws, L->ws_bytes, stream);
  auto dhz_buf = [&](int d, int which) { return L->dhz_ws + ((long long)d * 2 + which) * Bt * H; };

  // describes the pointwise differentiation of processing step s for direction d
  auto fill_point = [&](GruBwdPoint& p, int d, int s) {
    const IpnGruBwdDir& D = L->dir[d];
    const int t = D.reverse ? T - 1 - s : s;
    const int in_slot = D.reverse ? t + 1 : t;
    p.H = H; p.act_dt = dt; p.row0 = L->row0; p.trow = (long long)t * Bt;
    p.gates = D.gates;
    p.gates_blocked = L->gates_persist;
    p.h_prev = slot_ptr(D.hseq, in_slot, Bt, H, dt);
    p.dY = L->dY; p.ld_dy = L->ld_dy; p.y_col0 = D.y_col0; p.mask = L->mask; p.ld_mask = L->ld_mask;
    p.mask_scale = L->mask_scale;
    p.dh_n = (s == T - 1) ? D.dh_n : nullptr;
    p.ld_dhn = D.ld_dhn;
    p.dP = D.dP; p.dGn = D.dGn;
    p.dhz_out = dhz_buf(d, s & 1);
  };

  // start of the chain: pointwise for the last processed step
  {
    GruBwdPoint p0, p1;
    fill_point(p0, 0, T - 1);
    if (L->ndir > 1) fill_point(p1, 1, T - 1); else p1 = p0;
    constexpr int W = 4;
    const long long work = (long long)((L->nrows + W - 1) / W) * H;
    dim3 grid(cdiv(work, 256), 1, L->ndir);
    ProfScope prof("gru_bwd_pointwise", 0.0, 0.0, stream);
    gru_bwd_point_kernel<W><<<grid, 256, 0, stream>>>(p0, p1, L->nrows);
    IPN_LAUNCH_CHECK();
  }

  auto fill_epi = [&](GruBwdParams& e, int d, int s) {
    const IpnGruBwdDir& D = L->dir[d];
    e.dhz_in = dhz_buf(d, s & 1);
    e.is_first_step = (s == 0);
    e.dh0 = D.dh0; e.dh0_dt = D.dh0_dt; e.ld_dh0 = D.ld_dh0; e.dh0_selu = D.dh0_selu;
    e.h0 = slot_ptr(D.hseq, D.reverse ? T : 0, Bt, H, dt);
    if (s > 0) fill_point(e.pw, d, s - 1);
    else { memset(&e.pw, 0, sizeof(e.pw)); e.pw.H = H; e.pw.act_dt = dt; e.pw.row0 = L->row0; }
  };

  if (L->core == IPN_CORE_SIMT) {
    SimtBatch<EpiGruBwd> b;
    memset(&b, 0, sizeof(b));
    b.split_k = 1;
    for (int s = T - 1; s >= 0; --s) {
      for (int d = 0; d < L->ndir; ++d) {
        const IpnGruBwdDir& D = L->dir[d];
        const int t = D.reverse ? T - 1 - s : s;
        SimtProblem<EpiGruBwd>& P = b.p[d];
        P.nseg = 2; P.M = L->nrows; P.N = H; P.gate_stride = 0; P.in_dt = dt;
        HostOperand a0{D.dP, 3LL * H, 0, (long long)T * Bt, t * Bt + L->row0, 0};
        HostOperand w0{D.w_hh, H, 1, H, 0, 0};
        fill_simt_seg(P.seg[0], a0, w0, 2 * H, dt);
        HostOperand a1{D.dGn, H, 0, (long long)T * Bt, t * Bt + L->row0, 0};
        HostOperand w1{D.w_hh, H, 1, H, 0, 2LL * H};
        fill_simt_seg(P.seg[1], a1, w1, H, dt);
        fill_epi(P.epi, d, s);
      }
      IPN_PROPAGATE(launch_simt<EpiGruBwd>(b, L->ndir, L->nrows, H, stream, "gru_step_bwd_simt"));
    }
    return IPN_OK;
  }

  static const int bwd_cfg = getenv("IPN_GRU_BWD_CFG") ? atoi(getenv("IPN_GRU_BWD_CFG")) : 0;
  auto run = [&](auto cfg_tag) -> int {
    using Cfg = decltype(cfg_tag);   // W_hh read MN-major (columns = hidden units on the TMEM lanes)
    using EpiB = EpiGruBwdT<IPN_BF16>;
    UmmaBatch<EpiB> b;
    memset(&b, 0, sizeof(b));
    b.split_k = 1;
    for (int d = 0; d < L->ndir; ++d) {
      const IpnGruBwdDir& D = L->dir[d];
      UmmaProblem<EpiB>& P = b.p[d];
      P.nseg = 2; P.M = L->nrows; P.N = H; P.gate_stride = 0;
      HostOperand a0{D.dP, 3LL * H, 0, (long long)T * Bt, 0, 0};
      HostOperand w0{D.w_hh, H, 1, H, 0, 0};
      IPN_PROPAGATE(fill_umma_seg(P.seg[0], a0, w0, 2 * H, Cfg::BR));
      HostOperand a1{D.dGn, H, 0, (long long)T * Bt, 0, 0};
      HostOperand w1{D.w_hh, H, 1, H, 0, 2LL * H};
      IPN_PROPAGATE(fill_umma_seg(P.seg[1], a1, w1, H, Cfg::BR));
    }
    for (int s = T - 1; s >= 0; --s) {
      for (int d = 0; d < L->ndir; ++d) {
        const int t = L->dir[d].reverse ? T - 1 - s : s;
        b.p[d].seg[0].x_c1 = (int)(t * Bt + L->row0);
        b.p[d].seg[1].x_c1 = (int)(t * Bt + L->row0);
        fill_epi(b.p[d].epi, d, s);
      }
      IPN_PROPAGATE((launch_umma<Cfg, EpiB>(b, L->ndir, L->nrows, H, stream, "gru_step_bwd_umma")));
    }
    return IPN_OK;
  };
  if (bwd_cfg == 0) return run(UmmaCfg<1, 128, true, false>{});            // 2 CTAs/SM (96-register cap)
  if (bwd_cfg == 2) return run(UmmaCfg<1, 256, true, false, 200>{});       // 256-row tiles, 1 CTA/SM
  if (bwd_cfg == 3) return run(UmmaCfg<1, 128, true, false, 200, 16>{});   // 16 epilogue warps
  return run(UmmaCfg<1, 128, true, false, 200>{});                         // 128-row tiles, 1 CTA/SM, no spills
}
