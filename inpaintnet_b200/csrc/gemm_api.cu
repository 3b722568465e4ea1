// ipn_gemm: universal GEMM + fused epilogue (see include/inpaintnet_b200.h).
#include "launch.cuh"
#include <stdlib.h>

namespace ipn {

static int fill_linear_epi(EpiLinear::Params& e, const IpnGemm* g) {
  e.out = g->out;
  e.out_dt = g->out_dt;
  e.ld_out = g->ld_out;
  e.use_rowmap = g->use_rowmap;
  e.rowmap = g->rowmap;
  e.split_cols = g->split_cols;
  e.split_stride = g->split_stride;
  e.bias = g->bias;
  e.act = g->act;
  e.alpha = g->alpha;
  e.mul_src = g->mul_src;
  e.mul_dt = g->mul_dt;
  e.ld_mul = g->ld_mul;
  e.mul_mode = g->mul_mode;
  e.mul_scale = g->mul_scale;
  e.accumulate = g->accumulate;
  return IPN_OK;
}

// API operands: A = activations [M,K] (output rows), B = weights [N,K] (output columns).
template <int BR, bool TW, bool TX>
static int gemm_umma(const IpnGemm* g, int split_k, cudaStream_t stream) {
  static const int persist = getenv("IPN_GEMM_PERSIST") ? atoi(getenv("IPN_GEMM_PERSIST")) : 1;
  using Cfg = UmmaCfg<1, BR, TW, TX>;
  UmmaBatch<EpiLinear> b;
  memset(&b, 0, sizeof(b));
  b.split_k = split_k;
  UmmaProblem<EpiLinear>& P = b.p[0];
  P.nseg = g->nseg;
  P.M = g->M;
  P.N = g->N;
  P.gate_stride = 0;
  // CTA pairs (256-column x BR-row tiles, X tile shared between the two SMs) when there is more than one column tile
  const bool pair = persist == 1 && BR >= 128 && g->N > UMMA_BC;
  for (int s = 0; s < g->nseg; ++s) {
    const IpnGemmSeg& sg = g->seg[s];
    HostOperand x{sg.A, sg.lda, sg.transA, g->M, 0, 0};
    HostOperand w{sg.B, sg.ldb, sg.transB, g->N, 0, 0};
    IPN_PROPAGATE(fill_umma_seg(P.seg[s], x, w, sg.K, pair ? BR / 2 : BR));   // a pair's CTA stages half of the X rows
  }
  fill_linear_epi(P.epi, g);
  const char* tag = TX ? "gemm_umma_tn_wgrad" : (TW ? "gemm_umma_nn_dgrad" : "gemm_umma_nt");
  if constexpr (BR >= 128) {
    if (pair)
      return launch_umma_persist<UmmaPCfg<BR, TW, TX, true>, EpiLinear>(b, 1, g->M, g->N, stream, tag, 0, g->max_ctas);
  }
  if (persist) return launch_umma_persist<UmmaPCfg<BR, TW, TX, false>, EpiLinear>(b, 1, g->M, g->N, stream, tag, 0, g->max_ctas);
  return launch_umma<Cfg, EpiLinear>(b, 1, g->M, g->N, stream, tag);
}

static int pick_split_k(const IpnGemm* g, int tile_m, int tile_n, int bk) {
  if (g->split_k > 0) return g->split_k;
  if (g->accumulate != IPN_ATOMIC_ADD) return 1;
  long long tiles = (long long)cdiv(g->M, tile_m) * cdiv(g->N, tile_n);
  long long chunks = 0;
  for (int s = 0; s < g->nseg; ++s) chunks += cdiv(g->seg[s].K, bk);
  int split = 1;
  while (tiles * split < 2 * 148 && chunks / (split * 2) >= 8 && split < 64) split *= 2;
  return split;
}

}  // namespace ipn

using namespace ipn;

extern "C" int ipn_gemm(const IpnGemm* g, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  IPN_REQUIRE(g != nullptr, IPN_ERR_ARG, "ipn_gemm: null descriptor");
  IPN_PROPAGATE(ensure_device());
  IPN_REQUIRE(g->M > 0 && g->N > 0, IPN_ERR_ARG, "ipn_gemm: empty problem M=%d N=%d", g->M, g->N);
  IPN_REQUIRE(g->nseg == 1 || g->nseg == 2, IPN_ERR_ARG, "ipn_gemm: nseg must be 1 or 2");
  IPN_REQUIRE(g->out != nullptr, IPN_ERR_ARG, "ipn_gemm: null output");
  for (int s = 0; s < g->nseg; ++s)
    IPN_REQUIRE(g->seg[s].A && g->seg[s].B && g->seg[s].K > 0, IPN_ERR_ARG, "ipn_gemm: bad segment %d", s);
  IPN_REQUIRE(g->accumulate != IPN_ATOMIC_ADD || g->out_dt == IPN_F32, IPN_ERR_ARG,
              "ipn_gemm: atomic accumulation needs an fp32 output");
  IPN_REQUIRE(g->split_k <= 1 || g->accumulate == IPN_ATOMIC_ADD, IPN_ERR_ARG, "ipn_gemm: split_k needs atomic accumulate");
  IPN_REQUIRE(g->split_cols == 0 || g->split_cols % 16 == 0, IPN_ERR_ARG, "ipn_gemm: split_cols must be a multiple of 16");
  IPN_REQUIRE(!(g->split_k > 1 && (g->bias || g->act != IPN_ACT_NONE)), IPN_ERR_ARG,
              "ipn_gemm: bias/activation cannot be combined with split_k");

  if (g->core == IPN_CORE_SIMT) {
    SimtBatch<EpiLinear> b;
    memset(&b, 0, sizeof(b));
    b.split_k = pick_split_k(g, SIMT_BM, SIMT_BN, SIMT_BK);
    if (b.split_k > 1) IPN_REQUIRE(!g->bias && g->act == IPN_ACT_NONE, IPN_ERR_ARG, "split_k with bias/act");
    SimtProblem<EpiLinear>& P = b.p[0];
    P.nseg = g->nseg;
    P.M = g->M;
    P.N = g->N;
    P.gate_stride = 0;
    P.in_dt = g->in_dt;
    for (int s = 0; s < g->nseg; ++s) {
      const IpnGemmSeg& sg = g->seg[s];
      HostOperand a{sg.A, sg.lda, sg.transA, g->M, 0, 0};
      HostOperand bb{sg.B, sg.ldb, sg.transB, g->N, 0, 0};
      fill_simt_seg(P.seg[s], a, bb, sg.K, g->in_dt);
    }
    fill_linear_epi(P.epi, g);
    return launch_simt<EpiLinear>(b, 1, g->M, g->N, stream, "gemm_simt");
  }

  IPN_REQUIRE(g->core == IPN_CORE_UMMA, IPN_ERR_ARG, "ipn_gemm: unknown core %d", g->core);
  IPN_REQUIRE(g->in_dt == IPN_BF16, IPN_ERR_ARG, "ipn_gemm: the tcgen05 core takes bf16 operands");
  const int tA = g->seg[0].transA, tB = g->seg[0].transB;
  for (int s = 1; s < g->nseg; ++s)
    IPN_REQUIRE(g->seg[s].transA == tA && g->seg[s].transB == tB, IPN_ERR_ARG, "ipn_gemm: segments must share layouts");
  // rows per tile: 256 for tall problems, 128 otherwise, 64 for short ones (more CTAs)
  const int br = (tA ? (g->M >= 1024 ? 256 : 128) : (g->M >= 4096 ? 256 : (g->M > 64 ? 128 : 64)));
  const int split = pick_split_k(g, br, UMMA_BC, UMMA_BK);
  if (split > 1) IPN_REQUIRE(!g->bias && g->act == IPN_ACT_NONE, IPN_ERR_ARG, "split_k with bias/act");
  if (!tA && !tB) {
    if (br == 64) return gemm_umma<64, false, false>(g, split, stream);
    if (br == 256) return gemm_umma<256, false, false>(g, split, stream);
    return gemm_umma<128, false, false>(g, split, stream);
  }
  if (!tA && tB) {
    if (br == 256) return gemm_umma<256, true, false>(g, split, stream);
    return gemm_umma<128, true, false>(g, split, stream);
  }
  if (tA && tB) {
    if (br == 256) return gemm_umma<256, true, true>(g, split, stream);
    return gemm_umma<128, true, true>(g, split, stream);
  }
  ipn::set_error("ipn_gemm: layout transA=1,transB=0 is not provided by the tcgen05 core");
  return IPN_ERR_ARG;
}

// Blocked GRU input projection: see EpiBlockedP (epilogues.cuh) and include/inpaintnet_b200.h
extern "C" int ipn_gru_inproj_blocked(const IpnGruInproj* q, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  IPN_REQUIRE(q != nullptr && q->X && q->w_ih && q->b_ih && q->b_hh && q->out, IPN_ERR_ARG, "gru_inproj_blocked: null pointer");
  IPN_PROPAGATE(ensure_device());
  IPN_REQUIRE(q->rows > 0 && q->rows % 128 == 0 && q->H % 8 == 0 && q->H > 0 && q->K > 0, IPN_ERR_ARG,
              "gru_inproj_blocked: rows must be a multiple of 128 and H of 8 (rows=%lld H=%d)", q->rows, q->H);
  IPN_REQUIRE(q->rows < (1LL << 31), IPN_ERR_ARG, "gru_inproj_blocked: too many rows");
  constexpr int BR = 256;
  const int N3 = 3 * q->H;
  const bool pair = (q->rows / 128) % 2 == 0;
  UmmaBatch<EpiBlockedP> b;
  memset(&b, 0, sizeof(b));
  b.split_k = 1;
  UmmaProblem<EpiBlockedP>& P = b.p[0];
  P.nseg = 1;
  P.M = N3;                 // "rows" of the tile = gate units (X side, MMA N)
  P.N = (int)q->rows;       // "columns" = activation rows (W side, TMEM lanes)
  P.gate_stride = 0;
  HostOperand x{q->w_ih, q->ldw, 0, N3, 0, 0};
  HostOperand w{q->X, q->ldx, 0, q->rows, 0, 0};
  IPN_PROPAGATE(fill_umma_seg(P.seg[0], x, w, q->K, pair ? BR / 2 : BR));
  P.epi.out = reinterpret_cast<uint4*>(q->out);
  P.epi.b_ih = q->b_ih;
  P.epi.b_hh = q->b_hh;
  P.epi.H = q->H;
  P.epi.ngates = 3;
  P.epi.half_mask = 3;   // r, z
  P.epi.bhh_mask = 3;
  if (pair) return launch_umma_persist<UmmaPCfg<BR, false, false, true>, EpiBlockedP>(b, 1, N3, (int)q->rows, stream, "gemm_umma_inproj_blocked", 1);
  return launch_umma_persist<UmmaPCfg<BR, false, false, false>, EpiBlockedP>(b, 1, N3, (int)q->rows, stream, "gemm_umma_inproj_blocked", 1);
}

// Blocked LSTM input projection (4 gates, up to two input segments): see include/inpaintnet_b200.h
extern "C" int ipn_lstm_inproj_blocked(const IpnLstmInproj* q, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  IPN_REQUIRE(q != nullptr && q->X && q->w_ih && q->b_ih && q->b_hh && q->out, IPN_ERR_ARG, "lstm_inproj_blocked: null pointer");
  IPN_PROPAGATE(ensure_device());
  IPN_REQUIRE(q->rows > 0 && q->rows % 128 == 0 && q->H % 8 == 0 && q->H > 0 && q->K > 0, IPN_ERR_ARG,
              "lstm_inproj_blocked: rows must be a multiple of 128 and H of 8 (rows=%lld H=%d)", q->rows, q->H);
  IPN_REQUIRE(q->rows < (1LL << 31), IPN_ERR_ARG, "lstm_inproj_blocked: too many rows");
  IPN_REQUIRE(q->X2 == nullptr || (q->w_ih2 != nullptr && q->K2 > 0), IPN_ERR_ARG, "lstm_inproj_blocked: bad second segment");
  constexpr int BR = 256;
  const int N4 = 4 * q->H;
  const bool pair = (q->rows / 128) % 2 == 0;
  UmmaBatch<EpiBlockedP> b;
  memset(&b, 0, sizeof(b));
  b.split_k = 1;
  UmmaProblem<EpiBlockedP>& P = b.p[0];
  P.nseg = q->X2 != nullptr ? 2 : 1;
  P.M = N4;
  P.N = (int)q->rows;
  P.gate_stride = 0;
  {
    HostOperand x{q->w_ih, q->ldw, 0, N4, 0, 0};
    HostOperand w{q->X, q->ldx, 0, q->rows, 0, 0};
    IPN_PROPAGATE(fill_umma_seg(P.seg[0], x, w, q->K, pair ? BR / 2 : BR));
  }
  if (q->X2 != nullptr) {
    HostOperand x{q->w_ih2, q->ldw2, 0, N4, 0, 0};
    HostOperand w{q->X2, q->ldx2, 0, q->rows, 0, 0};
    IPN_PROPAGATE(fill_umma_seg(P.seg[1], x, w, q->K2, pair ? BR / 2 : BR));
  }
  P.epi.out = reinterpret_cast<uint4*>(q->out);
  P.epi.b_ih = q->b_ih;
  P.epi.b_hh = q->b_hh;
  P.epi.H = q->H;
  P.epi.ngates = 4;
  P.epi.half_mask = 1 | 2 | 8;   // i, f, o
  P.epi.bhh_mask = 15;
  if (pair) return launch_umma_persist<UmmaPCfg<BR, false, false, true>, EpiBlockedP>(b, 1, N4, (int)q->rows, stream, "gemm_umma_inproj_blocked", 1);
  return launch_umma_persist<UmmaPCfg<BR, false, false, false>, EpiBlockedP>(b, 1, N4, (int)q->rows, stream, "gemm_umma_inproj_blocked", 1);
}
