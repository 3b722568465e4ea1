// LSTM layer forward / backward drivers (AnticipationRNN baseline: uni-directional, zero initial state).
#include "launch.cuh"

namespace ipn {

template <int W>
__global__ void lstm_bwd_point_kernel(LstmBwdPoint p, int nrows) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int rchunks = (nrows + W - 1) / W;
  if (idx >= (long long)rchunks * p.H) return;
  const int col = (int)(idx % p.H);
  const int row0 = (int)(idx / p.H) * W;
  float dh[W];
#pragma unroll
  for (int i = 0; i < W; ++i) dh[i] = 0.f;
  lstm_bwd_pointwise<W>(p, col, row0, min(W, nrows - row0), dh);
}

int lstm_persist_fwd(const IpnLstmLayer* L, int s_begin, int s_end, cudaStream_t stream);   // lstm_persist.cu
int lstm_persist_bwd(const IpnLstmLayerBwd* L, cudaStream_t stream);

static int check_lstm(int core, int act_dt, int T, int B, int H) {
  IPN_REQUIRE(core == IPN_CORE_SIMT || core == IPN_CORE_UMMA, IPN_ERR_ARG, "lstm: unknown core %d", core);
  IPN_REQUIRE(core != IPN_CORE_UMMA || act_dt == IPN_BF16, IPN_ERR_ARG, "lstm: the tcgen05 core needs bf16 activations");
  IPN_REQUIRE(T > 0 && B > 0 && H > 0, IPN_ERR_ARG, "lstm: bad sizes");
  IPN_REQUIRE(core != IPN_CORE_UMMA || H % 8 == 0, IPN_ERR_ALIGN, "lstm: tcgen05 core needs H %% 8 == 0");
  return IPN_OK;
}

}  // namespace ipn

using namespace ipn;

extern "C" int ipn_lstm_layer_fwd(const IpnLstmLayer* L, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  IPN_REQUIRE(L != nullptr, IPN_ERR_ARG, "lstm_layer_fwd: null descriptor");
  IPN_PROPAGATE(ensure_device());
  IPN_PROPAGATE(check_lstm(L->core, L->act_dt, L->T, L->B, L->H));
  IPN_REQUIRE(L->w_hh && L->P && L->hseq && L->cseq, IPN_ERR_ARG, "lstm_layer_fwd: null pointer");
  const int T = L->T, H = L->H, dt = L->act_dt;
  const long long B = L->B;
  const long long es = dt == IPN_BF16 ? 2 : 4;
  const int s_begin = L->s_begin, s_end = L->s_end > 0 ? L->s_end : T;
  IPN_REQUIRE(0 <= s_begin && s_begin < s_end && s_end <= T, IPN_ERR_ARG, "lstm_layer_fwd: bad step range");
  IPN_REQUIRE(!L->table || L->tok_scalar, IPN_ERR_ARG, "lstm_layer_fwd: table without token");
  if (L->P_blocked) return lstm_persist_fwd(L, s_begin, s_end, stream);
  IPN_REQUIRE(!L->gates_blocked || (ipn_lstm_persist_eligible(L->core, L->act_dt, L->B, L->H) && L->gates != nullptr), IPN_ERR_ARG,
              "lstm_layer_fwd: gates_blocked needs a shape the persistent kernels accept (bf16, H 128/256, B %% 128 == 0)");
  auto fill_epi = [&](EpiLstmFwd::Params& e, int s) {
    e.H = H; e.act_dt = dt; e.trow = s * B;
    e.P = L->P; e.ldP = L->ldP; e.b_hh = L->b_hh;
    e.c_prev = L->cseq + s * B * H;
    e.c_out = L->cseq + (s + 1) * B * H;
    e.h_out = reinterpret_cast<char*>(L->hseq) + (s + 1) * B * H * es;
    e.gates = L->gates; e.y = L->y; e.ld_y = L->ld_y; e.y_col0 = L->y_col0;
    e.ytrow = (L->y_reverse_time ? (T - 1 - s) : s) * B;
    e.table = L->table; e.ld_table = L->ld_table; e.tok_scalar = L->tok_scalar;
    e.gates_blocked = L->gates_blocked;
  };
  if (L->core == IPN_CORE_SIMT) {
    SimtBatch<EpiLstmFwd> b;
    memset(&b, 0, sizeof(b));
    b.split_k = 1;
    for (int s = s_begin; s < s_end; ++s) {
      SimtProblem<EpiLstmFwd>& P = b.p[0];
      P.nseg = 1; P.M = (int)B; P.N = H; P.gate_stride = H; P.in_dt = dt;
      HostOperand a{L->hseq, H, 0, (T + 1) * B, s * B, 0};
      HostOperand w{L->w_hh, H, 0, 4LL * H, 0, 0};
      fill_simt_seg(P.seg[0], a, w, H, dt);
      fill_epi(P.epi, s);
      IPN_PROPAGATE(launch_simt<EpiLstmFwd>(b, 1, (int)B, H, stream, "lstm_step_fwd_simt"));
    }
    return IPN_OK;
  }
  using Cfg = UmmaCfg<4, 64, false, false, 200>;   // 128 units x 64 rows x 4 gates: 256 TMEM columns
  UmmaBatch<EpiLstmFwd> b;
  memset(&b, 0, sizeof(b));
  b.split_k = 1;
  UmmaProblem<EpiLstmFwd>& P = b.p[0];
  P.nseg = 1; P.M = (int)B; P.N = H; P.gate_stride = H;
  HostOperand a{L->hseq, H, 0, (T + 1) * B, 0, 0};
  HostOperand w{L->w_hh, H, 0, 4LL * H, 0, 0};
  IPN_PROPAGATE(fill_umma_seg(P.seg[0], a, w, H, Cfg::BR));
  for (int s = s_begin; s < s_end; ++s) {
    P.seg[0].x_c1 = (int)(s * B);
    fill_epi(P.epi, s);
    IPN_PROPAGATE((launch_umma<Cfg, EpiLstmFwd>(b, 1, (int)B, H, stream, "lstm_step_fwd_umma")));
  }
  return IPN_OK;
}

extern "C" int ipn_lstm_layer_bwd(const IpnLstmLayerBwd* L, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  IPN_REQUIRE(L != nullptr, IPN_ERR_ARG, "lstm_layer_bwd: null descriptor");
  IPN_PROPAGATE(ensure_device());
  IPN_PROPAGATE(check_lstm(L->core, L->act_dt, L->T, L->B, L->H));
  if (L->gates_persist) {
    IPN_REQUIRE(L->w_hh && L->gates && L->dP, IPN_ERR_ARG, "lstm_layer_bwd: null pointer");
    return lstm_persist_bwd(L, stream);
  }
  IPN_REQUIRE(L->w_hh && L->hseq && L->cseq && L->gates && L->dP && L->ws, IPN_ERR_ARG, "lstm_layer_bwd: null pointer");
  const int T = L->T, H = L->H, dt = L->act_dt;
  const long long B = L->B;
  auto fill_point = [&](LstmBwdPoint& p, int s) {
    p.H = H; p.act_dt = dt; p.trow = s * B;
    p.gates = L->gates;
    p.c_prev = L->cseq + s * B * H;
    p.c_cur = L->cseq + (s + 1) * B * H;
    p.dY = L->dY; p.ld_dy = L->ld_dy; p.y_col0 = L->y_col0;
    p.dytrow = (L->y_reverse_time ? (T - 1 - s) : s) * B;
    p.dc_in = (s == T - 1) ? nullptr : L->ws + ((s + 1) & 1) * B * H;
    p.dc_out = L->ws + (s & 1) * B * H;
    p.dP = L->dP;
  };
  {
    LstmBwdPoint p;
    fill_point(p, T - 1);
    constexpr int W = 4;
    const long long work = ((B + W - 1) / W) * H;
    lstm_bwd_point_kernel<W><<<cdiv(work, 256), 256, 0, stream>>>(p, (int)B);
    IPN_LAUNCH_CHECK();
  }
  if (L->core == IPN_CORE_SIMT) {
    SimtBatch<EpiLstmBwd> b;
    memset(&b, 0, sizeof(b));
    b.split_k = 1;
    for (int s = T - 1; s >= 1; --s) {
      SimtProblem<EpiLstmBwd>& P = b.p[0];
      P.nseg = 1; P.M = (int)B; P.N = H; P.gate_stride = 0; P.in_dt = dt;
      HostOperand a{L->dP, 4LL * H, 0, T * B, s * B, 0};
      HostOperand w{L->w_hh, H, 1, H, 0, 0};
      fill_simt_seg(P.seg[0], a, w, 4 * H, dt);
      P.epi.is_first_step = 0;
      fill_point(P.epi.pw, s - 1);
      IPN_PROPAGATE(launch_simt<EpiLstmBwd>(b, 1, (int)B, H, stream, "lstm_step_bwd_simt"));
    }
    return IPN_OK;
  }
  using Cfg = UmmaCfg<1, 128, true, false>;
  UmmaBatch<EpiLstmBwd> b;
  memset(&b, 0, sizeof(b));
  b.split_k = 1;
  UmmaProblem<EpiLstmBwd>& P = b.p[0];
  P.nseg = 1; P.M = (int)B; P.N = H; P.gate_stride = 0;
  HostOperand a{L->dP, 4LL * H, 0, T * B, 0, 0};
  HostOperand w{L->w_hh, H, 1, H, 0, 0};
  IPN_PROPAGATE(fill_umma_seg(P.seg[0], a, w, 4 * H, Cfg::BR));
  for (int s = T - 1; s >= 1; --s) {
    P.seg[0].x_c1 = (int)(s * B);
    P.epi.is_first_step = 0;
    fill_point(P.epi.pw, s - 1);
    IPN_PROPAGATE((launch_umma<Cfg, EpiLstmBwd>(b, 1, (int)B, H, stream, "lstm_step_bwd_umma")));
  }
  return IPN_OK;
}
