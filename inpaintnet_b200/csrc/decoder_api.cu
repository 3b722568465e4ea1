// Tick decoder, argmax-feedback mode: host loop over the 24 serial ticks (MeasureVAE/decoder.py:473-529).
#include "runtime.h"

using namespace ipn;

namespace ipn {
bool gru_persist_fwd_shape_ok(const IpnGruLayer* L);
long long gru_persist_fwd_ws_bytes(const IpnGruLayer* L);
// tick_persist.cu: the whole decode as one persistent kernel
bool tick_persist_shape_ok(const IpnTickDecode* p);
int tick_persist_decode(const IpnTickDecode* p, void* ws, long long ws_bytes, cudaStream_t stream);
}

// the two per-tick GRU layer descriptors (row window = beat i of the batch, one step)
static void tick_layers(const IpnTickDecode* p, IpnGruLayer& L0, IpnGruLayer& L1) {
  const int B = p->B, H = p->H;
  memset(&L0, 0, sizeof(L0));
  L0.core = p->core; L0.act_dt = p->act_dt; L0.T = 6; L0.B_total = 4 * B; L0.H = H; L0.ndir = 1;
  L0.dir[0] = p->l0;
  L0.dir[0].tok = p->tokprev;
  L0.gates_blocked = p->gates_blocked;
  L0.y = p->yt0; L0.ld_y = H; L0.mask = p->mask; L0.ld_mask = H; L0.mask_scale = p->mask_scale;
  L0.row0 = 0; L0.nrows = B; L0.s_begin = 0; L0.s_end = 1;
  L1 = L0;
  L1.dir[0] = p->l1;
  L1.dir[0].P = p->Pt1; L1.dir[0].ldP = 3LL * H; L1.dir[0].P_bcast = 0;
  L1.y = p->yt1; L1.mask = nullptr;
}

static bool tick_persist_ok(const IpnTickDecode* p) {
  if (p->B <= 0 || p->H <= 0 || p->B % 128 != 0) return false;
  if (p->l0.gates != nullptr && !p->gates_blocked) return false;   // the persistent kernel only writes the blocked layout
  IpnGruLayer L0, L1;
  tick_layers(p, L0, L1);
  L1.dir[0].P_blocked = 1;
  return gru_persist_fwd_shape_ok(&L0) && gru_persist_fwd_shape_ok(&L1);
}

extern "C" long long ipn_tick_decode_ws_bytes(const IpnTickDecode* p) {
  if (p == nullptr) return 0;
  if (!tick_persist_ok(p)) return tick_persist_shape_ok(p) ? 128LL * 3 * p->H * 2 : 0;
  IpnGruLayer L0, L1;   // both per-tick layer calls share the workspace (stream ordered): the larger requirement
  tick_layers(p, L0, L1);
  L1.dir[0].P_blocked = 1;
  const long long a = gru_persist_fwd_ws_bytes(&L0), b = gru_persist_fwd_ws_bytes(&L1);
  return a > b ? a : b;
}

extern "C" int ipn_tick_decode_argmax(const IpnTickDecode* p, void* stream) {
  IPN_REQUIRE(p != nullptr, IPN_ERR_ARG, "tick_decode: null descriptor");
  IPN_PROPAGATE(ensure_device());
  IPN_REQUIRE(p->B > 0 && p->H > 0 && p->V > 0, IPN_ERR_ARG, "tick_decode: bad sizes");
  IPN_REQUIRE(p->yt0 && p->yt1 && p->w_ih1 && p->b_ih1 && p->Pt1 && p->w_v && p->b_v && p->weights && p->tokprev,
              IPN_ERR_ARG, "tick_decode: null pointer");
  const int B = p->B, H = p->H, V = p->V, dt = p->act_dt;
  const long long B4 = 4LL * B;
  const long long es = dt == IPN_BF16 ? 2 : 4;

  // one launch for all 24 ticks where the 4-CTA-per-tile grid is co-resident (batches up to 4224 measures)
  if (p->ws != nullptr && tick_persist_shape_ok(p) && p->ws_bytes >= 128LL * 3 * H * 2)
    return tick_persist_decode(p, p->ws, p->ws_bytes, reinterpret_cast<cudaStream_t>(stream));

  IpnGruLayer L0, L1;
  tick_layers(p, L0, L1);
  const bool persist = p->ws != nullptr && tick_persist_ok(p) && p->ws_bytes >= ipn_tick_decode_ws_bytes(p);
  IpnGruInproj ip;   // persistent path: layer-1 input projection written in the layer kernel's blocked layout
  memset(&ip, 0, sizeof(ip));
  if (persist) {
    L0.ws = p->ws; L0.ws_bytes = p->ws_bytes;
    L1.ws = p->ws; L1.ws_bytes = p->ws_bytes;
    L1.dir[0].P_blocked = 1;
    ip.ldx = H; ip.rows = B; ip.K = H; ip.w_ih = p->w_ih1; ip.ldw = H; ip.b_ih = p->b_ih1; ip.b_hh = p->l1.b_hh; ip.H = H;
  }

  IpnGemm gp;  // layer-1 input projection of one tick
  memset(&gp, 0, sizeof(gp));
  gp.core = p->core; gp.in_dt = dt; gp.M = B; gp.N = 3 * H; gp.nseg = 1;
  gp.seg[0].lda = H; gp.seg[0].B = p->w_ih1; gp.seg[0].ldb = H; gp.seg[0].K = H;
  gp.out_dt = dt; gp.ld_out = 3LL * H; gp.bias = p->b_ih1; gp.act = IPN_ACT_NONE; gp.alpha = 1.f;
  IpnGemm gv = gp;  // vocabulary projection + ReLU
  gv.N = V; gv.seg[0].B = p->w_v; gv.bias = p->b_v; gv.act = IPN_ACT_RELU;
  gv.out_dt = IPN_F32; gv.ld_out = 24LL * V;

  IpnRowMap wmap{1 << 30, 1 << 30, 0, 0, 24LL * V};  // row b -> b*24V
  IpnRowMap smap{1 << 30, 1 << 30, 0, 0, 24};        // row b -> b*24
  if (p->use_maps) { wmap = p->wmap; smap = p->smap; }
  gv.use_rowmap = 1;
  gv.rowmap = wmap;

  for (int t = 0; t < 24; ++t) {
    const int i = t / 6, j = t % 6;
    const long long r0 = (long long)j * B4 + (long long)i * B;  // first row of this tick in time-ordered buffers
    L0.row0 = i * B; L0.nrows = B; L0.s_begin = j; L0.s_end = j + 1;
    IPN_PROPAGATE(ipn_gru_layer_fwd(&L0, stream));
    if (persist) {
      ip.X = reinterpret_cast<const char*>(p->yt0) + r0 * H * es;
      ip.out = reinterpret_cast<char*>(p->Pt1) + r0 * 3 * H * es;   // r0 % 128 == 0: tile-aligned in the blocked layout too
      IPN_PROPAGATE(ipn_gru_inproj_blocked(&ip, stream));
    } else {
      gp.seg[0].A = reinterpret_cast<const char*>(p->yt0) + r0 * H * es;
      gp.out = reinterpret_cast<char*>(p->Pt1) + r0 * 3 * H * es;
      IPN_PROPAGATE(ipn_gemm(&gp, stream));
    }
    L1.row0 = i * B; L1.nrows = B; L1.s_begin = j; L1.s_end = j + 1;
    IPN_PROPAGATE(ipn_gru_layer_fwd(&L1, stream));
    gv.seg[0].A = reinterpret_cast<const char*>(p->yt1) + r0 * H * es;
    gv.out = p->weights + (long long)t * V;
    IPN_PROPAGATE(ipn_gemm(&gv, stream));
    int* tok_next = nullptr;
    if (t < 23) {
      const int i2 = (t + 1) / 6, j2 = (t + 1) % 6;
      tok_next = p->tokprev + (long long)j2 * B4 + (long long)i2 * B;
    }
    IPN_PROPAGATE(ipn_argmax_rows(p->weights + (long long)t * V, B, V, &wmap, tok_next,
                                  p->samples ? p->samples + t : nullptr, &smap, stream));
  }
  return IPN_OK;
}
