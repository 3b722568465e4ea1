// Epilogue functors shared by the tcgen05 core and the fp32 SIMT core.
//
// "Lanes = output columns": both cores hand a functor ONE output column `col` and W consecutive output
// rows row0..row0+W-1 (acc[g][i] = accumulator of gate g at (row0+i, col)).  In the tcgen05 core a TMEM
// lane is an output column, so the 32 lanes of a warp touch 32 CONSECUTIVE columns of the same row:
// every global access of the epilogue is a coalesced 64-byte (bf16) or 128-byte (fp32) segment, with no
// shared-memory staging, for any number of input/output arrays.  Per-column constants (biases, constant
// input projection) are loaded once per thread in col_init().
//
// Each functor first issues all loads of its W rows (memory-level parallelism), then computes and stores.
#pragma once
#include "common.cuh"

namespace ipn {

// =============================================================================================
// Linear: out = alpha * act(acc + bias) * mul     (+ row map, column split, accumulate modes)
// =============================================================================================
struct EpiLinear {
  template <class PP>
  static __device__ __forceinline__ unsigned long long* dbg_buf(const PP&) { return nullptr; }
  static constexpr int G = 1;
  struct Params {
    void* out;
    int out_dt;
    long long ld_out;
    int use_rowmap;
    IpnRowMap rowmap;
    int split_cols;
    long long split_stride;
    const float* bias;
    int act;
    float alpha;
    const void* mul_src;
    int mul_dt;
    long long ld_mul;
    int mul_mode;
    float mul_scale;
    int accumulate;
  };
  struct Col {
    float bias;
    char* base;  // output base pointer after the column split
    int c;       // column inside the split block
  };
  static __device__ __forceinline__ void col_init(const Params& p, int col, Col& cc) {
    cc.bias = p.bias != nullptr ? p.bias[col] : 0.f;
    cc.base = reinterpret_cast<char*>(p.out);
    cc.c = col;
    if (p.split_cols > 0) {
      const int q = col / p.split_cols;
      cc.c = col - q * p.split_cols;
      cc.base += q * p.split_stride * (p.out_dt == IPN_BF16 ? 2 : 4);
    }
  }
  // epilogues without a separate load phase: empty preload (see EpiGruFwdT for the staged form)
  static constexpr int NARR = 0;
  template <int W> struct Pre {};
  template <int W>
  static __device__ __forceinline__ void preload(const Params&, const Col&, int, int, int, Pre<W>&) {}
  template <int W>
  static __device__ __forceinline__ void applyT(const Params& p, const Col& cc, int col, int row0, int nv,
                                                float (&acc)[G][W], const Pre<W>&) { applyT<W>(p, cc, col, row0, nv, acc); }
  template <int W>
  static __device__ __forceinline__ void applyT(const Params& p, const Col& cc, int col, int row0, int nv,
                                                float (&acc)[1][W]) {
    // offsets first (one hoisted branch), then all loads, then math + stores; rows beyond nv are clamped
    long long off[W];
    if (p.use_rowmap) {
#pragma unroll
      for (int i = 0; i < W; ++i) off[i] = map_row(p.rowmap, row0 + min(i, nv - 1)) + cc.c;
    } else {
#pragma unroll
      for (int i = 0; i < W; ++i) off[i] = (long long)(row0 + min(i, nv - 1)) * p.ld_out + cc.c;
    }
    float m[W];
#pragma unroll
    for (int i = 0; i < W; ++i) m[i] = p.alpha;
    if (p.mul_mode == IPN_MUL_KEEP_MASK) {
      unsigned char mb[W];
#pragma unroll
      for (int i = 0; i < W; ++i)
        mb[i] = reinterpret_cast<const unsigned char*>(p.mul_src)[(long long)(row0 + min(i, nv - 1)) * p.ld_mul + col];
#pragma unroll
      for (int i = 0; i < W; ++i) m[i] = mb[i] ? p.alpha * p.mul_scale : 0.f;
    } else if (p.mul_mode != IPN_MUL_NONE) {
      float y[W];
      if (p.mul_dt == IPN_BF16) {
#pragma unroll
        for (int i = 0; i < W; ++i) y[i] = ld_t<IPN_BF16>(p.mul_src, (long long)(row0 + min(i, nv - 1)) * p.ld_mul + col, 0);
      } else {
#pragma unroll
        for (int i = 0; i < W; ++i) y[i] = ld_t<IPN_F32>(p.mul_src, (long long)(row0 + min(i, nv - 1)) * p.ld_mul + col, 0);
      }
#pragma unroll
      for (int i = 0; i < W; ++i)
        m[i] *= (p.mul_mode == IPN_MUL_SELU_GRAD) ? selu_grad_from_out(y[i]) : (y[i] > 0.f ? 1.f : 0.f);
    }
    float v[W];
    if (p.act == IPN_ACT_NONE) {
#pragma unroll
      for (int i = 0; i < W; ++i) v[i] = (acc[0][i] + cc.bias) * m[i];
    } else if (p.act == IPN_ACT_RELU) {
#pragma unroll
      for (int i = 0; i < W; ++i) v[i] = fmaxf(acc[0][i] + cc.bias, 0.f) * m[i];
    } else {
#pragma unroll
      for (int i = 0; i < W; ++i) v[i] = selu_f(acc[0][i] + cc.bias) * m[i];
    }
    if (p.accumulate == IPN_STORE) {
      if (p.out_dt == IPN_BF16) {
#pragma unroll
        for (int i = 0; i < W; ++i)
          if (i < nv) st_t<IPN_BF16>(cc.base, off[i], v[i], 0);
      } else {
#pragma unroll
        for (int i = 0; i < W; ++i)
          if (i < nv) st_t<IPN_F32>(cc.base, off[i], v[i], 0);
      }
    } else if (p.accumulate == IPN_ATOMIC_ADD) {
#pragma unroll
      for (int i = 0; i < W; ++i)
        if (i < nv) atomicAdd(reinterpret_cast<float*>(cc.base) + off[i], v[i]);
    } else {
      float old[W];
#pragma unroll
      for (int i = 0; i < W; ++i) old[i] = ld_act(cc.base, off[i], p.out_dt);
#pragma unroll
      for (int i = 0; i < W; ++i)
        if (i < nv) st_act(cc.base, off[i], old[i] + v[i], p.out_dt);
    }
  }
};

// =============================================================================================
// Blocked GRU input projection (persistent GEMM with the operand roles swapped: TMEM lanes = activation
// rows R, accumulator columns = gate units n in [0, 3H)): one call = 8 consecutive units of one row = one
// 16-byte vector of the blocked layout vec16(R, g, u) = ((R/128*3 + g)*(H/8) + u/8)*128 + R%128 that the
// persistent GRU layer kernel reads (gru_persist.cuh).  Folds b_ih, the r/z part of b_hh and the 0.5 of
// sigmoid(x) = 0.5 tanh(0.5 x) + 0.5, exactly like gru_prep_p_kernel.
// =============================================================================================
struct EpiBlockedP {
  template <class PP>
  static __device__ __forceinline__ unsigned long long* dbg_buf(const PP&) { return nullptr; }
  static constexpr int G = 1;
  static constexpr int NARR = 0;
  struct Params {
    uint4* out;
    const float* b_ih;
    const float* b_hh;
    int H;
    int ngates;     // 3 (GRU) or 4 (LSTM) arrays per row tile
    int half_mask;  // bit g: gate g is a sigmoid gate evaluated as 0.5 tanh(0.5 x) + 0.5 -> stored pre-halved
    int bhh_mask;   // bit g: b_hh of gate g is folded in (GRU: r, z only -- b_hn stays inside r * (.))
  };
  struct Col {};
  static __device__ __forceinline__ void col_init(const Params&, int, Col&) {}
  template <int W>
  static __device__ __forceinline__ void applyT(const Params& p, const Col&, int col, int row0, int nv, float (&acc)[1][W]) {
    static_assert(W == 8, "one 16-byte vector per call");
    const int H = p.H;
    const int g = row0 / H, u = row0 - g * H;
    const float4 b0 = *reinterpret_cast<const float4*>(p.b_ih + row0), b1 = *reinterpret_cast<const float4*>(p.b_ih + row0 + 4);
    float f[8] = {acc[0][0] + b0.x, acc[0][1] + b0.y, acc[0][2] + b0.z, acc[0][3] + b0.w,
                  acc[0][4] + b1.x, acc[0][5] + b1.y, acc[0][6] + b1.z, acc[0][7] + b1.w};
    if ((p.bhh_mask >> g) & 1) {
      const float4 c0 = *reinterpret_cast<const float4*>(p.b_hh + row0), c1 = *reinterpret_cast<const float4*>(p.b_hh + row0 + 4);
      f[0] += c0.x; f[1] += c0.y; f[2] += c0.z; f[3] += c0.w;
      f[4] += c1.x; f[5] += c1.y; f[6] += c1.z; f[7] += c1.w;
    }
    if ((p.half_mask >> g) & 1) {
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] *= 0.5f;
    }
    uint4 o;
    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) h2[k] = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
    const long long R = col;
    p.out[(((R >> 7) * p.ngates + g) * (H >> 3) + (u >> 3)) * 128 + (R & 127)] = o;
  }
};

// =============================================================================================
// GRU forward step:  acc[g] = (h_prev W_hh^T)[row, g*H + col]   g in {r, z, n}
// =============================================================================================
struct GruFwdParams {
    int H, act_dt, row0;
    long long trow;  // t * B_total: first row of this timestep in time-ordered buffers
    const void* P;
    long long ldP;
    int P_bcast;
    const float* table;
    long long ld_table;
    const int* tok;
    const float* pvec;
    const float* b_hh;
    const void* h_prev;  // slot base [B_total, H]
    void* h_out;         // slot base
    void* gates;         // [T*B_total, 4H] base, nullable
    int gates_blocked;   // write the blocked 5-array layout of the persistent kernels instead (gru_persist.cuh)
    void* y;
    long long ld_y;
    int y_col0;
    const unsigned char* mask;
    long long ld_mask;
    float mask_scale;
    void* final_out;
    int final_dt;
    long long ld_final;
    int final_col0;
    int dbg;  // diagnostics (IPN_DBG_EPI): bit0 skip loads, bit1 skip stores
    int stage;  // 1: the tcgen05 kernel stages the inputs through shared memory (alignment checked on the host)
    unsigned long long* dbg_buf;  // per-CTA phase timestamps (IPN_DBG_TIMING), normally null
  __device__ __forceinline__ float mask_scale_eff() const { return mask != nullptr ? mask_scale : 1.f; }
};

template <int DT, bool STAGE = false>
struct EpiGruFwdT {
  static constexpr int G = 3;
  using Params = GruFwdParams;
  static __device__ __forceinline__ unsigned long long* dbg_buf(const Params& p) { return p.dbg_buf; }
  struct Col {
    float br, bz, bn;  // b_hh
    float cr, cz, cn;  // constant part of the input projection (pvec)
  };
  static __device__ __forceinline__ void col_init(const Params& p, int col, Col& cc) {
    const int H = p.H;
    cc.br = p.b_hh[col]; cc.bz = p.b_hh[H + col]; cc.bn = p.b_hh[2 * H + col];
    cc.cr = cc.cz = cc.cn = 0.f;
    if (p.pvec != nullptr) { cc.cr = p.pvec[col]; cc.cz = p.pvec[H + col]; cc.cn = p.pvec[2 * H + col]; }
  }
  // Load phase, separated from the math so that the tcgen05 kernel can issue the loads of the NEXT 8-row
  // phase before finishing the current one (software pipelining) and the first phase before the
  // accumulators are even ready.  All loads of the W rows are issued array by array with no consumer in
  // between; rows beyond nv re-read the last valid row instead of branching.
  template <int W>
  struct Pre {
    float pr[W], pz[W], pn[W], hp[W];
    unsigned mask_bits;
  };
  template <int W>
  static __device__ __forceinline__ void preload(const Params& p, const Col& cc, int col, int row0, int nv, Pre<W>& q) {
    const int H = p.H, dt = p.act_dt;
    long long R[W], TR[W];
#pragma unroll
    for (int i = 0; i < W; ++i) {
      R[i] = p.row0 + row0 + min(i, nv - 1);
      TR[i] = p.trow + R[i];
    }
#pragma unroll
    for (int i = 0; i < W; ++i) { q.pr[i] = cc.cr; q.pz[i] = cc.cz; q.pn[i] = cc.cn; }
    q.mask_bits = 0xffffffffu;
    const bool ld_on = !(p.dbg & 1);
    float a[W], b[W], c[W];
    if (p.P != nullptr && ld_on) {
#pragma unroll
      for (int i = 0; i < W; ++i) {
        const long long o = (p.P_bcast ? R[i] : TR[i]) * p.ldP + col;
        a[i] = ld_t<DT>(p.P, o, dt); b[i] = ld_t<DT>(p.P, o + H, dt); c[i] = ld_t<DT>(p.P, o + 2 * H, dt);
      }
    } else {
#pragma unroll
      for (int i = 0; i < W; ++i) { a[i] = 0.f; b[i] = 0.f; c[i] = 0.f; }
    }
#pragma unroll
    for (int i = 0; i < W; ++i) q.hp[i] = ld_on ? ld_t<DT>(p.h_prev, R[i] * H + col, dt) : 0.f;
    unsigned char mb[W];
    const bool has_mask = p.y != nullptr && p.mask != nullptr && ld_on;
    if (has_mask) {
#pragma unroll
      for (int i = 0; i < W; ++i) mb[i] = p.mask[TR[i] * p.ld_mask + p.y_col0 + col];
    }
    if (p.table != nullptr && ld_on) {
      int tk[W];
#pragma unroll
      for (int i = 0; i < W; ++i) tk[i] = p.tok[TR[i]];
      float tr[W], tz[W], tn[W];
#pragma unroll
      for (int i = 0; i < W; ++i) {
        const long long o = (long long)tk[i] * p.ld_table + col;
        tr[i] = p.table[o]; tz[i] = p.table[o + H]; tn[i] = p.table[o + 2 * H];
      }
#pragma unroll
      for (int i = 0; i < W; ++i) { q.pr[i] += tr[i]; q.pz[i] += tz[i]; q.pn[i] += tn[i]; }
    }
#pragma unroll
    for (int i = 0; i < W; ++i) { q.pr[i] += a[i]; q.pz[i] += b[i]; q.pn[i] += c[i]; }
    if (has_mask) {
      unsigned bits = 0;
#pragma unroll
      for (int i = 0; i < W; ++i) bits |= (mb[i] ? 1u : 0u) << i;
      q.mask_bits = bits;
    }
  }
  // ---- shared-memory staging (tcgen05 kernel): array order P(r,z,n) | table(r,z,n) | h_prev | mask
  static constexpr int NARR = STAGE ? 8 : 0;
  static constexpr int NARR_MAX = 8;
  static __device__ __forceinline__ bool stage_on(const Params& p) { return p.stage != 0 && !(p.dbg & 1); }
  static __device__ __forceinline__ int stage_arrays(const Params& p, StageArr (&a)[NARR_MAX]) {
    const int H = p.H;
    const long long R0 = p.row0, TR0 = p.trow + p.row0;
    int n = 0;
    if (p.P != nullptr) {
#pragma unroll
      for (int g = 0; g < 3; ++g)
        a[n++] = StageArr{reinterpret_cast<const char*>(p.P) + (long long)g * H * 2, p.ldP * 2, p.P_bcast ? R0 : TR0, nullptr, 0, 2};
    }
    if (p.table != nullptr) {
#pragma unroll
      for (int g = 0; g < 3; ++g)
        a[n++] = StageArr{reinterpret_cast<const char*>(p.table) + (long long)g * H * 4, p.ld_table * 4, 0, p.tok, TR0, 4};
    }
    a[n++] = StageArr{reinterpret_cast<const char*>(p.h_prev), (long long)H * 2, R0, nullptr, 0, 2};
    if (p.y != nullptr && p.mask != nullptr)
      a[n++] = StageArr{reinterpret_cast<const char*>(p.mask) + p.y_col0, p.ld_mask, TR0, nullptr, 0, 1};
    return n;
  }
  // fills Pre from a staged chunk: `chunk` = base of the chunk buffer, arrays laid out in stage_arrays order as
  // [16 rows][128 cols]; lc = column inside the tile, rr0 = first row inside the chunk
  template <int W>
  static __device__ __forceinline__ void preload_smem(const Params& p, const Col& cc, const char* chunk, int lc, int rr0,
                                                      Pre<W>& q) {
#pragma unroll
    for (int i = 0; i < W; ++i) { q.pr[i] = cc.cr; q.pz[i] = cc.cz; q.pn[i] = cc.cn; }
    q.mask_bits = 0xffffffffu;
    const char* s = chunk;
    if (p.P != nullptr) {
      const __nv_bfloat16* a0 = reinterpret_cast<const __nv_bfloat16*>(s);
      const __nv_bfloat16* a1 = a0 + 16 * 128;
      const __nv_bfloat16* a2 = a1 + 16 * 128;
#pragma unroll
      for (int i = 0; i < W; ++i) {
        const int o = (rr0 + i) * 128 + lc;
        q.pr[i] += __bfloat162float(a0[o]); q.pz[i] += __bfloat162float(a1[o]); q.pn[i] += __bfloat162float(a2[o]);
      }
      s += 3 * 16 * 128 * 2;
    }
    if (p.table != nullptr) {
      const float* a0 = reinterpret_cast<const float*>(s);
      const float* a1 = a0 + 16 * 128;
      const float* a2 = a1 + 16 * 128;
#pragma unroll
      for (int i = 0; i < W; ++i) {
        const int o = (rr0 + i) * 128 + lc;
        q.pr[i] += a0[o]; q.pz[i] += a1[o]; q.pn[i] += a2[o];
      }
      s += 3 * 16 * 128 * 4;
    }
    {
      const __nv_bfloat16* a0 = reinterpret_cast<const __nv_bfloat16*>(s);
#pragma unroll
      for (int i = 0; i < W; ++i) q.hp[i] = __bfloat162float(a0[(rr0 + i) * 128 + lc]);
      s += 16 * 128 * 2;
    }
    if (p.y != nullptr && p.mask != nullptr) {
      const unsigned char* a0 = reinterpret_cast<const unsigned char*>(s);
      unsigned bits = 0;
#pragma unroll
      for (int i = 0; i < W; ++i) bits |= (a0[(rr0 + i) * 128 + lc] ? 1u : 0u) << i;
      q.mask_bits = bits;
    }
  }
  // gate math + stores
  template <int W>
  static __device__ __forceinline__ void applyT(const Params& p, const Col& cc, int col, int row0, int nv,
                                                float (&acc)[3][W], const Pre<W>& q) {
    const int H = p.H, dt = p.act_dt;
#pragma unroll
    for (int i = 0; i < W; ++i) {
      const float r = sigmoid_t<DT>(q.pr[i] + acc[0][i] + cc.br);
      const float z = sigmoid_t<DT>(q.pz[i] + acc[1][i] + cc.bz);
      const float hn = acc[2][i] + cc.bn;
      const float n = tanh_t<DT>(q.pn[i] + r * hn);
      const float h = (1.f - z) * n + z * q.hp[i];
      if (i < nv && !(p.dbg & 2)) {
        const long long R = p.row0 + row0 + i, TR = p.trow + R;
        st_t<DT>(p.h_out, R * H + col, h, dt);
        if (p.gates != nullptr) {
          if (p.gates_blocked) {
            const long long gs = (long long)H * 128;
            const long long go = ((((TR >> 7) * 5) * (H >> 3) + (col >> 3)) * 128 + (TR & 127)) * 8 + (col & 7);
            st_t<DT>(p.gates, go, r, dt); st_t<DT>(p.gates, go + gs, z, dt); st_t<DT>(p.gates, go + 2 * gs, n, dt);
            st_t<DT>(p.gates, go + 3 * gs, hn, dt); st_t<DT>(p.gates, go + 4 * gs, q.hp[i], dt);
          } else {
            const long long go = TR * 4 * H + col;
            st_t<DT>(p.gates, go, r, dt); st_t<DT>(p.gates, go + H, z, dt);
            st_t<DT>(p.gates, go + 2 * H, n, dt); st_t<DT>(p.gates, go + 3 * H, hn, dt);
          }
        }
        if (p.y != nullptr) st_t<DT>(p.y, TR * p.ld_y + p.y_col0 + col, ((q.mask_bits >> i) & 1u) ? h * p.mask_scale_eff() : 0.f, dt);
        if (p.final_out != nullptr) st_act(p.final_out, R * p.ld_final + p.final_col0 + col, h, p.final_dt);
      }
    }
  }
  // direct form (no staging): loads issued array by array straight into the math registers
  template <int W>
  static __device__ __forceinline__ void applyT(const Params& p, const Col& cc, int col, int row0, int nv,
                                                float (&acc)[3][W]) {
    const int H = p.H, dt = p.act_dt;
    long long R[W], TR[W];
#pragma unroll
    for (int i = 0; i < W; ++i) {
      R[i] = p.row0 + row0 + min(i, nv - 1);
      TR[i] = p.trow + R[i];
    }
    float pr[W], pz[W], pn[W], hp[W], mk[W];
#pragma unroll
    for (int i = 0; i < W; ++i) { pr[i] = 0.f; pz[i] = 0.f; pn[i] = 0.f; mk[i] = 1.f; }
    const bool ld_on = !(p.dbg & 1);
    if (p.P != nullptr && ld_on) {
#pragma unroll
      for (int i = 0; i < W; ++i) {
        const long long o = (p.P_bcast ? R[i] : TR[i]) * p.ldP + col;
        pr[i] = ld_t<DT>(p.P, o, dt); pz[i] = ld_t<DT>(p.P, o + H, dt); pn[i] = ld_t<DT>(p.P, o + 2 * H, dt);
      }
    }
#pragma unroll
    for (int i = 0; i < W; ++i) hp[i] = ld_on ? ld_t<DT>(p.h_prev, R[i] * H + col, dt) : 0.f;
    if (p.y != nullptr && p.mask != nullptr && ld_on) {
      unsigned char mb[W];
#pragma unroll
      for (int i = 0; i < W; ++i) mb[i] = p.mask[TR[i] * p.ld_mask + p.y_col0 + col];
#pragma unroll
      for (int i = 0; i < W; ++i) mk[i] = mb[i] ? p.mask_scale : 0.f;
    }
    if (p.table != nullptr && ld_on) {
      int tk[W];
#pragma unroll
      for (int i = 0; i < W; ++i) tk[i] = p.tok[TR[i]];
      float tr[W], tz[W], tn[W];
#pragma unroll
      for (int i = 0; i < W; ++i) {
        const long long o = (long long)tk[i] * p.ld_table + col;
        tr[i] = p.table[o]; tz[i] = p.table[o + H]; tn[i] = p.table[o + 2 * H];
      }
#pragma unroll
      for (int i = 0; i < W; ++i) { pr[i] += tr[i]; pz[i] += tz[i]; pn[i] += tn[i]; }
    }
#pragma unroll
    for (int i = 0; i < W; ++i) {
      const float r = sigmoid_t<DT>(pr[i] + cc.cr + acc[0][i] + cc.br);
      const float z = sigmoid_t<DT>(pz[i] + cc.cz + acc[1][i] + cc.bz);
      const float hn = acc[2][i] + cc.bn;
      const float n = tanh_t<DT>(pn[i] + cc.cn + r * hn);
      const float h = (1.f - z) * n + z * hp[i];
      if (i < nv && !(p.dbg & 2)) {
        st_t<DT>(p.h_out, R[i] * H + col, h, dt);
        if (p.gates != nullptr) {
          if (p.gates_blocked) {
            const long long gs = (long long)H * 128;
            const long long go = ((((TR[i] >> 7) * 5) * (H >> 3) + (col >> 3)) * 128 + (TR[i] & 127)) * 8 + (col & 7);
            st_t<DT>(p.gates, go, r, dt); st_t<DT>(p.gates, go + gs, z, dt); st_t<DT>(p.gates, go + 2 * gs, n, dt);
            st_t<DT>(p.gates, go + 3 * gs, hn, dt); st_t<DT>(p.gates, go + 4 * gs, hp[i], dt);
          } else {
            const long long go = TR[i] * 4 * H + col;
            st_t<DT>(p.gates, go, r, dt); st_t<DT>(p.gates, go + H, z, dt);
            st_t<DT>(p.gates, go + 2 * H, n, dt); st_t<DT>(p.gates, go + 3 * H, hn, dt);
          }
        }
        if (p.y != nullptr) st_t<DT>(p.y, TR[i] * p.ld_y + p.y_col0 + col, h * mk[i], dt);
        if (p.final_out != nullptr) st_act(p.final_out, R[i] * p.ld_final + p.final_col0 + col, h, p.final_dt);
      }
    }
  }
};
using EpiGruFwd = EpiGruFwdT<-1>;

// =============================================================================================
// GRU backward.  Pointwise part for one timestep (shared by the GEMM epilogue and the standalone
// kernel that starts the chain).
// =============================================================================================
struct GruBwdPoint {
  int H, act_dt, row0;
  long long trow;       // t * B_total of the step whose gates are differentiated
  const void* gates;    // [T*B_total, 4H], or the blocked 5-array layout of the persistent forward kernel
  int gates_blocked;
  const void* h_prev;   // slot base of the state that ENTERED this step
  const void* dY;       // time-ordered, nullable
  long long ld_dy;
  int y_col0;
  const unsigned char* mask;
  long long ld_mask;
  float mask_scale;
  const float* dh_n;    // nullable, [rows, ld_dhn] indexed by slot row
  long long ld_dhn;
  void* dP;             // [T*B_total, 3H]
  void* dGn;            // [T*B_total, H]
  float* dhz_out;       // [B_total, H] fp32: dh * z for the next (earlier) step
};

template <int DT, int W>
__device__ __forceinline__ void gru_bwd_pointwise(const GruBwdPoint& p, int col, int row0, int nv,
                                                  const float (&dh_in)[W]) {
  const int H = p.H, dt = p.act_dt;
  long long R[W], TR[W];
#pragma unroll
  for (int i = 0; i < W; ++i) {
    R[i] = p.row0 + row0 + min(i, nv - 1);
    TR[i] = p.trow + R[i];
  }
  // ---- phase 1: all loads, array by array (see EpiGruFwdT)
  float r[W], z[W], n[W], hn[W], hp[W], dy[W], dn_[W];
  // gate array stride; blocked layout (gru_persist.cuh): vec16(R, a, u) = ((R/128*5 + a)*(H/8) + u/8)*128 + R%128
  const long long gs = p.gates_blocked ? (long long)H * 128 : (long long)H;
#pragma unroll
  for (int i = 0; i < W; ++i) {
    const long long go = p.gates_blocked
                             ? ((((TR[i] >> 7) * 5) * (H >> 3) + (col >> 3)) * 128 + (TR[i] & 127)) * 8 + (col & 7)
                             : TR[i] * 4 * H + col;
    r[i] = ld_t<DT>(p.gates, go, dt); z[i] = ld_t<DT>(p.gates, go + gs, dt);
    n[i] = ld_t<DT>(p.gates, go + 2 * gs, dt); hn[i] = ld_t<DT>(p.gates, go + 3 * gs, dt);
  }
#pragma unroll
  for (int i = 0; i < W; ++i) hp[i] = ld_t<DT>(p.h_prev, R[i] * H + col, dt);
#pragma unroll
  for (int i = 0; i < W; ++i) { dy[i] = 0.f; dn_[i] = 0.f; }
  if (p.dY != nullptr) {
#pragma unroll
    for (int i = 0; i < W; ++i) dy[i] = ld_t<DT>(p.dY, TR[i] * p.ld_dy + p.y_col0 + col, dt);
    if (p.mask != nullptr) {
      unsigned char mb[W];
#pragma unroll
      for (int i = 0; i < W; ++i) mb[i] = p.mask[TR[i] * p.ld_mask + p.y_col0 + col];
#pragma unroll
      for (int i = 0; i < W; ++i) dy[i] = mb[i] ? dy[i] * p.mask_scale : 0.f;
    }
  }
  if (p.dh_n != nullptr) {
#pragma unroll
    for (int i = 0; i < W; ++i) dn_[i] = p.dh_n[R[i] * p.ld_dhn + col];
  }
  // ---- phase 2
#pragma unroll
  for (int i = 0; i < W; ++i) {
    const float dh = dh_in[i] + dy[i] + dn_[i];
    const float dn = dh * (1.f - z[i]) * (1.f - n[i] * n[i]);
    const float dz = dh * (hp[i] - n[i]) * z[i] * (1.f - z[i]);
    const float dr = dn * hn[i] * r[i] * (1.f - r[i]);
    if (i < nv) {
      st_t<DT>(p.dP, TR[i] * 3 * H + col, dr, dt);
      st_t<DT>(p.dP, TR[i] * 3 * H + H + col, dz, dt);
      st_t<DT>(p.dP, TR[i] * 3 * H + 2 * H + col, dn, dt);
      st_t<DT>(p.dGn, TR[i] * H + col, dn * r[i], dt);
      p.dhz_out[R[i] * H + col] = dh * z[i];
    }
  }
}

// GEMM epilogue: acc = ([dP_r, dP_z | dGn] W_hh)[row, col] of step s  ->  dh wrt the state that entered
// step s; then either emit dh0 (first step of the chain) or differentiate the previous step's gates.
struct GruBwdParams {
    const float* dhz_in;  // [B_total, H] fp32, nullable
    int is_first_step;    // this GEMM produced the gradient wrt h0
    void* dh0;
    int dh0_dt;
    long long ld_dh0;
    int dh0_selu;
    const void* h0;  // slot base of h0 (act_dt), for SELU'
    GruBwdPoint pw;  // describes step s-1 (unused when is_first_step)
};

template <int DT>
struct EpiGruBwdT {
  template <class PP>
  static __device__ __forceinline__ unsigned long long* dbg_buf(const PP&) { return nullptr; }
  static constexpr int G = 1;
  using Params = GruBwdParams;
  struct Col {};
  static __device__ __forceinline__ void col_init(const Params&, int, Col&) {}
  // epilogues without a separate load phase: empty preload (see EpiGruFwdT for the staged form)
  static constexpr int NARR = 0;
  template <int W> struct Pre {};
  template <int W>
  static __device__ __forceinline__ void preload(const Params&, const Col&, int, int, int, Pre<W>&) {}
  template <int W>
  static __device__ __forceinline__ void applyT(const Params& p, const Col& cc, int col, int row0, int nv,
                                                float (&acc)[G][W], const Pre<W>&) { applyT<W>(p, cc, col, row0, nv, acc); }
  template <int W>
  static __device__ __forceinline__ void applyT(const Params& p, const Col&, int col, int row0, int nv,
                                                float (&acc)[1][W]) {
    const int H = p.pw.H;
    float dh[W];
#pragma unroll
    for (int i = 0; i < W; ++i) dh[i] = acc[0][i];
    if (p.dhz_in != nullptr) {
      float t[W];
#pragma unroll
      for (int i = 0; i < W; ++i) t[i] = p.dhz_in[(long long)(p.pw.row0 + row0 + min(i, nv - 1)) * H + col];
#pragma unroll
      for (int i = 0; i < W; ++i) dh[i] += t[i];
    }
    if (p.is_first_step) {
      if (p.dh0 != nullptr) {
#pragma unroll
        for (int i = 0; i < W; ++i) {
          if (i < nv) {
            const long long R = p.pw.row0 + row0 + i;
            float v = dh[i];
            if (p.dh0_selu) v *= selu_grad_from_out(ld_t<DT>(p.h0, R * H + col, p.pw.act_dt));
            st_act(p.dh0, R * p.ld_dh0 + col, v, p.dh0_dt);
          }
        }
      }
      return;
    }
    gru_bwd_pointwise<DT, W>(p.pw, col, row0, nv, dh);
  }
};
using EpiGruBwd = EpiGruBwdT<-1>;

// =============================================================================================
// LSTM forward step: acc[g] = (h_prev W_hh^T)[row, g*H + col], g in {i, f, g, o}
// =============================================================================================
struct EpiLstmFwd {
  template <class PP>
  static __device__ __forceinline__ unsigned long long* dbg_buf(const PP&) { return nullptr; }
  static constexpr int G = 4;
  struct Params {
    int H, act_dt;
    long long trow;
    const void* P;
    long long ldP;
    const float* b_hh;
    const float* c_prev;  // [B,H] fp32
    float* c_out;
    void* h_out;
    void* gates;  // [T*B, 4H] post-activation (i, f, g, o)
    void* y;
    long long ld_y;
    int y_col0;
    long long ytrow;  // first row of this step in y (differs from trow when the output is time-flipped)
    const float* table;
    long long ld_table;
    const int* tok_scalar;
    int gates_blocked;  // write `gates` in the persistent kernels' blocked 5-array layout (i, f, g, o, c_t; bf16)
  };
  struct Col {
    float b[4];
  };
  static __device__ __forceinline__ void col_init(const Params& p, int col, Col& cc) {
#pragma unroll
    for (int g = 0; g < 4; ++g) cc.b[g] = p.b_hh != nullptr ? p.b_hh[g * p.H + col] : 0.f;
    if (p.table != nullptr) {
      const long long o = (long long)(*p.tok_scalar) * p.ld_table + col;
#pragma unroll
      for (int g = 0; g < 4; ++g) cc.b[g] += p.table[o + (long long)g * p.H];
    }
  }
  // epilogues without a separate load phase: empty preload (see EpiGruFwdT for the staged form)
  static constexpr int NARR = 0;
  template <int W> struct Pre {};
  template <int W>
  static __device__ __forceinline__ void preload(const Params&, const Col&, int, int, int, Pre<W>&) {}
  template <int W>
  static __device__ __forceinline__ void applyT(const Params& p, const Col& cc, int col, int row0, int nv,
                                                float (&acc)[G][W], const Pre<W>&) { applyT<W>(p, cc, col, row0, nv, acc); }
  template <int W>
  static __device__ __forceinline__ void applyT(const Params& p, const Col& cc, int col, int row0, int nv,
                                                float (&acc)[4][W]) {
    const int H = p.H, dt = p.act_dt;
    float pre[4][W], cp[W];
#pragma unroll
    for (int i = 0; i < W; ++i) {
      cp[i] = 0.f;
#pragma unroll
      for (int g = 0; g < 4; ++g) pre[g][i] = 0.f;
      if (i < nv) {
        const long long R = row0 + i, TR = p.trow + R;
#pragma unroll
        for (int g = 0; g < 4; ++g) pre[g][i] = ld_act(p.P, TR * p.ldP + (long long)g * H + col, dt);
        cp[i] = p.c_prev[R * H + col];
      }
    }
#pragma unroll
    for (int i = 0; i < W; ++i) {
      if (i < nv) {
        const long long R = row0 + i, TR = p.trow + R;
        const float gi = sigmoid_acc(pre[0][i] + acc[0][i] + cc.b[0]);
        const float gf = sigmoid_acc(pre[1][i] + acc[1][i] + cc.b[1]);
        const float gg = tanhf(pre[2][i] + acc[2][i] + cc.b[2]);
        const float go = sigmoid_acc(pre[3][i] + acc[3][i] + cc.b[3]);
        const float c = gf * cp[i] + gi * gg;
        const float h = go * tanhf(c);
        p.c_out[R * H + col] = c;
        st_act(p.h_out, R * H + col, h, dt);
        if (p.gates != nullptr && p.gates_blocked) {
          // element (row TR, array a, unit col) of the blocked layout (gru_persist.cuh), so that the backward pass
          // can run the persistent cluster kernel (lstm_persist.cu) on a layer whose forward ran tick by tick
          __nv_bfloat16* gb = reinterpret_cast<__nv_bfloat16*>(p.gates);
          const long long vpr = H >> 3;
          const long long e0 = ((((TR >> 7) * 5) * vpr + (col >> 3)) * 128 + (TR & 127)) * 8 + (col & 7);
          const long long as = vpr * 128 * 8;
          gb[e0] = __float2bfloat16_rn(gi); gb[e0 + as] = __float2bfloat16_rn(gf);
          gb[e0 + 2 * as] = __float2bfloat16_rn(gg); gb[e0 + 3 * as] = __float2bfloat16_rn(go);
          gb[e0 + 4 * as] = __float2bfloat16_rn(c);
        } else if (p.gates != nullptr) {
          const long long o = TR * 4 * H + col;
          st_act(p.gates, o, gi, dt); st_act(p.gates, o + H, gf, dt);
          st_act(p.gates, o + 2 * H, gg, dt); st_act(p.gates, o + 3 * H, go, dt);
        }
        if (p.y != nullptr) st_act(p.y, (p.ytrow + R) * p.ld_y + p.y_col0 + col, h, dt);
      }
    }
  }
};

// LSTM backward pointwise for one step: given dh (total gradient wrt h_t) and dc_in (gradient wrt c_t
// arriving from step t+1), writes dP[t] (gradient wrt the 4 pre-activations) and dc_out = dc * f.
struct LstmBwdPoint {
  int H, act_dt;
  long long trow;
  const void* gates;
  const float* c_prev;  // c_{t-1}
  const float* c_cur;   // c_t
  const void* dY;
  long long ld_dy;
  int y_col0;
  long long dytrow;    // first row of this step in dY
  const float* dc_in;  // nullable
  float* dc_out;
  void* dP;  // [T*B, 4H]
};

template <int W>
__device__ __forceinline__ void lstm_bwd_pointwise(const LstmBwdPoint& p, int col, int row0, int nv,
                                                   const float (&dh_in)[W]) {
  const int H = p.H, dt = p.act_dt;
  float dh[W], gi[W], gf[W], gg[W], go[W], cp[W], cc[W], dci[W];
#pragma unroll
  for (int i = 0; i < W; ++i) {
    dh[i] = dh_in[i]; gi[i] = gf[i] = gg[i] = go[i] = cp[i] = cc[i] = dci[i] = 0.f;
    if (i < nv) {
      const long long R = row0 + i, TR = p.trow + R;
      if (p.dY != nullptr) dh[i] += ld_act(p.dY, (p.dytrow + R) * p.ld_dy + p.y_col0 + col, dt);
      const long long o = TR * 4 * H + col;
      gi[i] = ld_act(p.gates, o, dt); gf[i] = ld_act(p.gates, o + H, dt);
      gg[i] = ld_act(p.gates, o + 2 * H, dt); go[i] = ld_act(p.gates, o + 3 * H, dt);
      cp[i] = p.c_prev[R * H + col]; cc[i] = p.c_cur[R * H + col];
      if (p.dc_in != nullptr) dci[i] = p.dc_in[R * H + col];
    }
  }
#pragma unroll
  for (int i = 0; i < W; ++i) {
    if (i < nv) {
      const long long R = row0 + i, TR = p.trow + R;
      const float tc = tanhf(cc[i]);
      const float dc = dci[i] + dh[i] * go[i] * (1.f - tc * tc);
      const long long o = TR * 4 * H + col;
      st_act(p.dP, o, dc * gg[i] * gi[i] * (1.f - gi[i]), dt);
      st_act(p.dP, o + H, dc * cp[i] * gf[i] * (1.f - gf[i]), dt);
      st_act(p.dP, o + 2 * H, dc * gi[i] * (1.f - gg[i] * gg[i]), dt);
      st_act(p.dP, o + 3 * H, dh[i] * tc * go[i] * (1.f - go[i]), dt);
      p.dc_out[R * H + col] = dc * gf[i];
    }
  }
}

// GEMM epilogue: acc = (dP[s] W_hh)[row, col] = gradient wrt h_{s-1} from the recurrence.
struct EpiLstmBwd {
  template <class PP>
  static __device__ __forceinline__ unsigned long long* dbg_buf(const PP&) { return nullptr; }
  static constexpr int G = 1;
  struct Params {
    int is_first_step;  // nothing earlier to differentiate
    LstmBwdPoint pw;    // step s-1
  };
  struct Col {};
  static __device__ __forceinline__ void col_init(const Params&, int, Col&) {}
  // epilogues without a separate load phase: empty preload (see EpiGruFwdT for the staged form)
  static constexpr int NARR = 0;
  template <int W> struct Pre {};
  template <int W>
  static __device__ __forceinline__ void preload(const Params&, const Col&, int, int, int, Pre<W>&) {}
  template <int W>
  static __device__ __forceinline__ void applyT(const Params& p, const Col& cc, int col, int row0, int nv,
                                                float (&acc)[G][W], const Pre<W>&) { applyT<W>(p, cc, col, row0, nv, acc); }
  template <int W>
  static __device__ __forceinline__ void applyT(const Params& p, const Col&, int col, int row0, int nv,
                                                float (&acc)[1][W]) {
    if (p.is_first_step) return;
    float dh[W];
#pragma unroll
    for (int i = 0; i < W; ++i) dh[i] = acc[0][i];
    lstm_bwd_pointwise<W>(p.pw, col, row0, nv, dh);
  }
};

}  // namespace ipn
