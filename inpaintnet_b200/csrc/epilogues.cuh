// Epilogue functors shared by the tcgen05 core (W = 16 columns per call, thread == output row)
// and the fp32 SIMT core (W = 4).  acc[g][i] is the accumulator of gate g, column col0 + i.
#pragma once
#include "common.cuh"

namespace ipn {

// =============================================================================================
// Linear: out = alpha * act(acc + bias) * mul     (+ row map, column split, accumulate modes)
// =============================================================================================
struct EpiLinear {
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

  template <int W>
  static __device__ __forceinline__ void apply(const Params& p, int row, int col0, int nvalid, float (&acc)[1][W]) {
    float v[W];
#pragma unroll
    for (int i = 0; i < W; ++i) {
      float x = acc[0][i];
      if (p.bias != nullptr && i < nvalid) x += p.bias[col0 + i];
      v[i] = apply_act(x, p.act) * p.alpha;
    }
    if (p.mul_mode != IPN_MUL_NONE) {
      const long long mo = (long long)row * p.ld_mul + col0;
      if (p.mul_mode == IPN_MUL_KEEP_MASK) {
        const unsigned char* m = reinterpret_cast<const unsigned char*>(p.mul_src) + mo;
#pragma unroll
        for (int i = 0; i < W; ++i)
          if (i < nvalid) v[i] *= (m[i] ? p.mul_scale : 0.f);
      } else {
        float y[W];
        ld_act_n<W>(p.mul_src, mo, p.mul_dt, vec_ok(p.mul_src, p.ld_mul, p.mul_dt) && (col0 % 8 == 0), nvalid, y);
#pragma unroll
        for (int i = 0; i < W; ++i)
          v[i] *= (p.mul_mode == IPN_MUL_SELU_GRAD) ? selu_grad_from_out(y[i]) : (y[i] > 0.f ? 1.f : 0.f);
      }
    }
    int c = col0;
    char* base = reinterpret_cast<char*>(p.out);
    if (p.split_cols > 0) {
      const int q = col0 / p.split_cols;
      c = col0 - q * p.split_cols;
      base += q * p.split_stride * (p.out_dt == IPN_BF16 ? 2 : 4);
    }
    const long long off = (p.use_rowmap ? map_row(p.rowmap, row) : (long long)row * p.ld_out) + c;
    if (p.accumulate == IPN_STORE) {
      const bool vec = (reinterpret_cast<uintptr_t>(base) % 16 == 0) && (off % 8 == 0);
      st_act_n<W>(base, off, p.out_dt, vec, nvalid, v);
    } else if (p.accumulate == IPN_ATOMIC_ADD) {
      float* o = reinterpret_cast<float*>(base) + off;
#pragma unroll
      for (int i = 0; i < W; ++i)
        if (i < nvalid) atomicAdd(o + i, v[i]);
    } else {
#pragma unroll
      for (int i = 0; i < W; ++i)
        if (i < nvalid) st_act(base, off + i, ld_act(base, off + i, p.out_dt) + v[i], p.out_dt);
    }
  }
};

// =============================================================================================
// GRU forward step:  acc[g] = (h_prev W_hh^T)[row, g*H + col]   g in {r, z, n}
// =============================================================================================
struct EpiGruFwd {
  static constexpr int G = 3;
  struct Params {
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
  };

  template <int W>
  static __device__ __forceinline__ void apply(const Params& p, int row, int col0, int nvalid, float (&acc)[3][W]) {
    const int H = p.H, dt = p.act_dt;
    const long long R = p.row0 + row;   // row inside the slot
    const long long TR = p.trow + R;    // row in time-ordered buffers
    const bool al = (col0 % 8 == 0) && (H % 8 == 0);
    float pre[3][W];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int i = 0; i < W; ++i) pre[g][i] = 0.f;
    if (p.P != nullptr) {
      const bool v = al && vec_ok(p.P, p.ldP, dt);
      const long long PR = p.P_bcast ? R : TR;
#pragma unroll
      for (int g = 0; g < 3; ++g) ld_act_n<W>(p.P, PR * p.ldP + (long long)g * H + col0, dt, v, nvalid, pre[g]);
    }
    if (p.table != nullptr) {
      const long long tk = p.tok[TR];
      const bool v = al && vec_ok(p.table, p.ld_table, IPN_F32);
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        float t[W];
        ld_act_n<W>(p.table, tk * p.ld_table + (long long)g * H + col0, IPN_F32, v, nvalid, t);
#pragma unroll
        for (int i = 0; i < W; ++i) pre[g][i] += t[i];
      }
    }
    float hp[W];
    ld_act_n<W>(p.h_prev, R * H + col0, dt, al && vec_ok(p.h_prev, H, dt), nvalid, hp);
    float r[W], z[W], n[W], hn[W], h[W];
#pragma unroll
    for (int i = 0; i < W; ++i) {
      const int c = col0 + (i < nvalid ? i : 0);
      float pr = pre[0][i], pz = pre[1][i], pn = pre[2][i];
      if (p.pvec != nullptr) { pr += p.pvec[c]; pz += p.pvec[H + c]; pn += p.pvec[2 * H + c]; }
      r[i] = sigmoid_acc(pr + acc[0][i] + p.b_hh[c]);
      z[i] = sigmoid_acc(pz + acc[1][i] + p.b_hh[H + c]);
      hn[i] = acc[2][i] + p.b_hh[2 * H + c];
      n[i] = tanhf(pn + r[i] * hn[i]);
      h[i] = (1.f - z[i]) * n[i] + z[i] * hp[i];
    }
    st_act_n<W>(p.h_out, R * H + col0, dt, al && vec_ok(p.h_out, H, dt), nvalid, h);
    if (p.gates != nullptr) {
      const bool v = al && vec_ok(p.gates, 4 * H, dt);
      const long long go = TR * 4 * H + col0;
      st_act_n<W>(p.gates, go, dt, v, nvalid, r);
      st_act_n<W>(p.gates, go + H, dt, v, nvalid, z);
      st_act_n<W>(p.gates, go + 2 * H, dt, v, nvalid, n);
      st_act_n<W>(p.gates, go + 3 * H, dt, v, nvalid, hn);
    }
    if (p.y != nullptr) {
      float yv[W];
      if (p.mask != nullptr) {
        const unsigned char* m = p.mask + TR * p.ld_mask + p.y_col0 + col0;
#pragma unroll
        for (int i = 0; i < W; ++i) yv[i] = (i < nvalid && m[i]) ? h[i] * p.mask_scale : 0.f;
      } else {
#pragma unroll
        for (int i = 0; i < W; ++i) yv[i] = h[i];
      }
      st_act_n<W>(p.y, TR * p.ld_y + p.y_col0 + col0, dt, al && vec_ok(p.y, p.ld_y, dt) && (p.y_col0 % 8 == 0),
                  nvalid, yv);
    }
    if (p.final_out != nullptr)
      st_act_n<W>(p.final_out, R * p.ld_final + p.final_col0 + col0, p.final_dt,
                  al && vec_ok(p.final_out, p.ld_final, p.final_dt) && (p.final_col0 % 8 == 0), nvalid, h);
  }
};

// =============================================================================================
// GRU backward.  Pointwise part for one timestep (shared by the GEMM epilogue and the standalone
// kernel that starts the chain).
// =============================================================================================
struct GruBwdPoint {
  int H, act_dt, row0;
  long long trow;       // t * B_total of the step whose gates are differentiated
  const void* gates;    // [T*B_total, 4H]
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

template <int W>
__device__ __forceinline__ void gru_bwd_pointwise(const GruBwdPoint& p, int row, int col0, int nvalid,
                                                  const float (&dh_in)[W]) {
  const int H = p.H, dt = p.act_dt;
  const long long R = p.row0 + row, TR = p.trow + R;
  const bool al = (col0 % 8 == 0) && (H % 8 == 0);
  float dh[W];
#pragma unroll
  for (int i = 0; i < W; ++i) dh[i] = dh_in[i];
  if (p.dY != nullptr) {
    float dy[W];
    ld_act_n<W>(p.dY, TR * p.ld_dy + p.y_col0 + col0, dt, al && vec_ok(p.dY, p.ld_dy, dt) && (p.y_col0 % 8 == 0),
                nvalid, dy);
    if (p.mask != nullptr) {
      const unsigned char* m = p.mask + TR * p.ld_mask + p.y_col0 + col0;
#pragma unroll
      for (int i = 0; i < W; ++i) dh[i] += (i < nvalid && m[i]) ? dy[i] * p.mask_scale : 0.f;
    } else {
#pragma unroll
      for (int i = 0; i < W; ++i) dh[i] += dy[i];
    }
  }
  if (p.dh_n != nullptr) {
#pragma unroll
    for (int i = 0; i < W; ++i)
      if (i < nvalid) dh[i] += p.dh_n[R * p.ld_dhn + col0 + i];
  }
  float r[W], z[W], n[W], hn[W], hp[W];
  const bool vg = al && vec_ok(p.gates, 4 * H, dt);
  const long long go = TR * 4 * H + col0;
  ld_act_n<W>(p.gates, go, dt, vg, nvalid, r);
  ld_act_n<W>(p.gates, go + H, dt, vg, nvalid, z);
  ld_act_n<W>(p.gates, go + 2 * H, dt, vg, nvalid, n);
  ld_act_n<W>(p.gates, go + 3 * H, dt, vg, nvalid, hn);
  ld_act_n<W>(p.h_prev, R * H + col0, dt, al && vec_ok(p.h_prev, H, dt), nvalid, hp);
  float dr[W], dz[W], dn[W], dgn[W], dhz[W];
#pragma unroll
  for (int i = 0; i < W; ++i) {
    dn[i] = dh[i] * (1.f - z[i]) * (1.f - n[i] * n[i]);
    dz[i] = dh[i] * (hp[i] - n[i]) * z[i] * (1.f - z[i]);
    dr[i] = dn[i] * hn[i] * r[i] * (1.f - r[i]);
    dgn[i] = dn[i] * r[i];
    dhz[i] = dh[i] * z[i];
  }
  const bool vp = al && vec_ok(p.dP, 3 * H, dt);
  st_act_n<W>(p.dP, TR * 3 * H + col0, dt, vp, nvalid, dr);
  st_act_n<W>(p.dP, TR * 3 * H + H + col0, dt, vp, nvalid, dz);
  st_act_n<W>(p.dP, TR * 3 * H + 2 * H + col0, dt, vp, nvalid, dn);
  st_act_n<W>(p.dGn, TR * H + col0, dt, al && vec_ok(p.dGn, H, dt), nvalid, dgn);
  st_act_n<W>(p.dhz_out, R * H + col0, IPN_F32, al && vec_ok(p.dhz_out, H, IPN_F32), nvalid, dhz);
}

// GEMM epilogue: acc = ([dP_r, dP_z | dGn] W_hh)[row, col] of step s  ->  dh wrt the state that entered
// step s; then either emit dh0 (first step of the chain) or differentiate the previous step's gates.
struct EpiGruBwd {
  static constexpr int G = 1;
  struct Params {
    const float* dhz_in;  // [B_total, H] fp32, nullable
    int is_first_step;    // this GEMM produced the gradient wrt h0
    void* dh0;
    int dh0_dt;
    long long ld_dh0;
    int dh0_selu;
    const void* h0;  // slot base of h0 (act_dt), for SELU'
    GruBwdPoint pw;  // describes step s-1 (unused when is_first_step)
  };

  template <int W>
  static __device__ __forceinline__ void apply(const Params& p, int row, int col0, int nvalid, float (&acc)[1][W]) {
    const int H = p.pw.H;
    const long long R = p.pw.row0 + row;
    float dh[W];
#pragma unroll
    for (int i = 0; i < W; ++i) dh[i] = acc[0][i];
    if (p.dhz_in != nullptr) {
      float t[W];
      ld_act_n<W>(p.dhz_in, R * H + col0, IPN_F32, (col0 % 4 == 0) && vec_ok(p.dhz_in, H, IPN_F32), nvalid, t);
#pragma unroll
      for (int i = 0; i < W; ++i) dh[i] += t[i];
    }
    if (p.is_first_step) {
      if (p.dh0 != nullptr) {
        if (p.dh0_selu) {
          float h0[W];
          ld_act_n<W>(p.h0, R * H + col0, p.pw.act_dt, false, nvalid, h0);
#pragma unroll
          for (int i = 0; i < W; ++i) dh[i] *= selu_grad_from_out(h0[i]);
        }
        st_act_n<W>(p.dh0, R * p.ld_dh0 + col0, p.dh0_dt, false, nvalid, dh);
      }
      return;
    }
    gru_bwd_pointwise<W>(p.pw, row, col0, nvalid, dh);
  }
};

// =============================================================================================
// LSTM forward step: acc[g] = (h_prev W_hh^T)[row, g*H + col], g in {i, f, g, o}
// =============================================================================================
struct EpiLstmFwd {
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
  };
  template <int W>
  static __device__ __forceinline__ void apply(const Params& p, int row, int col0, int nvalid, float (&acc)[4][W]) {
    const int H = p.H, dt = p.act_dt;
    const long long R = row, TR = p.trow + row;
    const bool al = (col0 % 8 == 0) && (H % 8 == 0);
    float pre[4][W];
    const bool v = al && vec_ok(p.P, p.ldP, dt);
#pragma unroll
    for (int g = 0; g < 4; ++g) ld_act_n<W>(p.P, TR * p.ldP + (long long)g * H + col0, dt, v, nvalid, pre[g]);
    float cp[W];
    ld_act_n<W>(p.c_prev, R * H + col0, IPN_F32, al && vec_ok(p.c_prev, H, IPN_F32), nvalid, cp);
    float gi[W], gf[W], gg[W], go[W], c[W], h[W];
#pragma unroll
    for (int i = 0; i < W; ++i) {
      const int cc = col0 + (i < nvalid ? i : 0);
      float bi = 0.f, bf = 0.f, bg = 0.f, bo = 0.f;
      if (p.b_hh != nullptr) { bi = p.b_hh[cc]; bf = p.b_hh[H + cc]; bg = p.b_hh[2 * H + cc]; bo = p.b_hh[3 * H + cc]; }
      gi[i] = sigmoid_acc(pre[0][i] + acc[0][i] + bi);
      gf[i] = sigmoid_acc(pre[1][i] + acc[1][i] + bf);
      gg[i] = tanhf(pre[2][i] + acc[2][i] + bg);
      go[i] = sigmoid_acc(pre[3][i] + acc[3][i] + bo);
      c[i] = gf[i] * cp[i] + gi[i] * gg[i];
      h[i] = go[i] * tanhf(c[i]);
    }
    st_act_n<W>(p.c_out, R * H + col0, IPN_F32, al && vec_ok(p.c_out, H, IPN_F32), nvalid, c);
    st_act_n<W>(p.h_out, R * H + col0, dt, al && vec_ok(p.h_out, H, dt), nvalid, h);
    if (p.gates != nullptr) {
      const bool vg = al && vec_ok(p.gates, 4 * H, dt);
      const long long o = TR * 4 * H + col0;
      st_act_n<W>(p.gates, o, dt, vg, nvalid, gi);
      st_act_n<W>(p.gates, o + H, dt, vg, nvalid, gf);
      st_act_n<W>(p.gates, o + 2 * H, dt, vg, nvalid, gg);
      st_act_n<W>(p.gates, o + 3 * H, dt, vg, nvalid, go);
    }
    if (p.y != nullptr)
      st_act_n<W>(p.y, TR * p.ld_y + p.y_col0 + col0, dt, al && vec_ok(p.y, p.ld_y, dt) && (p.y_col0 % 8 == 0), nvalid, h);
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
  const float* dc_in;  // nullable
  float* dc_out;
  void* dP;  // [T*B, 4H]
};

template <int W>
__device__ __forceinline__ void lstm_bwd_pointwise(const LstmBwdPoint& p, int row, int col0, int nvalid,
                                                   const float (&dh_in)[W]) {
  const int H = p.H, dt = p.act_dt;
  const long long R = row, TR = p.trow + row;
  const bool al = (col0 % 8 == 0) && (H % 8 == 0);
  float dh[W];
#pragma unroll
  for (int i = 0; i < W; ++i) dh[i] = dh_in[i];
  if (p.dY != nullptr) {
    float dy[W];
    ld_act_n<W>(p.dY, TR * p.ld_dy + p.y_col0 + col0, dt, al && vec_ok(p.dY, p.ld_dy, dt) && (p.y_col0 % 8 == 0),
                nvalid, dy);
#pragma unroll
    for (int i = 0; i < W; ++i) dh[i] += dy[i];
  }
  float gi[W], gf[W], gg[W], go[W], cp[W], cc[W], dci[W];
  const bool vg = al && vec_ok(p.gates, 4 * H, dt);
  const long long o = TR * 4 * H + col0;
  ld_act_n<W>(p.gates, o, dt, vg, nvalid, gi);
  ld_act_n<W>(p.gates, o + H, dt, vg, nvalid, gf);
  ld_act_n<W>(p.gates, o + 2 * H, dt, vg, nvalid, gg);
  ld_act_n<W>(p.gates, o + 3 * H, dt, vg, nvalid, go);
  const bool vf = al && vec_ok(p.c_prev, H, IPN_F32);
  ld_act_n<W>(p.c_prev, R * H + col0, IPN_F32, vf, nvalid, cp);
  ld_act_n<W>(p.c_cur, R * H + col0, IPN_F32, vf, nvalid, cc);
  if (p.dc_in != nullptr) ld_act_n<W>(p.dc_in, R * H + col0, IPN_F32, vf, nvalid, dci);
  else {
#pragma unroll
    for (int i = 0; i < W; ++i) dci[i] = 0.f;
  }
  float di[W], df[W], dg[W], dob[W], dco[W];
#pragma unroll
  for (int i = 0; i < W; ++i) {
    const float tc = tanhf(cc[i]);
    const float dc = dci[i] + dh[i] * go[i] * (1.f - tc * tc);
    dob[i] = dh[i] * tc * go[i] * (1.f - go[i]);
    di[i] = dc * gg[i] * gi[i] * (1.f - gi[i]);
    df[i] = dc * cp[i] * gf[i] * (1.f - gf[i]);
    dg[i] = dc * gi[i] * (1.f - gg[i] * gg[i]);
    dco[i] = dc * gf[i];
  }
  const bool vp = al && vec_ok(p.dP, 4 * H, dt);
  st_act_n<W>(p.dP, o, dt, vp, nvalid, di);
  st_act_n<W>(p.dP, o + H, dt, vp, nvalid, df);
  st_act_n<W>(p.dP, o + 2 * H, dt, vp, nvalid, dg);
  st_act_n<W>(p.dP, o + 3 * H, dt, vp, nvalid, dob);
  st_act_n<W>(p.dc_out, R * H + col0, IPN_F32, vf, nvalid, dco);
}

// GEMM epilogue: acc = (dP[s] W_hh)[row, col] = gradient wrt h_{s-1} from the recurrence.
struct EpiLstmBwd {
  static constexpr int G = 1;
  struct Params {
    int is_first_step;  // nothing earlier to differentiate
    LstmBwdPoint pw;    // step s-1
  };
  template <int W>
  static __device__ __forceinline__ void apply(const Params& p, int row, int col0, int nvalid, float (&acc)[1][W]) {
    if (p.is_first_step) return;
    float dh[W];
#pragma unroll
    for (int i = 0; i < W; ++i) dh[i] = acc[0][i];
    lstm_bwd_pointwise<W>(p.pw, row, col0, nvalid, dh);
  }
};

}  // namespace ipn
