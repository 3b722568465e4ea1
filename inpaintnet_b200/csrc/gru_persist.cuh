// Persistent GRU layer kernels (forward and backward): one launch runs ALL timesteps of a layer.
//
// A CTA (or, with the column split, a cluster of 2 / 4 CTAs that share the tile and own 1/2 / 1/4 of the
// hidden units each) owns a tile of 128 batch rows of one direction for the whole sequence, so the serial
// chain needs no grid-wide synchronisation: the only cross-step hand-off is inside the CTA / cluster (TMA
// store of h_t -> TMA reload as the next step's A operand, tracked per 64-unit k-block with mbarriers;
// across the CTAs of a cluster the barriers are signalled remotely with release/acquire at cluster scope).
// "Lanes = batch rows": MMA M = 128 rows (TMEM lanes), N = 3 gates x 64 hidden units per chunk
// (forward) so one thread holds r, z, n of 16 consecutive units of ITS row per tcgen05.ld.  All
// per-element traffic of the epilogue uses the BLOCKED layout below (16-byte vectors, fully
// coalesced 512-byte warp accesses, no shared-memory staging); only the tensors that TMA-fed GEMMs
// consume (h sequence, y, dP, dGn) are row-major and leave through a swizzled staging tile + TMA store.
//
// Blocked layout of a logical [R rows, A arrays, H units] bf16 tensor (R multiple of 128):
//     vec16(R, a, u) = ((R/128 * A + a) * (H/8) + u/8) * 128 + R%128        (index of a 16-byte vector)
// i.e. for fixed (array, 8-unit group) the 128 rows of a tile are contiguous.
#pragma once
#include "runtime.h"
#include "ptx.cuh"
#include <cuda_bf16.h>

namespace ipn {

constexpr int GP_ROWS = 128;                      // batch rows per CTA
constexpr int GP_CH = 64;                         // hidden units per chunk / per k-block
constexpr int GP_KB_BYTES = GP_ROWS * 128;        // one A k-block: 128 rows x 64 bf16 (SWIZZLE_128B)
constexpr int GP_THREADS = 640;                   // W producer, MMA, store, A loader, 16 epilogue warps
constexpr int GP_GATE_ARRAYS = 5;                 // saved per step: r, z, n, W_hn h + b_hn, h_prev

struct GruPersistFwdDir {
  alignas(64) CUtensorMap tmW;  // W_hh [3H, H], box 64 x 64
  alignas(64) CUtensorMap tmH;  // hseq [(T+1)*Bt, H], box 64 x 128 (loads and stores)
  alignas(64) CUtensorMap tmY;  // y [T*Bt, ld_y], box 64 x 128 (stores)
  const uint4* Pblk;            // blocked [T*Bt, 3, H] input projection, nullable
  const uint4* ftab;            // token-table mode: folded bf16 table [rows, 3H]; P(t, row) = ftab[tok[t, row]] (Pblk null)
  const int* tok;               // token ids [T*Bt], time-ordered rows
  uint4* gates;                 // blocked [T*Bt, 5, H], nullable
  const float* b_hh;
  const float* pvec;            // nullable
  int reverse, y_col0, has_y;
  int p_t_stride;               // 128-row tile index of (time tt, tile x) in Pblk = tt * p_t_stride + p_t0 + x
  long long p_t0;
};
struct GruPersistFwd {
  GruPersistFwdDir d[2];
  int T, H, Bt;
  int row0, s_begin, s_end;   // window of this call: rows [row0, row0 + 128*gridDim.x), processing steps [s_begin, s_end)
  int dbg;                    // diagnostics (IPN_GPF_DBG): 1 no P loads, 2 no gate stores, 4 no gate math, 8 no W loads
  unsigned long long* timing; // per-CTA wait-cycle counters (ipn_dbg_set_timing_buffer), normally null
  unsigned long long* trace;  // cycle stamps of CTA (0,0), steps 8 and 9: [role][step - 8][event] x 32 (tests/dev/persist_time.py)
};

struct GruPersistBwdDir {
  alignas(64) CUtensorMap tmW;    // W_hh [3H, H] as {64 n, 3H k, H/64 n-blocks}: box {64, 64, n-blocks per stage} (MN-major B)
  alignas(64) CUtensorMap tmDP;   // dP [T*Bt, 3H], box 64 x 128 (stores and A-operand loads)
  alignas(64) CUtensorMap tmDG;   // dGn [T*Bt, H], box 64 x 128
  const uint4* gates;             // blocked [T*Bt, 5, H]
  const uint4* dYblk;             // blocked [T*Bt, 1, H] (mask already applied), nullable
  const float* dh_n;              // nullable, row-major fp32 [Bt, ld_dhn]
  long long ld_dhn;
  void* dh0;                      // nullable
  long long ld_dh0;
  const __nv_bfloat16* h0;        // row-major [Bt, H]: the initial state (SELU' of it when dh0_selu)
  int dh0_dt, dh0_selu, reverse, pad_;
};
struct GruPersistBwd {
  GruPersistBwdDir d[2];
  int T, H, Bt;
  int nbs;                        // 64-wide n-blocks of W_hh staged per CTA and stage
  int dbg;
  int pad_;
  unsigned long long* timing;
};

}  // namespace ipn

namespace ipn {
// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 x = __bfloat1622float2(h2[k]);
    f[2 * k] = x.x;
    f[2 * k + 1] = x.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int k = 0; k < 4; ++k) h2[k] = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
  return u;
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 u;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(addr));
  return u;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& u) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}
// streaming 16-byte global accesses (each element is touched once per step)
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 u;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
               : "l"(p));
  return u;
}
__device__ __forceinline__ void stg_stream(uint4* p, const uint4& u) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w)
               : "memory");
}
// 256-bit loads (LDG.E.256 on sm_100)
__device__ __forceinline__ void ldg_nc_32B(const void* p, uint4& a, uint4& b) {     // read-only data
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
__device__ __forceinline__ void ldg_cg_32B(const void* p, uint4& a, uint4& b) {     // data written earlier in this kernel (L2)
  asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p) : "memory");
}
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }


extern unsigned long long* g_dbg_timing;

// mbarrier wait that adds the waited cycles to a counter when diagnostics are on
__device__ __forceinline__ void wait_acc(uint64_t* bar, uint32_t parity, bool on, long long& acc) {
  if (on) {
    const long long t0 = clock64();
    ptx::mbar_wait(bar, parity);
    acc += clock64() - t0;
  } else {
    ptx::mbar_wait(bar, parity);
  }
}

}  // namespace ipn
