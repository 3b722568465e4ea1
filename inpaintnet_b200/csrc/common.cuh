// Shared device/host helpers for the inpaintnet_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/inpaintnet_b200.h"

namespace ipn {

// ---------------------------------------------------------------------------------------------
// error plumbing (C-ABI: int status + ipn_last_error())
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define IPN_CHECK_CUDA(expr)                                                                     \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      ipn::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));      \
      return IPN_ERR_CUDA;                                                                       \
    }                                                                                            \
  } while (0)

#define IPN_REQUIRE(cond, code, ...)                                                             \
  do {                                                                                           \
    if (!(cond)) {                                                                               \
      ipn::set_error(__VA_ARGS__);                                                               \
      return (code);                                                                             \
    }                                                                                            \
  } while (0)

#define IPN_PROPAGATE(expr)                                                                      \
  do {                                                                                           \
    int _s = (expr);                                                                             \
    if (_s != IPN_OK) return _s;                                                                 \
  } while (0)

// ---------------------------------------------------------------------------------------------
// dtype-erased element access. dt: IPN_F32 (0) or IPN_BF16 (1). Branches are warp uniform.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float ld_act(const void* p, long long i, int dt) {
  return dt == IPN_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i])
                        : reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void st_act(void* p, long long i, float v, int dt) {
  if (dt == IPN_BF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(p)[i] = v;
}

// W contiguous elements starting at element index i. `vec` says 16-byte accesses are legal
// (base pointer 16B aligned, i multiple of 8 elements for bf16 / 4 for fp32).
template <int W>
__device__ __forceinline__ void ld_act_vec(const void* p, long long i, int dt, bool vec, float (&out)[W]) {
  if (dt == IPN_BF16) {
    const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(p) + i;
    if (vec && (W % 8 == 0)) {
#pragma unroll
      for (int c = 0; c < W / 8; ++c) {
        uint4 u = reinterpret_cast<const uint4*>(q)[c];
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float2 f = __bfloat1622float2(h2[k]);
          out[c * 8 + 2 * k] = f.x;
          out[c * 8 + 2 * k + 1] = f.y;
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < W; ++k) out[k] = __bfloat162float(q[k]);
    }
  } else {
    const float* q = reinterpret_cast<const float*>(p) + i;
    if (vec && (W % 4 == 0)) {
#pragma unroll
      for (int c = 0; c < W / 4; ++c) {
        float4 f = reinterpret_cast<const float4*>(q)[c];
        out[c * 4] = f.x; out[c * 4 + 1] = f.y; out[c * 4 + 2] = f.z; out[c * 4 + 3] = f.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < W; ++k) out[k] = q[k];
    }
  }
}

// nvalid: number of leading elements that may be touched (column guard); vec path needs nvalid == W.
template <int W>
__device__ __forceinline__ void ld_act_n(const void* p, long long i, int dt, bool vec, int nvalid, float (&out)[W]) {
  if (nvalid >= W) { ld_act_vec<W>(p, i, dt, vec, out); return; }
#pragma unroll
  for (int k = 0; k < W; ++k) out[k] = (k < nvalid) ? ld_act(p, i + k, dt) : 0.f;
}

template <int W>
__device__ __forceinline__ void st_act_vec(void* p, long long i, int dt, bool vec, const float (&v)[W]) {
  if (dt == IPN_BF16) {
    __nv_bfloat16* q = reinterpret_cast<__nv_bfloat16*>(p) + i;
    if (vec && (W % 8 == 0)) {
#pragma unroll
      for (int c = 0; c < W / 8; ++c) {
        uint4 u;
        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int k = 0; k < 4; ++k) h2[k] = __floats2bfloat162_rn(v[c * 8 + 2 * k], v[c * 8 + 2 * k + 1]);
        reinterpret_cast<uint4*>(q)[c] = u;
      }
    } else {
#pragma unroll
      for (int k = 0; k < W; ++k) q[k] = __float2bfloat16_rn(v[k]);
    }
  } else {
    float* q = reinterpret_cast<float*>(p) + i;
    if (vec && (W % 4 == 0)) {
#pragma unroll
      for (int c = 0; c < W / 4; ++c)
        reinterpret_cast<float4*>(q)[c] = make_float4(v[c * 4], v[c * 4 + 1], v[c * 4 + 2], v[c * 4 + 3]);
    } else {
#pragma unroll
      for (int k = 0; k < W; ++k) q[k] = v[k];
    }
  }
}

template <int W>
__device__ __forceinline__ void st_act_n(void* p, long long i, int dt, bool vec, int nvalid, const float (&v)[W]) {
  if (nvalid >= W) { st_act_vec<W>(p, i, dt, vec, v); return; }
#pragma unroll
  for (int k = 0; k < W; ++k) if (k < nvalid) st_act(p, i + k, v[k], dt);
}

__host__ __device__ __forceinline__ bool vec_ok(const void* p, long long ld, int dt) {
  const long long per16 = (dt == IPN_BF16) ? 8 : 4;
  return p != nullptr && (reinterpret_cast<uintptr_t>(p) % 16 == 0) && (ld % per16 == 0);
}

// ---------------------------------------------------------------------------------------------
// math
// ---------------------------------------------------------------------------------------------
#define IPN_SELU_ALPHA 1.6732632423543772848170429916717f
#define IPN_SELU_SCALE 1.0507009873554804934193349852946f

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float selu_f(float x) {
  return IPN_SELU_SCALE * (x > 0.f ? x : IPN_SELU_ALPHA * (expf(x) - 1.f));
}
// derivative of SELU expressed through its OUTPUT y:  y>0 -> scale ; else y + scale*alpha
__device__ __forceinline__ float selu_grad_from_out(float y) {
  return y > 0.f ? IPN_SELU_SCALE : (y + IPN_SELU_SCALE * IPN_SELU_ALPHA);
}
__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == IPN_ACT_SELU) return selu_f(x);
  if (act == IPN_ACT_RELU) return fmaxf(x, 0.f);
  return x;
}

// compile-time dtype variants (DT = IPN_BF16 / IPN_F32, or -1 = decide at run time from `rt`).  With a
// compile-time dtype there is no branch between a load and its conversion, so the compiler can issue all
// loads of an epilogue phase back to back (memory-level parallelism) instead of serialising them.
template <int DT>
__device__ __forceinline__ float ld_t(const void* p, long long i, int rt) {
  if constexpr (DT == IPN_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  else if constexpr (DT == IPN_F32) return reinterpret_cast<const float*>(p)[i];
  else return ld_act(p, i, rt);
}
template <int DT>
__device__ __forceinline__ void st_t(void* p, long long i, float v, int rt) {
  if constexpr (DT == IPN_BF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else if constexpr (DT == IPN_F32) reinterpret_cast<float*>(p)[i] = v;
  else st_act(p, i, v, rt);
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// bf16 mode: one MUFU op per gate (outputs are rounded to bf16 anyway); fp32 mode: accurate libm paths
template <int DT>
__device__ __forceinline__ float sigmoid_t(float x) {
  if constexpr (DT == IPN_BF16) return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f);
  else return sigmoid_acc(x);
}
template <int DT>
__device__ __forceinline__ float tanh_t(float x) {
  if constexpr (DT == IPN_BF16) return tanh_approx(x);
  else return tanhf(x);
}

// One input array of an epilogue that the tcgen05 kernel stages through shared memory with bulk copies
// (rows of 128 output columns): element (row, col) of the problem lives at
//   ptr + src_row * row_stride_bytes + col * eb,   src_row = gather ? gather[row + gather_add] : row + row_add
struct StageArr {
  const char* ptr;
  long long row_stride_bytes;
  long long row_add;
  const int* gather;
  long long gather_add;
  int eb;  // bytes per element (1, 2 or 4)
};

// 3-level affine row map: off = (r / g1) * s1 + ((r % g1) / g2) * s2 + (r % g2) * s3
__device__ __forceinline__ long long map_row(const IpnRowMap& m, int r) {
  const int a = r / m.g1, rem = r - a * m.g1;
  const int b = rem / m.g2, c = rem - b * m.g2;
  return (long long)a * m.s1 + (long long)b * m.s2 + (long long)c * m.s3;
}

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace ipn
