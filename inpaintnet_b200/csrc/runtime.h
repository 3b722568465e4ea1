// Host-side runtime shared by the API translation units: error text, device check, launch counter,
// TMA descriptor cache.
#pragma once
#include "common.cuh"

namespace ipn {

int ensure_device();  // IPN_OK when the current device is sm_100
void count_launch(int n = 1);

// Optional per-kernel-class timing with CUDA events on the launching stream (bench.py's roofline
// numbers).  Disabled by default: prof_begin/prof_end are no-ops unless ipn_prof_enable(1) was called.
void prof_begin(const char* tag, double flops, double bytes, cudaStream_t stream);
void prof_end(cudaStream_t stream);
struct ProfScope {
  cudaStream_t s;
  ProfScope(const char* tag, double flops, double bytes, cudaStream_t stream) : s(stream) {
    prof_begin(tag, flops, bytes, stream);
  }
  ~ProfScope() { prof_end(s); }
};

// bf16, 2-D, SWIZZLE_128B tensor map over a row-major matrix: `inner` contiguous elements per row,
// `outer` rows, row stride `ld` elements; box = {64, box_outer}.  Cached by value of all arguments.
int get_tensor_map(CUtensorMap* out, const void* ptr, unsigned long long inner, unsigned long long outer,
                   long long ld, unsigned box_outer);

// bf16, 3-D, SWIZZLE_128B tensor map: dims {d0 (contiguous), d1, d2}, strides ld1/ld2 in elements, box {64, box1, box2}.
int get_tensor_map_3d(CUtensorMap* out, const void* ptr, unsigned long long d0, unsigned long long d1, unsigned long long d2,
                      long long ld1, long long ld2, unsigned box1, unsigned box2);

#define IPN_LAUNCH_CHECK()                                                                        \
  do {                                                                                            \
    cudaError_t _e = cudaGetLastError();                                                          \
    if (_e != cudaSuccess) {                                                                      \
      ipn::set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return IPN_ERR_CUDA;                                                                        \
    }                                                                                             \
    ipn::count_launch();                                                                          \
  } while (0)

}  // namespace ipn
