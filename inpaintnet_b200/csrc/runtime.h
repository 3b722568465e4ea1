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
