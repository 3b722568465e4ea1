// dev microbenchmark: tcgen05.mma issue rate for bf16 SS-mode tiles, 1-CTA and CTA-pair, on all SMs.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_rate tests/dev/mma_rate.cu -I inpaintnet_b200/csrc -lcuda
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace ipn;

template <bool PAIR>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, int kstages, int a_tmem, int nacc, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? ptx::cluster_ctarank() : 0;
  for (int i = threadIdx.x; i < (kstages * (16384 + 32768)) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0) {
    if (lane == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
    __syncwarp();
    if (PAIR) { ptx::tmem_alloc_pair<512>(&slot); ptx::tmem_relinquish_pair(); }
    else { ptx::tmem_alloc<512>(&slot); ptx::tmem_relinquish(); }
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  if (PAIR) ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0 && rank == 0) {
    // whole warp runs the loop (uniform address math), one elected lane issues
    const uint32_t idesc = ptx::make_idesc_bf16(PAIR ? 256 : 128, N, 0, 0);
    const long long t0 = clock64();
    const uint32_t a_base = ptx::smem_u32(smem), b_base = ptx::smem_u32(smem + kstages * 16384);
    int st = 0, acc = 0;
    for (int it = 0; it < iters; ++it) {
      const uint64_t da0 = ptx::make_smem_desc(a_base + st * 16384, 16, 1024), db0 = ptx::make_smem_desc(b_base + st * 32768, 16, 1024);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t da = da0 + (uint64_t)(kk * 2), db = db0 + (uint64_t)(kk * 2);
        const uint32_t d = tmem + (uint32_t)(acc * N);
        if (ptx::elect_one()) {
          if (a_tmem) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(tmem + 448u + kk * 8), "l"(db),
                         "r"(idesc), "r"(1u) : "memory");
          } else if (PAIR) ptx::umma_bf16_pair(d, da, db, idesc, 1u);
          else ptx::umma_bf16(d, da, db, idesc, 1u);
        }
        if (++acc == nacc) acc = 0;
      }
      if (++st == kstages) st = 0;
    }
    if (ptx::elect_one()) { if (PAIR) ptx::umma_commit_pair(&bar); else ptx::umma_commit(&bar); }
    __syncwarp();
    ptx::mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  }
  __syncthreads();
  if (PAIR) ptx::cluster_sync_all();
  if (warp == 0) {
    ptx::tc_fence_after();
    if (PAIR) ptx::tmem_dealloc_pair<512>(tmem); else ptx::tmem_dealloc<512>(tmem);
  }
}

template <bool PAIR>
static void run(int N, int kstages, int a_tmem, int nblk, int nacc = 1) {
  const int iters = 2000;
  unsigned long long* out;
  cudaMalloc(&out, nblk * 8);
  cudaMemset(out, 0, nblk * 8);
  const int smem = kstages * (16384 + 32768);
  auto kern = rate_kernel<PAIR>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nblk); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, N, iters, kstages, a_tmem, nacc, out);
    if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); exit(1); }
  }
  unsigned long long h[256];
  cudaMemcpy(h, out, nblk * 8, cudaMemcpyDeviceToHost);
  double s = 0; int n = 0; unsigned long long mx = 0;
  for (int i = 0; i < nblk; ++i) if (h[i]) { s += h[i]; ++n; if (h[i] > mx) mx = h[i]; }
  const double per = s / n / (iters * 4.0);
  const double floor_c = (PAIR ? 256.0 : 128.0) * N / (256.0 * (PAIR ? 2 : 1));
  printf("%s N=%3d nacc=%d a_tmem=%d blocks=%3d : %.1f cycles/MMA (max CTA %.1f)  floor %.0f  -> %.0f%% of tensor peak\n",
         PAIR ? "pair M=256" : "1cta M=128", N, nacc, a_tmem, nblk, per, mx / (iters * 4.0), floor_c, 100.0 * floor_c / per);
  cudaFree(out);
}

int main() {
  const int nblk = 148;
  for (int nacc : {1, 2, 4}) {
    for (int N : {64, 96, 128, 192, 256}) if (nacc * N <= 448) run<false>(N, 2, 0, nblk, nacc);
    for (int N : {64, 128, 192, 256}) if (nacc * N <= 448) run<true>(N, 2, 0, nblk, nacc);
  }
  for (int nacc : {1, 2}) for (int N : {64, 128, 192}) if (nacc * N <= 448) run<false>(N, 2, 1, nblk, nacc);
  return 0;
}
