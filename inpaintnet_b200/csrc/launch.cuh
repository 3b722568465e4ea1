// Host-side launch helpers for the two GEMM cores.
#pragma once
#include "runtime.h"
#include "gemm_simt.cuh"
#include "gemm_umma.cuh"
#include "epilogues.cuh"

namespace ipn {

// Operand of a GEMM segment as the host sees it.
//   rows_total x K logical matrix; trans == 0: row-major [rows_total, K] (K contiguous), ld = row stride
//                                  trans == 1: row-major [K, rows_total] (rows contiguous), ld = row stride
//   row_off: first logical row used by this problem (slot offsets etc.); k_off: first K index.
struct HostOperand {
  const void* ptr;
  long long ld;
  int trans;
  long long rows_total;
  long long row_off;
  long long k_off;
};

inline void fill_simt_seg(SimtSeg& s, const HostOperand& a, const HostOperand& b, int K, int in_dt) {
  const long long es = (in_dt == IPN_BF16) ? 2 : 4;
  const long long aoff = a.trans ? a.k_off * a.ld + a.row_off : a.row_off * a.ld + a.k_off;
  const long long boff = b.trans ? b.k_off * b.ld + b.row_off : b.row_off * b.ld + b.k_off;
  s.A = reinterpret_cast<const char*>(a.ptr) + aoff * es;
  s.lda = a.ld;
  s.transA = a.trans;
  s.B = reinterpret_cast<const char*>(b.ptr) + boff * es;
  s.ldb = b.ld;
  s.transB = b.trans;
  s.K = K;
}

// x = activation-side operand (output rows), w = weight-side operand (output columns / TMEM lanes).
// box_rows_x: rows per TMA box of a K-major X operand (= BR of the kernel config)
inline int fill_umma_seg(UmmaSeg& s, const HostOperand& x, const HostOperand& w, int K, int box_rows_x) {
  s.K = K;
  if (!w.trans) {
    IPN_PROPAGATE(get_tensor_map(&s.tmW, w.ptr, (unsigned long long)(w.k_off + K), (unsigned long long)w.rows_total,
                                 w.ld, UMMA_BC));
    s.w_c0 = (int)w.k_off;
    s.w_c1 = (int)w.row_off;
  } else {
    IPN_PROPAGATE(get_tensor_map(&s.tmW, w.ptr, (unsigned long long)w.rows_total, (unsigned long long)(w.k_off + K),
                                 w.ld, 64));
    s.w_c0 = (int)w.row_off;
    s.w_c1 = (int)w.k_off;
  }
  if (!x.trans) {
    IPN_PROPAGATE(get_tensor_map(&s.tmX, x.ptr, (unsigned long long)(x.k_off + K), (unsigned long long)x.rows_total,
                                 x.ld, (unsigned)box_rows_x));
    s.x_c0 = (int)x.k_off;
    s.x_c1 = (int)x.row_off;
  } else {
    IPN_PROPAGATE(get_tensor_map(&s.tmX, x.ptr, (unsigned long long)x.rows_total, (unsigned long long)(x.k_off + K),
                                 x.ld, 64));
    s.x_c0 = (int)x.row_off;
    s.x_c1 = (int)x.k_off;
  }
  return IPN_OK;
}

template <class P>
inline double batch_flops(const P* p, int nprob, int G) {
  double f = 0;
  for (int i = 0; i < nprob; ++i) {
    double k = 0;
    for (int s = 0; s < p[i].nseg; ++s) k += p[i].seg[s].K;
    f += 2.0 * p[i].M * (double)p[i].N * G * k;
  }
  return f;
}

template <class Cfg, class Epi>
int launch_umma(const UmmaBatch<Epi>& batch, int nprob, int maxM, int maxN, cudaStream_t stream,
                const char* tag = "umma") {
  ProfScope prof(tag, batch_flops(batch.p, nprob, Epi::G), 0.0, stream);
  static bool configured = false;
  auto kern = umma_gemm_kernel<Cfg, Epi>;
  if (!configured) {
    IPN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  dim3 grid(cdiv(maxN, UMMA_BC), cdiv(maxM, Cfg::BR), nprob * batch.split_k);
  IPN_REQUIRE(grid.y <= 65535, IPN_ERR_ARG, "too many row tiles (%u) for one launch", grid.y);
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(batch);
  IPN_LAUNCH_CHECK();
  return IPN_OK;
}

template <class Cfg, class Epi>
int launch_umma_persist(const UmmaBatch<Epi>& batch, int nprob, int maxM, int maxN, cudaStream_t stream,
                        const char* tag = "umma", int row_fastest = 0, int max_ctas = 0) {
  ProfScope prof(tag, batch_flops(batch.p, nprob, Epi::G), 0.0, stream);
  static bool configured = false;
  static int sms = 148;
  auto kern = umma_gemm_persist_kernel<Cfg, Epi>;
  if (!configured) {
    IPN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    configured = true;
  }
  const int ntx = cdiv(maxN, Cfg::PAIR ? 2 * UMMA_BC : UMMA_BC), nty = cdiv(maxM, Cfg::BR);
  const long long ntiles = (long long)ntx * nty * nprob * batch.split_k;
  IPN_REQUIRE(ntiles < (1LL << 30), IPN_ERR_ARG, "too many tiles");
  const int per = Cfg::PAIR ? 2 : 1;
  const int ctas = (max_ctas > 0 && max_ctas < sms) ? (max_ctas < per ? per : max_ctas) : sms;
  const long long workers = ntiles < ctas / per ? ntiles : ctas / per;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(workers * per), 1, 1);
  cfg.blockDim = dim3(Cfg::THREADS, 1, 1);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = per;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  IPN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, batch, ntx, nty, (int)ntiles, row_fastest));
  IPN_LAUNCH_CHECK();
  return IPN_OK;
}

template <class Epi>
int launch_simt(const SimtBatch<Epi>& batch, int nprob, int maxM, int maxN, cudaStream_t stream,
                const char* tag = "simt") {
  ProfScope prof(tag, batch_flops(batch.p, nprob, Epi::G), 0.0, stream);
  dim3 grid(cdiv(maxN, SIMT_BN), cdiv(maxM, SIMT_BM), nprob * batch.split_k);
  simt_gemm_kernel<Epi><<<grid, 256, 0, stream>>>(batch);
  IPN_LAUNCH_CHECK();
  return IPN_OK;
}

}  // namespace ipn
