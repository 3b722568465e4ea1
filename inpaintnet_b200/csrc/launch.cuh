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

// box_rows_b: rows per TMA box of a K-major B operand (= BNG of the kernel config)
inline int fill_umma_seg(UmmaSeg& s, const HostOperand& a, const HostOperand& b, int K, int box_rows_b) {
  s.K = K;
  if (!a.trans) {
    IPN_PROPAGATE(get_tensor_map(&s.tmA, a.ptr, (unsigned long long)(a.k_off + K), (unsigned long long)a.rows_total,
                                 a.ld, UMMA_BM));
    s.a_c0 = (int)a.k_off;
    s.a_c1 = (int)a.row_off;
  } else {
    IPN_PROPAGATE(get_tensor_map(&s.tmA, a.ptr, (unsigned long long)a.rows_total, (unsigned long long)(a.k_off + K),
                                 a.ld, 64));
    s.a_c0 = (int)a.row_off;
    s.a_c1 = (int)a.k_off;
  }
  if (!b.trans) {
    IPN_PROPAGATE(get_tensor_map(&s.tmB, b.ptr, (unsigned long long)(b.k_off + K), (unsigned long long)b.rows_total,
                                 b.ld, (unsigned)box_rows_b));
    s.b_c0 = (int)b.k_off;
    s.b_c1 = (int)b.row_off;
  } else {
    IPN_PROPAGATE(get_tensor_map(&s.tmB, b.ptr, (unsigned long long)b.rows_total, (unsigned long long)(b.k_off + K),
                                 b.ld, 64));
    s.b_c0 = (int)b.row_off;
    s.b_c1 = (int)b.k_off;
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
  dim3 grid(cdiv(maxN, Cfg::BNG), cdiv(maxM, UMMA_BM), nprob * batch.split_k);
  kern<<<grid, UMMA_THREADS, Cfg::SMEM_BYTES, stream>>>(batch);
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
