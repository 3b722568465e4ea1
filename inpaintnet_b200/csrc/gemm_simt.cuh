// fp32 CUDA-core GEMM core ("fp32 mode"): same problem statement and epilogue functors as the
// tcgen05 core, any dtype/layout/alignment, used for exact-parity runs and for tiny problems.
//   D[m, g, n] = sum_seg sum_k A_seg[m, k] * B_seg[g*gate_stride + n, k]
#pragma once
#include "common.cuh"

namespace ipn {

struct SimtSeg {
  const void* A;
  long long lda;
  int transA;
  const void* B;
  long long ldb;
  int transB;
  int K;
};

template <class Epi>
struct SimtProblem {
  SimtSeg seg[2];
  int nseg;
  int M, N;
  int gate_stride;
  int in_dt;
  typename Epi::Params epi;
};

template <class Epi>
struct SimtBatch {
  SimtProblem<Epi> p[2];
  int split_k;
};

constexpr int SIMT_BM = 64, SIMT_BN = 64, SIMT_BK = 16;

template <class Epi>
__global__ void __launch_bounds__(256) simt_gemm_kernel(const __grid_constant__ SimtBatch<Epi> batch) {
  constexpr int G = Epi::G;
  __shared__ float As[SIMT_BK][SIMT_BM + 4];
  __shared__ float Bs[G][SIMT_BK][SIMT_BN + 4];
  const int prob = blockIdx.z / batch.split_k;
  const int ksplit = blockIdx.z - prob * batch.split_k;
  const SimtProblem<Epi>& P = batch.p[prob];
  const int n0 = blockIdx.x * SIMT_BN, m0 = blockIdx.y * SIMT_BM;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int dt = P.in_dt;

  float acc[G][4][4];
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[g][i][j] = 0.f;

  const int chunks0 = (P.seg[0].K + SIMT_BK - 1) / SIMT_BK;
  const int chunks_total = chunks0 + (P.nseg > 1 ? (P.seg[1].K + SIMT_BK - 1) / SIMT_BK : 0);
  const int per = (chunks_total + batch.split_k - 1) / batch.split_k;
  const int kc_begin = ksplit * per, kc_end = min(chunks_total, kc_begin + per);

  for (int kc = kc_begin; kc < kc_end; ++kc) {
    const int si = kc >= chunks0 ? 1 : 0;
    const SimtSeg& S = P.seg[si];
    const int k0 = (si ? kc - chunks0 : kc) * SIMT_BK;
    // ---- load A tile (64 x 16) ----
#pragma unroll
    for (int it = 0; it < (SIMT_BM * SIMT_BK) / 256; ++it) {
      const int idx = it * 256 + tid;
      int m, k;
      if (S.transA) { k = idx / SIMT_BM; m = idx % SIMT_BM; } else { m = idx / SIMT_BK; k = idx % SIMT_BK; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < P.M && gk < S.K)
        v = ld_act(S.A, S.transA ? (long long)gk * S.lda + gm : (long long)gm * S.lda + gk, dt);
      As[k][m] = v;
    }
    // ---- load B tiles (G x 64 x 16) ----
#pragma unroll
    for (int g = 0; g < G; ++g) {
#pragma unroll
      for (int it = 0; it < (SIMT_BN * SIMT_BK) / 256; ++it) {
        const int idx = it * 256 + tid;
        int n, k;
        if (S.transB) { k = idx / SIMT_BN; n = idx % SIMT_BN; } else { n = idx / SIMT_BK; k = idx % SIMT_BK; }
        const int gn = n0 + n, gk = k0 + k;
        float v = 0.f;
        if (gn < P.N && gk < S.K) {
          const long long r = (long long)g * P.gate_stride + gn;
          v = ld_act(S.B, S.transB ? (long long)gk * S.ldb + r : r * S.ldb + gk, dt);
        }
        Bs[g][k][n] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SIMT_BK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[g][k][tx * 4]);
        const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[g][i][j] = fmaf(a[i], b[j], acc[g][i][j]);
      }
    }
    __syncthreads();
  }

  const int col0 = n0 + tx * 4;
  if (col0 >= P.N) return;
  const int row0 = m0 + ty * 4;
  if (row0 >= P.M) return;
  const int nv = min(4, P.M - row0);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int col = col0 + j;
    if (col >= P.N) break;
    typename Epi::Col cc;
    Epi::col_init(P.epi, col, cc);
    float a[G][4];
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int i = 0; i < 4; ++i) a[g][i] = acc[g][i][j];
    Epi::template applyT<4>(P.epi, cc, col, row0, nv, a);
  }
}

}  // namespace ipn
