// HBM-bound helper kernels: token reshapes, embedding gather / scatter-add, argmax, Philox RNG,
// reparameterisation, fused CE+KL loss (fwd+bwd), fused flat Adam, bf16 weight pack, column sums.
#include "runtime.h"

namespace ipn {

// ---------------------------------------------------------------------------------------------
// tokens
// ---------------------------------------------------------------------------------------------
__global__ void tokens_time_major_kernel(const long long* tok, int B, int T, int V, int* out, int* flag) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * T) return;
  const int t = (int)(i / B), b = (int)(i % B);
  const long long v = tok[(long long)b * T + t];
  if (v < 0 || v >= V) { if (flag) *flag = 1; out[i] = 0; } else out[i] = (int)v;
}

__global__ void dec_prev_tokens_kernel(const long long* tok, int B, int V, int* out, int* flag) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * 24) return;
  const int j = (int)(i / (4LL * B));
  const int rem = (int)(i % (4LL * B));
  const int ib = rem / B, b = rem % B;
  const int t = 6 * ib + j;
  int v = V;
  if (t > 0) {
    const long long x = tok[(long long)b * 24 + t - 1];
    if (x < 0 || x >= V) { if (flag) *flag = 1; v = 0; } else v = (int)x;
  }
  out[i] = v;
}

__global__ void embed_rows_kernel(const float* emb, int E, const int* tok, long long rows, void* out, int out_dt,
                                  long long ld_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * ld_out) return;
  const long long r = i / ld_out;
  const int c = (int)(i % ld_out);
  const float v = c < E ? emb[(long long)tok[r] * E + c] : 0.f;
  st_act(out, i, v, out_dt);
}

__global__ void embed_grad_kernel(const void* dX, int dx_dt, long long ld_dx, const int* tok, long long rows, int E,
                                  int V, float* demb, int skip_id, float* dskip) {
  extern __shared__ float sm[];  // (V + 1) * E
  const int n = (V + 1) * E;
  for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const long long per = (rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * per, r1 = min(rows, r0 + per);
  for (long long i = r0 * E + threadIdx.x; i < r1 * E; i += blockDim.x) {
    const long long r = i / E;
    const int e = (int)(i % E);
    int tk = tok[r];
    if (tk < 0) continue;       // "no token" rows (zero input) receive no gradient
    if (tk == skip_id) tk = V;  // extra slot
    atomicAdd(&sm[tk * E + e], ld_act(dX, r * ld_dx + e, dx_dt));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = sm[i];
    if (v == 0.f) continue;
    if (i < V * E) { if (i / E != skip_id) atomicAdd(&demb[i], v); }
    else if (dskip) atomicAdd(&dskip[i - V * E], v);
  }
}

__global__ void argmax_rows_kernel(const float* logits, int rows, int V, IpnRowMap rm, int use_rm, int* tok_out,
                                   long long* samples_out, IpnRowMap sm_, int use_sm) {
  const int warp = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* p = logits + (use_rm ? map_row(rm, warp) : (long long)warp * V);
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int c = lane; c < V; c += 32) {
    const float v = p[c];
    if (v > best) { best = v; bi = c; }  // strict > keeps the lowest index within a lane
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (lane == 0) {
    if (bi == 0x7fffffff) bi = 0;  // all NaN row
    if (tok_out) tok_out[warp] = bi;
    if (samples_out) samples_out[use_sm ? map_row(sm_, warp) : warp] = bi;
  }
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

__global__ void rng_keep_mask_kernel(unsigned long long seed, unsigned long long offset, long long n, float p_drop,
                                     unsigned char* out) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // 16 outputs per thread
  if (q * 16 >= n) return;
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  const unsigned long long c = offset + (unsigned long long)q * 4;
  // thresholds on 8-bit lanes would be too coarse; use 4 calls x 4 words = 16 uniform 32-bit values
  unsigned char m[16];
  const uint32_t thr = (uint32_t)fminf(4294967295.f, p_drop * 4294967296.f);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const unsigned long long cc = c + k;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)cc, (uint32_t)(cc >> 32), 0x1f123bb5u, 0u), key);
    m[4 * k + 0] = r.x >= thr; m[4 * k + 1] = r.y >= thr; m[4 * k + 2] = r.z >= thr; m[4 * k + 3] = r.w >= thr;
  }
  const long long base = q * 16;
  if (base + 16 <= n && (reinterpret_cast<uintptr_t>(out) % 16 == 0)) {
    *reinterpret_cast<uint4*>(out + base) = *reinterpret_cast<const uint4*>(m);
  } else {
    for (int k = 0; k < 16 && base + k < n; ++k) out[base + k] = m[k];
  }
}

__global__ void rng_normal_kernel(unsigned long long seed, unsigned long long offset, long long n, float* out) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // 4 outputs per thread
  if (q * 4 >= n) return;
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  const unsigned long long cc = offset + (unsigned long long)q;
  const uint4 r = philox4x32_10(make_uint4((uint32_t)cc, (uint32_t)(cc >> 32), 0x5eedbeefu, 1u), key);
  const float u0 = ((float)r.x + 0.5f) * 2.3283064365386963e-10f, u1 = ((float)r.y + 0.5f) * 2.3283064365386963e-10f;
  const float u2 = ((float)r.z + 0.5f) * 2.3283064365386963e-10f, u3 = ((float)r.w + 0.5f) * 2.3283064365386963e-10f;
  const float ra = sqrtf(-2.f * logf(u0)), rb = sqrtf(-2.f * logf(u2));
  float s0, c0, s1, c1;
  sincosf(6.283185307179586f * u1, &s0, &c0);
  sincosf(6.283185307179586f * u3, &s1, &c1);
  const float v[4] = {ra * c0, ra * s0, rb * c1, rb * s1};
  for (int k = 0; k < 4 && q * 4 + k < n; ++k) out[q * 4 + k] = v[k];
}

// ---------------------------------------------------------------------------------------------
// reparameterisation
// ---------------------------------------------------------------------------------------------
__global__ void reparam_fwd_kernel(const float* mu, const float* ls, const float* eps, long long n, float* z,
                                   void* z_act, int act_dt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = mu[i] + expf(ls[i]) * eps[i];
  if (z) z[i] = v;
  if (z_act) st_act(z_act, i, v, act_dt);
}
__global__ void reparam_bwd_kernel(const void* dz, int dz_dt, const float* ls, const float* eps, long long n,
                                   void* dmu, void* dls, int out_dt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float g = ld_act(dz, i, dz_dt);
  st_act(dmu, i, ld_act(dmu, i, out_dt) + g, out_dt);
  st_act(dls, i, ld_act(dls, i, out_dt) + g * eps[i] * expf(ls[i]), out_dt);
}

// ---------------------------------------------------------------------------------------------
// fused CE (+ accuracy) forward/backward: one warp per row, grid-stride
// ---------------------------------------------------------------------------------------------
__global__ void ce_kernel(IpnCeKl p) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int gw = blockIdx.x * wpb + (threadIdx.x >> 5);
  const int nw = gridDim.x * wpb;
  float ce_acc = 0.f, ok_acc = 0.f;
  const float inv_rows = p.grad_scale / (float)p.rows;
  for (int r = gw; r < p.rows; r += nw) {
    const float* w = p.weights + (long long)r * p.V;
    const int tgt = (int)p.targets[r];
    float mx = -INFINITY;
    int mi = 0x7fffffff;
    for (int c = lane; c < p.V; c += 32) {
      const float v = w[c];
      if (v > mx) { mx = v; mi = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
      if (ov > mx || (ov == mx && oi < mi)) { mx = ov; mi = oi; }
    }
    float se = 0.f;
    for (int c = lane; c < p.V; c += 32) se += expf(w[c] - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    const float lse = mx + logf(se);
    const bool tgt_ok = tgt >= 0 && tgt < p.V;
    if (lane == 0) {
      ce_acc += lse - (tgt_ok ? w[tgt] : 0.f);
      ok_acc += (mi == tgt) ? 1.f : 0.f;
    }
    if (p.dlogits != nullptr) {
      const long long ro = p.use_drow ? map_row(p.drow, r) : (long long)r * p.ld_dl;
      for (int c = lane; c < p.ld_dl; c += 32) {
        float g = 0.f;
        if (c < p.V) {
          const float v = w[c];
          g = (expf(v - lse) - (c == tgt ? 1.f : 0.f)) * inv_rows;
          if (p.relu_mask && !(v > 0.f)) g = 0.f;
        }
        st_act(p.dlogits, ro + c, g, p.dl_dt);
      }
    }
  }
  __shared__ float s_ce[32], s_ok[32];
  if (lane == 0) { s_ce[threadIdx.x >> 5] = ce_acc; s_ok[threadIdx.x >> 5] = ok_acc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < wpb; ++i) { a += s_ce[i]; b += s_ok[i]; }
    atomicAdd(&p.scalars[0], a);
    atomicAdd(&p.scalars[2], b);
  }
}

__global__ void kl_kernel(IpnCeKl p) {
  const long long n = (long long)p.Bz * p.Z;
  float acc = 0.f;
  const float gs = p.grad_scale * p.beta / (float)p.Bz;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float m = p.mu[i], s = p.log_std[i];
    const float e2 = expf(2.f * s);
    acc += 0.5f * (e2 + m * m - 1.f) - s;
    if (p.dmu) st_act(p.dmu, i, gs * m, p.dz_dt);
    if (p.dls) st_act(p.dls, i, gs * (e2 - 1.f), p.dz_dt);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float s_acc[32];
  if ((threadIdx.x & 31) == 0) s_acc[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) a += s_acc[i];
    atomicAdd(&p.scalars[1], a);
  }
}

// ---------------------------------------------------------------------------------------------
// Adam (flat arena) + NaN guard
// ---------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr_bc1, float inv_sqrt_bc2, float b1, float b2,
                            float eps, float gscale, int* nan_flag) {
  const long long n4 = n / 4;
  bool bad = false;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pa = reinterpret_cast<float*>(&pp);
    const float* ga = reinterpret_cast<const float*>(&gg);
    float* ma = reinterpret_cast<float*>(&mm);
    float* va = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gr = ga[k] * gscale;
      ma[k] = b1 * ma[k] + (1.f - b1) * gr;
      va[k] = b2 * va[k] + (1.f - b2) * gr * gr;
      pa[k] -= lr_bc1 * ma[k] / (sqrtf(va[k]) * inv_sqrt_bc2 + eps);
      bad |= (pa[k] != pa[k]);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (long long i = n4 * 4; i < n; ++i) {
      const float gr = g[i] * gscale;
      m[i] = b1 * m[i] + (1.f - b1) * gr;
      v[i] = b2 * v[i] + (1.f - b2) * gr * gr;
      p[i] -= lr_bc1 * m[i] / (sqrtf(v[i]) * inv_sqrt_bc2 + eps);
      bad |= (p[i] != p[i]);
    }
  }
  if (bad && nan_flag) *nan_flag = 1;
}

// ---------------------------------------------------------------------------------------------
// bf16 weight pack (table driven), column sums, dtype conversion
// ---------------------------------------------------------------------------------------------
__global__ void pack_bf16_kernel(const IpnPackItem* items) {
  const IpnPackItem it = items[blockIdx.y];
  const long long total = (long long)it.rows * it.ld_dst;
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(it.dst);
  // fast path (every large matrix): 8 elements per thread, two 16-byte loads -> one 16-byte store, 32-bit index math
  if ((it.cols & 7) == 0 && (it.ld_dst & 7) == 0 && (it.ld_src & 3) == 0 && total < (1LL << 31) &&
      (reinterpret_cast<uintptr_t>(it.src) & 15) == 0 && (reinterpret_cast<uintptr_t>(it.dst) & 15) == 0) {
    const unsigned vpr = (unsigned)it.ld_dst >> 3, nvec = (unsigned)(total >> 3), cv = (unsigned)it.cols >> 3;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += gridDim.x * blockDim.x) {
      const unsigned r = i / vpr, c = i - r * vpr;
      uint4 o = make_uint4(0, 0, 0, 0);
      if (c < cv) {
        const float4* sp = reinterpret_cast<const float4*>(it.src + (long long)r * it.ld_src + c * 8);
        const float4 a = sp[0], b = sp[1];
        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&o);
        h2[0] = __floats2bfloat162_rn(a.x, a.y); h2[1] = __floats2bfloat162_rn(a.z, a.w);
        h2[2] = __floats2bfloat162_rn(b.x, b.y); h2[3] = __floats2bfloat162_rn(b.z, b.w);
      }
      reinterpret_cast<uint4*>(dst)[i] = o;
    }
    return;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / it.ld_dst;
    const int c = (int)(i % it.ld_dst);
    dst[i] = __float2bfloat16_rn(c < it.cols ? it.src[r * it.ld_src + c] : 0.f);
  }
}

// column sums: block = 16 column-groups of 8 columns x 16 row lanes; 16-byte loads when aligned
__global__ void colsum_kernel(const void* X, int dt, long long ld, long long rows, int cols, float* out, float* out2,
                              int cols2) {
  __shared__ float sm[16][129];
  const int cg = threadIdx.x & 15, ry = threadIdx.x >> 4;
  const int c0 = blockIdx.x * 128 + cg * 8;
  const long long per = (rows + gridDim.y - 1) / gridDim.y;
  const long long r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  if (c0 < cols) {
    const int nvalid = min(8, cols - c0);
    const bool vec = vec_ok(X, ld, dt) && nvalid == 8;
    long long r = r0 + ry;
    if (vec && dt == IPN_BF16) {
      // four independent 16-byte loads in flight per thread (the loop is otherwise latency bound)
      const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(X);
      for (; r + 48 < r1; r += 64) {
        uint4 u[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) u[j] = *reinterpret_cast<const uint4*>(xb + (r + 16 * j) * ld + c0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u[j]);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 f = __bfloat1622float2(h2[k]);
            acc[2 * k] += f.x;
            acc[2 * k + 1] += f.y;
          }
        }
      }
    }
    for (; r < r1; r += 16) {
      float v[8];
      ld_act_n<8>(X, r * ld + c0, dt, vec, nvalid, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += v[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) sm[ry][cg * 8 + k] = acc[k];
  __syncthreads();
  if (threadIdx.x < 128) {
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c < cols) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) t += sm[i][threadIdx.x];
      atomicAdd(&out[c], t);
      if (out2 != nullptr && c < cols2) atomicAdd(&out2[c], t);
    }
  }
}

__global__ void convert_2d_kernel(const void* src, int sdt, long long lds, void* dst, int ddt, long long ldd,
                                  long long rows, int cols) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const long long r = i / cols;
  const int c = (int)(i % cols);
  st_act(dst, r * ldd + c, ld_act(src, r * lds + c, sdt), ddt);
}


__global__ void gather_cols_kernel(const float* table, int E, const int* idx, long long idx_stride, long long rows,
                                   void* out, int out_dt, long long ld_out, int col0, const float* row_scale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * E) return;
  const long long r = i / E;
  const int e = (int)(i % E);
  const int k = idx[r * idx_stride];
  float v = k >= 0 ? table[(long long)k * E + e] : 0.f;
  if (row_scale != nullptr) v *= row_scale[r];
  st_act(out, r * ld_out + col0 + e, v, out_dt);
}

__global__ void fill_i32_kernel(int* dst, long long n, int value) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = value;
}

// bf16 fast path: one 16-byte vector (8 columns) per thread and slot, fp32 sums, 16-byte store
__global__ void sum_slots_vec_kernel(const __nv_bfloat16* X, long long ld, int slots, long long rows, int cols,
                                     __nv_bfloat16* out, long long ld_out) {
  const int vpr = cols >> 3;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * vpr) return;
  const long long r = i / vpr;
  const int c = (int)(i - r * vpr) << 3;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (int s0 = 0; s0 < slots; s0 += 6) {
    uint4 u[6];
#pragma unroll
    for (int j = 0; j < 6; ++j)
      if (s0 + j < slots) u[j] = *reinterpret_cast<const uint4*>(X + ((long long)(s0 + j) * rows + r) * ld + c);
#pragma unroll
    for (int j = 0; j < 6; ++j)
      if (s0 + j < slots) {
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u[j]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __bfloat1622float2(h2[k]);
          acc[2 * k] += f.x;
          acc[2 * k + 1] += f.y;
        }
      }
  }
  uint4 o;
  __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int k = 0; k < 4; ++k) o2[k] = __floats2bfloat162_rn(acc[2 * k], acc[2 * k + 1]);
  *reinterpret_cast<uint4*>(out + r * ld_out + c) = o;
}

__global__ void sum_slots_kernel(const void* X, int dt, long long ld, int slots, long long rows, int cols, void* out,
                                 long long ld_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const long long r = i / cols;
  const int c = (int)(i % cols);
  float acc = 0.f;
  for (int s = 0; s < slots; ++s) acc += ld_act(X, ((long long)s * rows + r) * ld + c, dt);
  st_act(out, r * ld_out + c, acc, dt);
}

__global__ void dlogits_relayout_kernel(const float* dw, const float* w, int B, int V, void* out, int out_dt,
                                        long long ld_out, IpnRowMap map, int use_map) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over (row_out, c)
  if (i >= 24LL * B * ld_out) return;
  const long long ro = i / ld_out;
  const int c = (int)(i % ld_out);
  const int j = (int)(ro / (4LL * B));
  const int rem = (int)(ro % (4LL * B));
  const int ib = rem / B, b = rem % B;
  const int t = 6 * ib + j;
  float g = 0.f;
  if (c < V) {
    const long long src = (use_map ? map_row(map, b) : (long long)b * 24 * V) + (long long)t * V + c;
    g = (w[src] > 0.f) ? dw[src] : 0.f;
  }
  st_act(out, i, g, out_dt);
}

}  // namespace ipn

static inline long long imin(long long a, long long b) { return a < b ? a : b; }
using namespace ipn;
#define STREAM reinterpret_cast<cudaStream_t>(stream_)

extern "C" {

int ipn_tokens_time_major(const long long* tok64, int B, int T, int V, int* out32, int* range_flag, void* stream_) {
  IPN_PROPAGATE(ensure_device());
  ProfScope prof("tokens", 0.0, (double)(12.0 * (double)B * T), STREAM);
  IPN_REQUIRE(tok64 && out32 && B > 0 && T > 0, IPN_ERR_ARG, "tokens_time_major: bad args");
  tokens_time_major_kernel<<<cdiv((long long)B * T, 256), 256, 0, STREAM>>>(tok64, B, T, V, out32, range_flag);
  IPN_LAUNCH_CHECK();
  return IPN_OK;
}

int ipn_dec_prev_tokens(const long long* tok64, int B, int V, int* out32, int* range_flag, void* stream_) {
  IPN_PROPAGATE(ensure_device());
  ProfScope prof("tokens", 0.0, (double)(12.0 * 24.0 * (double)B), STREAM);
  IPN_REQUIRE(tok64 && out32 && B > 0, IPN_ERR_ARG, "dec_prev_tokens: bad args");
  dec_prev_tokens_kernel<<<cdiv((long long)B * 24, 256), 256, 0, STREAM>>>(tok64, B, V, out32, range_flag);
  IPN_LAUNCH_CHECK();
  return IPN_OK;
}

int ipn_embed_rows(const float* emb, int E, const int* tok, long long rows, void* out, int out_dt, long long ld_out,
                   void* stream_) {
  IPN_PROPAGATE(ensure_device());
  ProfScope prof("embed_rows", 0.0, (double)((double)rows * ld_out * (out_dt == IPN_BF16 ? 2.0 : 4.0)), STREAM);
  IPN_REQUIRE(emb && tok && out && rows > 0 && ld_out >= E, IPN_ERR_ARG, "embed_rows: bad args");
  embed_rows_kernel<<<cdiv(rows * ld_out, 256), 256, 0, STREAM>>>(emb, E, tok, rows, out, out_dt, ld_out);
  IPN_LAUNCH_CHECK();
  return IPN_OK;
}

int ipn_embed_grad(const void* dX, int dx_dt, long long ld_dx, const int* tok, long long rows, int E, int V,
                   float* demb, int skip_id, float* dskip, void* stream_) {
  IPN_PROPAGATE(ensure_device());
  ProfScope prof("embed_grad", 0.0, (double)((double)rows * E * (dx_dt == IPN_BF16 ? 2.0 : 4.0)), STREAM);
  IPN_REQUIRE(dX && tok && demb && rows > 0, IPN_ERR_ARG, "embed_grad: bad args");
  const int smem = (V + 1) * E * (int)sizeof(float);
  IPN_REQUIRE(smem <= 48 * 1024, IPN_ERR_ARG, "embed_grad: table too large for shared memory");
  const int blocks = (int)imin(296, (rows + 255) / 256);
  embed_grad_kernel<<<blocks, 256, smem, STREAM>>>(dX, dx_dt, ld_dx, tok, rows, E, V, demb, skip_id, dskip);
  IPN_LAUNCH_CHECK();
  return IPN_OK;
}

int ipn_argmax_rows(const float* logits, int rows, int V, const IpnRowMap* rowmap, int* tok_out,
                    long long* samples_out, const IpnRowMap* samples_map, void* stream_) {
  IPN_PROPAGATE(ensure_device());
  ProfScope prof("argmax_rows", 0.0, (double)((double)rows * V * 4.0), STREAM);
  IPN_REQUIRE(logits && rows > 0 && V > 0, IPN_ERR_ARG, "argmax_rows: bad args");
  IpnRowMap z{1, 1, 0, 0, 0};
  argmax_rows_kernel<<<cdiv((long long)rows * 32, 256), 256, 0, STREAM>>>(
      logits, rows, V, rowmap ? *rowmap : z, rowmap != nullptr, tok_out, samples_out, samples_map ? *samples_map : z,
      samples_map != nullptr);
  IPN_LAUNCH_CHECK();
  return IPN_OK;
}

int ipn_rng_keep_mask(unsigned long long seed, unsigned long long offset, long long n, float p_drop, unsigned char* out,
                      void* stream_) {
  IPN_PROPAGATE(ensure_device());
  ProfScope prof("rng_keep_mask", 0.0, (double)((double)n), STREAM);
  IPN_REQUIRE(out && n > 0 && p_drop >= 0.f && p_drop < 1.f, IPN_ERR_ARG, "rng_keep_mask: bad args");
  rng_keep_mask_kernel<<<cdiv(cdiv(n, 16), 256), 256, 0, STREAM>>>(seed, offset, n, p_drop, out);
  IPN_LAUNCH_CHECK();
  return IPN_OK;
}

int ipn_rng_normal(unsigned long long seed, unsigned long long offset, long long n, float* out, void* stream_) {
  IPN_PROPAGATE(ensure_device());
  ProfScope prof("rng_normal", 0.0, (double)(4.0 * (double)n), STREAM);
  IPN_REQUIRE(out && n > 0, IPN_ERR_ARG, "rng_normal: bad args");
  rng_normal_kernel<<<cdiv(cdiv(n, 4), 256), 256, 0, STREAM>>>(seed, offset, n, out);
  IPN_LAUNCH_CHECK();
  return IPN_OK;
}

int ipn_reparam_fwd(const float* mu, const float* log_std, const float* eps, long long n, float* z, void* z_act,
                    int act_dt, void* stream_) {
  IPN_PROPAGATE(ensure_device());
  IPN_REQUIRE(mu && log_std && eps && n > 0, IPN_ERR_ARG, "reparam_fwd: bad args");
  reparam_fwd_kernel<<<cdiv(n, 256), 256, 0, STREAM>>>(mu, log_std, eps, n, z, z_act, act_dt);
  IPN_LAUNCH_CHECK();
  return IPN_OK;
}

int ipn_reparam_bwd(const void* dz, int dz_dt, const float* log_std, const float* eps, long long n, void* dmu,
                    void* dls, int out_dt, void* stream_) {
  IPN_PROPAGATE(ensure_device());
  IPN_REQUIRE(dz && log_std && eps && dmu && dls && n > 0, IPN_ERR_ARG, "reparam_bwd: bad args");
  reparam_bwd_kernel<<<cdiv(n, 256), 256, 0, STREAM>>>(dz, dz_dt, log_std, eps, n, dmu, dls, out_dt);
  IPN_LAUNCH_CHECK();
  return IPN_OK;
}

int ipn_ce_kl(const IpnCeKl* p, void* stream_) {
  IPN_PROPAGATE(ensure_device());
  ProfScope prof("ce_kl_fwd_bwd", 0.0, (double)((p ? (double)p->rows * p->V * (4.0 + (p->dlogits ? (p->dl_dt == IPN_BF16 ? 2.0 : 4.0) : 0.0)) + (double)p->rows * 8.0 + (double)p->Bz * p->Z * 16.0 : 0.0)), STREAM);
  IPN_REQUIRE(p && p->weights && p->targets && p->scalars && p->rows > 0 && p->V > 0, IPN_ERR_ARG, "ce_kl: bad args");
  IPN_REQUIRE(!p->dlogits || p->ld_dl >= p->V, IPN_ERR_ARG, "ce_kl: ld_dl < V");
  const int blocks = (int)imin(148 * 8, cdiv(p->rows, 8));
  ce_kernel<<<blocks, 256, 0, STREAM>>>(*p);
  IPN_LAUNCH_CHECK();
  if (p->mu && p->log_std) {
    const long long n = (long long)p->Bz * p->Z;
    kl_kernel<<<(int)imin(296, cdiv(n, 256)), 256, 0, STREAM>>>(*p);
    IPN_LAUNCH_CHECK();
  }
  return IPN_OK;
}

int ipn_adam_step(float* p, const float* g, float* m, float* v, long long n, int step, float lr, float beta1,
                  float beta2, float eps, float grad_scale, int* nan_flag, void* stream_) {
  IPN_PROPAGATE(ensure_device());
  ProfScope prof("adam_fused", 0.0, (double)(28.0 * (double)n), STREAM);
  IPN_REQUIRE(p && g && m && v && n > 0 && step >= 1, IPN_ERR_ARG, "adam_step: bad args");
  IPN_REQUIRE(((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16 == 0, IPN_ERR_ALIGN,
              "adam_step: arenas must be 16B aligned");
  const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  const int blocks = (int)imin(148 * 8, cdiv(n / 4 + 1, 256));
  adam_kernel<<<blocks, 256, 0, STREAM>>>(p, g, m, v, n, (float)(lr / bc1), (float)(1.0 / sqrt(bc2)), beta1, beta2, eps,
                                          grad_scale, nan_flag);
  IPN_LAUNCH_CHECK();
  return IPN_OK;
}

int ipn_pack_bf16(const IpnPackItem* items_dev, int n, int max_rows, int max_ld_dst, void* stream_) {
  IPN_PROPAGATE(ensure_device());
  ProfScope prof("pack_bf16", 0.0, (double)(0), STREAM);
  IPN_REQUIRE(items_dev && n > 0, IPN_ERR_ARG, "pack_bf16: bad args");
  const long long total = (long long)max_rows * max_ld_dst;
  dim3 grid((unsigned)imin(96, cdiv(total, 8 * 256)), n);
  pack_bf16_kernel<<<grid, 256, 0, STREAM>>>(items_dev);
  IPN_LAUNCH_CHECK();
  return IPN_OK;
}

int ipn_colsum2(const void* X, int dt, long long ld, long long rows, int cols, float* out, float* out2, int cols2,
                void* stream_) {
  IPN_PROPAGATE(ensure_device());
  ProfScope prof("colsum", 0.0, (double)((double)rows * cols * (dt == IPN_BF16 ? 2.0 : 4.0)), STREAM);
  IPN_REQUIRE(X && out && rows > 0 && cols > 0, IPN_ERR_ARG, "colsum: bad args");
  const int gx = cdiv(cols, 128);
  const long long want = cdiv(8 * 148, gx);
  dim3 grid(gx, (unsigned)imin(cdiv(rows, 64), want < 1 ? 1 : want));
  colsum_kernel<<<grid, 256, 0, STREAM>>>(X, dt, ld, rows, cols, out, out2, cols2);
  IPN_LAUNCH_CHECK();
  return IPN_OK;
}

int ipn_colsum(const void* X, int dt, long long ld, long long rows, int cols, float* out, void* stream_) {
  return ipn_colsum2(X, dt, ld, rows, cols, out, nullptr, 0, stream_);
}

int ipn_convert_2d(const void* src, int src_dt, long long ld_src, void* dst, int dst_dt, long long ld_dst,
                   long long rows, int cols, void* stream_) {
  IPN_PROPAGATE(ensure_device());
  ProfScope prof("convert_2d", 0.0, (double)((double)rows * cols * 6.0), STREAM);
  IPN_REQUIRE(src && dst && rows > 0 && cols > 0, IPN_ERR_ARG, "convert_2d: bad args");
  convert_2d_kernel<<<cdiv(rows * cols, 256), 256, 0, STREAM>>>(src, src_dt, ld_src, dst, dst_dt, ld_dst, rows, cols);
  IPN_LAUNCH_CHECK();
  return IPN_OK;
}

int ipn_gather_cols(const float* table, int E, const int* idx, long long idx_stride, long long rows, void* out,
                    int out_dt, long long ld_out, int col0, const float* row_scale, void* stream_) {
  IPN_PROPAGATE(ensure_device());
  IPN_REQUIRE(table && idx && out && rows > 0 && E > 0 && col0 >= 0 && col0 + E <= ld_out, IPN_ERR_ARG, "gather_cols: bad args");
  gather_cols_kernel<<<cdiv(rows * E, 256), 256, 0, STREAM>>>(table, E, idx, idx_stride, rows, out, out_dt, ld_out, col0,
                                                              row_scale);
  IPN_LAUNCH_CHECK();
  return IPN_OK;
}

int ipn_fill_i32(int* dst, long long n, int value, void* stream_) {
  IPN_PROPAGATE(ensure_device());
  IPN_REQUIRE(dst && n > 0, IPN_ERR_ARG, "fill_i32: bad args");
  fill_i32_kernel<<<cdiv(n, 256), 256, 0, STREAM>>>(dst, n, value);
  IPN_LAUNCH_CHECK();
  return IPN_OK;
}

int ipn_sum_slots(const void* X, int dt, long long ld, int slots, long long rows, int cols, void* out, long long ld_out,
                  void* stream_) {
  IPN_PROPAGATE(ensure_device());
  ProfScope prof("sum_slots", 0.0, (double)((double)(slots + 1) * rows * cols * (dt == IPN_BF16 ? 2.0 : 4.0)), STREAM);
  IPN_REQUIRE(X && out && slots > 0 && rows > 0 && cols > 0, IPN_ERR_ARG, "sum_slots: bad args");
  if (dt == IPN_BF16 && cols % 8 == 0 && ld % 8 == 0 && ld_out % 8 == 0 && reinterpret_cast<uintptr_t>(X) % 16 == 0 &&
      reinterpret_cast<uintptr_t>(out) % 16 == 0) {
    sum_slots_vec_kernel<<<cdiv(rows * (cols / 8), 256), 256, 0, STREAM>>>(reinterpret_cast<const __nv_bfloat16*>(X), ld, slots, rows,
                                                                         cols, reinterpret_cast<__nv_bfloat16*>(out), ld_out);
  } else {
    sum_slots_kernel<<<cdiv(rows * cols, 256), 256, 0, STREAM>>>(X, dt, ld, slots, rows, cols, out, ld_out);
  }
  IPN_LAUNCH_CHECK();
  return IPN_OK;
}

int ipn_dlogits_relayout(const float* dweights, const float* weights, int B, int V, void* out, int out_dt,
                         long long ld_out, void* stream_) {
  IPN_PROPAGATE(ensure_device());
  ProfScope prof("dlogits_relayout", 0.0, (double)(24.0 * B * (8.0 * V + ld_out * (out_dt == IPN_BF16 ? 2.0 : 4.0))), STREAM);
  IPN_REQUIRE(dweights && weights && out && B > 0 && V > 0 && ld_out >= V, IPN_ERR_ARG, "dlogits_relayout: bad args");
  IpnRowMap z{1, 1, 0, 0, 0};
  dlogits_relayout_kernel<<<cdiv(24LL * B * ld_out, 256), 256, 0, STREAM>>>(dweights, weights, B, V, out, out_dt, ld_out, z, 0);
  IPN_LAUNCH_CHECK();
  return IPN_OK;
}

int ipn_dlogits_relayout_mapped(const float* dweights, const float* weights, int B, int V, const IpnRowMap* map,
                                void* out, int out_dt, long long ld_out, void* stream_) {
  IPN_PROPAGATE(ensure_device());
  IPN_REQUIRE(dweights && weights && out && map && B > 0 && V > 0 && ld_out >= V, IPN_ERR_ARG, "dlogits_relayout: bad args");
  dlogits_relayout_kernel<<<cdiv(24LL * B * ld_out, 256), 256, 0, STREAM>>>(dweights, weights, B, V, out, out_dt, ld_out, *map, 1);
  IPN_LAUNCH_CHECK();
  return IPN_OK;
}

}  // extern "C"
