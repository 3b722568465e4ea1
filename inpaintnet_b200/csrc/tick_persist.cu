// Persistent argmax tick decoder: all serial ticks of MeasureVAE/decoder.py:473-529 (argmax feedback, the path of
// every train=False call and of the non-teacher-forced training steps) in ONE launch.
//
// Per tick t the reference runs, on the B rows of the batch:
//     x      = table[token_{t-1}] + BeatProj                  (decoder.py:496-503: embedding + beat conditioning)
//     h0_t   = GRU layer 0 (x, h0_{t-1});  y0_t = dropout(h0_t)
//     h1_t   = GRU layer 1 (y0_t, h1_{t-1})                   (decoder.py:505: rnn_tick, 2 layers)
//     w_t    = ReLU(W_v h1_t + b_v); token_t = first-max argmax(w_t)   (decoder.py:507-519)
// Every operation is row-local, so a 128-row tile never talks to another tile.  A cluster of 4 CTAs owns one tile
// (32 tiles -> 128 SMs at 4096 measures); CTA `rank` owns hidden units [rank*H/4, (rank+1)*H/4) of BOTH layers.
// W_hh0, W_ih1, W_hh1 and W_v stream from L2 through one TMA ring (the weights of a tick are 4.7 MB, far beyond
// shared memory) together with the 128-row x H activation tile that is the A operand of the phase:
//
//   phase   A operand   B operand          accumulators (TMEM columns, per 64-unit chunk b of this CTA)
//   A(t)    h0_{t-1}    W_hh0 [r z n]      b*256 + [0,192)
//   Bh(t)   h1_{t-1}    W_hh1 [r z n]      b*256 + [64,256)   = [r | z | n_h]
//   Bx(t)   y0_t        W_ih1 [n r z]      b*256 + [0,192)    = [n_x | r | z]   (r, z accumulate on top of Bh)
//   V(t)    h1_t        W_v                [192,256)          (64 logits; every CTA of the cluster computes them)
//
// issued in the order  A(0) Bh(0) | Bx(t) A(t+1) V(t) Bh(t+1) | ...  so that the tensor pipe runs A(t+1) while the
// layer-1 epilogue of tick t and the exchange of h1_t are in flight.  h slices travel between the 4 CTAs as in the
// column split of the layer kernel (gru_persist.cu): TMA store to the time-major history (which the backward pass
// needs anyway), one cluster fence + relaxed remote mbarrier arrives, TMA reload from L2.  The token never leaves the
// SM: each epilogue thread keeps the argmax of its row in a register for the next tick's table gather.
#include "gru_persist.cuh"
#include <stdlib.h>

namespace ipn {

constexpr int TK_CS = 4;                       // CTAs per row tile
constexpr int TK_WST_BYTES = 3 * GP_CH * 128;  // one W stage: 3 gates x 64 units x 64 k bf16 = 24 KB
constexpr int TK_WSTAGES = 3;
constexpr int TK_NBAR = 96;
constexpr int TK_NFLOAT = 5 * 128 + 64;        // b_hn0, 0.5 b_r1, 0.5 b_z1, b_in1, b_hn1 (own units) + b_v

struct TickPersist {
  alignas(64) CUtensorMap tmW0;    // W_hh0 as {k, unit, gate}: box 64 x 64 x 3
  alignas(64) CUtensorMap tmW1;    // W_hh1, same
  alignas(64) CUtensorMap tmWxn;   // W_ih1: box 64 x 64 x 1 (gate n)
  alignas(64) CUtensorMap tmWxrz;  // W_ih1: box 64 x 64 x 2 (gates r, z)
  alignas(64) CUtensorMap tmWv;    // W_v [V, H]: box 64 x 64 (rows >= V read as zero)
  alignas(64) CUtensorMap tmH0;    // layer-0 history [(tpb+1)*B4, H], box 64 x 128 (loads and stores)
  alignas(64) CUtensorMap tmH1;    // layer-1 history
  alignas(64) CUtensorMap tmY0;    // yt0 [tpb*B4, H]: layer-0 output after dropout
  alignas(64) CUtensorMap tmY1;    // yt1
  const __nv_bfloat16* hseq0;
  const __nv_bfloat16* hseq1;
  const uint4* ftab;               // folded token table, bf16 [rows, 3H] (r,z halved)
  const uint4* BPblk;              // blocked [B4, 3, H]: beat projection + biases (r,z halved)
  uint4* gates0;                   // blocked [tpb*B4, 5, H], nullable (both or none)
  uint4* gates1;
  const unsigned char* mask;       // [tpb*B4, H] keep mask, nullable
  const float* b_hh0;
  const float* b_ih1;
  const float* b_hh1;
  const float* b_v;
  float* weights;
  long long* samples;              // nullable
  int* tokprev;                    // [tpb*B4] decoder order; rows of tick 0 set by the caller
  IpnRowMap wmap, smap;
  float mask_scale;
  int B, H, V, nticks, tpb;
  int dbg;                         // diagnostics (IPN_TICK_DBG): 1 no weight loads after tick 0, 2 no MMAs
  unsigned long long* timing;
};

static inline int tk_smem_bytes(int H, bool stream) {
  const int nch = (H / 64) / TK_CS;
  const int ring = stream ? (nch == 2 ? 3 : 5) * (GP_KB_BYTES + nch * TK_WST_BYTES)   // {A k-block + W tiles} stages
                          : (H / 64) * GP_KB_BYTES + TK_WSTAGES * TK_WST_BYTES;        // resident A tile + W ring
  return ring + GP_KB_BYTES + TK_NFLOAT * 4 + 4 * 128 * 8 + TK_NBAR * 8 + 16;
}

enum { PH_A = 0, PH_BH = 1, PH_BX = 2, PH_V = 3 };

// the phase program: step 0 = A(0), 1 = Bh(0), then per tick g: Bx(g), A(g+1), V(g), Bh(g+1)
__device__ __forceinline__ bool tk_phase(int step, int NT, int& kind, int& t) {
  if (step < 2) { kind = step == 0 ? PH_A : PH_BH; t = 0; return true; }
  const int g = (step - 2) >> 2, r = (step - 2) & 3;
  if (g >= NT) return false;
  kind = r == 0 ? PH_BX : r == 1 ? PH_A : r == 2 ? PH_V : PH_BH;
  t = (r & 1) ? g + 1 : g;
  return true;
}

__device__ __forceinline__ uint4 ldg_cg(const void* p) {   // L2 read (the data was written by a TMA store of this kernel)
  uint4 u;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p) : "memory");
  return u;
}

// (256-bit gathers: ldg_nc_32B / ldg_cg_32B in gru_persist.cuh -- a gather costs one LSU wavefront per lane and
// instruction, 32 distinct 128-byte lines per warp, so halving the instruction count halves its time)
__device__ __forceinline__ void ld8f(const float* sp, float (&f)[8]) {   // 32-byte aligned shared-memory vector
  const float4 a = *reinterpret_cast<const float4*>(sp), b = *reinterpret_cast<const float4*>(sp + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// STREAM (default): no resident A tile.  A ring stage holds one 64-wide k-block of the A operand AND the weight tile of
// every chunk of this CTA for that k-block (16 + NCH x 24 KB; 3 stages at H = 512), filled by two producers that
// arrive on the same stage barrier (weights run ahead, the A k-block waits for the exchange), consumed by 4 MMAs per
// chunk.  That is 1.1 kcycles of MMA work per stage and 3.3 k in flight against the ring's ~2.8 kcycle round trip: the
// three products of a tick stream at the SM's L2 -> shared-memory rate instead of the ring's round trip (DESIGN.md
// section 4.3 (ii)).  The two chunk accumulators of a phase complete together and are drained one after the other
// by all 16 epilogue warps.  STREAM = false is the first form of the kernel (resident A tile,
// weights alone in a 3-stage ring, chunks one after the other), kept selectable (IPN_TICK_STREAM=0).
template <bool SAVE, bool STREAM>
__global__ void __launch_bounds__(GP_THREADS, 1) tick_decode_persist_kernel(const __grid_constant__ TickPersist p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int H = p.H, KB = H >> 6, NT = p.nticks, TPB = p.tpb, B = p.B;
  const long long B4 = (long long)(NT / TPB) * B;
  const int NCH = KB / TK_CS;                               // 64-unit chunks of this CTA (1 or 2)
  const int tile_x = (int)blockIdx.x / TK_CS;
  const int rbase = tile_x * GP_ROWS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = ptx::cluster_ctarank();
  const int c_lo = (int)crank * NCH, c_hi = c_lo + NCH;
  const bool masked = p.mask != nullptr;
  // diagnostics: cycle stamps of the first cluster's rank-0 CTA for ticks 8 and 9 (tests/dev/tick_persist_counters.py)
  const bool trace_on = p.timing != nullptr && blockIdx.x == 0;
  auto tr = [&](int role, int t, int ev) {
    if (trace_on && t >= 8 && t < 10) p.timing[65536 / 2 + (role * 2 + (t - 8)) * 32 + ev] = (unsigned long long)clock64();
  };

  const int SST_BYTES = GP_KB_BYTES + NCH * TK_WST_BYTES;   // STREAM: bytes of a ring stage
  const int SSTAGES = NCH == 2 ? 3 : 5;
  const int VKB = NCH == 2 ? 2 : 1;                          // k-blocks of the vocabulary projection per ring stage
  // STREAM: k-blocks are consumed in PRODUCTION order -- every CTA of the cluster stores its first chunk, then its second,
  // so chunks {0,2,4,6} of a new h arrive before {1,3,5,7}: i-th k-block of a phase = chunk kperm(i)
  const int KPB = KB / NCH;                                  // k-blocks per arrival batch
  auto kperm = [&](int i) { return (i % KPB) * NCH + i / KPB; };
  uint8_t* sA = smem;
  uint8_t* sW = STREAM ? smem : sA + KB * GP_KB_BYTES;
  uint8_t* sStg = STREAM ? smem + SSTAGES * SST_BYTES : sW + TK_WSTAGES * TK_WST_BYTES;
  float* sF = reinterpret_cast<float*>(sStg + GP_KB_BYTES);
  float* sBn0 = sF;             // [128] b_hn of layer 0, own units
  float* sHbr = sF + 128;       // 0.5 (b_ir + b_hr) layer 1
  float* sHbz = sF + 256;
  float* sBin = sF + 384;
  float* sBhn = sF + 512;
  float* sBv = sF + 640;        // [64]
  float* sBest = sF + TK_NFLOAT;                            // [4][128]
  int* sIdx = reinterpret_cast<int*>(sBest + 4 * 128);      // [4][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sIdx + 4 * 128);
  uint64_t* w_full = bars;            // [3]
  uint64_t* w_empty = bars + 3;       // [3]
  uint64_t* a_full = bars + 6;        // [8]
  uint64_t* a_free = bars + 14;       // [8]
  uint64_t* tmem_full = bars + 22;    // [2]
  uint64_t* tmem_empty = bars + 24;   // [2]
  uint64_t* lg_full = bars + 26;
  uint64_t* lg_empty = bars + 27;
  uint64_t* stg_ready = bars + 28;
  uint64_t* stg_free = bars + 29;
  uint64_t* s_full = bars + 84;       // [6] STREAM ring: both producers arrived and their bytes landed
  uint64_t* s_empty = bars + 90;      // [6]
  uint64_t* E0 = bars + 32;           // [2][8]: chunk kb of h0_t is in global memory (set = t & 1)
  uint64_t* E1 = bars + 48;           // [2][8]: h1_t
  uint64_t* EY = bars + 64;           // [2][8]: y0_t (only with a dropout mask)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 80);

  if ((ptx::smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("inpaintnet_b200: tick_decode_persist: shared memory base not 1024-byte aligned\n");
    __trap();
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < TK_WSTAGES; ++s) { ptx::mbar_init(&w_full[s], 1); ptx::mbar_init(&w_empty[s], 1); }
      for (int k = 0; k < 8; ++k) { ptx::mbar_init(&a_full[k], 1); ptx::mbar_init(&a_free[k], 1); }
      for (int b = 0; b < 2; ++b) { ptx::mbar_init(&tmem_full[b], 1); ptx::mbar_init(&tmem_empty[b], 16); }
      ptx::mbar_init(lg_full, 1);
      ptx::mbar_init(lg_empty, 16);
      ptx::mbar_init(stg_ready, 16);
      ptx::mbar_init(stg_free, 1);
      for (int k = 0; k < 6; ++k) { ptx::mbar_init(&s_full[k], 2); ptx::mbar_init(&s_empty[k], 1); }
      for (int k = 0; k < 16; ++k) { ptx::mbar_init(&E0[k], 1); ptx::mbar_init(&E1[k], 1); ptx::mbar_init(&EY[k], 1); }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc<512>(tmem_slot);
    ptx::tmem_relinquish();
  } else if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.tmW0); ptx::prefetch_tmap(&p.tmW1); ptx::prefetch_tmap(&p.tmWxn); ptx::prefetch_tmap(&p.tmWxrz);
    ptx::prefetch_tmap(&p.tmWv); ptx::prefetch_tmap(&p.tmH0); ptx::prefetch_tmap(&p.tmH1); ptx::prefetch_tmap(&p.tmY0);
    ptx::prefetch_tmap(&p.tmY1);
  }
  for (int u = threadIdx.x; u < NCH * 64; u += GP_THREADS) {
    const int g = c_lo * 64 + u;   // global hidden unit
    sBn0[u] = p.b_hh0[2 * H + g];
    sHbr[u] = 0.5f * (p.b_ih1[g] + p.b_hh1[g]);
    sHbz[u] = 0.5f * (p.b_ih1[H + g] + p.b_hh1[H + g]);
    sBin[u] = p.b_ih1[2 * H + g];
    sBhn[u] = p.b_hh1[2 * H + g];
  }
  for (int u = threadIdx.x; u < 64; u += GP_THREADS) sBv[u] = u < p.V ? p.b_v[u] : 0.f;
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();   // every CTA's barriers are initialised before any remote arrive
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // A-operand source of a phase: (tensor map id, row coordinate) packed into one key; equal keys = no refill
  auto a_key = [&](int kind, int t) -> unsigned {
    const int i = t / TPB, j = t - i * TPB;
    const unsigned rowc = (unsigned)(i * B + rbase);
    const unsigned B4u = (unsigned)B4;
    switch (kind) {
      case PH_A: return (0u << 28) | ((unsigned)j * B4u + rowc);
      case PH_BH: return (1u << 28) | ((unsigned)j * B4u + rowc);
      case PH_BX: return masked ? ((2u << 28) | ((unsigned)j * B4u + rowc)) : ((0u << 28) | ((unsigned)(j + 1) * B4u + rowc));
      default: return (1u << 28) | ((unsigned)(j + 1) * B4u + rowc);
    }
  };
  // the refill decision of the next executed phase after `step` (false at the end of the program)
  auto next_reloads = [&](int step, unsigned key) -> bool {
    int k2, t2;
    for (int s = step + 1;; ++s) {
      if (!tk_phase(s, NT, k2, t2)) return false;
      if (t2 >= NT) continue;
      return a_key(k2, t2) != key;
    }
  };

  if (warp < 4) {
    ptx::setmaxnreg_dec<56>();
    if (warp == 0) {
      // ===================== weight producer: every phase's B tiles through one ring =====================
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        int kind, t;
        if constexpr (STREAM) {
          for (int step = 0; tk_phase(step, NT, kind, t); ++step) {
            if (t >= NT) continue;
            const int kstep = (kind == PH_V) ? VKB : 1;   // V: VKB k-blocks per stage [A x VKB | W_v x VKB]
            for (int i = 0; i < KB; i += kstep) {
              const int kb = kperm(i);
              ptx::mbar_wait(&s_empty[stage], phase ^ 1);
              uint8_t* dst = smem + stage * SST_BYTES + GP_KB_BYTES;
              if ((p.dbg & 1) && t > 0) {
                ptx::mbar_arrive(&s_full[stage]);
              } else if (kind == PH_V) {
                ptx::mbar_arrive_expect_tx(&s_full[stage], (uint32_t)(VKB * 64 * 128));
                for (int x = 0; x < VKB; ++x)
                  ptx::tma_load_2d(smem + stage * SST_BYTES + VKB * GP_KB_BYTES + x * 64 * 128, &p.tmWv, &s_full[stage], kperm(i + x) * 64, 0);
              } else {
                ptx::mbar_arrive_expect_tx(&s_full[stage], (uint32_t)(NCH * TK_WST_BYTES));
                for (int ci = 0; ci < NCH; ++ci) {
                  const int c = c_lo + ci;
                  uint8_t* d = dst + ci * TK_WST_BYTES;
                  if (kind == PH_A) ptx::tma_load_3d(d, &p.tmW0, &s_full[stage], kb * 64, c * 64, 0);
                  else if (kind == PH_BH) ptx::tma_load_3d(d, &p.tmW1, &s_full[stage], kb * 64, c * 64, 0);
                  else {   // rows [n | r | z]
                    ptx::tma_load_3d(d, &p.tmWxn, &s_full[stage], kb * 64, c * 64, 2);
                    ptx::tma_load_3d(d + 64 * 128, &p.tmWxrz, &s_full[stage], kb * 64, c * 64, 0);
                  }
                }
              }
              if (++stage == SSTAGES) { stage = 0; phase ^= 1; }
            }
          }
        } else
        for (int step = 0; tk_phase(step, NT, kind, t); ++step) {
          if (t >= NT) continue;
          const int nch = kind == PH_V ? 1 : NCH;
          for (int ci = 0; ci < nch; ++ci)
            for (int kb = 0; kb < KB; ++kb) {
              const int c = c_lo + ci;
              ptx::mbar_wait(&w_empty[stage], phase ^ 1);
              uint8_t* dst = sW + stage * TK_WST_BYTES;
              if ((p.dbg & 1) && t > 0) {
                ptx::mbar_arrive(&w_full[stage]);
              } else if (kind == PH_V) {
                ptx::mbar_arrive_expect_tx(&w_full[stage], 64 * 128);
                ptx::tma_load_2d(dst, &p.tmWv, &w_full[stage], kb * 64, 0);
              } else {
                ptx::mbar_arrive_expect_tx(&w_full[stage], TK_WST_BYTES);
                if (kind == PH_A) ptx::tma_load_3d(dst, &p.tmW0, &w_full[stage], kb * 64, c * 64, 0);
                else if (kind == PH_BH) ptx::tma_load_3d(dst, &p.tmW1, &w_full[stage], kb * 64, c * 64, 0);
                else {   // rows [n | r | z]
                  ptx::tma_load_3d(dst, &p.tmWxn, &w_full[stage], kb * 64, c * 64, 2);
                  ptx::tma_load_3d(dst + 64 * 128, &p.tmWxrz, &w_full[stage], kb * 64, c * 64, 0);
                }
              }
              if (++stage == TK_WSTAGES) { stage = 0; phase ^= 1; }
            }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      constexpr uint32_t id192 = ptx::make_idesc_bf16(128, 192, 0, 0);
      constexpr uint32_t id128 = ptx::make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t id64 = ptx::make_idesc_bf16(128, 64, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int n_load = 0;              // refills of the A tile so far
      int n_write[2] = {0, 0};     // fresh accumulations started in TMEM buffer b
      unsigned prev_key = 0xffffffffu;
      const bool tm = p.timing != nullptr && (p.dbg & 16);   // wait counters cost ~150 cycles per clock64: off unless asked
      long long w_te = 0, w_af = 0, w_wf = 0, w_lg = 0, w_is = 0, w_cm = 0, w_kb = 0;
      const long long t_begin = clock64();
      unsigned long long gt0;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));
      const uint64_t descA0 = ptx::make_smem_desc(ptx::smem_u32(sA), 16, 1024);
      const uint64_t descW0 = ptx::make_smem_desc(ptx::smem_u32(sW), 16, 1024);
      int kind, t;
      if constexpr (STREAM) {
        for (int step = 0; tk_phase(step, NT, kind, t); ++step) {
          if (t >= NT) continue;
          const int nch = kind == PH_V ? 1 : NCH;
          const int ttr = kind == PH_BX || kind == PH_V ? t : t - 1;
          if (kind == PH_A || kind == PH_BH) {
            for (int b = 0; b < nch; ++b) {
              wait_acc(&tmem_empty[b], (uint32_t)((n_write[b] & 1) ^ 1), tm, w_te);
              ++n_write[b];
            }
            if (kind == PH_BH && t > 0) wait_acc(lg_empty, (uint32_t)((t - 1) & 1), tm, w_lg);
          } else if (kind == PH_V && t == NT - 1) {
            wait_acc(&tmem_empty[0], (uint32_t)((n_write[0] & 1) ^ 1), tm, w_te);   // last tick: no A(t+1) has waited for it
          }
          ptx::tc_fence_after();
          if (lane == 0) tr(0, ttr, kind * 4);
          const int kstep = (kind == PH_V) ? VKB : 1;
          for (int kb = 0; kb < KB; kb += kstep) {   // kb: position in the phase's k-block sequence (chunk kperm(kb))
            wait_acc(&s_full[stage], phase, tm, w_wf);
            ptx::tc_fence_after();
            if (lane == 0 && kb == 0) tr(0, ttr, kind * 4 + 1);
            const uint64_t da0 = descA0 + (uint64_t)((stage * SST_BYTES) >> 4);
            if (ptx::elect_one()) {
              if (kind == PH_V) {
                for (int x = 0; x < VKB; ++x) {
                  const uint64_t dax = da0 + (uint64_t)((x * GP_KB_BYTES) >> 4);
                  const uint64_t dwx = da0 + (uint64_t)((VKB * GP_KB_BYTES + x * 64 * 128) >> 4);
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk) {
                    if (p.dbg & 2) break;
                    ptx::umma_bf16(tmem_base + 192, dax + (uint64_t)(kk * 2), dwx + (uint64_t)(kk * 2), id64, (kb + x > 0 || kk > 0) ? 1u : 0u);
                  }
                }
              } else
              for (int ci = 0; ci < nch; ++ci) {
                const uint64_t dw0 = da0 + (uint64_t)((GP_KB_BYTES + ci * TK_WST_BYTES) >> 4);
                const uint32_t dbase = tmem_base + (uint32_t)(ci * 256);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                  if (p.dbg & 2) break;
                  const uint64_t da = da0 + (uint64_t)(kk * 2), dw = dw0 + (uint64_t)(kk * 2);
                  const uint32_t acc = (kb > 0 || kk > 0) ? 1u : 0u;
                  if (kind == PH_A) ptx::umma_bf16(dbase, da, dw, id192, acc);
                  else if (kind == PH_BH) ptx::umma_bf16(dbase + 64, da, dw, id192, acc);
                  else if (kind == PH_V) ptx::umma_bf16(tmem_base + 192, da, dw, id64, acc);
                  else if (acc) ptx::umma_bf16(dbase, da, dw, id192, 1u);
                  else {   // first Bx MMA of the chunk: n_x starts fresh, r and z continue on top of Bh
                    ptx::umma_bf16(dbase, da, dw, id64, 0u);
                    ptx::umma_bf16(dbase + 64, da, dw + (uint64_t)((64 * 128) >> 4), id128, 1u);
                  }
                }
              }
              ptx::umma_commit(&s_empty[stage]);
            }
            __syncwarp();
            if (++stage == SSTAGES) { stage = 0; phase ^= 1; }
          }
          if (ptx::elect_one()) {
            if (kind == PH_A || kind == PH_BX) {
              for (int b = 0; b < nch; ++b) ptx::umma_commit(&tmem_full[b]);
            } else if (kind == PH_V) {
              ptx::umma_commit(lg_full);
            }
          }
          __syncwarp();
          if (lane == 0) tr(0, ttr, kind * 4 + 2);
        }
      } else
      for (int step = 0; tk_phase(step, NT, kind, t); ++step) {
        if (t >= NT) continue;
        const unsigned key = a_key(kind, t);
        const bool reload = key != prev_key;
        prev_key = key;
        const bool nxt = next_reloads(step, key);
        const int nch = kind == PH_V ? 1 : NCH;
        for (int ci = 0; ci < nch; ++ci) {
          const int b = ci;
          if (kind == PH_A || kind == PH_BH) {
            wait_acc(&tmem_empty[b], (uint32_t)((n_write[b] & 1) ^ 1), tm, w_te);
            ++n_write[b];
            if (kind == PH_BH && b == 0 && t > 0) wait_acc(lg_empty, (uint32_t)((t - 1) & 1), tm, w_lg);
          } else if (kind == PH_V && t == NT - 1) {
            wait_acc(&tmem_empty[0], (uint32_t)((n_write[0] & 1) ^ 1), tm, w_te);   // last tick: no A(t+1) has waited for it
          }
          ptx::tc_fence_after();
          if (lane == 0 && ci == 0) tr(0, kind == PH_BX || kind == PH_V ? t : t - 1, kind * 4);   // TMEM free: phase starts
          const uint32_t dbase = tmem_base + (uint32_t)(b * 256);
          for (int kb = 0; kb < KB; ++kb) {
            const long long tk0 = tm ? clock64() : 0;
            if (reload && ci == 0) wait_acc(&a_full[kb], (uint32_t)(n_load & 1), tm, w_af);
            wait_acc(&w_full[stage], phase, tm, w_wf);
            if (!(p.dbg & 4)) ptx::tc_fence_after();
            if (lane == 0 && ci == 0 && kb == 0) tr(0, kind == PH_BX || kind == PH_V ? t : t - 1, kind * 4 + 1);   // first operands landed
            const uint64_t da0 = descA0 + (uint64_t)((kb * GP_KB_BYTES) >> 4);
            const uint64_t dw0 = descW0 + (uint64_t)((stage * TK_WST_BYTES) >> 4);
            if (ptx::elect_one()) {
              const long long ti0 = tm ? clock64() : 0;
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                if (p.dbg & 2) break;
                const uint64_t da = da0 + (uint64_t)(kk * 2), dw = dw0 + (uint64_t)(kk * 2);
                const uint32_t acc = (kb > 0 || kk > 0) ? 1u : 0u;
                if (kind == PH_A) ptx::umma_bf16(dbase, da, dw, id192, acc);
                else if (kind == PH_BH) ptx::umma_bf16(dbase + 64, da, dw, id192, acc);
                else if (kind == PH_V) ptx::umma_bf16(tmem_base + 192, da, dw, id64, acc);
                else if (acc) ptx::umma_bf16(dbase, da, dw, id192, 1u);
                else {   // first Bx MMA of the chunk: n_x starts fresh, r and z continue on top of Bh
                  ptx::umma_bf16(dbase, da, dw, id64, 0u);
                  ptx::umma_bf16(dbase + 64, da, dw + (uint64_t)((64 * 128) >> 4), id128, 1u);
                }
              }
              const long long ti1 = tm ? clock64() : 0;
              ptx::umma_commit(&w_empty[stage]);
              if (ci == nch - 1 && nxt) ptx::umma_commit(&a_free[kb]);
              if (tm) { w_is += ti1 - ti0; w_cm += clock64() - ti1; }
            }
            if (!(p.dbg & 8)) __syncwarp();
            if (++stage == TK_WSTAGES) { stage = 0; phase ^= 1; }
            if (tm) w_kb += clock64() - tk0;
          }
          if (ptx::elect_one()) {
            if (kind == PH_A || kind == PH_BX) ptx::umma_commit(&tmem_full[b]);
            else if (kind == PH_V) ptx::umma_commit(lg_full);
          }
          __syncwarp();
          if (lane == 0 && ci == nch - 1) tr(0, kind == PH_BX || kind == PH_V ? t : t - 1, kind * 4 + 2);          // last MMA issued
        }
        if (reload) ++n_load;
      }
      if (p.timing != nullptr && lane == 0) {
        unsigned long long* o = p.timing + (long long)blockIdx.x * 16;
        unsigned long long gt1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
        o[5] = gt1 - gt0;   // ns: with o[0], the SM clock the kernel actually ran at
        o[0] = (unsigned long long)(clock64() - t_begin); o[1] = w_te; o[2] = w_af; o[3] = w_wf; o[4] = w_lg; o[13] = w_is; o[14] = w_cm; o[15] = w_kb;
      }
    } else if (warp == 2) {
      // ===================== store warp: staging tile -> history slot / layer output, then the exchange signal ======
      if (lane == 0) {
        int n_tile = 0;
        auto put = [&](int kind, int t, int c) {   // kind: 0 = y0 tile (masked), 1 = h0 tile, 2 = h1 tile
          const int i = t / TPB, j = t - i * TPB;
          const int r_out = (int)((long long)(j + 1) * B4) + i * B + rbase;   // history slot j+1
          const int r_y = (int)((long long)j * B4) + i * B + rbase;           // time-ordered row of the tick
          ptx::mbar_wait(stg_ready, (uint32_t)(n_tile & 1));
          ++n_tile;
          tr(2, t, kind * 8 + (c - c_lo) * 4);
          uint64_t* e;
          if (kind == 0) {
            ptx::tma_store_2d(&p.tmY0, sStg, c * 64, r_y);
            e = EY;
          } else if (kind == 1) {
            ptx::tma_store_2d(&p.tmH0, sStg, c * 64, r_out);
            if (!masked) ptx::tma_store_2d(&p.tmY0, sStg, c * 64, r_y);
            e = E0;
          } else {
            ptx::tma_store_2d(&p.tmH1, sStg, c * 64, r_out);
            ptx::tma_store_2d(&p.tmY1, sStg, c * 64, r_y);
            e = E1;
          }
          ptx::bulk_commit();
          ptx::bulk_wait_read0();
          ptx::mbar_arrive(stg_free);
          ptx::bulk_wait0();
          tr(2, t, kind * 8 + (c - c_lo) * 4 + 1);
          uint64_t* hs = e + (t & 1) * 8 + c;
          ptx::mbar_arrive(hs);
          ptx::fence_acq_rel_cluster();
          tr(2, t, kind * 8 + (c - c_lo) * 4 + 2);
#pragma unroll
          for (uint32_t pr = 1; pr < (uint32_t)TK_CS; ++pr) ptx::mbar_arrive_remote_relaxed(hs, (crank + pr) % (uint32_t)TK_CS);
          tr(2, t, kind * 8 + (c - c_lo) * 4 + 3);
        };
        auto put_l0 = [&](int t) {
          for (int c = c_lo; c < c_hi; ++c) {
            if (masked) put(0, t, c);
            put(1, t, c);
          }
        };
        put_l0(0);
        for (int t = 0; t < NT; ++t) {
          for (int c = c_lo; c < c_hi; ++c) put(2, t, c);
          if (t + 1 < NT) put_l0(t + 1);
        }
      }
    } else {
      // ===================== A loader =====================
      if (lane == 0) {
        int n_load = 0;
        unsigned prev_key = 0xffffffffu;
        const bool tm = p.timing != nullptr && (p.dbg & 16);
        long long w_fr = 0, w_ex = 0;
        int kind, t;
        int stage = 0;
        uint32_t phase = 0;
        int acquired[3] = {-1, -1, -1};   // last tick whose E0 / E1 / EY exchange this thread has acquired (all chunks)
        for (int step = 0; tk_phase(step, NT, kind, t); ++step) {
          if (t >= NT) continue;
          const unsigned key = a_key(kind, t);
          if (!STREAM && key == prev_key) continue;   // STREAM: the k-blocks of every phase pass through the ring
          prev_key = key;
          const int j = t % TPB;
          // which exchange the source waits for: (barrier array, tick); none for the initial state of a beat
          uint64_t* e = nullptr;
          int te = t;
          if (kind == PH_A) { if (j > 0) { e = E0; te = t - 1; } }
          else if (kind == PH_BH) { if (j > 0) { e = E1; te = t - 1; } }
          else if (kind == PH_BX) e = masked ? EY : E0;
          else e = E1;
          const CUtensorMap* tmap = (key >> 28) == 0 ? &p.tmH0 : (key >> 28) == 1 ? &p.tmH1 : &p.tmY0;
          const int coord = (int)(key & 0x0fffffffu);
          const int ttr = kind == PH_BX || kind == PH_V ? t : t - 1;
          const int kstep = (STREAM && kind == PH_V) ? VKB : 1;
          if constexpr (STREAM) {
            // one acquire + proxy fence per ARRIVAL BATCH (first chunks of the 4 CTAs, then their second chunks) instead
            // of per k-block (each cost ~1.3 kcycles of this thread), and none when an earlier phase already acquired
            // this exchange (A(t+1) after Bx(t) without a mask, Bh(t+1) after V(t))
            const int which = e == E0 ? 0 : e == E1 ? 1 : 2;
            const bool need = e != nullptr && acquired[which] != te;
            for (int i = 0; i < KB; ++i) {
              const int kb = kperm(i);
              if (i % kstep == 0) wait_acc(&s_empty[stage], phase ^ 1, tm, w_fr);
              if (i == 0) tr(3, ttr, kind * 4);
              if (need && i % KPB == 0) {
                const uint32_t par = (uint32_t)((te >> 1) & 1);
                const long long t0 = tm ? clock64() : 0;
                for (int x = 0; x < KPB; ++x) {
                  const int kx = kperm(i + x);
                  uint64_t* hs = e + (te & 1) * 8 + kx;
                  if (kx < c_lo || kx >= c_hi) ptx::mbar_wait_cluster(hs, par);
                  else ptx::mbar_wait(hs, par);
                }
                if (tm) w_ex += clock64() - t0;
                ptx::fence_proxy_async_all();
              }
              if (i % kstep == 0) ptx::mbar_arrive_expect_tx(&s_full[stage], (uint32_t)(kstep * GP_KB_BYTES));
              ptx::tma_load_2d(smem + stage * SST_BYTES + (i % kstep) * GP_KB_BYTES, tmap, &s_full[stage], kb * 64, coord);
              if (i % kstep == kstep - 1 && ++stage == SSTAGES) { stage = 0; phase ^= 1; }
              if (i == 0) tr(3, ttr, kind * 4 + 1);
            }
            if (need) acquired[which] = te;
          } else
          for (int kb = 0; kb < KB; ++kb) {
            if (n_load > 0) wait_acc(&a_free[kb], (uint32_t)((n_load - 1) & 1), tm, w_fr);
            if (kb == 0) tr(3, ttr, kind * 4);
            if (e != nullptr) {
              uint64_t* hs = e + (te & 1) * 8 + kb;
              const uint32_t par = (uint32_t)((te >> 1) & 1);
              const long long t0 = tm ? clock64() : 0;
              if (kb < c_lo || kb >= c_hi) ptx::mbar_wait_cluster(hs, par);
              else ptx::mbar_wait(hs, par);
              if (tm) w_ex += clock64() - t0;
              ptx::fence_proxy_async_all();
            }
            ptx::mbar_arrive_expect_tx(&a_full[kb], GP_KB_BYTES);
            ptx::tma_load_2d(sA + kb * GP_KB_BYTES, tmap, &a_full[kb], kb * 64, coord);
            if (kb == 0) tr(3, ttr, kind * 4 + 1);
          }
          tr(3, ttr, kind * 4 + 2);
          ++n_load;
        }
        if (tm) {
          unsigned long long* o = p.timing + (long long)blockIdx.x * 16;
          o[6] = w_fr; o[7] = w_ex;
        }
      }
    }
  } else {
    // ===================== epilogue warps 4..19: warp = (TMEM lane quadrant q, 16-unit sub-chunk) =====================
    ptx::setmaxnreg_inc<104>();
    const int q = warp & 3;
    const int sub = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const int vpr = H >> 3;
    const uint32_t sStg_u = ptx::smem_u32(sStg);
    const uint32_t sw = (uint32_t)(row & 7);
    const long long astride = (long long)vpr * 128;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool tmt = p.timing != nullptr && threadIdx.x == 128;   // trace stamps of one thread
    const bool tm = tmt && (p.dbg & 16);
    long long w_tf = 0, w_sf = 0, w_lf = 0, w_hp = 0;
    const long long te0 = clock64();
    int n_full[2] = {0, 0};   // tmem_full phases consumed per buffer
    int n_tile = 0;           // staging tiles written
    int tokv = __ldg(p.tokprev + rbase + row);   // token fed to tick 0 (decoder order row of (j = 0, i = 0))

    // write one 128 x 64 tile (this thread: 16 units of its row) and hand it to the store warp
    auto stage_tile = [&](const uint4& v0, const uint4& v1) {
      wait_acc(stg_free, (uint32_t)((n_tile & 1) ^ 1), tm, w_sf);
      ++n_tile;
      st_shared_v4(sStg_u + row * 128 + (((uint32_t)(sub * 2) ^ sw) << 4), v0);
      st_shared_v4(sStg_u + row * 128 + (((uint32_t)(sub * 2 + 1) ^ sw) << 4), v1);
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(stg_ready);
    };
    // h_{t-1} of this thread's 16 units: own slice of the history slot this CTA stored one tick earlier
    auto load_hp = [&](const __nv_bfloat16* hseq, uint64_t* e, int t, int c, uint4 (&hp)[2]) {
      const int i = t / TPB, j = t - i * TPB;
      if (j > 0) wait_acc(e + ((t - 1) & 1) * 8 + c, (uint32_t)(((t - 1) >> 1) & 1), tm, w_hp);
      const __nv_bfloat16* src = hseq + ((long long)j * B4 + (long long)i * B + rbase + row) * H + c * 64 + sub * 16;
      if (p.dbg & 128) { hp[0] = make_uint4(0, 0, 0, 0); hp[1] = hp[0]; return; }
      ldg_cg_32B(src, hp[0], hp[1]);
    };

    auto l0_epilogue = [&](int t) {
      const int i = t / TPB, j = t - i * TPB;
      const long long R0 = (long long)j * B4 + (long long)i * B + rbase;
      const long long rt = R0 >> 7;
      const long long rtb = ((long long)i * B + rbase) >> 7;
      for (int ci = 0; ci < NCH; ++ci) {
        const int c = c_lo + ci, b = ci;
        uint4 pv[3][2], hp[2];
        {
          const uint4* tb = p.ftab + (long long)tokv * (3 * vpr) + c * 8 + sub * 2;
          const uint4* pb = p.BPblk + (rtb * 3 * vpr + c * 8 + sub * 2) * 128 + row;
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            if (p.dbg & 64) { pv[g][0] = make_uint4(0, 0, 0, 0); pv[g][1] = pv[g][0]; continue; }
            ldg_nc_32B(tb + g * vpr, pv[g][0], pv[g][1]);   // 32 contiguous bytes per gate
          }
#pragma unroll
          for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int v = 0; v < 2; ++v) {
              float a[8], bb[8];
              unpack8(pv[g][v], a);
              unpack8((p.dbg & 32) ? make_uint4(0, 0, 0, 0) : ldg_stream(pb + (g * vpr + v) * 128), bb);
#pragma unroll
              for (int k = 0; k < 8; ++k) a[k] += bb[k];
              pv[g][v] = pack8(a);
            }
        }
        load_hp(p.hseq0, E0, t, c, hp);
        if (tmt) tr(1, t - 1, 16 + ci * 4);       // L0(t) is traced with tick t-1 (it follows V(t-1))
        uint2 mk[2] = {make_uint2(0x01010101u, 0x01010101u), make_uint2(0x01010101u, 0x01010101u)};
        if (masked) {
          const unsigned char* mp = p.mask + (R0 + row) * H + c * 64 + sub * 16;
          const uint4 m16 = *reinterpret_cast<const uint4*>(mp);   // 16 keep bytes = this thread's 16 units
          mk[0] = make_uint2(m16.x, m16.y);
          mk[1] = make_uint2(m16.z, m16.w);
        }
        wait_acc(&tmem_full[b], (uint32_t)(n_full[b] & 1), tm, w_tf);
        ++n_full[b];
        ptx::tc_fence_after();
        if (tmt) tr(1, t - 1, 16 + ci * 4 + 1);
        const uint32_t tacc = tlane + (uint32_t)(b * 256 + sub * 16);
        uint4* gp = SAVE ? p.gates0 + ((rt * GP_GATE_ARRAYS) * vpr + c * 8 + sub * 2) * 128 + row : nullptr;
        uint4 hpk[2], ypk[2];
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          float acc[3][8];
#pragma unroll
          for (int g = 0; g < 3; ++g) ptx::tmem_ld8(tacc + (uint32_t)(g * 64 + v * 8), acc[g]);
          float pr[8], pz[8], pn[8], hpf[8], rr[8], zz[8], nn[8], hn[8], hh[8];
          unpack8(pv[0][v], pr);
          unpack8(pv[1][v], pz);
          unpack8(pv[2][v], pn);
          unpack8(hp[v], hpf);
          const float4 b0 = *reinterpret_cast<const float4*>(sBn0 + ci * 64 + sub * 16 + v * 8);
          const float4 b1 = *reinterpret_cast<const float4*>(sBn0 + ci * 64 + sub * 16 + v * 8 + 4);
          const float bn[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          ptx::tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float r = fmaf(0.5f, tanh_fast(fmaf(0.5f, acc[0][k], pr[k])), 0.5f);
            const float z = fmaf(0.5f, tanh_fast(fmaf(0.5f, acc[1][k], pz[k])), 0.5f);
            const float g = acc[2][k] + bn[k];
            const float nv = tanh_fast(fmaf(r, g, pn[k]));
            rr[k] = r; zz[k] = z; hn[k] = g; nn[k] = nv;
            hh[k] = fmaf(z, hpf[k] - nv, nv);
          }
          hpk[v] = pack8(hh);
          if (masked) {   // dropout on the bf16 layer output, as the separate pass of the layer kernel does it
            float y[8];
            unpack8(hpk[v], y);
            const unsigned char* mb = reinterpret_cast<const unsigned char*>(&mk[v]);
#pragma unroll
            for (int k = 0; k < 8; ++k) y[k] = mb[k] ? y[k] * p.mask_scale : 0.f;
            ypk[v] = pack8(y);
          }
          if (SAVE) {
            stg_stream(gp + v * 128, pack8(rr));
            stg_stream(gp + astride + v * 128, pack8(zz));
            stg_stream(gp + 2 * astride + v * 128, pack8(nn));
            stg_stream(gp + 3 * astride + v * 128, pack8(hn));
            stg_stream(gp + 4 * astride + v * 128, hp[v]);
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tmem_empty[b]);
        if (tmt) tr(1, t - 1, 16 + ci * 4 + 2);
        if (masked) stage_tile(ypk[0], ypk[1]);   // y0 first: layer 1 of this tick waits for it
        stage_tile(hpk[0], hpk[1]);
        if (tmt) tr(1, t - 1, 16 + ci * 4 + 3);
      }
    };

    auto l1_epilogue = [&](int t) {
      const int i = t / TPB, j = t - i * TPB;
      const long long R0 = (long long)j * B4 + (long long)i * B + rbase;
      const long long rt = R0 >> 7;
      for (int ci = 0; ci < NCH; ++ci) {
        const int c = c_lo + ci, b = ci;
        uint4 hp[2];
        load_hp(p.hseq1, E1, t, c, hp);
        if (tmt) tr(1, t, ci * 4);
        wait_acc(&tmem_full[b], (uint32_t)(n_full[b] & 1), tm, w_tf);
        ++n_full[b];
        ptx::tc_fence_after();
        if (tmt) tr(1, t, ci * 4 + 1);
        const uint32_t tacc = tlane + (uint32_t)(b * 256 + sub * 16);   // [n_x | r | z | n_h] x 64
        uint4* gp = SAVE ? p.gates1 + ((rt * GP_GATE_ARRAYS) * vpr + c * 8 + sub * 2) * 128 + row : nullptr;
        uint4 hpk[2];
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          float anx[8], ar[8], az[8], anh[8];
          ptx::tmem_ld8(tacc + (uint32_t)(v * 8), anx);
          ptx::tmem_ld8(tacc + (uint32_t)(64 + v * 8), ar);
          ptx::tmem_ld8(tacc + (uint32_t)(128 + v * 8), az);
          ptx::tmem_ld8(tacc + (uint32_t)(192 + v * 8), anh);
          float hpf[8], rr[8], zz[8], nn[8], hn[8], hh[8], br[8], bz[8], bi[8], bh[8];
          unpack8(hp[v], hpf);
          const int u0 = ci * 64 + sub * 16 + v * 8;
          ld8f(sHbr + u0, br);
          ld8f(sHbz + u0, bz);
          ld8f(sBin + u0, bi);
          ld8f(sBhn + u0, bh);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float r = fmaf(0.5f, tanh_fast(fmaf(0.5f, ar[k], br[k])), 0.5f);
            const float z = fmaf(0.5f, tanh_fast(fmaf(0.5f, az[k], bz[k])), 0.5f);
            const float g = anh[k] + bh[k];
            const float nv = tanh_fast(fmaf(r, g, anx[k] + bi[k]));
            rr[k] = r; zz[k] = z; hn[k] = g; nn[k] = nv;
            hh[k] = fmaf(z, hpf[k] - nv, nv);
          }
          hpk[v] = pack8(hh);
          if (SAVE) {
            stg_stream(gp + v * 128, pack8(rr));
            stg_stream(gp + astride + v * 128, pack8(zz));
            stg_stream(gp + 2 * astride + v * 128, pack8(nn));
            stg_stream(gp + 3 * astride + v * 128, pack8(hn));
            stg_stream(gp + 4 * astride + v * 128, hp[v]);
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tmem_empty[b]);
        if (tmt) tr(1, t, ci * 4 + 2);
        stage_tile(hpk[0], hpk[1]);
        if (tmt) tr(1, t, ci * 4 + 3);
      }
    };

    // logits of tick t: ReLU(acc + b_v), written to the API tensor by the cluster's first CTA; first-max argmax
    auto v_epilogue = [&](int t) {
      wait_acc(lg_full, (uint32_t)(t & 1), tm, w_lf);
      ptx::tc_fence_after();
      if (tmt) tr(1, t, 8);
      float a[16];
      ptx::tmem_ld16(tlane + (uint32_t)(192 + sub * 16), a);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(lg_empty);
      float best = -INFINITY;
      int bi = 0x7fffffff;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int col = sub * 16 + k;
        a[k] = fmaxf(a[k] + sBv[col], 0.f);
        if (col < p.V && a[k] > best) { best = a[k]; bi = col; }
      }
      if (crank == 0) {
        float* wp = p.weights + map_row(p.wmap, rbase + row) + (long long)t * p.V + sub * 16;
        if ((p.V & 3) == 0 && (reinterpret_cast<uintptr_t>(wp) & 15) == 0) {
#pragma unroll
          for (int k = 0; k < 16; k += 4)
            if (sub * 16 + k < p.V) *reinterpret_cast<float4*>(wp + k) = make_float4(a[k], a[k + 1], a[k + 2], a[k + 3]);
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k)
            if (sub * 16 + k < p.V) wp[k] = a[k];
        }
      }
      sBest[sub * 128 + row] = best;
      sIdx[sub * 128 + row] = bi;
      ptx::named_bar_sync(1, 512);
      best = sBest[row];
      bi = sIdx[row];
#pragma unroll
      for (int s = 1; s < 4; ++s) {
        const float ov = sBest[s * 128 + row];
        const int oi = sIdx[s * 128 + row];
        if (ov > best) { best = ov; bi = oi; }   // strict >: the lowest index wins ties
      }
      if (bi == 0x7fffffff) bi = 0;
      tokv = bi;
      if (tmt) tr(1, t, 9);
      if (crank == 0 && sub == 0) {
        if (p.samples != nullptr) p.samples[map_row(p.smap, rbase + row) + t] = bi;
        if (t + 1 < NT) {
          const int i2 = (t + 1) / TPB, j2 = (t + 1) - i2 * TPB;
          p.tokprev[(long long)j2 * B4 + (long long)i2 * B + rbase + row] = bi;
        }
      }
    };

    l0_epilogue(0);
    for (int t = 0; t < NT; ++t) {
      l1_epilogue(t);
      v_epilogue(t);
      if (t + 1 < NT) l0_epilogue(t + 1);
    }
    if (tm) {
      unsigned long long* o = p.timing + (long long)blockIdx.x * 16;
      o[8] = (unsigned long long)(clock64() - te0); o[9] = w_tf; o[10] = w_sf; o[11] = w_lf; o[12] = w_hp;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();   // peers' shared memory and barriers stay alive until every CTA of the cluster is done
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
bool persist_enabled();
int gru_fold_table(const float* table, long long ld_table, int rows, const float* b_hh, int H, void* out, cudaStream_t stream);

static bool al16(const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; }

// Row tiles are independent, so a batch beyond one wave (a 4-CTA cluster grid places 132 CTAs = 33 tiles) could run as
// successive waves of clusters inside the same launch -- measured at 32768 rows (an inpainting batch): 11.0 ms, the
// same as the per-tick launches, whose full-width kernels are the better shape once latency no longer matters.  So the
// default takes this path up to one wave; IPN_TICK_PERSIST_MAXB moves the limit.
bool tick_persist_shape_ok(const IpnTickDecode* p) {
  const int on = getenv("IPN_TICK_PERSIST") ? atoi(getenv("IPN_TICK_PERSIST")) : 1;   // read per call: tests switch it
  if (!on || !persist_enabled()) return false;
  if (p->core != IPN_CORE_UMMA || p->act_dt != IPN_BF16) return false;
  if (p->H != 256 && p->H != 512) return false;
  const long long maxb = getenv("IPN_TICK_PERSIST_MAXB") ? atoll(getenv("IPN_TICK_PERSIST_MAXB")) : 33LL * GP_ROWS;
  if (p->B <= 0 || p->B % GP_ROWS != 0 || p->B > maxb) return false;
  if (p->V <= 0 || p->V > 64) return false;
  const IpnGruDir& A = p->l0;
  const IpnGruDir& Bd = p->l1;
  // layer 0: blocked broadcast beat projection + gathered token-table row (the two-term form of the layer kernel)
  if (!(A.table != nullptr && A.tok != nullptr && A.P != nullptr && A.P_blocked && A.P_bcast && A.pvec == nullptr &&
        A.table_rows > 0 && A.table_rows <= 128 && A.ld_table % 4 == 0 && al16(A.table) && al16(A.P)))
    return false;
  if (A.reverse || Bd.reverse || Bd.pvec != nullptr || Bd.table != nullptr) return false;
  if ((A.gates != nullptr) != (Bd.gates != nullptr)) return false;
  if (A.gates != nullptr && (!p->gates_blocked || !al16(A.gates) || !al16(Bd.gates))) return false;
  if (p->mask != nullptr && reinterpret_cast<uintptr_t>(p->mask) % 16 != 0) return false;
  if (!al16(p->yt0) || !al16(p->yt1) || !al16(p->w_ih1) || !al16(p->w_v)) return false;
  if (reinterpret_cast<uintptr_t>(A.hseq) % 32 != 0 || reinterpret_cast<uintptr_t>(Bd.hseq) % 32 != 0) return false;   // 256-bit h_prev loads
  if (A.tok != p->tokprev) return false;
  if (7LL * 4 * p->B >= (1LL << 28)) return false;
  return true;
}

int tick_persist_decode(const IpnTickDecode* p, void* ws, long long ws_bytes, cudaStream_t stream) {
  const int B = p->B, H = p->H, V = p->V;
  const long long B4 = 4LL * B;
  IPN_REQUIRE(ws != nullptr && reinterpret_cast<uintptr_t>(ws) % 32 == 0 && ws_bytes >= 128LL * 3 * H * 2, IPN_ERR_ARG,
              "tick_persist_decode: workspace too small or not 32-byte aligned");
  TickPersist q;
  memset(&q, 0, sizeof(q));
  q.B = B; q.H = H; q.V = V; q.nticks = 24; q.tpb = 6;
  IPN_PROPAGATE(get_tensor_map_3d(&q.tmW0, p->l0.w_hh, (unsigned long long)H, (unsigned long long)H, 3ULL, H, (long long)H * H, 64, 3));
  IPN_PROPAGATE(get_tensor_map_3d(&q.tmW1, p->l1.w_hh, (unsigned long long)H, (unsigned long long)H, 3ULL, H, (long long)H * H, 64, 3));
  IPN_PROPAGATE(get_tensor_map_3d(&q.tmWxn, p->w_ih1, (unsigned long long)H, (unsigned long long)H, 3ULL, H, (long long)H * H, 64, 1));
  IPN_PROPAGATE(get_tensor_map_3d(&q.tmWxrz, p->w_ih1, (unsigned long long)H, (unsigned long long)H, 3ULL, H, (long long)H * H, 64, 2));
  IPN_PROPAGATE(get_tensor_map(&q.tmWv, p->w_v, (unsigned long long)H, (unsigned long long)V, H, 64));
  IPN_PROPAGATE(get_tensor_map(&q.tmH0, p->l0.hseq, (unsigned long long)H, (unsigned long long)(7 * B4), H, GP_ROWS));
  IPN_PROPAGATE(get_tensor_map(&q.tmH1, p->l1.hseq, (unsigned long long)H, (unsigned long long)(7 * B4), H, GP_ROWS));
  IPN_PROPAGATE(get_tensor_map(&q.tmY0, p->yt0, (unsigned long long)H, (unsigned long long)(6 * B4), H, GP_ROWS));
  IPN_PROPAGATE(get_tensor_map(&q.tmY1, p->yt1, (unsigned long long)H, (unsigned long long)(6 * B4), H, GP_ROWS));
  q.hseq0 = reinterpret_cast<const __nv_bfloat16*>(p->l0.hseq);
  q.hseq1 = reinterpret_cast<const __nv_bfloat16*>(p->l1.hseq);
  // token table folded to bf16 (r, z halved; the biases live in the blocked beat projection)
  IPN_PROPAGATE(gru_fold_table(p->l0.table, p->l0.ld_table, p->l0.table_rows, nullptr, H, ws, stream));
  q.ftab = reinterpret_cast<const uint4*>(ws);
  q.BPblk = reinterpret_cast<const uint4*>(p->l0.P);
  q.gates0 = reinterpret_cast<uint4*>(p->l0.gates);
  q.gates1 = reinterpret_cast<uint4*>(p->l1.gates);
  q.mask = p->mask;
  q.mask_scale = p->mask_scale;
  q.b_hh0 = p->l0.b_hh;
  q.b_ih1 = p->b_ih1;
  q.b_hh1 = p->l1.b_hh;
  q.b_v = p->b_v;
  q.weights = p->weights;
  q.samples = p->samples;
  q.tokprev = p->tokprev;
  const IpnRowMap wdef{1 << 30, 1 << 30, 0, 0, 24LL * V};
  const IpnRowMap sdef{1 << 30, 1 << 30, 0, 0, 24};
  q.wmap = p->use_maps ? p->wmap : wdef;
  q.smap = p->use_maps ? p->smap : sdef;
  q.timing = g_dbg_timing;
  q.dbg = getenv("IPN_TICK_DBG") ? atoi(getenv("IPN_TICK_DBG")) : 0;
  const bool save = p->l0.gates != nullptr;
  const int ntw = B / GP_ROWS;
  const bool stream_form = getenv("IPN_TICK_STREAM") ? atoi(getenv("IPN_TICK_STREAM")) != 0 : true;
  const int smem = tk_smem_bytes(H, stream_form);
  auto launch = [&](auto kern, bool* configured) -> int {
    if (!*configured) {
      const int a = tk_smem_bytes(512, stream_form), b = tk_smem_bytes(256, stream_form);
      IPN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, a > b ? a : b));
      *configured = true;
    }
    // algorithmic work: three H x 3H products and the vocabulary projection per row and tick;
    // bytes: saved gates of both layers, both histories and layer outputs, the logits
    const double rows = 24.0 * B;
    ProfScope prof("tick_decode_persist", 2.0 * rows * (9.0 * H * H + (double)V * H),
                   rows * (H * 2.0 * (4 + (save ? 2 * GP_GATE_ARRAYS : 0)) + V * 4.0), stream);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(ntw * TK_CS, 1, 1);
    cfg.blockDim = dim3(GP_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = TK_CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    IPN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, q));
    IPN_LAUNCH_CHECK();
    return IPN_OK;
  };
  static bool cfgd[4] = {false, false, false, false};
  if (stream_form) {
    if (save) IPN_PROPAGATE(launch(tick_decode_persist_kernel<true, true>, &cfgd[0]));
    else IPN_PROPAGATE(launch(tick_decode_persist_kernel<false, true>, &cfgd[1]));
  } else {
    if (save) IPN_PROPAGATE(launch(tick_decode_persist_kernel<true, false>, &cfgd[2]));
    else IPN_PROPAGATE(launch(tick_decode_persist_kernel<false, false>, &cfgd[3]));
  }
  return IPN_OK;
}

}  // namespace ipn
