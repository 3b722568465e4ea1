/*
 * inpaintnet_b200 -- C-ABI of the B200 (sm_100a) implementation of InpaintNet's data-parallel hot
 * path: the batched GRU/LSTM recurrences of MeasureVAE, LatentRNN and AnticipationRNN.
 *
 * The reference (ashispati/InpaintNet) is pure Python and delegates this path to torch library
 * modules; the "FFI" a maintainer would bind is therefore the set of torch ops listed in
 * SURVEY.md section 2.2.  Each entry point below names the reference call site(s) it replaces
 * (paths relative to the reference tree).  Plain pointers + sizes only, no torch types; every
 * pointer is a DEVICE pointer unless the name ends in `_host`; `stream` is a cudaStream_t.
 * All functions are stream-ordered, do not synchronise the host, own no memory across calls
 * (except a cache of TMA descriptors keyed by pointer/shape) and return IPN_OK or an error code
 * whose text is available from ipn_last_error().
 *
 * There is NO CPU fallback: on a host without an sm_100 device every compute entry point
 * returns IPN_ERR_ARCH.
 */
#ifndef INPAINTNET_B200_H_
#define INPAINTNET_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define IPN_ABI_VERSION 3

/* status codes */
#define IPN_OK 0
#define IPN_ERR_ARG 1       /* bad shape / null pointer / unsupported combination  -> AssertionError / ValueError */
#define IPN_ERR_ALIGN 2     /* pointer or leading dimension not aligned for the tensor-core path */
#define IPN_ERR_ARCH 3      /* no sm_100 device */
#define IPN_ERR_CUDA 4      /* CUDA runtime / driver error (launch failure, ...) */
#define IPN_ERR_RANGE 5     /* token id out of range (reference: MeasureVAE/decoder.py:34-45) */
#define IPN_ERR_NAN 6       /* NaN in parameters (reference: MeasureVAE/encoder.py:111-116) */

/* element types of activations / GEMM operands */
#define IPN_F32 0
#define IPN_BF16 1
#define IPN_U8 2

/* activation fused into a GEMM epilogue */
#define IPN_ACT_NONE 0
#define IPN_ACT_SELU 1
#define IPN_ACT_RELU 2

/* GEMM core.  SIMT = fp32 CUDA-core path ("fp32 mode": exact parity bar).  UMMA = tcgen05
 * tensor-core path with TMEM accumulators fed by TMA (bf16 operands, fp32 accumulate). */
#define IPN_CORE_SIMT 0
#define IPN_CORE_UMMA 1

/* epilogue multiplier modes (IpnGemm.mul_mode) */
#define IPN_MUL_NONE 0
#define IPN_MUL_SELU_GRAD 1 /* multiply by d SELU / d pre, computed from the saved SELU OUTPUT in mul_src */
#define IPN_MUL_RELU_GRAD 2 /* multiply by (mul_src > 0) */
#define IPN_MUL_KEEP_MASK 3 /* multiply by u8 keep-mask * mul_scale (dropout) */

/* accumulate modes (IpnGemm.accumulate) */
#define IPN_STORE 0
#define IPN_ATOMIC_ADD 1 /* fp32 out only; required for split_k > 1 */
#define IPN_RMW_ADD 2    /* out += value, non atomic (split_k must be 1) */

const char* ipn_last_error(void);
int ipn_abi_version(void);
/* sizeof() of every descriptor struct, in declaration order (binding self-check); returns the count. */
int ipn_struct_sizes(int* out_host, int n);
/* IPN_OK when device `dev` is sm_100; fills *sm_count when non-null. */
int ipn_device_check(int dev, int* sm_count_host);
/* number of kernels this library has launched since load (bench.py's gpu_launches). */
long long ipn_launch_count(void);

/* Per-kernel-class timing with CUDA events on the launching stream (used by bench.py for the
 * roofline numbers; off by default).  ipn_prof_report synchronises the device and writes one line per
 * kernel class: tag \t launches \t total_ms \t algorithmic_flops \t algorithmic_bytes. */
void ipn_prof_enable(int on);
/* diagnostics: when non-null, every CTA of the GRU forward step kernel writes 8 globaltimer stamps
 * (start, setup done, TMA issued, MMA issued, accumulators ready, first epilogue warp done, all done, -) */
void ipn_dbg_set_timing_buffer(void* dev_ptr);
int ipn_prof_report(char* buf_host, int cap);

/* 3-level affine row map: off = (r / g1) * s1 + ((r % g1) / g2) * s2 + (r % g2) * s3 (elements) */
typedef struct {
  int g1, g2;
  long long s1, s2, s3;
} IpnRowMap;

/* ------------------------------------------------------------------------------------------
 * ipn_gemm: D[M,N] = epilogue( sum_s A_s[M,K_s] . B_s[N,K_s]^T )
 * replaces nn.Linear (+SELU/ReLU) and its autograd (addmm / mm): MeasureVAE/encoder.py:40-52,
 * 130-131; MeasureVAE/decoder.py:335-338,350-358,369-372,403-405,495,499;
 * LatentRNN/latent_rnn.py:83,233; AnticipationRNN/anticipation_rnn_gauss_reg_model.py:134-140,
 * 388-400; and the hoisted all-timestep input projections W_ih.x of every torch.nn.GRU/LSTM call
 * (SURVEY.md section 8(c')).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* A; /* logical [M,K]: transA=0 -> A[m*lda+k] ; transA=1 -> A[k*lda+m] */
  long long lda;
  int transA;
  const void* B; /* logical [N,K]: transB=0 -> B[n*ldb+k] ; transB=1 -> B[k*ldb+n] */
  long long ldb;
  int transB;
  int K;
} IpnGemmSeg;

typedef struct {
  int core;  /* IPN_CORE_* */
  int in_dt; /* operand element type; UMMA requires IPN_BF16 */
  int M, N;
  int nseg; /* 1 or 2 */
  IpnGemmSeg seg[2];
  void* out;
  int out_dt;
  long long ld_out;
  int use_rowmap; /* 0: off = row*ld_out ; 1: off = rowmap(row) */
  IpnRowMap rowmap;
  int split_cols; /* >0: column block q = col / split_cols goes to out + q*split_stride, col %= split_cols */
  long long split_stride;
  const float* bias; /* [N] or null */
  int act;
  float alpha; /* value = alpha * act(acc + bias) * mul */
  const void* mul_src;
  int mul_dt;
  long long ld_mul;
  int mul_mode;
  float mul_scale;
  int accumulate;
  int split_k;
  int max_ctas; /* 0 = one CTA per SM; >0 caps the persistent grid so that the GEMM can share the GPU with a kernel
                   that holds the other SMs (weight-gradient GEMMs on a side stream under a persistent GRU layer kernel) */
} IpnGemm;
int ipn_gemm(const IpnGemm* g, void* stream);

/* Hoisted all-timestep GRU input projection written directly in the layout the persistent layer kernel
 * reads (so no relayout pass is needed): out = blocked bf16 [rows, 3, H] of
 *   r, z: 0.5 * (X W_ih^T + b_ih + b_hh)      n: X W_ih^T + b_ih
 * Pass the result as IpnGruDir.P with P_blocked = 1.  rows % 128 == 0; tcgen05 core (bf16 operands) only. */
typedef struct {
  const void* X; /* [rows, K] bf16, row stride ldx */
  long long ldx;
  long long rows;
  int K;
  const void* w_ih; /* [3H, K] bf16, row stride ldw */
  long long ldw;
  const float* b_ih;
  const float* b_hh;
  int H;
  void* out;
} IpnGruInproj;
int ipn_gru_inproj_blocked(const IpnGruInproj* p, void* stream);

/* The same for an LSTM layer (4 gates [i; f; g; o]) and for an input that is the concatenation of up to two matrices
 * (AnticipationRNN generation stack: [shifted note embedding | constraint output], arnn_model.py:375):
 *   out = blocked bf16 [rows, 4, H] of   i, f, o: 0.5 * (X W_ih^T [+ X2 W_ih2^T] + b_ih + b_hh)     g: the same without the 0.5
 * (the sigmoid gates are evaluated as 0.5 tanh(0.5 x) + 0.5).  Pass the result as IpnLstmLayer.P with P_blocked = 1. */
typedef struct {
  const void* X; /* [rows, K] bf16, row stride ldx */
  long long ldx;
  int K;
  const void* w_ih; /* [4H, K] bf16, row stride ldw */
  long long ldw;
  const void* X2;   /* optional second segment [rows, K2] */
  long long ldx2;
  int K2;
  const void* w_ih2; /* [4H, K2] */
  long long ldw2;
  long long rows;
  const float* b_ih;
  const float* b_hh;
  int H;
  void* out;
} IpnLstmInproj;
int ipn_lstm_inproj_blocked(const IpnLstmInproj* p, void* stream);

/* ------------------------------------------------------------------------------------------
 * GRU layer (all directions of ONE layer), forward.  replaces torch.nn.GRU.forward:
 * MeasureVAE/encoder.py:125; MeasureVAE/decoder.py:470,498; LatentRNN/latent_rnn.py:188,190,
 * 231,249.  Gate row order [r; z; n] (torch).  The input projection is NOT computed here: it is
 * supplied as any sum of (a) a hoisted matrix P, (b) a row gathered from `table` by token id,
 * (c) a constant vector `pvec`; b_ih must already be folded into one of them.
 * Rows: a "slot" holds B_total rows; the call processes rows [row0, row0+nrows) of each slot.
 * hseq: [(T+1)*B_total, H]. forward dir: slot 0 = h0, slot t+1 = h_t. reverse dir: slot T = h0,
 * slot t = h_t.  gates: opaque buffer of T*B_total*ipn_gru_gates_cols(H) elements saved for backward
 * (nullable): (r, z, n, W_hn h + b_hn [, h_prev]); its internal layout is private to the fwd/bwd pair.
 * ws: optional workspace of ipn_gru_layer_fwd_ws_bytes() bytes.  When it is supplied and the shape is
 * eligible (tcgen05 core, H % 64 == 0, H <= 512, B_total % 128 == 0, full row and step range) the whole
 * layer runs as ONE persistent kernel (all timesteps, W_hh streamed from L2, h_t handed from step to step
 * inside the CTA); otherwise one fused kernel is launched per timestep.
 * y: layer output rows t*B_total+b, this direction at columns [y_col0, y_col0+H), multiplied by
 * keep-mask*mask_scale when mask != null (inter-layer dropout, train mode).
 * final_out: h after the last processed step, at columns [final_col0, +H) of row b (nullable).
 * Steps s in [s_begin, s_end) are processed (time t = s, or T-1-s for the reverse direction).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* w_hh; /* [3H,H] in_dt: bf16 for UMMA, fp32 for SIMT */
  const float* b_hh;
  const void* P; /* act_dt, [T*B_total, ldP]; or [B_total, ldP] reused at every step when P_bcast != 0 */
  long long ldP;
  int P_bcast;
  const float* table;
  long long ld_table;
  const int* tok; /* [T*B_total] int32, time-ordered rows */
  const float* pvec;
  void* hseq;
  void* gates;
  int reverse;
  int y_col0;
  int final_col0;
  void* final_out_dir; /* when non-null: this direction writes its final state here instead of IpnGruLayer.final_out */
  int final_dir_dt;
  long long ld_final_dir;
  int P_blocked; /* P was produced by ipn_gru_inproj_blocked (persistent kernel only; pvec must be null; a table is
                    allowed only in the two-term form P_bcast = 1, table_rows in (0, 128]: P is then ONE blocked
                    [B_total, 3, H] tile set reused at every step -- the per-beat part of the tick GRU's layer-0
                    projection, b_ih and b_hh folded in -- and the table row of tok is added in the epilogue) */
  int table_rows; /* rows of `table` (0 = unknown).  Known and <= 128 with no other input-projection term: the
                     persistent kernel's blocked P is gathered from a folded bf16 copy of the table (HBM write bound) */
} IpnGruDir;

typedef struct {
  int core, act_dt;
  int T, B_total, H;
  int row0, nrows;
  int s_begin, s_end;
  int ndir;
  IpnGruDir dir[2];
  void* y;
  long long ld_y;
  const unsigned char* mask;
  long long ld_mask;
  float mask_scale;
  void* final_out;
  int final_dt;
  long long ld_final;
  void* ws; /* nullable */
  long long ws_bytes;
  int gates_blocked; /* per-step kernels only: write `gates` in the persistent kernels' private layout so that the
                        backward pass can run ipn_gru_layer_bwd with gates_persist = 1 (needs ipn_gru_persist_eligible) */
} IpnGruLayer;
int ipn_gru_layer_fwd(const IpnGruLayer* p, void* stream);
/* bytes of workspace with which ipn_gru_layer_fwd runs the persistent kernel; 0 = not eligible */
long long ipn_gru_layer_fwd_ws_bytes(const IpnGruLayer* p);
/* elements per (timestep, batch row) of the `gates` buffer */
int ipn_gru_gates_cols(int H);
/* 1 when a layer of this shape can use the persistent kernels' saved-gates layout */
int ipn_gru_persist_eligible(int core, int act_dt, int B_total, int H);

/* GRU layer backward (BPTT, reverse-time).  replaces the autograd of torch.nn.GRU reached from
 * utils/trainer.py:150 (loss.backward()).  Produces, time-ordered: dP [T*B_total,3H] (gradient wrt
 * the input projection, = gradient wrt b_ih rows) and dGn [T*B_total,H] (gradient wrt W_hn h + b_hn);
 * the weight gradients are then hoisted GEMMs over all timesteps (ipn_gemm with transA/transB).
 * dY: gradient wrt the layer output y (same layout/mask as forward), nullable.
 * dh_n: fp32 gradient wrt the final hidden [nrows rows, ld_dhn], nullable.
 * dh0: gradient wrt the initial hidden, optionally multiplied by SELU'(h0) (dh0_selu) so that it is
 * directly the gradient wrt the pre-activation of the Linear+SELU that produced h0
 * (MeasureVAE/decoder.py:392-410). */
typedef struct {
  const void* w_hh;
  const void* hseq;
  const void* gates;
  void* dP;
  void* dGn;
  const float* dh_n;
  long long ld_dhn;
  void* dh0;
  int dh0_dt;
  long long ld_dh0;
  int dh0_selu;
  int reverse;
  int y_col0;
} IpnGruBwdDir;

typedef struct {
  int core, act_dt;
  int T, B_total, H;
  int row0, nrows;
  int ndir;
  IpnGruBwdDir dir[2];
  const void* dY;
  long long ld_dy;
  const unsigned char* mask;
  long long ld_mask;
  float mask_scale;
  float* dhz_ws; /* workspace fp32 [ndir * 2 * B_total * H] */
  void* ws;      /* nullable; see ipn_gru_layer_bwd_ws_bytes */
  long long ws_bytes;
  int gates_persist; /* 1 when `gates` was written by the persistent forward kernel (ipn_gru_layer_fwd called with a
                        workspace on an eligible shape), 0 when it was written by the per-step kernels */
} IpnGruLayerBwd;
int ipn_gru_layer_bwd(const IpnGruLayerBwd* p, void* stream);
/* workspace for the persistent backward kernel; 0 = not eligible (the per-step kernels then read either
 * gates layout, selected by gates_persist). */
long long ipn_gru_layer_bwd_ws_bytes(const IpnGruLayerBwd* p);

/* ------------------------------------------------------------------------------------------
 * LSTM layer forward/backward (uni-directional, zero initial state), gate rows [i; f; g; o].
 * replaces torch.nn.LSTM: AnticipationRNN/anticipation_rnn_gauss_reg_model.py:14-39 (called from
 * :237,:294,:324,:382,:470).  P must hold W_ih x + b_ih (+ b_hh may be passed separately).
 * hseq: [(T+1)*B, H] slot 0 = zeros; cseq likewise (fp32 cell state); gates: [T*B, 4H] post-activation.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int core, act_dt;
  int T, B, H;
  const void* w_hh; /* [4H,H] */
  const float* b_hh;
  const void* P;
  long long ldP;
  void* hseq;
  float* cseq;
  void* gates;
  void* y;
  long long ld_y;
  int y_col0;
  int y_reverse_time; /* write y of step s at time slot T-1-s (constraint stack runs on the flipped sequence) */
  int s_begin, s_end; /* steps processed; s_end == 0 means T */
  const float* table; /* optional [*, 4H] rows added to the input projection, selected by *tok_scalar (same row for */
  long long ld_table; /* the whole batch: AnticipationRNN no-teacher-forcing feedback, arnn_model.py:252-256)    */
  const int* tok_scalar;
  int P_blocked; /* P was produced by ipn_lstm_inproj_blocked: the steps [s_begin, s_end) run as ONE persistent cluster
                    kernel (needs ipn_lstm_persist_eligible; table must be null).  W_hh stays resident in the shared
                    memory of a thread-block cluster (one 64-unit gate-column slice per CTA), h_t is exchanged through
                    distributed shared memory once per step.  `gates` then has ipn_lstm_gates_cols(H) columns in a
                    layout private to the fwd/bwd pair (i, f, g, o, c_t), b_hh must already be folded into P, and
                    cseq is only touched at slots s_begin (read) and s_end (written). */
  int gates_blocked; /* per-step kernels only: write `gates` (ipn_lstm_gates_cols(H, 1) columns) in the persistent
                        kernels' private layout, so that ipn_lstm_layer_bwd can run with gates_persist = 1 on a layer
                        whose forward had to run tick by tick (token feedback); needs ipn_lstm_persist_eligible */
} IpnLstmLayer;
int ipn_lstm_layer_fwd(const IpnLstmLayer* p, void* stream);
/* 1 when an LSTM layer of this shape can run the persistent cluster kernels (tcgen05 core, bf16, H = 128 or 256,
 * B % 128 == 0) */
int ipn_lstm_persist_eligible(int core, int act_dt, int B, int H);
/* elements per (timestep, batch row) of `gates`: 4H for the per-step kernels, 5H for the persistent ones */
int ipn_lstm_gates_cols(int H, int persistent);

typedef struct {
  int core, act_dt;
  int T, B, H;
  const void* w_hh;
  const void* hseq;
  const float* cseq;
  const void* gates;
  const void* dY;
  long long ld_dy;
  int y_col0;
  void* dP; /* [T*B,4H] gradient wrt pre-activations */
  float* ws; /* workspace fp32 [3 * B * H] (per-step kernels only) */
  int y_reverse_time;
  int gates_persist; /* 1: `gates` was written by the persistent forward kernel (IpnLstmLayer.P_blocked): the whole
                        reverse-time chain runs as ONE persistent cluster kernel (dY must be non-null, ld_dy % 8 == 0;
                        hseq / cseq / ws are not read) */
} IpnLstmLayerBwd;
int ipn_lstm_layer_bwd(const IpnLstmLayerBwd* p, void* stream);

/* ------------------------------------------------------------------------------------------
 * token / embedding helpers.  replace F.embedding + autograd (MeasureVAE/encoder.py:93-102;
 * MeasureVAE/decoder.py:519) and the int64 batch reshapes of MeasureVAE/vae_trainer.py:42-55.
 * ------------------------------------------------------------------------------------------ */
/* tok64 [B,T] -> out32 [T,B]; sets *flag (device int) to 1 when an id is outside [0,V). */
int ipn_tokens_time_major(const long long* tok64, int B, int T, int V, int* out32, int* range_flag, void* stream);
/* decoder order: out32[j*4B + i*B + b] = (t==0 ? V : tok64[b, t-1]),  t = 6i+j; T fixed to 24 */
int ipn_dec_prev_tokens(const long long* tok64, int B, int V, int* out32, int* range_flag, void* stream);
/* out[r, 0:E] = emb[tok[r], 0:E], zero padded to ld_out columns */
int ipn_embed_rows(const float* emb, int E, const int* tok, long long rows, void* out, int out_dt, long long ld_out,
                   void* stream);
/* demb[tok[r], e] += dX[r, e]  (rows with tok[r] == skip_id go to dskip[e] instead when dskip != null) */
int ipn_embed_grad(const void* dX, int dx_dt, long long ld_dx, const int* tok, long long rows, int E, int V,
                   float* demb, int skip_id, float* dskip, void* stream);
/* first-maximum (lowest index) argmax over V columns of fp32 rows addressed through `rowmap`;
 * writes int32 ids to tok_out[r] (nullable) and int64 ids to samples_out[map2(r)] (nullable).
 * replaces probs.topk(k=1) at MeasureVAE/decoder.py:511 (tie rule documented in DESIGN.md). */
int ipn_argmax_rows(const float* logits, int rows, int V, const IpnRowMap* rowmap, int* tok_out,
                    long long* samples_out, const IpnRowMap* samples_map, void* stream);

/* out[r, col0 + e] = table[idx[r*idx_stride], e] * (row_scale ? row_scale[r] : 1) for e < E; rows with a negative
 * index give zeros.  Builds concatenated embedding inputs (arnn_model.py:437-532) column block by column block. */
int ipn_gather_cols(const float* table, int E, const int* idx, long long idx_stride, long long rows, void* out,
                    int out_dt, long long ld_out, int col0, const float* row_scale, void* stream);
/* i32 fill */
int ipn_fill_i32(int* dst, long long n, int value, void* stream);
/* out[r, c] = sum_s X[s*rows + r, c]  (act dtype in/out): reduces the 6 tick slots of the decoder */
int ipn_sum_slots(const void* X, int dt, long long ld, int slots, long long rows, int cols, void* out, long long ld_out,
                  void* stream);
/* gradient wrt the returned weights tensor (B,24,V) fp32 -> internal decoder row order
 * [(j*4B + i*B + b), ld_out] (t = 6i+j), ReLU mask (weights > 0) applied, zero padded, act dtype. */
int ipn_dlogits_relayout(const float* dweights, const float* weights, int B, int V, void* out, int out_dt,
                         long long ld_out, void* stream);
/* same with the (b, t) element of the source tensors at  map(b) + t*V  instead of (b*24 + t)*V */
int ipn_dlogits_relayout_mapped(const float* dweights, const float* weights, int B, int V, const IpnRowMap* map,
                                void* out, int out_dt, long long ld_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Tick decoder, argmax-feedback mode: the 24 serial ticks of MeasureVAE/decoder.py:473-529 when
 * no teacher forcing is used (always the case for train=False): per tick
 *   tick-GRU layer 0 step (input = table[prev token] + BeatProj) -> dropout -> layer-1 input
 *   projection -> tick-GRU layer 1 step -> ReLU(W_v h + b_v) -> first-max argmax -> next token.
 * Buffers use the decoder row order: slot j (tick within beat) x 4B rows (i*B + b).
 * l0/l1 describe the two layers as for ipn_gru_layer_fwd (T = 6, B_total = 4B); l0.tok must point to
 * `tokprev` whose rows [0,B) the caller has set to V (the x_0 row of the table).
 * With a workspace (ws), bf16, H in {256, 512}, V <= 64, B a multiple of 128 up to 4224, and layer 0 in the
 * two-term form (blocked broadcast P + token table) the whole decode is ONE persistent kernel launch
 * (csrc/tick_persist.cu); `Pt1` is then not written (the layer-1 input product never leaves tensor memory).
 * Otherwise: 5 launches per tick.  Same outputs and saved state either way.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int core, act_dt;
  int B, H, V;
  IpnGruDir l0, l1;
  void* yt0; /* [24B, H] layer-0 output after dropout */
  void* yt1; /* [24B, H] layer-1 output */
  const unsigned char* mask; /* [24B, H] keep mask or null */
  float mask_scale;
  const void* w_ih1; /* [3H, H] in_dt */
  const float* b_ih1;
  void* Pt1; /* [24B, 3H] */
  const void* w_v; /* [V, H] in_dt */
  const float* b_v;
  float* weights;     /* (B, 24, V) fp32 */
  long long* samples; /* (B, 1, 24) int64 */
  int* tokprev;       /* [24B] int32, decoder order */
  int use_maps;       /* 0: row b of tick t -> weights + (b*24 + t)*V, samples + b*24 + t */
  IpnRowMap wmap;     /* else: weights + wmap(b) + t*V   and   samples + smap(b) + t */
  IpnRowMap smap;
  int gates_blocked;  /* see IpnGruLayer.gates_blocked */
  void* ws;           /* nullable: ipn_tick_decode_ws_bytes() bytes; with it (and an eligible shape) every tick's two  */
  long long ws_bytes; /* GRU steps run the persistent layer kernel on the tick's row window instead of the per-step one */
} IpnTickDecode;
int ipn_tick_decode_argmax(const IpnTickDecode* p, void* stream);
long long ipn_tick_decode_ws_bytes(const IpnTickDecode* p);

/* ------------------------------------------------------------------------------------------
 * random numbers (Philox4x32-10, counter = element index): dropout keep-masks and N(0,1) noise.
 * replace the RNG inside torch.nn.GRU dropout and Normal.rsample (MeasureVAE/measure_vae.py:119).
 * ------------------------------------------------------------------------------------------ */
int ipn_rng_keep_mask(unsigned long long seed, unsigned long long offset, long long n, float p_drop,
                      unsigned char* out, void* stream);
int ipn_rng_normal(unsigned long long seed, unsigned long long offset, long long n, float* out, void* stream);

/* z = mu + exp(log_std) * eps  (MeasureVAE/measure_vae.py:119); writes fp32 z and an act_dt copy. */
int ipn_reparam_fwd(const float* mu, const float* log_std, const float* eps, long long n, float* z, void* z_act,
                    int act_dt, void* stream);
/* dmu += dz ; dls += dz * eps * exp(log_std); dz given in dz_dt. Outputs written in out_dt (RMW add). */
int ipn_reparam_bwd(const void* dz, int dz_dt, const float* log_std, const float* eps, long long n, void* dmu,
                    void* dls, int out_dt, void* stream);

/* ------------------------------------------------------------------------------------------
 * fused loss: mean cross-entropy over ReLU'd logits + beta * KL, forward AND backward in one pass.
 * replaces utils/trainer.py:271-306,344-376 and MeasureVAE/vae_trainer.py:128-139.
 * weights: fp32 [B, T, V] (the tensor the module returns).  targets: int64 [B, T].
 * Row order of dlogits (the internal gradient buffer): given by `drow` map applied to r = b*T+t.
 * dlogits = grad_scale * (softmax - onehot) / (B*T), zero where weights == 0 is NOT applied here
 * (weights are post-ReLU; the ReLU mask is applied by the consumer via IPN_MUL_RELU_GRAD or here
 * when relu_mask != 0).  mu/log_std nullable (no KL term: LatentRNN loss).
 * out_host-visible scalars are accumulated on device in `scalars` (fp32[4]: ce_sum, kl_sum,
 * n_correct, unused) which the caller zeroes first.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const float* weights;
  const long long* targets;
  int rows; /* B*T */
  int V;
  void* dlogits; /* nullable */
  int dl_dt;
  long long ld_dl;
  int use_drow;
  IpnRowMap drow;
  int relu_mask;
  float grad_scale;
  const float* mu;
  const float* log_std;
  int Bz, Z;
  float beta;
  void* dmu; /* nullable; out_dt */
  void* dls;
  int dz_dt;
  float* scalars;
} IpnCeKl;
int ipn_ce_kl(const IpnCeKl* p, void* stream);

/* ------------------------------------------------------------------------------------------
 * fused multi-tensor Adam over a flat parameter arena (torch.optim.Adam defaults,
 * utils/trainer.py:32-35,165-177) + NaN guard (MeasureVAE/encoder.py:111-116): nan_flag (device
 * int) is set when a new parameter value is NaN.  step is 1-based.  grad_scale multiplies g first
 * (1/world_size after a sum all-reduce).
 * ------------------------------------------------------------------------------------------ */
int ipn_adam_step(float* p, const float* g, float* m, float* v, long long n, int step, float lr, float beta1,
                  float beta2, float eps, float grad_scale, int* nan_flag, void* stream);

/* table-driven fp32 -> bf16 shadow copy of 2-D parameter blocks (the "weight pack"). */
typedef struct {
  const float* src;
  long long ld_src;
  void* dst; /* bf16 */
  long long ld_dst;
  int rows, cols; /* columns [cols, ld_dst) of dst are zero filled */
} IpnPackItem;
/* items: DEVICE array of n descriptors */
int ipn_pack_bf16(const IpnPackItem* items_dev, int n, int max_rows, int max_ld_dst, void* stream);

/* out[n] += sum_r X[r, n]  (bias gradients) */
int ipn_colsum(const void* X, int dt, long long ld, long long rows, int cols, float* out, void* stream);
/* same, and additionally out2[n] += the same sums for n < cols2 (b_hh shares its r,z part with b_ih) */
int ipn_colsum2(const void* X, int dt, long long ld, long long rows, int cols, float* out, float* out2, int cols2,
                void* stream);
/* dst[i] = (dt) src[i]; generic dtype conversion, 2-D with leading dims */
int ipn_convert_2d(const void* src, int src_dt, long long ld_src, void* dst, int dst_dt, long long ld_dst,
                   long long rows, int cols, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* INPAINTNET_B200_H_ */
