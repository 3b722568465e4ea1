"""Thin Python wrappers over the C-ABI.  Device pointers are passed as plain ints (tensor.data_ptr()
plus byte offsets); the caller keeps the owning tensors alive.  PyTorch is only the allocator and
the stream provider here."""
import ctypes as C

import torch

from . import _lib as L
from ._lib import (F32, BF16, U8, ACT_NONE, ACT_SELU, ACT_RELU, CORE_SIMT, CORE_UMMA, MUL_NONE, MUL_SELU_GRAD,
                   MUL_RELU_GRAD, MUL_KEEP_MASK, STORE, ATOMIC_ADD, RMW_ADD)

ES = {F32: 4, BF16: 2, U8: 1}
TORCH_DT = {F32: torch.float32, BF16: torch.bfloat16}


class Precision:
    """Numerical mode of the hot path.
    fp32: CUDA-core GEMMs, fp32 activations (exact-parity mode: 1e-3 / bit-exact argmax bar).
    bf16: tcgen05 tensor-core GEMMs on bf16 operands, fp32 accumulate, fp32 master weights."""

    def __init__(self, name):
        assert name in ("fp32", "bf16"), name
        self.name = name
        self.core = CORE_SIMT if name == "fp32" else CORE_UMMA
        self.act = F32 if name == "fp32" else BF16
        self.tdt = TORCH_DT[self.act]
        self.es = ES[self.act]

    def __repr__(self):
        return f"Precision({self.name})"


def lib():
    return L.load()


def stream():
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t, what="tensor"):
    if not t.is_cuda:
        raise L.InpaintNetB200Error(
            f"{what} is on {t.device}: inpaintnet_b200 runs on sm_100a only and has no CPU fallback")


def ptr(t, elem_off=0):
    return t.data_ptr() + elem_off * t.element_size()


_gemm = L.Gemm()
GEMM_MAX_CTAS = 0   # >0 while GEMMs are being issued on a side stream next to a persistent layer kernel
GEMM_MAIN_MAX_CTAS = 0   # >0: grid cap of every other GEMM (micro-batches pipelined on two streams share the SMs)


def gemm(core, in_dt, M, N, segs, out, out_dt, ld_out, bias=0, act=ACT_NONE, alpha=1.0, rowmap=None,
         split_cols=0, split_stride=0, mul=None, accumulate=STORE, split_k=0):
    """segs: list of (A_ptr, lda, transA, B_ptr, ldb, transB, K).  mul: (ptr, dt, ld, mode, scale)."""
    g = _gemm
    g.core, g.in_dt, g.M, g.N, g.nseg = core, in_dt, M, N, len(segs)
    for i, s in enumerate(segs):
        sg = g.seg[i]
        sg.A, sg.lda, sg.transA, sg.B, sg.ldb, sg.transB, sg.K = s
    g.out, g.out_dt, g.ld_out = out, out_dt, ld_out
    if rowmap is None:
        g.use_rowmap = 0
    else:
        g.use_rowmap = 1
        g.rowmap.g1, g.rowmap.g2, g.rowmap.s1, g.rowmap.s2, g.rowmap.s3 = rowmap
    g.split_cols, g.split_stride = split_cols, split_stride
    g.bias, g.act, g.alpha = bias or None, act, alpha
    if mul is None:
        g.mul_src, g.mul_mode = None, MUL_NONE
    else:
        g.mul_src, g.mul_dt, g.ld_mul, g.mul_mode, g.mul_scale = mul
    g.accumulate, g.split_k, g.max_ctas = accumulate, split_k, GEMM_MAX_CTAS or GEMM_MAIN_MAX_CTAS
    L.check(lib().ipn_gemm(C.byref(g), stream()))


def gates_cols(H):
    """elements per (timestep, row) of the opaque `gates` buffer of ipn_gru_layer_fwd/bwd"""
    return lib().ipn_gru_gates_cols(H)


def persist_eligible(prec, B_total, H):
    return bool(lib().ipn_gru_persist_eligible(prec.core, prec.act, B_total, H))


def _workspace(nbytes):
    """Transient device workspace of a persistent layer kernel (stream-ordered: the caching allocator reuses
    the block only for later work on the same stream)."""
    if nbytes <= 0:
        return None
    return torch.empty(nbytes, dtype=torch.uint8, device="cuda")


def gru_inproj_blocked(X, ldx, rows, K, w_ih, ldw, b_ih, b_hh, H, out):
    """out (blocked bf16 [rows,3,H]) = folded input projection of a GRU layer direction; see ipn_gru_inproj_blocked"""
    q = L.GruInproj()
    q.X, q.ldx, q.rows, q.K, q.w_ih, q.ldw, q.b_ih, q.b_hh, q.H, q.out = X, ldx, rows, K, w_ih, ldw, b_ih, b_hh, H, out
    L.check(lib().ipn_gru_inproj_blocked(C.byref(q), stream()))


def gru_dir(w_hh, b_hh, hseq, gates=0, P=0, ldP=0, P_bcast=0, table=0, ld_table=0, tok=0, pvec=0, reverse=0,
            y_col0=0, final_col0=0, final_out=0, final_dt=F32, ld_final=0, P_blocked=0, table_rows=0):
    d = L.GruDir()
    d.P_blocked, d.table_rows = P_blocked, table_rows
    d.final_out_dir, d.final_dir_dt, d.ld_final_dir = final_out or None, final_dt, ld_final
    d.w_hh, d.b_hh, d.P, d.ldP, d.P_bcast = w_hh, b_hh, P or None, ldP, P_bcast
    d.table, d.ld_table, d.tok, d.pvec = table or None, ld_table, tok or None, pvec or None
    d.hseq, d.gates, d.reverse, d.y_col0, d.final_col0 = hseq, gates or None, reverse, y_col0, final_col0
    return d


def gru_layer_fwd(prec, T, B_total, H, dirs, y=0, ld_y=0, mask=0, ld_mask=0, mask_scale=1.0, final_out=0,
                  final_dt=F32, ld_final=0, row0=0, nrows=None, s_begin=0, s_end=None, persistent=True):
    p = L.GruLayer()
    p.core, p.act_dt, p.T, p.B_total, p.H = prec.core, prec.act, T, B_total, H
    p.row0, p.nrows = row0, B_total if nrows is None else nrows
    p.s_begin, p.s_end = s_begin, T if s_end is None else s_end
    p.ndir = len(dirs)
    for i, d in enumerate(dirs):
        p.dir[i] = d
    p.y, p.ld_y, p.mask, p.ld_mask, p.mask_scale = y or None, ld_y, mask or None, ld_mask, mask_scale
    p.final_out, p.final_dt, p.ld_final = final_out or None, final_dt, ld_final
    ws = _workspace(lib().ipn_gru_layer_fwd_ws_bytes(C.byref(p))) if persistent else None
    p.ws, p.ws_bytes = (ws.data_ptr(), ws.numel()) if ws is not None else (None, 0)
    L.check(lib().ipn_gru_layer_fwd(C.byref(p), stream()))
    return ws is not None


def gru_bwd_dir(w_hh, hseq, gates, dP, dGn, dh_n=0, ld_dhn=0, dh0=0, dh0_dt=F32, ld_dh0=0, dh0_selu=0, reverse=0,
                y_col0=0):
    d = L.GruBwdDir()
    d.w_hh, d.hseq, d.gates, d.dP, d.dGn = w_hh, hseq, gates, dP, dGn
    d.dh_n, d.ld_dhn, d.dh0, d.dh0_dt, d.ld_dh0, d.dh0_selu = dh_n or None, ld_dhn, dh0 or None, dh0_dt, ld_dh0, dh0_selu
    d.reverse, d.y_col0 = reverse, y_col0
    return d


def gru_layer_bwd(prec, T, B_total, H, dirs, dhz_ws, dY=0, ld_dy=0, mask=0, ld_mask=0, mask_scale=1.0, row0=0,
                  nrows=None, persistent=False):
    """persistent: what the matching gru_layer_fwd call returned (the persistent forward kernel writes the
    saved-gates buffer in its own blocked layout)."""
    p = L.GruLayerBwd()
    p.core, p.act_dt, p.T, p.B_total, p.H = prec.core, prec.act, T, B_total, H
    p.row0, p.nrows, p.ndir = row0, B_total if nrows is None else nrows, len(dirs)
    for i, d in enumerate(dirs):
        p.dir[i] = d
    p.dY, p.ld_dy, p.mask, p.ld_mask, p.mask_scale, p.dhz_ws = dY or None, ld_dy, mask or None, ld_mask, mask_scale, dhz_ws
    p.gates_persist = 1 if persistent else 0
    ws = _workspace(lib().ipn_gru_layer_bwd_ws_bytes(C.byref(p))) if persistent else None
    p.ws, p.ws_bytes = (ws.data_ptr(), ws.numel()) if ws is not None else (None, 0)
    L.check(lib().ipn_gru_layer_bwd(C.byref(p), stream()))


def lstm_persist_eligible(prec, B, H):
    return bool(lib().ipn_lstm_persist_eligible(prec.core, prec.act, B, H))


def lstm_gates_cols(H, persistent):
    return lib().ipn_lstm_gates_cols(H, 1 if persistent else 0)


def lstm_inproj_blocked(X, ldx, K, w_ih, ldw, rows, b_ih, b_hh, H, out, X2=0, ldx2=0, K2=0, w_ih2=0, ldw2=0):
    """Blocked bf16 [rows, 4, H] input projection of an LSTM layer (up to two input segments)."""
    p = L.LstmInproj()
    p.X, p.ldx, p.K, p.w_ih, p.ldw = X, ldx, K, w_ih, ldw
    p.X2, p.ldx2, p.K2, p.w_ih2, p.ldw2 = X2 or None, ldx2, K2, w_ih2 or None, ldw2
    p.rows, p.b_ih, p.b_hh, p.H, p.out = rows, b_ih, b_hh, H, out
    L.check(lib().ipn_lstm_inproj_blocked(C.byref(p), stream()))


def lstm_layer_fwd(prec, T, B, H, w_hh, b_hh, P, ldP, hseq, cseq, gates=0, y=0, ld_y=0, y_col0=0, y_reverse_time=0,
                   s_begin=0, s_end=0, table=0, ld_table=0, tok_scalar=0, P_blocked=0, gates_blocked=0):
    p = L.LstmLayer()
    p.P_blocked, p.gates_blocked = P_blocked, gates_blocked
    p.y_reverse_time, p.s_begin, p.s_end = y_reverse_time, s_begin, s_end
    p.table, p.ld_table, p.tok_scalar = table or None, ld_table, tok_scalar or None
    p.core, p.act_dt, p.T, p.B, p.H = prec.core, prec.act, T, B, H
    p.w_hh, p.b_hh, p.P, p.ldP, p.hseq, p.cseq = w_hh, b_hh or None, P, ldP, hseq, cseq
    p.gates, p.y, p.ld_y, p.y_col0 = gates or None, y or None, ld_y, y_col0
    L.check(lib().ipn_lstm_layer_fwd(C.byref(p), stream()))


def lstm_layer_bwd(prec, T, B, H, w_hh, hseq, cseq, gates, dY, ld_dy, y_col0, dP, ws, y_reverse_time=0, gates_persist=0):
    p = L.LstmLayerBwd()
    p.y_reverse_time, p.gates_persist = y_reverse_time, gates_persist
    p.core, p.act_dt, p.T, p.B, p.H = prec.core, prec.act, T, B, H
    p.w_hh, p.hseq, p.cseq, p.gates, p.dY, p.ld_dy, p.y_col0, p.dP, p.ws = w_hh, hseq or None, cseq or None, gates, dY or None, ld_dy, y_col0, dP, ws or None
    L.check(lib().ipn_lstm_layer_bwd(C.byref(p), stream()))


def tick_decode_argmax(prec, B, H, V, l0, l1, yt0, yt1, mask, mask_scale, w_ih1, b_ih1, Pt1, w_v, b_v, weights,
                       samples, tokprev, wmap=None, smap=None, gates_blocked=False):
    p = L.TickDecode()
    p.gates_blocked = 1 if gates_blocked else 0
    if wmap is not None:
        p.use_maps = 1
        p.wmap.g1, p.wmap.g2, p.wmap.s1, p.wmap.s2, p.wmap.s3 = wmap
        p.smap.g1, p.smap.g2, p.smap.s1, p.smap.s2, p.smap.s3 = smap
    p.core, p.act_dt, p.B, p.H, p.V = prec.core, prec.act, B, H, V
    p.l0, p.l1 = l0, l1
    p.yt0, p.yt1, p.mask, p.mask_scale = yt0, yt1, mask or None, mask_scale
    p.w_ih1, p.b_ih1, p.Pt1, p.w_v, p.b_v = w_ih1, b_ih1, Pt1, w_v, b_v
    p.weights, p.samples, p.tokprev = weights, samples or None, tokprev
    ws = _workspace(lib().ipn_tick_decode_ws_bytes(C.byref(p)))
    p.ws, p.ws_bytes = (ws.data_ptr(), ws.numel()) if ws is not None else (None, 0)
    L.check(lib().ipn_tick_decode_argmax(C.byref(p), stream()))


def tokens_time_major(tok64, B, T, V, out32, flag=0):
    L.check(lib().ipn_tokens_time_major(tok64, B, T, V, out32, flag or None, stream()))


def dec_prev_tokens(tok64, B, V, out32, flag=0):
    L.check(lib().ipn_dec_prev_tokens(tok64, B, V, out32, flag or None, stream()))


def embed_rows(emb, E, tok, rows, out, out_dt, ld_out):
    L.check(lib().ipn_embed_rows(emb, E, tok, rows, out, out_dt, ld_out, stream()))


def embed_grad(dX, dx_dt, ld_dx, tok, rows, E, V, demb, skip_id=-1, dskip=0):
    L.check(lib().ipn_embed_grad(dX, dx_dt, ld_dx, tok, rows, E, V, demb, skip_id, dskip or None, stream()))


def gather_cols(table, E, idx, idx_stride, rows, out, out_dt, ld_out, col0, row_scale=0):
    L.check(lib().ipn_gather_cols(table, E, idx, idx_stride, rows, out, out_dt, ld_out, col0, row_scale or None, stream()))


def argmax_rows(logits, rows, V, tok_out=0, samples_out=0):
    L.check(lib().ipn_argmax_rows(logits, rows, V, None, tok_out or None, samples_out or None, None, stream()))


def fill_i32(dst, n, value):
    L.check(lib().ipn_fill_i32(dst, n, value, stream()))


def sum_slots(X, dt, ld, slots, rows, cols, out, ld_out):
    L.check(lib().ipn_sum_slots(X, dt, ld, slots, rows, cols, out, ld_out, stream()))


def dlogits_relayout(dweights, weights, B, V, out, out_dt, ld_out, bmap=None):
    if bmap is None:
        L.check(lib().ipn_dlogits_relayout(dweights, weights, B, V, out, out_dt, ld_out, stream()))
    else:
        m = L.RowMap(*bmap)
        L.check(lib().ipn_dlogits_relayout_mapped(dweights, weights, B, V, C.byref(m), out, out_dt, ld_out, stream()))


def rng_keep_mask(seed, offset, n, p_drop, out):
    L.check(lib().ipn_rng_keep_mask(seed, offset, n, p_drop, out, stream()))


def rng_normal(seed, offset, n, out):
    L.check(lib().ipn_rng_normal(seed, offset, n, out, stream()))


def reparam_fwd(mu, log_std, eps, n, z, z_act, act_dt):
    L.check(lib().ipn_reparam_fwd(mu, log_std, eps, n, z or None, z_act or None, act_dt, stream()))


def reparam_bwd(dz, dz_dt, log_std, eps, n, dmu, dls, out_dt):
    L.check(lib().ipn_reparam_bwd(dz, dz_dt, log_std, eps, n, dmu, dls, out_dt, stream()))


def ce_kl(weights, targets, rows, V, scalars, dlogits=0, dl_dt=F32, ld_dl=0, drow=None, relu_mask=0, grad_scale=1.0,
          mu=0, log_std=0, Bz=0, Z=0, beta=0.0, dmu=0, dls=0, dz_dt=F32):
    p = L.CeKl()
    p.weights, p.targets, p.rows, p.V = weights, targets, rows, V
    p.dlogits, p.dl_dt, p.ld_dl = dlogits or None, dl_dt, ld_dl
    if drow is None:
        p.use_drow = 0
    else:
        p.use_drow = 1
        p.drow.g1, p.drow.g2, p.drow.s1, p.drow.s2, p.drow.s3 = drow
    p.relu_mask, p.grad_scale = relu_mask, grad_scale
    p.mu, p.log_std, p.Bz, p.Z, p.beta = mu or None, log_std or None, Bz, Z, beta
    p.dmu, p.dls, p.dz_dt, p.scalars = dmu or None, dls or None, dz_dt, scalars
    L.check(lib().ipn_ce_kl(C.byref(p), stream()))


def adam_step(p, g, m, v, n, step, lr, beta1, beta2, eps, grad_scale=1.0, nan_flag=0):
    L.check(lib().ipn_adam_step(p, g, m, v, n, step, lr, beta1, beta2, eps, grad_scale, nan_flag or None, stream()))


def pack_bf16(items_dev, n, max_rows, max_ld):
    L.check(lib().ipn_pack_bf16(items_dev, n, max_rows, max_ld, stream()))


def colsum(X, dt, ld, rows, cols, out, out2=0, cols2=0):
    L.check(lib().ipn_colsum2(X, dt, ld, rows, cols, out, out2 or None, cols2, stream()))


def convert_2d(src, src_dt, ld_src, dst, dst_dt, ld_dst, rows, cols):
    L.check(lib().ipn_convert_2d(src, src_dt, ld_src, dst, dst_dt, ld_dst, rows, cols, stream()))


def launch_count():
    return lib().ipn_launch_count()


def prof_enable(on):
    lib().ipn_prof_enable(1 if on else 0)


def prof_report():
    """-> {tag: dict(launches, ms, flops, bytes)}; synchronises the device and clears the records."""
    buf = C.create_string_buffer(1 << 16)
    lib().ipn_prof_report(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        tag, n, ms, fl, by = line.split("\t")
        out[tag] = dict(launches=int(n), ms=float(ms), flops=float(fl), bytes=float(by))
    return out
