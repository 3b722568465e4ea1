"""ctypes binding of the C-ABI in include/inpaintnet_b200.h (no torch types cross this boundary)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libinpaintnet_b200.so")

OK, ERR_ARG, ERR_ALIGN, ERR_ARCH, ERR_CUDA, ERR_RANGE, ERR_NAN = range(7)
F32, BF16, U8 = 0, 1, 2
ACT_NONE, ACT_SELU, ACT_RELU = 0, 1, 2
CORE_SIMT, CORE_UMMA = 0, 1
MUL_NONE, MUL_SELU_GRAD, MUL_RELU_GRAD, MUL_KEEP_MASK = 0, 1, 2, 3
STORE, ATOMIC_ADD, RMW_ADD = 0, 1, 2

vp, ll, i32, f32 = C.c_void_p, C.c_longlong, C.c_int, C.c_float


class RowMap(C.Structure):
    _fields_ = [("g1", i32), ("g2", i32), ("s1", ll), ("s2", ll), ("s3", ll)]


class GemmSeg(C.Structure):
    _fields_ = [("A", vp), ("lda", ll), ("transA", i32), ("B", vp), ("ldb", ll), ("transB", i32), ("K", i32)]


class Gemm(C.Structure):
    _fields_ = [("core", i32), ("in_dt", i32), ("M", i32), ("N", i32), ("nseg", i32), ("seg", GemmSeg * 2),
                ("out", vp), ("out_dt", i32), ("ld_out", ll), ("use_rowmap", i32), ("rowmap", RowMap),
                ("split_cols", i32), ("split_stride", ll), ("bias", vp), ("act", i32), ("alpha", f32),
                ("mul_src", vp), ("mul_dt", i32), ("ld_mul", ll), ("mul_mode", i32), ("mul_scale", f32),
                ("accumulate", i32), ("split_k", i32), ("max_ctas", i32)]


class GruDir(C.Structure):
    _fields_ = [("w_hh", vp), ("b_hh", vp), ("P", vp), ("ldP", ll), ("P_bcast", i32), ("table", vp),
                ("ld_table", ll), ("tok", vp), ("pvec", vp), ("hseq", vp), ("gates", vp), ("reverse", i32),
                ("y_col0", i32), ("final_col0", i32), ("final_out_dir", vp), ("final_dir_dt", i32), ("ld_final_dir", ll),
                ("P_blocked", i32), ("table_rows", i32)]


class GruInproj(C.Structure):
    _fields_ = [("X", vp), ("ldx", ll), ("rows", ll), ("K", i32), ("w_ih", vp), ("ldw", ll), ("b_ih", vp), ("b_hh", vp),
                ("H", i32), ("out", vp)]


class LstmInproj(C.Structure):
    _fields_ = [("X", vp), ("ldx", ll), ("K", i32), ("w_ih", vp), ("ldw", ll), ("X2", vp), ("ldx2", ll), ("K2", i32),
                ("w_ih2", vp), ("ldw2", ll), ("rows", ll), ("b_ih", vp), ("b_hh", vp), ("H", i32), ("out", vp)]


class GruLayer(C.Structure):
    _fields_ = [("core", i32), ("act_dt", i32), ("T", i32), ("B_total", i32), ("H", i32), ("row0", i32),
                ("nrows", i32), ("s_begin", i32), ("s_end", i32), ("ndir", i32), ("dir", GruDir * 2), ("y", vp),
                ("ld_y", ll), ("mask", vp), ("ld_mask", ll), ("mask_scale", f32), ("final_out", vp),
                ("final_dt", i32), ("ld_final", ll), ("ws", vp), ("ws_bytes", ll), ("gates_blocked", i32)]


class GruBwdDir(C.Structure):
    _fields_ = [("w_hh", vp), ("hseq", vp), ("gates", vp), ("dP", vp), ("dGn", vp), ("dh_n", vp), ("ld_dhn", ll),
                ("dh0", vp), ("dh0_dt", i32), ("ld_dh0", ll), ("dh0_selu", i32), ("reverse", i32), ("y_col0", i32)]


class GruLayerBwd(C.Structure):
    _fields_ = [("core", i32), ("act_dt", i32), ("T", i32), ("B_total", i32), ("H", i32), ("row0", i32),
                ("nrows", i32), ("ndir", i32), ("dir", GruBwdDir * 2), ("dY", vp), ("ld_dy", ll), ("mask", vp),
                ("ld_mask", ll), ("mask_scale", f32), ("dhz_ws", vp), ("ws", vp), ("ws_bytes", ll), ("gates_persist", i32)]


class LstmLayer(C.Structure):
    _fields_ = [("core", i32), ("act_dt", i32), ("T", i32), ("B", i32), ("H", i32), ("w_hh", vp), ("b_hh", vp),
                ("P", vp), ("ldP", ll), ("hseq", vp), ("cseq", vp), ("gates", vp), ("y", vp), ("ld_y", ll),
                ("y_col0", i32), ("y_reverse_time", i32), ("s_begin", i32), ("s_end", i32), ("table", vp),
                ("ld_table", ll), ("tok_scalar", vp), ("P_blocked", i32), ("gates_blocked", i32)]


class LstmLayerBwd(C.Structure):
    _fields_ = [("core", i32), ("act_dt", i32), ("T", i32), ("B", i32), ("H", i32), ("w_hh", vp), ("hseq", vp),
                ("cseq", vp), ("gates", vp), ("dY", vp), ("ld_dy", ll), ("y_col0", i32), ("dP", vp), ("ws", vp),
                ("y_reverse_time", i32), ("gates_persist", i32)]


class CeKl(C.Structure):
    _fields_ = [("weights", vp), ("targets", vp), ("rows", i32), ("V", i32), ("dlogits", vp), ("dl_dt", i32),
                ("ld_dl", ll), ("use_drow", i32), ("drow", RowMap), ("relu_mask", i32), ("grad_scale", f32),
                ("mu", vp), ("log_std", vp), ("Bz", i32), ("Z", i32), ("beta", f32), ("dmu", vp), ("dls", vp),
                ("dz_dt", i32), ("scalars", vp)]


class PackItem(C.Structure):
    _fields_ = [("src", vp), ("ld_src", ll), ("dst", vp), ("ld_dst", ll), ("rows", i32), ("cols", i32)]


class TickDecode(C.Structure):
    _fields_ = [("core", i32), ("act_dt", i32), ("B", i32), ("H", i32), ("V", i32), ("l0", GruDir), ("l1", GruDir),
                ("yt0", vp), ("yt1", vp), ("mask", vp), ("mask_scale", f32), ("w_ih1", vp), ("b_ih1", vp),
                ("Pt1", vp), ("w_v", vp), ("b_v", vp), ("weights", vp), ("samples", vp), ("tokprev", vp),
                ("use_maps", i32), ("wmap", RowMap), ("smap", RowMap), ("gates_blocked", i32), ("ws", vp), ("ws_bytes", ll)]


STRUCTS_IN_ORDER = [RowMap, GemmSeg, Gemm, GruInproj, LstmInproj, GruDir, GruLayer, GruBwdDir, GruLayerBwd, LstmLayer, LstmLayerBwd, CeKl,
                    PackItem, TickDecode]

# name -> (restype, argtypes); every symbol declared in include/inpaintnet_b200.h
SYMBOLS = {
    "ipn_last_error": (C.c_char_p, []),
    "ipn_abi_version": (i32, []),
    "ipn_struct_sizes": (i32, [C.POINTER(i32), i32]),
    "ipn_device_check": (i32, [i32, C.POINTER(i32)]),
    "ipn_launch_count": (ll, []),
    "ipn_prof_enable": (None, [i32]),
    "ipn_dbg_set_timing_buffer": (None, [vp]),
    "ipn_prof_report": (i32, [C.c_char_p, i32]),
    "ipn_gemm": (i32, [C.POINTER(Gemm), vp]),
    "ipn_gru_inproj_blocked": (i32, [C.POINTER(GruInproj), vp]),
    "ipn_gru_layer_fwd": (i32, [C.POINTER(GruLayer), vp]),
    "ipn_gru_layer_bwd": (i32, [C.POINTER(GruLayerBwd), vp]),
    "ipn_gru_layer_fwd_ws_bytes": (ll, [C.POINTER(GruLayer)]),
    "ipn_gru_layer_bwd_ws_bytes": (ll, [C.POINTER(GruLayerBwd)]),
    "ipn_gru_gates_cols": (i32, [i32]),
    "ipn_gru_persist_eligible": (i32, [i32, i32, i32, i32]),
    "ipn_lstm_inproj_blocked": (i32, [C.POINTER(LstmInproj), vp]),
    "ipn_lstm_persist_eligible": (i32, [i32, i32, i32, i32]),
    "ipn_lstm_gates_cols": (i32, [i32, i32]),
    "ipn_lstm_layer_fwd": (i32, [C.POINTER(LstmLayer), vp]),
    "ipn_lstm_layer_bwd": (i32, [C.POINTER(LstmLayerBwd), vp]),
    "ipn_tokens_time_major": (i32, [vp, i32, i32, i32, vp, vp, vp]),
    "ipn_dec_prev_tokens": (i32, [vp, i32, i32, vp, vp, vp]),
    "ipn_embed_rows": (i32, [vp, i32, vp, ll, vp, i32, ll, vp]),
    "ipn_embed_grad": (i32, [vp, i32, ll, vp, ll, i32, i32, vp, i32, vp, vp]),
    "ipn_argmax_rows": (i32, [vp, i32, i32, C.POINTER(RowMap), vp, vp, C.POINTER(RowMap), vp]),
    "ipn_fill_i32": (i32, [vp, ll, i32, vp]),
    "ipn_gather_cols": (i32, [vp, i32, vp, ll, ll, vp, i32, ll, i32, vp, vp]),
    "ipn_sum_slots": (i32, [vp, i32, ll, i32, ll, i32, vp, ll, vp]),
    "ipn_dlogits_relayout": (i32, [vp, vp, i32, i32, vp, i32, ll, vp]),
    "ipn_dlogits_relayout_mapped": (i32, [vp, vp, i32, i32, C.POINTER(RowMap), vp, i32, ll, vp]),
    "ipn_tick_decode_argmax": (i32, [C.POINTER(TickDecode), vp]),
    "ipn_tick_decode_ws_bytes": (ll, [C.POINTER(TickDecode)]),
    "ipn_rng_keep_mask": (i32, [C.c_ulonglong, C.c_ulonglong, ll, f32, vp, vp]),
    "ipn_rng_normal": (i32, [C.c_ulonglong, C.c_ulonglong, ll, vp, vp]),
    "ipn_reparam_fwd": (i32, [vp, vp, vp, ll, vp, vp, i32, vp]),
    "ipn_reparam_bwd": (i32, [vp, i32, vp, vp, ll, vp, vp, i32, vp]),
    "ipn_ce_kl": (i32, [C.POINTER(CeKl), vp]),
    "ipn_adam_step": (i32, [vp, vp, vp, vp, ll, i32, f32, f32, f32, f32, f32, vp, vp]),
    "ipn_pack_bf16": (i32, [vp, i32, i32, i32, vp]),
    "ipn_colsum": (i32, [vp, i32, ll, ll, i32, vp, vp]),
    "ipn_colsum2": (i32, [vp, i32, ll, ll, i32, vp, vp, i32, vp]),
    "ipn_convert_2d": (i32, [vp, i32, ll, vp, i32, ll, ll, i32, vp]),
}

_lib = None


class InpaintNetB200Error(RuntimeError):
    pass


def load():
    """Loads the CUDA library; raises loudly when it has not been built (there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise InpaintNetB200Error(
            f"{LIB_PATH} is missing: build it with `python -m inpaintnet_b200.build` "
            "(inpaintnet_b200 has no CPU or PyTorch fallback path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    n = len(STRUCTS_IN_ORDER)
    buf = (i32 * n)()
    got = lib.ipn_struct_sizes(buf, n)
    if got != n or any(buf[k] != C.sizeof(s) for k, s in enumerate(STRUCTS_IN_ORDER)):
        raise InpaintNetB200Error("ctypes struct layout does not match the C header: "
                                  f"{[buf[k] for k in range(n)]} vs {[C.sizeof(s) for s in STRUCTS_IN_ORDER]}")
    _lib = lib
    return lib


_EXC = {ERR_ARG: AssertionError, ERR_ALIGN: ValueError, ERR_ARCH: InpaintNetB200Error, ERR_CUDA: RuntimeError,
        ERR_RANGE: ValueError, ERR_NAN: ValueError}


def check(status):
    if status != OK:
        msg = load().ipn_last_error().decode("utf-8", "replace")
        raise _EXC.get(status, RuntimeError)(f"inpaintnet_b200 [{status}]: {msg}")
