"""ctypes binding of libpianobart_b200.so (include/pianobart_b200.h).

The library is the product path: if it is missing or fails to load this module raises -
there is no CPU or PyTorch fallback anywhere in the package.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PIANOBART_B200_LIB: developer override (A/B runs against another build of the same ABI); never a fallback
LIB_PATH = os.environ.get("PIANOBART_B200_LIB") or os.path.join(_HERE, "libpianobart_b200.so")

PB_GEMM_OUT_F32 = 1
PB_GEMM_GELU = 2
PB_GEMM_ATOMIC_ACC = 4
PB_GEMM_RES_F32 = 8
PB_GEMM_AUX_PREACT = 16
PB_GEMM_MUL_DGELU = 32
PB_GEMM_AUX_DGELU = 64
PB_GEMM_MUL_AUX = 128


class GemmDesc(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("b", C.c_void_p), ("c", C.c_void_p),
        ("bias", C.c_void_p), ("residual", C.c_void_p),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("a_mn_major", C.c_int), ("b_mn_major", C.c_int),
        ("lda", C.c_longlong), ("ldb", C.c_longlong), ("ldc", C.c_longlong), ("ldr", C.c_longlong),
        ("batch_h", C.c_int), ("batch_b", C.c_int),
        ("a_stride_h", C.c_longlong), ("a_stride_b", C.c_longlong),
        ("b_stride_h", C.c_longlong), ("b_stride_b", C.c_longlong),
        ("c_stride_h", C.c_longlong), ("c_stride_b", C.c_longlong),
        ("r_stride_h", C.c_longlong), ("r_stride_b", C.c_longlong),
        ("alpha", C.c_float), ("flags", C.c_int), ("split_k", C.c_int), ("causal", C.c_int),
        ("block_n", C.c_int), ("aux", C.c_void_p), ("ldaux", C.c_longlong), ("r_row_mod", C.c_int),
        ("drop_seed", C.c_void_p), ("drop_op", C.c_uint), ("drop_thresh", C.c_uint), ("drop_scale", C.c_float),
        ("cta_group", C.c_int),
    ]


class DropSite(C.Structure):
    _fields_ = [("seed", C.c_void_p), ("op", C.c_uint), ("thresh", C.c_uint), ("scale", C.c_float)]


class AttnDesc(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("o", C.c_void_p), ("dout", C.c_void_p),
        ("dq", C.c_void_p), ("dk", C.c_void_p), ("dv", C.c_void_p),
        ("ldq", C.c_longlong), ("ldk", C.c_longlong), ("ldv", C.c_longlong), ("ldo", C.c_longlong),
        ("lddo", C.c_longlong), ("lddq", C.c_longlong), ("lddk", C.c_longlong), ("lddv", C.c_longlong),
        ("lse", C.c_void_p), ("dvec", C.c_void_p), ("key_keep", C.c_void_p),
        ("B", C.c_int), ("H", C.c_int), ("Sq", C.c_int), ("Sk", C.c_int), ("hd", C.c_int), ("causal", C.c_int),
        ("scale", C.c_float),
    ]


class DecodeLayer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        'wqkv', 'bqkv', 'wo', 'bo', 'ln1_g', 'ln1_b', 'wqc', 'bqc', 'woc', 'boc', 'ln2_g', 'ln2_b', 'w1', 'b1', 'w2', 'b2',
        'ln3_g', 'ln3_b', 'self_k', 'self_v', 'cross_k', 'cross_v')]


class DecodePersistDesc(C.Structure):
    """include/pianobart_b200.h: pb_decode_persist_desc"""
    _fields_ = ([('layer', DecodeLayer * 8), ('n_layers', C.c_int), ('S_enc', C.c_int), ('S_max', C.c_int),
                 ('stop_when_done', C.c_int)] +
                [(n, C.c_void_p) for n in (
                    'emb_table', 'w_in', 'b_in', 'pos_table', 'lne_g', 'lne_b', 'w_heads', 'b_heads', 'enc_keep',
                    't_dev', 'cur_tok', 'result', 'sampled', 'done', 'n_written', 'uniforms', 'forced', 'logits_out',
                    'raw0', 'raw1', 'raw2', 'qkv', 'qc', 'ob', 'f1', 'part', 'logits_ll', 'tok_ll', 'epoch', 'error_flag', 'trace')] +
                [('dbg_flags', C.c_int)])


class DecodeBatchDesc(C.Structure):
    """include/pianobart_b200.h: pb_decode_batch_desc"""
    _fields_ = ([('layer', DecodeLayer * 8), ('n_layers', C.c_int), ('S_enc', C.c_int), ('S_max', C.c_int), ('B', C.c_int)] +
                [(n, C.c_void_p) for n in (
                    'emb_table', 'w_in', 'b_in', 'pos_table', 'lne_g', 'lne_b', 'w_heads', 'b_heads', 'enc_keep',
                    't_dev', 'cur_tok', 'result', 'sampled', 'done', 'n_written', 'uniforms', 'forced', 'logits',
                    'xemb', 'raw0', 'raw1', 'raw2', 'hn', 'qkv', 'qc', 'ob', 'f1', 'stats', 'barrier', 'error_flag', 'trace',
                    'part', 'part_cnt')])


class PBError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PBError(
                "libpianobart_b200.so not built (run `python -m pianobart_b200.build`); "
                "there is no fallback path")
        _lib = C.CDLL(LIB_PATH)
        _lib.pb_last_error.restype = C.c_char_p
        _lib.pb_launch_count.restype = C.c_longlong
        _lib.pb_reset_launch_count.restype = None
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise PBError("%s failed (%d): %s" % (what, rc, lib().pb_last_error().decode()))


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
