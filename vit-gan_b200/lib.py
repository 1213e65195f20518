"""ctypes binding of libvitgan_b200.so (C ABI declared in include/vitgan_b200.h).

The library is the product: if it is missing or fails to load, importing this module raises -- there is
no CPU or PyTorch fallback (the oracle under ``oracle/`` is test infrastructure and is never imported here).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvitgan_b200.so")

# enums (include/vitgan_b200.h)
F32, BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_TANH, ACT_SIN, ACT_SIGMOID, ACT_MUL_DGELU, ACT_MUL_DTANH, ACT_MUL_DSIN, ACT_MUL_DSIGMOID = range(9)
GEMM_AUTO, GEMM_SIMT, GEMM_TCGEN05 = -1, 0, 1
ATTN_DOT, ATTN_L2 = 0, 1

i32, i64, f32, vp = C.c_int, C.c_int64, C.c_float, C.c_void_p


class GemmArgs(C.Structure):
    _fields_ = [
        ("path", i32), ("ab_dtype", i32), ("c_dtype", i32), ("trans_a", i32), ("trans_b", i32),
        ("M", i32), ("N", i32), ("K", i32),
        ("A", vp), ("lda", i64), ("B", vp), ("ldb", i64), ("C", vp), ("ldc", i64),
        ("bias", vp), ("act", i32), ("act_param", f32),
        ("aux", vp), ("ldaux", i64), ("residual", vp), ("ldres", i64), ("c_pre", vp), ("ldpre", i64),
        ("c_row_group", i32), ("res_row_mod", i32), ("res_row_off", i32), ("accumulate", i32),
        ("a_rowsum", vp),
        ("ln_gamma", vp), ("ln_beta", vp), ("ln_out", vp), ("ld_ln", i64), ("ln_mean", vp), ("ln_rstd", vp), ("ln_eps", f32),
    ]


# name -> argtypes; every function returns int (vg_status) unless listed in _OTHER_RESTYPE
SIGNATURES = {
    "vg_version": [],
    "vg_last_error": [],
    "vg_device_is_sm100": [],
    "vg_gemm": [C.POINTER(GemmArgs), vp],
    "vg_gemm_set_trace": [vp],
    "vg_cast_scale": [vp, i32, vp, i32, i64, vp, vp, vp],
    "vg_colsum": [vp, i32, i64, i32, i64, vp, vp, i32, vp, vp],
    "vg_layernorm_fwd": [i32, i64, i32, vp, vp, vp, vp, vp, vp, f32, vp],
    "vg_layernorm_bwd": [i32, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp],
    "vg_layernorm_bwd_partials": [i32, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp],
    "vg_fold_partials": [vp, i32, i32, vp, vp, vp, vp, vp],
    "vg_sln_fwd": [i32, i64, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, f32, vp],
    "vg_sln_bwd": [i32, i64, i64, i32] + [vp] * 18,
    "vg_attention_fwd": [i32, i32, i32, i32, i32, i32, vp, vp, vp, i64, vp, i64, vp, f32, vp],
    "vg_attention_bwd": [i32, i32, i32, i32, i32, i32, vp, vp, vp, i64, vp, vp, i64, vp, vp, vp, vp, i64, f32, vp, vp],
    "vg_pack_pad": [vp, i32, i32, i32, i32, i32, vp, vp, vp, i32, vp],
    "vg_attention_path": [i32, i32, i32, i32, i32, i32],
    "vg_attention_set_trace": [vp],
    "vg_im2col_patches": [i32, i32, i32, i32, i32, vp, vp, vp],
    "vg_col2im_patches": [i32, i32, i32, i32, i32, vp, vp, vp],
    "vg_v1_tokens_fwd": [i32, i32, i32, i32, i32, i32, i32, vp, vp, vp],
    "vg_v1_tokens_bwd": [i32, i32, i32, i32, i32, i32, i32, vp, vp, vp],
    "vg_fill_rows": [i32, i32, i32, i32, i32, vp, vp, vp, vp],
    "vg_embed_bwd_split": [i32, i32, i32, i32, vp, vp, vp, vp, i32, vp],
    "vg_sigma_max": [vp, i32, i32, i32, vp, i32, vp, vp],
    "vg_add_inplace": [i32, vp, vp, i64, vp],
    "vg_broadcast_rows": [i32, vp, i64, i64, vp, i64, vp],
    "vg_act_backward": [i32, i64, vp, vp, i32, f32, vp, vp],
    "vg_softmax_ce": [vp, vp, i32, i32, i32, vp, vp, vp],
    "vg_bce": [vp, vp, i32, i32, vp, vp, vp],
    "vg_denorm_u8": [i32, vp, i64, vp, vp],
    "vg_copy_rows": [i32, i64, i32, vp, i64, vp, i64, vp],
    "vg_adam_step": [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, i32, f32, vp, vp, vp],
    "vg_selftest_tcgen05": [i32, i32, i32, i32, i32, f32, C.POINTER(f32)],
}
_OTHER_RESTYPE = {"vg_last_error": C.c_char_p}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C vit-gan_b200/csrc`). vitgan_b200 has no fallback path without its CUDA library.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header/library mismatch: fail loudly
        fn.argtypes = argtypes
        fn.restype = _OTHER_RESTYPE.get(name, C.c_int)
    return lib


lib = _load()


ERR_SHAPE, ERR_ALIGN, ERR_UNSUPPORTED, ERR_LAUNCH, ERR_ARG = -1, -2, -3, -4, -5      # vg_status (include/vitgan_b200.h)


class VitganError(RuntimeError):
    """`status` is the vg_status code.  `rejected` = the call was refused before anything was launched (shape / alignment /
    unsupported combination): only those may be answered by taking another path of THIS library; launch errors must surface."""

    def __init__(self, message, status=ERR_LAUNCH):
        super().__init__(message)
        self.status = status

    @property
    def rejected(self):
        return self.status in (ERR_SHAPE, ERR_ALIGN, ERR_UNSUPPORTED)


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib.vg_last_error()
        raise VitganError(f"{what or 'libvitgan_b200'} failed (status {rc}): {msg.decode() if msg else ''}", rc)
