"""ctypes binding of libkgeb200.so (include/kgeb200.h).  No CPU fallback: if the library is missing or cannot be
loaded this module raises, and every non-zero status from the C ABI becomes a KgebError carrying kgeb_last_error()."""
from __future__ import annotations

import ctypes as C
import os

# enum values of include/kgeb200.h
DISTMULT, TRANSE_L1, TRANSE_L2, COMPLEX, HOLE, ROTATE, PROTATE, MODE_L1, MODE_L2, HAKE = range(10)
HOLE_DIRECT, HOLE_FFT = 0, 1
FUSED_TRAIN_KINDS = (DISTMULT, TRANSE_L1, TRANSE_L2, COMPLEX)   # kinds the fused train kernel is instantiated for
POINTWISE_HINGE, POINTWISE_LOGISTIC, POINTWISE_SQUARE, PAIRWISE_HINGE, PAIRWISE_LOGISTIC, PAIRWISE_SQUARE, POINTWISE_BCE = range(7)


def is_pairwise(loss_kind):
    return PAIRWISE_HINGE <= loss_kind <= PAIRWISE_SQUARE

PAIRED, ALLPAIRS = 0, 1
SGD, ADAGRAD, LAZY_ADAM, ADAM_DENSE = range(4)
CORRUPT_BOTH, CORRUPT_HEAD, CORRUPT_TAIL, CORRUPT_EITHER = range(4)
SIDE_TAIL, SIDE_HEAD = 0, 1
TC_BF16, TC_BF16X3_EXACT = 0, 1
TC_F16X3_EXACT = TC_BF16X3_EXACT   # the exact mode's planes are fp16 since round 2; the first name is kept
ACT_KINDS = {None: 0, 'linear': 0, 'tanh': 1, 'relu': 2, 'sigmoid': 3}   # Keras activation names served by the literal Dense
COMM_ID_BYTES = 128
TIMED_RANK_TC, TIMED_RANK_REFINE, TIMED_RANK_TILE, TIMED_TRAIN, TIMED_OPT, TIMED_AUX = range(6)

LOSS_KINDS = {  # reference names, utils.py:321-330
    "pointwize_hinge": POINTWISE_HINGE, "pointwize_logistic": POINTWISE_LOGISTIC,
    "pointwize_square_loss": POINTWISE_SQUARE, "pairwize_hinge": PAIRWISE_HINGE,
    "pairwize_logistic": PAIRWISE_LOGISTIC, "pairwize_square_loss": PAIRWISE_SQUARE,
    "pointwize_cross_entropy": POINTWISE_BCE,
}
OPT_KINDS = {"sgd": SGD, "adagrad": ADAGRAD, "lazy_adam": LAZY_ADAM, "adam": ADAM_DENSE}
STATUS = {0: "OK", -1: "INVALID", -2: "CUDA", -3: "ALIGNMENT", -4: "UNSUPPORTED", -5: "WORKSPACE"}

LIB_PATH = os.environ.get("KGEB_LIB_PATH") or os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "lib", "libkgeb200.so"))  # env: experiment builds


class KgebError(RuntimeError):
    def __init__(self, status, text):
        super().__init__(f"libkgeb200: {STATUS.get(status, status)}: {text}")
        self.status = status


_p, _i, _i64, _u64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float

# name -> (restype, argtypes); mirrors include/kgeb200.h one to one
PROTOTYPES = {
    "kgeb_version": (_i, []),
    "kgeb_create": (_i, [C.POINTER(_p), _i, _i64, _i64, _i, _i, _i, _i, _i, _f]),
    "kgeb_destroy": (_i, [_p]),
    "kgeb_last_error": (C.c_char_p, [_p]),
    "kgeb_launch_count": (_i64, [_p]),
    "kgeb_num_sms": (_i, [_p]),
    "kgeb_set_kernel_timing": (_i, [_p, _i]),
    "kgeb_last_kernel_ms": (_i, [_p, _i, C.POINTER(_f)]),
    "kgeb_kernel_ms_sum": (_i, [_p, _i, C.POINTER(_f), C.POINTER(_i)]),
    "kgeb_set_score_param": (_i, [_p, _f]),
    "kgeb_set_hake_params": (_i, [_p, _f, _f, _f]),
    "kgeb_set_hake_call_rows": (_i, [_p, _i64]),
    "kgeb_set_hole_mode": (_i, [_p, _i]),
    "kgeb_filter_csr_workspace": (_i, [_p, _i64, _i64, _p]),
    "kgeb_filter_csr_count": (_i, [_p, _p, _i64, _p, _i64, _i, _p, _p, _i64, _p]),
    "kgeb_filter_csr_fill": (_i, [_p, _p, _i64, _i64, _i, _p, _p, _p, _p, _p]),
    "kgeb_set_dropout": (_i, [_p, _f, _u64, _u64, _i]),
    "kgeb_score_fwd": (_i, [_p, _p, _p, _p, _i64, _p, _p]),
    "kgeb_score_bwd": (_i, [_p, _p, _p, _p, _i64, _p, _p, _p, _p, _p, _p]),
    "kgeb_negatives": (_i, [_p, _p, _i64, _i, _u64, _u64, _i, _p, _p]),
    "kgeb_oversample": (_i, [_p, _p, _i64, _i64, _u64, _u64, _i, _p, _p]),
    "kgeb_loss_fwd_bwd": (_i, [_p, _p, _i64, _p, _i64, _i, _i, _i, _f, _f, _p, _p, _p, _p]),
    "kgeb_train_fwd_bwd": (_i, [_p, _p, _p, _p, _i64, _p, _i, _i, _f, _f, _i, _u64, _u64, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "kgeb_sortseg_workspace_bytes": (_i64, [_p, _i64, _i]),
    "kgeb_set_grad_sortseg": (_i, [_p, _p, _i64]),
    "kgeb_train_step_sortseg": (_i, [_p, _p, _p, _p, _p, _i64, _p, _i, _i, _f, _f, _i, _u64, _u64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _f, _f, _f, _f, _i64, _p]),
    "kgeb_l3_reg": (_i, [_p, _p, _i64, _i, _i, _f, _p, _p, _p, _p]),
    "kgeb_opt_update": (_i, [_p, _p, _p, _p, _i64, _i, _i, _p, _p, _i, _f, _f, _f, _f, _i64, _p]),
    "kgeb_opt_update2": (_i, [_p, _p, _p, _p, _i64, _i, _i, _p, _p, _p, _p, _p, _i64, _i, _i, _p, _p, _i, _f, _f, _f, _f, _i64, _p]),
    "kgeb_query_width": (_i, [_p, _i]),
    "kgeb_query_prep": (_i, [_p, _p, _i64, _i64, _p, _p, _i64, _i, _p, _p]),
    "kgeb_target_scores": (_i, [_p, _p, _i64, _i64, _p, _p, _i64, _i, _p, _p]),
    "kgeb_rank_1n": (_i, [_p, _p, _i64, _i64, _p, _p, _p, _i64, _i, _p, _p, _p, _p]),
    "kgeb_count_greater": (_i, [_p, _p, _i64, _i64, _p, _p]),
    "kgeb_rows_gather": (_i, [_p, _p, _i, _i, _p, _i64, _p, _p]),
    "kgeb_rows_scatter_add": (_i, [_p, _p, _i, _i, _p, _p, _i64, _p, _p]),
    "kgeb_grad_rows_export": (_i, [_p, _p, _p, _i64, _i, _i, _p, _p, _i64, _p, _i, _p]),
    "kgeb_comm_unique_id": (_i, [_p]),
    "kgeb_comm_version": (_i, []),
    "kgeb_comm_init": (_i, [_p, _p, _i, _i]),
    "kgeb_comm_destroy": (_i, [_p]),
    "kgeb_comm_info": (_i, [_p, C.POINTER(_i), C.POINTER(_i)]),
    "kgeb_allreduce_counts": (_i, [_p, _p, _i64, _p]),
    "kgeb_allreduce_f32": (_i, [_p, _p, _i64, _p]),
    "kgeb_allreduce_flags": (_i, [_p, _p, _i64, _p]),
    "kgeb_alltoall_rows": (_i, [_p, _p, C.POINTER(_i64), _p, C.POINTER(_i64), _i, _p]),
    "kgeb_peer_granularity": (_i64, [_i]),
    "kgeb_peer_create": (_i, [C.POINTER(_p), _i, _i, _i, _i64]),
    "kgeb_peer_destroy": (_i, [_p]),
    "kgeb_peer_last_error": (C.c_char_p, [_p]),
    "kgeb_peer_shard_bytes": (_i64, [_p]),
    "kgeb_peer_base": (_p, [_p]),
    "kgeb_peer_export_fd": (_i, [_p, C.POINTER(_i)]),
    "kgeb_peer_import_fd": (_i, [_p, _i, _i]),
    "kgeb_peer_complete": (_i, [_p]),
    "kgeb_peer_pull_add": (_i, [_p, _p, _i64, _i, _i, _i64, _i64, _i, _i, _p, _p]),
    "kgeb_rows_concat": (_i, [_p, _p, _i, _i, _p, _i, _i, _p, _i64, _p, _i, _p]),
    "kgeb_rows_concat_bwd": (_i, [_p, _p, _i, _i, _p, _i64, _p, _i, _p, _p]),
    "kgeb_sgemm": (_i, [_p, _i, _i, _i64, _i, _i, _p, _i, _p, _i, _p, _i, _i, _p]),
    "kgeb_dense_fwd": (_i, [_p, _p, _i, _i64, _i, _p, _i, _i, _p, _i, _p, _i, _p]),
    "kgeb_dense_act_fwd": (_i, [_p, _p, _i64, _i, _i, _p, _i, _p]),
    "kgeb_dense_act_bwd": (_i, [_p, _p, _p, _i64, _i, _i, _i, _p, _p]),
    "kgeb_rank_tc_packed_bytes": (_i64, [_p, _i64, _i]),
    "kgeb_rank_tc_pack_entities": (_i, [_p, _p, _i64, _i, _p, _i64, _p]),
    "kgeb_rank_tc_workspace_bytes": (_i64, [_p, _i64, _i, _i64]),
    "kgeb_rank_1n_tc": (_i, [_p, _p, _i64, _i64, _p, _p, _p, _p, _i64, _i, _p, _p, _i, _p, _i64, _i64, _p, _p]),
}

_lib = None


def lib():
    """The loaded library (loaded once).  Raises ImportError with the build hint if it is not there."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              f"or `make -C kge-keras_b200/csrc` (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)  # AttributeError here = header / library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(status, ctx=None):
    if status != 0:
        msg = lib().kgeb_last_error(ctx)
        raise KgebError(status, msg.decode() if msg else "")
