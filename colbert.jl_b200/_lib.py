"""ctypes binding of libcolbert_b200.so (include/colbert_b200.h).  This is the Python twin of
the `ccall` stubs in julia/ColBERTB200.jl.  Loading fails loudly when the library has not been
built: there is no CPU fallback anywhere in this package."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("COLBERT_B200_LIB") or os.path.join(_HERE, "lib", "libcolbert_b200.so")   # env override: A/B builds

CB_OK, CB_ERR_BAD_ARG, CB_ERR_DOMAIN, CB_ERR_CUDA, CB_ERR_OOM, CB_ERR_UNSUPPORTED, CB_ERR_BOUNDS = range(7)
CB_FLAG_DEVICE_POINTERS = 1
CB_FLAG_BORROW_RESIDUALS = 2

# name -> (restype, argtypes); exactly the symbols include/colbert_b200.h declares
_p = C.c_void_p
SIGNATURES = {
    "cb_version": (C.c_char_p, []),
    "cb_last_error": (C.c_char_p, []),
    "cb_device_count": (C.c_int32, []),
    "cb_index_create": (C.c_int32, [C.POINTER(_p), C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64,
                                    C.c_int64, _p, _p, _p, _p, _p, _p, _p, C.c_int64, C.c_int32]),
    "cb_index_open": (C.c_int32, [C.POINTER(_p), C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int64)]),
    "cb_jld2_read": (C.c_int32, [C.c_char_p, C.c_char_p, C.POINTER(C.c_int64), _p, C.c_int64]),
    "cb_index_destroy": (C.c_int32, [_p]),
    "cb_index_info": (C.c_int32, [_p, C.POINTER(C.c_int64)]),
    "cb_set_option": (C.c_int32, [_p, C.c_char_p, C.c_int64]),
    "cb_get_stat": (C.c_int32, [_p, C.c_char_p, C.POINTER(C.c_double)]),
    "cb_search_batch": (C.c_int32, [_p, _p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _p, _p, _p]),
    "cb_search_batch_device": (C.c_int32, [_p, _p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _p, _p, _p, _p]),
    "cb_probe_device": (C.c_int32, [_p, _p, C.c_int32, C.c_int32, C.c_int32, _p, _p]),
    "cb_search_batch_cells_device": (C.c_int32, [_p, _p, _p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _p, _p, _p, _p]),
    "cb_multi_create": (C.c_int32, [C.POINTER(_p), C.c_int32, C.POINTER(_p)]),
    "cb_multi_open": (C.c_int32, [C.POINTER(_p), C.c_char_p, C.c_int32, C.POINTER(C.c_int32)]),
    "cb_multi_destroy": (C.c_int32, [_p]),
    "cb_multi_info": (C.c_int32, [_p, C.POINTER(C.c_int32), C.POINTER(_p), C.c_int32]),
    "cb_multi_search_batch": (C.c_int32, [_p, _p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _p, _p, _p]),
    "cb_search_batch_plaid": (C.c_int32, [_p, _p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_int32, _p, _p, _p]),
    "cb_search_batch_plaid_device": (C.c_int32, [_p, _p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_int32, _p, _p, _p, _p]),
    "cb_probe": (C.c_int32, [_p, _p, C.c_int32, C.c_int32, C.c_int32, _p, _p]),
    "cb_retrieve": (C.c_int32, [_p, _p, C.c_int32, C.c_int32, _p, C.c_int64, C.POINTER(C.c_int64)]),
    "cb_decompress": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int64, _p, _p, _p, _p, C.c_int64, _p, _p, _p]),
    "cb_compress": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int64, _p, _p, _p, C.c_int64, _p, _p]),
    "cb_maxsim": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, _p, _p, C.c_int64, _p, C.c_int64, _p, C.c_int64, _p]),
    "cb_score_pids": (C.c_int32, [_p, _p, C.c_int32, _p, C.c_int64, _p]),
    "cb_debug_tc_operand": (C.c_int32, [_p, _p, C.c_int64, _p, _p, C.c_int64]),
    "cb_merge_topk": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, _p, _p, _p, _p]),
    "cb_merge_topk_device": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, _p, _p, _p, _p, _p]),
}


class ColBERTB200Error(RuntimeError):
    """Base class; `status` is the CB_ERR_* code."""
    status = None


class DimensionMismatch(ColBERTB200Error):
    status = CB_ERR_BAD_ARG


class DomainError(ColBERTB200Error):
    status = CB_ERR_DOMAIN


class CudaError(ColBERTB200Error):
    status = CB_ERR_CUDA


class OutOfMemory(ColBERTB200Error):
    status = CB_ERR_OOM


class Unsupported(ColBERTB200Error):
    status = CB_ERR_UNSUPPORTED


class BoundsError(ColBERTB200Error, IndexError):
    status = CB_ERR_BOUNDS


_BY_STATUS = {c.status: c for c in (DimensionMismatch, DomainError, CudaError, OutOfMemory, Unsupported, BoundsError)}

_lib = None


def load():
    """Loads the shared library (once).  Raises ImportError with build instructions when the
    .so is missing -- the product path never degrades to a CPU implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"or colbert.jl_b200/csrc/build.sh (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library ever diverge
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status):
    if status != CB_OK:
        msg = load().cb_last_error().decode("utf-8", "replace")
        raise _BY_STATUS.get(status, ColBERTB200Error)(msg)
