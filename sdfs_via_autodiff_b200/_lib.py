"""ctypes binding of libsdfs_b200.so (include/sdfs_b200.h).

The shared library is the only compute path of this package.  If it is missing or
cannot be loaded the import fails loudly -- there is no CPU or PyTorch fallback.
"""
import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsdfs_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "sdfs_b200.h")


class SdfsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsdfs_b200 error {code}: {msg}")
        self.code = code


def _point_at_nccl():
    """libsdfs_b200 dlopens NCCL lazily (multi-GPU contexts only).  Tell it where the NCCL wheel of this
    environment lives unless the caller already did; nothing is imported (torch stays optional)."""
    if os.environ.get("SDFS_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["SDFS_NCCL_LIB"] = cand
                return
    except Exception:
        pass


def _load():
    _point_at_nccl()
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found. Build it with `python sdfs_via_autodiff_b200/build.py` "
            "(needs nvcc). This package has no CPU fallback.")
    return C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)


lib = _load()

c_i64 = C.c_int64
c_f64 = C.c_double
c_vp = C.c_void_p
P = C.POINTER

_SIGS = {
    "sdfs_abi_version": (C.c_int, []),
    "sdfs_version_string": (C.c_char_p, []),
    "sdfs_ctx_create": (C.c_int, [C.c_int, P(c_vp)]),
    "sdfs_ctx_destroy": (C.c_int, [c_vp]),
    "sdfs_last_error": (C.c_char_p, [c_vp]),
    "sdfs_ctx_sync": (C.c_int, [c_vp]),
    "sdfs_ctx_device_sync": (C.c_int, [c_vp]),
    "sdfs_ctx_device": (C.c_int, [c_vp, P(C.c_int), P(C.c_int), P(C.c_size_t), P(C.c_size_t)]),
    "sdfs_ctx_launch_count": (c_i64, [c_vp]),
    "sdfs_prof_enable": (C.c_int, [c_vp, C.c_int]),
    "sdfs_prof_read": (C.c_int, [c_vp, P(c_f64), P(c_i64)]),
    "sdfs_timer_start": (C.c_int, [c_vp]),
    "sdfs_timer_stop_ms": (C.c_int, [c_vp, P(c_f64)]),
    "sdfs_malloc": (C.c_int, [c_vp, C.c_size_t, P(c_vp)]),
    "sdfs_free": (C.c_int, [c_vp, c_vp]),
    "sdfs_memset": (C.c_int, [c_vp, c_vp, C.c_int, C.c_size_t]),
    "sdfs_h2d": (C.c_int, [c_vp, c_vp, c_vp, C.c_size_t]),
    "sdfs_d2h": (C.c_int, [c_vp, c_vp, c_vp, C.c_size_t]),
    "sdfs_d2d": (C.c_int, [c_vp, c_vp, c_vp, C.c_size_t]),
    "sdfs_host_alloc_pinned": (C.c_int, [C.c_size_t, P(c_vp)]),
    "sdfs_host_free_pinned": (C.c_int, [c_vp]),
    "sdfs_fill_f64": (C.c_int, [c_vp, c_vp, c_f64, c_i64]),
    "sdfs_factors_build": (C.c_int, [c_vp, C.c_int, P(c_f64), P(C.c_int32), P(c_vp)]),
    "sdfs_factors_from_host": (C.c_int, [c_vp, C.c_int, P(c_f64), P(C.c_int32), P(c_vp), C.c_int, P(c_vp)]),
    "sdfs_factors_destroy": (C.c_int, [c_vp]),
    "sdfs_factors_count": (C.c_int, [c_vp, P(C.c_int)]),
    "sdfs_factors_array": (C.c_int, [c_vp, C.c_int, P(c_i64), P(c_vp)]),
    "sdfs_factors_loglinear": (C.c_int, [c_vp, P(c_f64), C.c_int, c_vp]),
    "sdfs_op_from_dense": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_f64, c_f64, P(c_vp)]),
    "sdfs_op_from_factors": (C.c_int, [c_vp, c_vp, C.c_int, P(c_vp)]),
    "sdfs_op_continuous": (C.c_int, [c_vp, C.c_int, P(c_f64), P(C.c_int32), P(c_f64), P(c_f64), P(c_f64), c_i64, P(c_vp)]),
    "sdfs_interp_points": (C.c_int, [c_vp, C.c_int, P(C.c_int32), P(c_f64), P(c_f64), c_vp, c_vp, c_i64, c_vp]),
    "sdfs_op_destroy": (C.c_int, [c_vp]),
    "sdfs_op_info": (C.c_int, [c_vp, P(c_i64), P(c_i64), P(c_i64), P(c_i64), P(c_f64), P(c_f64), P(C.c_int)]),
    "sdfs_op_arrays": (C.c_int, [c_vp, P(c_vp), P(c_vp), P(c_vp), P(c_vp)]),
    "sdfs_op_set_esdf": (C.c_int, [c_vp, c_vp]),
    "sdfs_op_set_preferences": (C.c_int, [c_vp, c_f64, c_f64, c_f64]),
    "sdfs_op_apply_T": (C.c_int, [c_vp, c_vp, c_vp]),
    "sdfs_op_apply_jvp": (C.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "sdfs_op_apply_P": (C.c_int, [c_vp, c_vp, c_vp]),
    "sdfs_op_bench_pass": (C.c_int, [c_vp, C.c_int, C.c_int, P(c_f64)]),
    "sdfs_op_sdf": (C.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "sdfs_op_sdf_rows": (C.c_int, [c_vp, c_vp, P(c_i64), c_i64, c_vp]),
    "sdfs_solve_sa": (C.c_int, [c_vp, c_vp, c_f64, c_i64, c_vp, P(c_i64), P(c_f64), c_vp, c_i64, c_i64]),
    "sdfs_solve_newton": (C.c_int, [c_vp, c_vp, c_f64, c_i64, C.c_int, c_f64, c_f64, C.c_int, c_i64, c_vp,
                                    P(c_i64), P(c_f64), P(c_f64), P(c_i64), c_i64, P(c_i64)]),
    "sdfs_solve_anderson": (C.c_int, [c_vp, c_vp, c_f64, c_i64, C.c_int, C.c_int, c_f64, c_f64, c_vp, P(c_i64), P(c_f64)]),
    "sdfs_sweep_solve_sa": (C.c_int, [c_vp, P(c_f64), c_i64, c_f64, c_f64, c_i64, c_vp, P(c_i64), P(c_f64)]),
    "sdfs_sweep_solve_newton": (C.c_int, [c_vp, P(c_f64), c_i64, c_f64, c_f64, c_i64, c_f64, c_f64, c_i64, c_vp,
                                          P(c_i64), P(c_f64), P(c_i64), P(c_i64)]),
    "sdfs_sweep_apply_T": (C.c_int, [c_vp, P(c_f64), c_i64, c_vp, c_vp]),
    "sdfs_sweep_set_form": (C.c_int, [c_vp, C.c_int]),
    "sdfs_comm_unique_id": (C.c_int, [c_vp]),
    "sdfs_comm_init": (C.c_int, [c_vp, C.c_int, C.c_int, c_vp]),
    "sdfs_comm_rank": (C.c_int, [c_vp, P(C.c_int), P(C.c_int)]),
    "sdfs_comm_allgather_f64": (C.c_int, [c_vp, c_vp, c_i64]),
    "sdfs_comm_barrier": (C.c_int, [c_vp]),
    "sdfs_comm_arena_export": (C.c_int, [c_vp, c_i64, c_vp]),
    "sdfs_comm_arena_import": (C.c_int, [c_vp, c_vp]),
    "sdfs_dlpack_export": (C.c_int, [c_vp, c_vp, C.c_int, P(c_i64), c_vp, c_vp, P(c_vp)]),
    "sdfs_dlpack_import": (C.c_int, [c_vp, P(c_vp), P(C.c_int), P(c_i64), P(C.c_int), P(c_i64)]),
    "sdfs_dlpack_call_deleter": (None, [c_vp]),
}


def declared_symbols():
    """Every function name declared in include/sdfs_b200.h."""
    text = open(HEADER, encoding="utf-8").read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sdfs_[A-Za-z0-9_]+)\s*\(", text)))


def _bind():
    missing = []
    for name, (res, args) in _SIGS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    if missing:
        raise ImportError(f"libsdfs_b200.so lacks symbols {missing}; rebuild it")


_bind()


def check(rc, ctx=None):
    if rc != 0:
        msg = lib.sdfs_last_error(ctx if ctx else None)
        raise SdfsError(rc, msg.decode() if msg else "unknown")
