"""Device context and device arrays (ctypes + DLPack; no torch, no cupy).

``DeviceArray`` is a C-contiguous fp64 tensor in HBM.  It implements the DLPack
protocol by hand (``__dlpack__`` / ``__dlpack_device__`` producer,
``from_dlpack`` consumer) so that arrays can be exchanged zero-copy with any
other CUDA library, and ``__array__`` so that ``np.asarray(x)`` downloads it.
Stands where ``jax.device_put`` / JAX device arrays stand in the reference
(ssy_wc_ratio.py:227).
"""
import ctypes as C
import threading

import numpy as np

from ._lib import lib, check, SdfsError

_pyapi = C.pythonapi
_pyapi.PyCapsule_New.restype = C.py_object
_pyapi.PyCapsule_New.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
_pyapi.PyCapsule_GetPointer.restype = C.c_void_p
_pyapi.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
_pyapi.PyCapsule_IsValid.restype = C.c_int
_pyapi.PyCapsule_IsValid.argtypes = [C.py_object, C.c_char_p]
_pyapi.PyCapsule_SetName.restype = C.c_int
_pyapi.PyCapsule_SetName.argtypes = [C.py_object, C.c_char_p]

_DLTENSOR = b"dltensor"
_USED = b"used_dltensor"
_live_exports = {}
_live_lock = threading.Lock()
_next_token = [1]


@C.CFUNCTYPE(None, C.c_void_p)
def _release_export(token):
    with _live_lock:
        _live_exports.pop(int(token or 0), None)


@C.CFUNCTYPE(None, C.c_void_p)
def _capsule_destructor(capsule_ptr):
    # called when a capsule nobody consumed is garbage collected
    cap = C.cast(capsule_ptr, C.py_object)
    if _pyapi.PyCapsule_IsValid(cap, _DLTENSOR):
        mt = _pyapi.PyCapsule_GetPointer(cap, _DLTENSOR)
        lib.sdfs_dlpack_call_deleter(mt)


def _release_pinned(ptr):
    lib.sdfs_host_free_pinned(C.c_void_p(ptr))


class Context:
    """One GPU, one stream.  ``Context.default()`` is created lazily on device
    ``LOCAL_RANK`` (or 0)."""

    _default = None

    def __init__(self, device=0):
        h = C.c_void_p()
        check(lib.sdfs_ctx_create(int(device), C.byref(h)))
        self.handle = h
        self.device = int(device)
        self.rank, self.nranks = 0, 1

    @classmethod
    def default(cls):
        if cls._default is None:
            import os
            cls._default = cls(int(os.environ.get("LOCAL_RANK", "0")))
        return cls._default

    def sync(self):
        check(lib.sdfs_ctx_sync(self.handle), self.handle)

    def device_info(self):
        dev, sms = C.c_int(), C.c_int()
        fr, tot = C.c_size_t(), C.c_size_t()
        check(lib.sdfs_ctx_device(self.handle, C.byref(dev), C.byref(sms), C.byref(fr), C.byref(tot)), self.handle)
        return dict(device=dev.value, sm_count=sms.value, free_bytes=fr.value, total_bytes=tot.value)

    @property
    def launch_count(self):
        return int(lib.sdfs_ctx_launch_count(self.handle))

    def prof_enable(self, max_launches):
        check(lib.sdfs_prof_enable(self.handle, int(max_launches)), self.handle)

    def prof_read(self):
        """(total device ms, launches) of the dense-pass kernel since prof_enable."""
        ms, n = C.c_double(), C.c_int64()
        check(lib.sdfs_prof_read(self.handle, C.byref(ms), C.byref(n)), self.handle)
        return ms.value, n.value

    def timer_start(self):
        check(lib.sdfs_timer_start(self.handle), self.handle)

    def timer_stop_ms(self):
        ms = C.c_double()
        check(lib.sdfs_timer_stop_ms(self.handle, C.byref(ms)), self.handle)
        return ms.value

    # -- arrays -----------------------------------------------------------
    def empty(self, shape):
        return DeviceArray._alloc(self, shape)

    def full(self, shape, value):
        a = DeviceArray._alloc(self, shape)
        check(lib.sdfs_fill_f64(self.handle, a.ptr, float(value), a.size), self.handle)
        return a

    def pinned_empty(self, shape):
        """Host ndarray backed by page-locked memory (cudaMallocHost): h2d/d2h copies of such
        arrays are true DMA transfers.  The memory is released when the array is collected."""
        shape = tuple(int(s) for s in (shape if hasattr(shape, "__len__") else (shape,)))
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        p = C.c_void_p()
        check(lib.sdfs_host_alloc_pinned(n * 8, C.byref(p)))
        # the owner of the memory is the ctypes buffer object: every ndarray view (slices, reshapes;
        # NumPy collapses .base to the frombuffer array, whose base is this buffer) keeps it alive, and
        # the page-locked allocation is released only when the buffer object itself is collected
        owner = type("PinnedBuffer", (C.c_double * max(n, 1),), {}).from_address(p.value)
        import weakref
        weakref.finalize(owner, _release_pinned, p.value)
        return np.frombuffer(owner, dtype=np.float64)[:n].reshape(shape)

    def asarray(self, x):
        """Host ndarray / DLPack producer / DeviceArray -> DeviceArray on this context."""
        if isinstance(x, DeviceArray):
            return x
        if hasattr(x, "__dlpack__") and not isinstance(x, np.ndarray):
            try:
                return DeviceArray.from_dlpack(x, self)
            except SdfsError:
                pass      # e.g. a CPU tensor: stage through the host below
        h = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
        a = DeviceArray._alloc(self, h.shape)
        check(lib.sdfs_h2d(self.handle, a.ptr, h.ctypes.data, h.nbytes), self.handle)
        return a


class DeviceArray:
    def __init__(self):
        raise TypeError("use Context.empty / Context.asarray / DeviceArray.from_dlpack")

    @classmethod
    def _alloc(cls, ctx, shape):
        self = object.__new__(cls)
        self.ctx = ctx
        self.shape = tuple(int(s) for s in (shape if hasattr(shape, "__len__") else (shape,)))
        self.size = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        p = C.c_void_p()
        check(lib.sdfs_malloc(ctx.handle, self.size * 8, C.byref(p)), ctx.handle)
        self.ptr = p
        self._owned = True
        self._foreign = None
        self._base = None
        return self

    @classmethod
    def _view(cls, ctx, ptr, shape, base):
        self = object.__new__(cls)
        self.ctx = ctx
        self.shape = tuple(int(s) for s in shape)
        self.size = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        self.ptr = C.c_void_p(ptr)
        self._owned = False
        self._foreign = None
        self._base = base
        return self

    def __del__(self):
        try:
            if getattr(self, "_owned", False) and self.ptr:
                lib.sdfs_free(self.ctx.handle, self.ptr)
            elif getattr(self, "_foreign", None):
                lib.sdfs_dlpack_call_deleter(self._foreign)
        except Exception:
            pass

    # -- numpy interop ------------------------------------------------------
    dtype = np.dtype(np.float64)

    @property
    def ndim(self):
        return len(self.shape)

    def numpy(self, out=None):
        """Download.  ``out`` (optional): a C-contiguous float64 host array of the same size,
        e.g. a pinned buffer from Context.pinned_empty."""
        if out is None:
            out = np.empty(self.shape, dtype=np.float64)
        elif out.size != self.size or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float64 array of the same size")
        check(lib.sdfs_d2h(self.ctx.handle, out.ctypes.data, self.ptr, out.nbytes), self.ctx.handle)
        return out

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)

    def fill(self, value):
        check(lib.sdfs_fill_f64(self.ctx.handle, self.ptr, float(value), self.size), self.ctx.handle)
        return self

    def __getitem__(self, key):
        """Contiguous slices along the leading axis (views, no copy)."""
        if not isinstance(key, slice) or not self.shape:
            raise TypeError("DeviceArray supports only contiguous leading-axis slices a[i:j]")
        i, j, step = key.indices(self.shape[0])
        if step != 1:
            raise TypeError("DeviceArray slices must have step 1")
        j = max(i, j)
        inner = int(np.prod(self.shape[1:], dtype=np.int64)) if len(self.shape) > 1 else 1
        return DeviceArray._view(self.ctx, self.ptr.value + i * inner * 8, (j - i,) + self.shape[1:], self)

    def reshape(self, *shape):
        if len(shape) == 1 and hasattr(shape[0], "__len__"):
            shape = tuple(shape[0])
        if -1 in shape:                          # one inferred axis, like NumPy
            known = int(np.prod([s for s in shape if s != -1], dtype=np.int64))
            if list(shape).count(-1) != 1 or known == 0 or self.size % known:
                raise ValueError(f"cannot reshape {self.shape} to {shape}")
            shape = tuple(self.size // known if s == -1 else s for s in shape)
        if int(np.prod(shape, dtype=np.int64)) != self.size:
            raise ValueError(f"cannot reshape {self.shape} to {shape}")
        return DeviceArray._view(self.ctx, self.ptr.value, shape, self)

    def copy(self):
        out = DeviceArray._alloc(self.ctx, self.shape)
        check(lib.sdfs_d2d(self.ctx.handle, out.ptr, self.ptr, self.size * 8), self.ctx.handle)
        return out

    def __repr__(self):
        return f"DeviceArray(shape={self.shape}, dtype=float64, device=cuda:{self.ctx.device})"

    # -- DLPack ---------------------------------------------------------------
    def __dlpack_device__(self):
        return (2, self.ctx.device)      # kDLCUDA

    def __dlpack__(self, stream=None, **kw):
        self.ctx.sync()                  # our work is complete before the consumer touches it
        with _live_lock:
            token = _next_token[0]
            _next_token[0] += 1
            _live_exports[token] = self  # keeps the memory alive until the consumer's deleter runs
        shape = (C.c_int64 * max(1, len(self.shape)))(*self.shape)
        mt = C.c_void_p()
        check(lib.sdfs_dlpack_export(self.ctx.handle, self.ptr, len(self.shape), shape, C.c_void_p(token),
                                     C.cast(_release_export, C.c_void_p), C.byref(mt)), self.ctx.handle)
        return _pyapi.PyCapsule_New(mt, _DLTENSOR, C.cast(_capsule_destructor, C.c_void_p))

    @classmethod
    def from_dlpack(cls, obj, ctx=None):
        ctx = ctx or Context.default()
        cap = obj.__dlpack__() if hasattr(obj, "__dlpack__") else obj
        if not _pyapi.PyCapsule_IsValid(cap, _DLTENSOR):
            raise ValueError("not a DLPack capsule (or already consumed)")
        mt = _pyapi.PyCapsule_GetPointer(cap, _DLTENSOR)
        ptr, ndim, dev, n = C.c_void_p(), C.c_int(), C.c_int(), C.c_int64()
        shape = (C.c_int64 * 8)()
        rc = lib.sdfs_dlpack_import(mt, C.byref(ptr), C.byref(ndim), shape, C.byref(dev), C.byref(n))
        check(rc)                        # on failure the capsule keeps ownership
        if dev.value != ctx.device:
            raise ValueError(f"tensor lives on cuda:{dev.value}, context on cuda:{ctx.device}")
        _pyapi.PyCapsule_SetName(cap, _USED)
        check(lib.sdfs_ctx_device_sync(ctx.handle), ctx.handle)   # producer's pending work is done
        self = cls._view(ctx, ptr.value, tuple(shape[i] for i in range(ndim.value)), None)
        self._foreign = C.c_void_p(mt)
        return self


def from_dlpack(obj, ctx=None):
    return DeviceArray.from_dlpack(obj, ctx)
