"""Operator handles: the device-resident form of the closure
``T = lambda w: T_ssy(w, shapes, params, arrays)`` (ssy_wc_ratio.py:230).

A ``WCOperator`` is callable (one T evaluation per call, so the reference's generic
``solver(f, x_init)`` signature keeps working) and is what the device-resident
solver loops dispatch on.
"""
import ctypes as C
import hashlib

import numpy as np

from ._lib import lib, check
from .device import Context, DeviceArray

MODEL_SSY, MODEL_GCY = 0, 1
STORAGE_DENSE, STORAGE_KRON = 0, 1
STORAGE_KRON_LOCAL = 4     # factor form kept whole on this rank even in a multi-rank context
_SSY_ARRAY_SHAPES = lambda s: [(s[0],), (s[0], s[0]), (s[1],), (s[1], s[1]), (s[2],), (s[2], s[2]),
                               (s[2], s[3]), (s[2], s[3], s[3]), (s[1],), (s[2],)]


def _gcy_array_shapes(s):
    nz, nzp, nhz, nhc, nhzp, nhl = s
    return [(nzp, nhz, nhzp, nz), (nzp, nhz, nhzp, nz, nz), (nhzp, nzp), (nhzp, nzp, nzp),
            (nhz,), (nhz, nhz), (nhz,), (nhc,), (nhc, nhc), (nhc,), (nhzp,), (nhzp, nhzp), (nhzp,),
            (nhl,), (nhl, nhl)]


def array_shapes(model, shapes):
    return _SSY_ARRAY_SHAPES(shapes) if model == MODEL_SSY else _gcy_array_shapes(shapes)


class Factors:
    """Discretised model on the device (Markov factor arrays in the reference's tuple order)."""

    def __init__(self, ctx, handle, model, shapes, params):
        self.ctx, self.handle, self.model = ctx, handle, model
        self.shapes, self.params = tuple(int(s) for s in shapes), tuple(float(p) for p in params)

    @classmethod
    def build(cls, model, params, shapes, ctx=None):
        """Device-side discretisation (Rouwenhorst chains computed on the GPU)."""
        ctx = ctx or Context.default()
        p = (C.c_double * len(params))(*[float(x) for x in params])
        s = (C.c_int32 * len(shapes))(*[int(x) for x in shapes])
        h = C.c_void_p()
        check(lib.sdfs_factors_build(ctx.handle, model, p, s, C.byref(h)), ctx.handle)
        return cls(ctx, h, model, shapes, params)

    @classmethod
    def from_host(cls, model, params, shapes, arrays, ctx=None):
        ctx = ctx or Context.default()
        want = array_shapes(model, shapes)
        if len(arrays) != len(want):
            raise ValueError(f"expected {len(want)} factor arrays, got {len(arrays)}")
        host = []
        for a, shp in zip(arrays, want):
            a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
            if a.shape != tuple(shp):
                raise ValueError(f"factor array has shape {a.shape}, expected {tuple(shp)}")
            host.append(a)
        ptrs = (C.c_void_p * len(host))(*[a.ctypes.data for a in host])
        p = (C.c_double * len(params))(*[float(x) for x in params])
        s = (C.c_int32 * len(shapes))(*[int(x) for x in shapes])
        h = C.c_void_p()
        check(lib.sdfs_factors_from_host(ctx.handle, model, p, s, ptrs, len(host), C.byref(h)), ctx.handle)
        return cls(ctx, h, model, shapes, params)

    def arrays(self):
        """Download the factor arrays as the NumPy tuple the reference's discretisers return."""
        out = []
        for i, shp in enumerate(array_shapes(self.model, self.shapes)):
            n, p = C.c_int64(), C.c_void_p()
            check(lib.sdfs_factors_array(self.handle, i, C.byref(n), C.byref(p)), self.ctx.handle)
            a = np.empty(shp, dtype=np.float64)
            assert a.size == n.value
            check(lib.sdfs_d2h(self.ctx.handle, a.ctypes.data, p, a.nbytes), self.ctx.handle)
            a.flags.writeable = False     # immutable like the reference's device arrays (lets cached_operator memoise digests)
            out.append(a)
        return tuple(out)

    def __del__(self):
        try:
            lib.sdfs_factors_destroy(self.handle)
        except Exception:
            pass


class WCOperator:
    """T w = 1 + β (a_row ⊙ P (a_col ⊙ w^θ))^(1/θ) on the device."""

    def __init__(self, ctx, handle, shapes, keep=()):
        self.ctx, self.handle = ctx, handle
        self.shapes = tuple(int(s) for s in shapes)
        self._keep = list(keep)          # device arrays / factors the handle borrows
        N, ld, rb, re = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        beta, theta, st = C.c_double(), C.c_double(), C.c_int()
        check(lib.sdfs_op_info(handle, C.byref(N), C.byref(ld), C.byref(rb), C.byref(re), C.byref(beta),
                               C.byref(theta), C.byref(st)), ctx.handle)
        self.N, self.ld, self.row_begin, self.row_end = N.value, ld.value, rb.value, re.value
        self.β, self.θ, self.storage = beta.value, theta.value, st.value
        if int(np.prod(self.shapes)) != self.N:
            raise ValueError(f"shapes {self.shapes} do not multiply to N={self.N}")

    # -- constructors ---------------------------------------------------------
    @classmethod
    def from_factors(cls, factors, storage="auto"):
        ctx = factors.ctx
        N = int(np.prod(factors.shapes))
        if storage == "auto":
            # up to 4096 states (P <= 134 MB) a dense P serves both solvers best: Newton's mat-vecs are one 18 us pass
            # over an L2-resident matrix (six factor-form modes cost 54 us on the reference's default GCY grid), and
            # successive approximation runs as a one-CTA kernel either way (P in registers up to 160 states, the
            # factor form in shared memory above: the operator keeps its factors).  Beyond that the factor form is
            # faster everywhere (3x at 10^4 states; dense P is impossible beyond ~1.5 10^5), so "auto" never
            # materialises a large P - ask for "dense" to get one.  In a multi-rank context the small dense P is
            # replicated, not row-sharded (cross-rank barriers would cost more than the work).
            storage = ("dense_replicated" if ctx.nranks > 1 else "dense") if N <= 4096 else "kron"
        st = {"dense": STORAGE_DENSE, "kron": STORAGE_KRON, "kron_local": STORAGE_KRON_LOCAL,
              "dense_replicated": 2}[storage]
        h = C.c_void_p()
        check(lib.sdfs_op_from_factors(ctx.handle, factors.handle, st, C.byref(h)), ctx.handle)
        return cls(ctx, h, factors.shapes, keep=[factors])

    @classmethod
    def from_dense(cls, P, a_row, a_col, β, θ, shapes=None, e_sdf=None, ctx=None, row_range=None):
        """Dense single-index operator from caller-supplied arrays (host arrays are uploaded;
        DeviceArray / DLPack inputs are used in place).  P is N x N row-major, or the
        (row_end-row_begin) x N row slice of a row-sharded rank."""
        ctx = ctx or Context.default()
        dP, da, dc = ctx.asarray(P), ctx.asarray(a_row), ctx.asarray(a_col)
        N = da.size
        rb, re = row_range if row_range is not None else (0, N)
        if dP.ndim != 2 or dP.shape[0] != re - rb or dP.shape[1] < N:
            raise ValueError(f"P has shape {dP.shape}, expected ({re - rb}, >={N})")
        h = C.c_void_p()
        check(lib.sdfs_op_from_dense(ctx.handle, dP.ptr, N, dP.shape[1], rb, re, da.ptr, dc.ptr, float(β),
                                     float(θ), C.byref(h)), ctx.handle)
        keep = [dP, da, dc]
        if e_sdf is not None:
            de = ctx.asarray(e_sdf)
            check(lib.sdfs_op_set_esdf(h, de.ptr), ctx.handle)
            keep.append(de)
        return cls(ctx, h, shapes if shapes is not None else (N,), keep=keep)

    def __del__(self):
        try:
            lib.sdfs_op_destroy(self.handle)
        except Exception:
            pass

    # -- applications -----------------------------------------------------------
    def _in(self, w):
        d = self.ctx.asarray(w)
        if d.size != self.N:
            raise ValueError(f"array of size {d.size} given to an operator with N={self.N}")
        return d

    def __call__(self, w, out=None):
        """One evaluation of T.  ``out`` (optional, an extension): a DeviceArray of N elements to write
        into instead of allocating the result (must not alias ``w``)."""
        if isinstance(w, _Probe):
            return _ProbeResult(self)
        d = self._in(w)
        if out is None:
            out = self.ctx.empty(self.shapes)
        elif not isinstance(out, DeviceArray) or out.size != self.N or out.ptr.value == d.ptr.value:
            raise ValueError("out must be a DeviceArray of N elements distinct from the input")
        check(lib.sdfs_op_apply_T(self.handle, d.ptr, out.ptr), self.ctx.handle)
        return out

    T = __call__

    def jvp(self, w, v):
        """J_T(w) v (analytic; replaces jax.jvp, solvers.py:87)."""
        dw, dv = self._in(w), self._in(v)
        out = self.ctx.empty(self.shapes)
        check(lib.sdfs_op_apply_jvp(self.handle, dw.ptr, dv.ptr, out.ptr), self.ctx.handle)
        return out

    def apply_P(self, x):
        d = self._in(x)
        out = self.ctx.empty(self.shapes)
        check(lib.sdfs_op_apply_P(self.handle, d.ptr, out.ptr), self.ctx.handle)
        return out

    def bench_pass(self, mode=0, reps=20):
        """Average device ms of `reps` back-to-back dense passes (diagnostic; see sdfs_b200.h)."""
        ms = C.c_double()
        check(lib.sdfs_op_bench_pass(self.handle, int(mode), int(reps), C.byref(ms)), self.ctx.handle)
        return ms.value

    def sdf(self, w):
        """(q_f, euler_residual): one-period risk-free price E[M'|x] and the Euler-equation
        residual β^θ s/(w-1)^θ - 1 of paper/autosdfs.tex:374-384."""
        d = self._in(w)
        qf, eu = self.ctx.empty(self.shapes), self.ctx.empty(self.shapes)
        check(lib.sdfs_op_sdf(self.handle, d.ptr, qf.ptr, eu.ptr), self.ctx.handle)
        return qf, eu

    def sdf_rows(self, w, rows):
        """Rows M̄(n, ·) of the SDF matrix for current states ``rows`` (flat indices)."""
        d = self._in(w)
        rows = np.ascontiguousarray(np.asarray(rows, dtype=np.int64).reshape(-1))
        out = self.ctx.empty((rows.size, self.N))
        check(lib.sdfs_op_sdf_rows(self.handle, d.ptr, rows.ctypes.data_as(C.POINTER(C.c_int64)), rows.size,
                                   out.ptr), self.ctx.handle)
        return out

    def set_preferences(self, γ, ψ, β):
        check(lib.sdfs_op_set_preferences(self.handle, float(γ), float(ψ), float(β)), self.ctx.handle)
        self.β, self.θ = float(β), (1 - γ) / (1 - 1 / ψ)

    def device_arrays(self):
        """(P, a_row, a_col, e_sdf) as DeviceArray views (None when not materialised)."""
        ps = [C.c_void_p() for _ in range(4)]
        check(lib.sdfs_op_arrays(self.handle, *[C.byref(p) for p in ps]), self.ctx.handle)
        nloc = self.row_end - self.row_begin
        shapes = [(nloc, self.ld), (self.N,), (self.N,), (self.N,)]
        return tuple(DeviceArray._view(self.ctx, p.value, s, self) if p.value else None
                     for p, s in zip(ps, shapes))


class _Probe:
    """Stand-in for ``w`` used to discover which operator a Python closure applies --
    the analogue of the abstract value JAX traces a jitted function with."""


class _ProbeResult:
    def __init__(self, op):
        self.op = op


def resolve_operator(f):
    """WCOperator behind ``f`` (an operator, or a closure such as
    ``lambda w: T_ssy(w, shapes, params, arrays)``), else None.

    The closure is called once with a probe object.  Only failures caused by the probe itself
    (a callable that does arithmetic on its argument: TypeError / AttributeError naming the
    probe) mean "not an operator of this package"; anything raised while the operator is being
    built or applied (SdfsError: out of memory, CUDA errors; ValueError: wrong array shapes or
    parameters; MemoryError ...) propagates unchanged."""
    if isinstance(f, WCOperator):
        return f
    try:
        r = f(_Probe())
    except (TypeError, AttributeError) as e:
        if "_Probe" in str(e):
            return None
        raise
    return r.op if isinstance(r, _ProbeResult) else None


_op_cache = {}
_digest_cache = {}      # id(array) -> (weakref, shape, digest): hash each factor array once, not per call


def _array_digest(a):
    import weakref
    if isinstance(a, np.ndarray) and not a.flags.writeable:
        ent = _digest_cache.get(id(a))
        if ent is not None and ent[0]() is a:
            return ent[1]
    d = hashlib.sha1(np.ascontiguousarray(np.asarray(a, dtype=np.float64)).tobytes()).digest()
    if isinstance(a, np.ndarray) and not a.flags.writeable:
        try:
            key = id(a)
            _digest_cache[key] = (weakref.ref(a, lambda _r, k=key: _digest_cache.pop(k, None)), d)
        except TypeError:
            pass
    return d


class _CachedOperator(WCOperator):
    """Operator shared through ``cached_operator``: keyed by (shapes, params, arrays), so it
    must keep computing what that key says -- preferences cannot be changed in place."""

    def set_preferences(self, γ, ψ, β):
        raise ValueError("this operator is shared through the T_ssy/T_gcy operator cache and is immutable; "
                         "build a private one with make_T_ssy / make_T_gcy (or WCOperator.from_factors) "
                         "to change preferences")


def cached_operator(model, shapes, params, arrays, storage="auto", ctx=None):
    """Operator for (shapes, params, arrays), built once per distinct input.  The discretisers of
    this package return read-only arrays, whose content digests are remembered, so repeated
    ``T_ssy(w, shapes, params, arrays)`` calls do not re-hash the factors."""
    ctx = ctx or Context.default()
    h = hashlib.sha1()
    h.update(repr((model, tuple(shapes), tuple(float(p) for p in params), storage, ctx.device)).encode())
    for a in arrays:
        h.update(_array_digest(a))
    key = h.hexdigest()
    op = _op_cache.get(key)
    if op is None:
        fac = Factors.from_host(model, params, shapes, arrays, ctx)
        op = WCOperator.from_factors(fac, storage)
        op.__class__ = _CachedOperator
        if len(_op_cache) >= 8:
            _op_cache.pop(next(iter(_op_cache)))
        _op_cache[key] = op
    return op
