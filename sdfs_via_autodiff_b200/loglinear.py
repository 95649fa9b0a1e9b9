"""Log-linear closed form of the wealth-consumption ratio -- host mirror of
``wc_loglinear_factory`` (/root/reference/code/ssy/ssy_model.py:86-156,
/root/reference/code/gcy/gcy_model.py:80-159) plus its evaluation on the whole
discretised grid ON THE DEVICE as an initial guess for the solvers
(w_init = exp(log-linear), cuts Newton outer iterations).

The constants come from one scalar root (the reference brackets it on [-20, 20] with
scipy's brentq; here a self-contained bisection to machine precision -- the two roots
differ by at most brentq's xtol = 2e-12).
"""
import ctypes as C
import math

from ._lib import lib, check
from .operator import Factors, MODEL_SSY, MODEL_GCY


def _bisect(f, lo, hi):
    flo, fhi = f(lo), f(hi)
    if flo == 0.0:
        return lo
    if fhi == 0.0:
        return hi
    if (flo < 0) == (fhi < 0):
        raise ValueError("f(a) and f(b) must have different signs")      # same failure mode as brentq
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        fm = f(mid)
        if fm == 0.0 or mid == lo or mid == hi:
            return mid
        if (fm < 0) == (flo < 0):
            lo, flo = mid, fm
        else:
            hi = mid
    return 0.5 * (lo + hi)


def _constants(β, ψ, θ, μ_c, ρ, ρ_λ, s_λ, φ_c, s_c, ρ_c, φ_z, s_z, ρ_z, extra=None):
    s_wc = 2 * φ_c ** 2 * s_c
    s_wx = 2 * φ_z ** 2 * s_z
    exp, log = math.exp, math.log

    def k1(x):
        return exp(x) / (1 + exp(x))

    def k0(x):
        return log(1 + exp(x)) - k1(x) * x

    def A1(x):
        return (1 - 1 / ψ) / (1 - k1(x) * ρ)

    def Aλ(x):
        return ρ_λ / (1 - k1(x) * ρ_λ)

    def Az(x):
        return (θ / 2) * (k1(x) * A1(x)) ** 2 / (1 - k1(x) * ρ_z)

    def Ac(x):
        return (θ / 2) * (1 - 1 / ψ) ** 2 / (1 - k1(x) * ρ_c)

    if extra is not None:
        ρ_π, ρ_ππ, φ_zπ, s_zπ, ρ_zπ = extra
        s_wxπ = 2 * φ_zπ ** 2 * s_zπ

        def Aπ(x):
            return k1(x) * (1 - 1 / ψ) * ρ_π / ((1 - k1(x) * ρ) * (1 - k1(x) * ρ_ππ))

        def Azπ(x):
            return (θ / 2) * (k1(x) * Aπ(x)) ** 2 / (1 - k1(x) * ρ_zπ)

    def A0(x):
        v = (log(β) + k0(x) + μ_c * (1 - 1 / ψ)
             + k1(x) * Az(x) * φ_z ** 2 * (1 - ρ_z)
             + k1(x) * Ac(x) * φ_c ** 2 * (1 - ρ_c))
        q = (k1(x) * Aλ(x) + 1) ** 2 * s_λ ** 2 + (k1(x) * Az(x) * s_wx) ** 2 + (k1(x) * Ac(x) * s_wc) ** 2
        if extra is not None:
            v += k1(x) * Azπ(x) * φ_zπ ** 2 * (1 - ρ_zπ)
            q += (k1(x) * Azπ(x) * s_wxπ) ** 2
        return (v + (θ / 2) * q) / (1 - k1(x))

    def fq_bar(x):
        v = x - A0(x) - Ac(x) * φ_c ** 2 - Az(x) * φ_z ** 2
        if extra is not None:
            v -= Azπ(x) * φ_zπ ** 2
        return v

    q = _bisect(fq_bar, -20.0, 20.0)
    c = dict(qbar=q, A0=A0(q), Ah_λ=Aλ(q), Ah_c=Ac(q), Ah_z=Az(q), Az=A1(q))
    if extra is not None:
        c.update(Ah_zπ=Azπ(q), Az_π=Aπ(q))
    return c


def ssy_factory(ssy):
    β, γ, ψ, μ_c, ρ, ϕ_z, ϕ_c, ρ_z, ρ_c, ρ_λ, s_z, s_c, s_λ = ssy.params
    c = _constants(β, ψ, ssy.θ, μ_c, ρ, ρ_λ, s_λ, ϕ_c, s_c, ρ_c, ϕ_z, s_z, ρ_z)

    def wc_loglinear(x):
        """Evaluates the log-linear solution (log w) at state (h_λ, h_c, h_z, z)."""
        h_λ, h_c, h_z, z = x
        sz = h_z * 2 * ϕ_z ** 2 + ϕ_z ** 2
        sc = h_c * 2 * ϕ_c ** 2 + ϕ_c ** 2
        return c["A0"] + c["Ah_λ"] * h_λ + c["Ah_c"] * sc + c["Ah_z"] * sz + c["Az"] * z
    wc_loglinear.coeffs = c
    return wc_loglinear


def gcy_factory(gcy):
    (β, ψ, γ, ρ_λ, s_λ, μ_c, φ_c, ρ, ρ_π, φ_z, ρ_c, s_c, ρ_z, s_z, ρ_ππ, φ_zπ, ρ_zπ, s_zπ) = gcy.params
    θ = (1 - γ) / (1 - 1 / ψ)
    c = _constants(β, ψ, θ, μ_c, ρ, ρ_λ, s_λ, φ_c, s_c, ρ_c, φ_z, s_z, ρ_z, extra=(ρ_π, ρ_ππ, φ_zπ, s_zπ, ρ_zπ))

    def wc_loglinear(x):
        """Evaluates the log-linear solution (log w) at state (h_λ, h_c, h_z, h_zπ, z, z_π)."""
        h_λ, h_c, h_z, h_zπ, z, z_π = x
        s_z_1 = h_z * 2 * φ_z ** 2 + φ_z ** 2
        s_c_1 = h_c * 2 * φ_c ** 2 + φ_c ** 2
        s_zπ_1 = h_zπ * 2 * φ_zπ ** 2 + φ_zπ ** 2
        return (c["A0"] + c["Ah_λ"] * h_λ + c["Ah_c"] * s_c_1 + c["Ah_z"] * s_z_1 + c["Az"] * z
                + c["Ah_zπ"] * s_zπ_1 + c["Az_π"] * z_π)
    wc_loglinear.coeffs = c
    return wc_loglinear


def loglinear_guess(model, shapes, factors=None, ctx=None, log=False):
    """exp(log-linear closed form) evaluated at every grid state on the device: a warm start for
    ``newton_solver`` / ``successive_approx``.  ``log=True`` returns log w instead."""
    is_gcy = hasattr(model, "ρ_ππ")
    kind = MODEL_GCY if is_gcy else MODEL_SSY
    fac = factors if factors is not None else Factors.build(kind, model.params, shapes, ctx)
    c = (gcy_factory if is_gcy else ssy_factory)(model).coeffs
    vals = [c["A0"], c["Ah_λ"], c["Ah_c"], c["Ah_z"], c["Az"], c.get("Ah_zπ", 0.0), c.get("Az_π", 0.0)]
    coeffs = (C.c_double * 7)(*vals)
    out = fac.ctx.empty(fac.shapes)
    check(lib.sdfs_factors_loglinear(fac.handle, coeffs, 0 if log else 1, out.ptr), fac.ctx.handle)
    return out
