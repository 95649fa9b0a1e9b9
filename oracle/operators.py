"""Oracle operators T (loops, factor form, dense single-index) and the analytic
JVP (test infrastructure only).

Reference lines followed:
  * loop forms: ssy_wc_ratio.py:159-199, gcy_wc_ratio.py:244-302
  * vectorised factor form: ssy_wc_ratio.py:82-149, gcy_wc_ratio.py:134-236
    (here evaluated by sum-factorisation, mathematically the same 2D-axis sum)
  * dense single-index P / H / T: temp_ssy.py:24-42,106,146,153-159
  * analytic Jacobian: temp_ssy.py:204-216
"""
import numpy as np


def _theta(γ, ψ):
    return (1 - γ) / (1 - 1 / ψ)


# --------------------------------------------------------------------------
# Plain nested-loop forms (slow, obviously correct; small shapes only)
# --------------------------------------------------------------------------
def T_ssy_loops(w, shapes, params, arrays):
    L, K, I, J = shapes
    β, γ, ψ, μ_c = params[0], params[1], params[2], params[3]
    h_λ, Q_λ, h_c, Q_c, h_z, Q_hz, z, z_Q, σ_c, σ_z = arrays
    θ = _theta(γ, ψ)
    out = np.empty(shapes)
    for l in range(L):
        for k in range(K):
            for i in range(I):
                for j in range(J):
                    acc = 0.0
                    row = np.exp(0.5 * ((1 - γ) * σ_c[k]) ** 2) * \
                        np.exp((1 - γ) * (μ_c + z[i, j]))
                    for lp in range(L):
                        col = np.exp(θ * h_λ[lp])
                        for kp in range(K):
                            for ip in range(I):
                                for jp in range(J):
                                    pr = Q_λ[l, lp] * Q_c[k, kp] * \
                                        Q_hz[i, ip] * z_Q[i, j, jp]
                                    acc += w[lp, kp, ip, jp] ** θ * col * row * pr
                    out[l, k, i, j] = 1 + β * acc ** (1 / θ)
    return out


def T_gcy_loops(w, shapes, params, arrays):
    n_z, n_zπ, n_hz, n_hc, n_hzπ, n_hλ = shapes
    β, ψ, γ, μ_c = params[0], params[1], params[2], params[5]
    (z, z_Q, zπ, zπ_Q, h_z, Q_hz, σ_z, h_c, Q_hc, σ_c,
     h_zπ, Q_hzπ, σ_zπ, h_λ, Q_hλ) = arrays
    θ = _theta(γ, ψ)
    out = np.empty(shapes)
    wθ = w ** θ
    col = np.exp(θ * h_λ)
    for iz in range(n_z):
      for izp in range(n_zπ):
        for ihz in range(n_hz):
          for ihc in range(n_hc):
            for ihzp in range(n_hzπ):
              for ihl in range(n_hλ):
                row = np.exp(0.5 * ((1 - γ) * σ_c[ihc]) ** 2) * \
                    np.exp((1 - γ) * (μ_c + z[izp, ihz, ihzp, iz]))
                acc = 0.0
                for jz in range(n_z):
                  for jzp in range(n_zπ):
                    for jhz in range(n_hz):
                      for jhc in range(n_hc):
                        for jhzp in range(n_hzπ):
                          for jhl in range(n_hλ):
                            pr = (z_Q[izp, ihz, ihzp, iz, jz] *
                                  zπ_Q[ihzp, izp, jzp] * Q_hz[ihz, jhz] *
                                  Q_hc[ihc, jhc] * Q_hzπ[ihzp, jhzp] *
                                  Q_hλ[ihl, jhl])
                            acc += wθ[jz, jzp, jhz, jhc, jhzp, jhl] * \
                                pr * col[jhl] * row
                out[iz, izp, ihz, ihc, ihzp, ihl] = 1 + β * acc ** (1 / θ)
    return out


# --------------------------------------------------------------------------
# Factor-structured ("Kronecker") operators: the reference's multi-index form,
# evaluated by sum-factorisation.  P never materialised.
# --------------------------------------------------------------------------
class KronSSY:
    """H = diag(a_row) (Q_lam (x) Q_c (x) B) diag(a_col), state order (l,k,i,j)."""

    def __init__(self, shapes, params, arrays):
        self.shapes = tuple(shapes)
        β, γ, ψ, μ_c = params[0], params[1], params[2], params[3]
        (self.h_λ, self.Q_λ, self.h_c, self.Q_c, self.h_z, self.Q_hz,
         self.z, self.z_Q, self.σ_c, self.σ_z) = [np.asarray(a) for a in arrays]
        self.β, self.γ, self.μ_c = β, γ, μ_c
        self.θ = _theta(γ, ψ)
        L, K, I, J = self.shapes
        a1 = np.exp(self.θ * self.h_λ)
        a2 = np.exp(0.5 * ((1 - γ) * self.σ_c) ** 2)
        a3 = np.exp((1 - γ) * (μ_c + self.z))
        self.a_col = np.broadcast_to(a1[:, None, None, None], self.shapes).copy()
        self.a_row = np.broadcast_to(a2[None, :, None, None] * a3[None, None],
                                     self.shapes).copy()

    def P_apply(self, V):
        """(P V)[l,k,i,j] for V indexed by next-period (l',k',i',j')."""
        R = np.einsum('ab,lkbj->lkaj', self.Q_hz, V)        # contract i'
        U = np.einsum('ijq,lkiq->lkij', self.z_Q, R)        # contract j'
        U = np.einsum('ab,lbij->laij', self.Q_c, U)         # contract k'
        U = np.einsum('ab,bkij->akij', self.Q_λ, U)         # contract l'
        return U

    def s(self, w):
        return self.a_row * self.P_apply(self.a_col * w ** self.θ)

    def T(self, w):
        return 1 + self.β * self.s(w) ** (1 / self.θ)

    def jvp(self, w, v):
        s = self.s(w)
        d = self.β * self.a_row * s ** ((1 - self.θ) / self.θ)
        return d * self.P_apply(self.a_col * w ** (self.θ - 1) * v)


class KronGCY:
    """State order (z, z_pi, h_z, h_c, h_zpi, h_lam); see gcy_wc_ratio.py:230."""

    def __init__(self, shapes, params, arrays):
        self.shapes = tuple(shapes)
        β, ψ, γ, μ_c = params[0], params[1], params[2], params[5]
        (self.z, self.z_Q, self.zπ, self.zπ_Q, self.h_z, self.Q_hz, self.σ_z,
         self.h_c, self.Q_hc, self.σ_c, self.h_zπ, self.Q_hzπ, self.σ_zπ,
         self.h_λ, self.Q_hλ) = [np.asarray(a) for a in arrays]
        self.β, self.γ, self.μ_c = β, γ, μ_c
        self.θ = _theta(γ, ψ)
        a1 = np.exp(self.θ * self.h_λ)
        a2 = np.exp(0.5 * ((1 - γ) * self.σ_c) ** 2)
        # z[i_zpi, i_hz, i_hzpi, i_z] -> axes (z, zpi, hz, hzpi)
        a3 = np.exp((1 - γ) * (μ_c + np.transpose(self.z, (3, 0, 1, 2))))
        sh = self.shapes
        self.a_col = np.broadcast_to(a1[None, None, None, None, None, :], sh).copy()
        self.a_row = np.broadcast_to(
            a3[:, :, :, None, :, None] * a2[None, None, None, :, None, None],
            sh).copy()

    def P_apply(self, V):
        # V[jz, jzp, jhz, jhc, jhzp, jhl]; contract the h modes first, then
        # z_pi (kernel indexed by current h_zpi, z_pi), then z.
        U = np.einsum('fF,abcdeF->abcdef', self.Q_hλ, V)
        U = np.einsum('eE,abcdEf->abcdef', self.Q_hzπ, U)
        U = np.einsum('dD,abcDef->abcdef', self.Q_hc, U)
        U = np.einsum('cC,abCdef->abcdef', self.Q_hz, U)
        U = np.einsum('ebB,aBcdef->abcdef', self.zπ_Q, U)
        U = np.einsum('bceaA,Abcdef->abcdef', self.z_Q, U)
        return U

    s = KronSSY.s
    T = KronSSY.T
    jvp = KronSSY.jvp


def T_ssy(w, shapes, params, arrays):
    return KronSSY(shapes, params, arrays).T(np.asarray(w, dtype=np.float64))


def T_gcy(w, shapes, params, arrays):
    return KronGCY(shapes, params, arrays).T(np.asarray(w, dtype=np.float64))


# --------------------------------------------------------------------------
# Dense single-index form: P (N x N), a_row, a_col, C-order flattening
# --------------------------------------------------------------------------
def dense_ssy(shapes, params, arrays):
    """Returns (P, a_row, a_col, beta, theta) with n = ((l*K+k)*I+i)*J+j."""
    op = KronSSY(shapes, params, arrays)
    L, K, I, J = shapes
    N = L * K * I * J
    B = np.einsum('ia,ijb->ijab', op.Q_hz, op.z_Q).reshape(I * J, I * J)
    P = np.kron(op.Q_λ, np.kron(op.Q_c, B))
    assert P.shape == (N, N)
    return P, op.a_row.reshape(N), op.a_col.reshape(N), op.β, op.θ


def dense_gcy(shapes, params, arrays):
    op = KronGCY(shapes, params, arrays)
    N = int(np.prod(shapes))
    P = np.einsum('bceaA,ebB,cC,dD,eE,fF->abcdefABCDEF', op.z_Q, op.zπ_Q,
                  op.Q_hz, op.Q_hc, op.Q_hzπ, op.Q_hλ,
                  optimize=True).reshape(N, N)
    return P, op.a_row.reshape(N), op.a_col.reshape(N), op.β, op.θ


def dense_T(w, P, a_row, a_col, β, θ):
    """temp_ssy.py:153-159 with H = diag(a_row) P diag(a_col)."""
    return 1 + β * (a_row * (P @ (a_col * w ** θ))) ** (1 / θ)


def dense_jvp(w, v, P, a_row, a_col, β, θ):
    """J_T(w) v from temp_ssy.py:204-216 (without the '- I')."""
    s = a_row * (P @ (a_col * w ** θ))
    return β * s ** ((1 - θ) / θ) * a_row * (P @ (a_col * w ** (θ - 1) * v))


# --------------------------------------------------------------------------
# Reference-faithful broadcast form restricted to a slab of current states.
# This is the arithmetic the reference actually performs (ssy_wc_ratio.py:116-148:
# a 2D-axis broadcast product H followed by a sum over the next-period axes), i.e.
# N^2 multiply-adds per evaluation with no sum-factorisation.  Used as the CPU
# baseline of bench.py on a bounded sample of (l, k) slabs.
# --------------------------------------------------------------------------
def ssy_broadcast_slab(op, w, l, k):
    """Tw[l, k, :, :] computed the reference's way for one (l, k) pair."""
    L, K, I, J = op.shapes
    γ, θ = op.γ, op.θ
    A1 = np.exp(θ * op.h_λ)[None, None, :, None, None, None]              # jn_h_lam
    A2 = np.exp(0.5 * ((1 - γ) * op.σ_c[k]) ** 2)                          # n_h_c (scalar here)
    A3 = np.exp((1 - γ) * (op.μ_c + op.z))[:, :, None, None, None, None]   # n_h_z, n_z
    Qλ = op.Q_λ[l][None, None, :, None, None, None]
    Qc = op.Q_c[k][None, None, None, :, None, None]
    Qhz = op.Q_hz[:, None, None, None, :, None]
    zQ = op.z_Q[:, :, None, None, None, :]
    H = A1 * A2 * A3 * Qλ * Qc * Qhz * zQ                                  # (I, J, L, K, I, J)
    Hwθ = np.sum(w[None, None] ** θ * H, axis=(2, 3, 4, 5))
    return 1 + op.β * Hwθ ** (1 / θ)
