"""SSY (Schorfheide-Song-Yaron) parameter object -- host mirror of
/root/reference/code/ssy/ssy_model.py:50-81: same keyword names, defaults,
``.θ`` and ``.params`` order (β, γ, ψ, μ_c, ρ, ϕ_z, ϕ_c, ρ_z, ρ_c, ρ_λ, s_z, s_c, s_λ).

State x = (h_λ, h_c, h_z, z), indexed (l, k, i, j).
"""
import numpy as np


class SSY:
    def __init__(self,
                 β=0.999, γ=8.89, ψ=1.97,
                 ρ=0.987, ρ_z=0.992, ρ_c=0.991, ρ_λ=0.959,
                 s_z=np.sqrt(0.0039), s_c=np.sqrt(0.0096), s_λ=0.0004,
                 μ_c=0.0016,
                 ϕ_z=0.215 * 0.0035 * np.sqrt(1 - 0.987 ** 2),
                 ϕ_c=1.00 * 0.0035):
        self.β, self.γ, self.ψ = β, γ, ψ
        self.μ_c, self.ϕ_z, self.ϕ_c = μ_c, ϕ_z, ϕ_c
        self.ρ, self.ρ_z, self.ρ_c, self.ρ_λ = ρ, ρ_z, ρ_c, ρ_λ
        self.s_z, self.s_c, self.s_λ = s_z, s_c, s_λ
        self.θ = (1 - γ) / (1 - 1 / ψ)
        self.params = β, γ, ψ, μ_c, ρ, ϕ_z, ϕ_c, ρ_z, ρ_c, ρ_λ, s_z, s_c, s_λ


def wc_loglinear_factory(ssy):
    """Constant terms of the log-linear approximation of the W/C ratio and a function that
    evaluates it (log w) at a state (h_λ, h_c, h_z, z) -- mirror of ssy_model.py:86-156."""
    from .loglinear import ssy_factory
    return ssy_factory(ssy)
