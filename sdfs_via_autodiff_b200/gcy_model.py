"""GCY (Gomez-Cram-Yaron) parameter object -- host mirror of
/root/reference/code/gcy/gcy_model.py:43-75: same keyword names, defaults and
``.params`` order (β, ψ, γ, ρ_λ, s_λ, μ_c, φ_c, ρ, ρ_π, φ_z, ρ_c, s_c, ρ_z, s_z,
ρ_ππ, φ_zπ, ρ_zπ, s_zπ).  Like the reference it has no ``.θ`` attribute.

State x = (z, z_π, h_z, h_c, h_zπ, h_λ).
"""


class GCY:
    def __init__(self,
                 β=0.9987, ψ=1.5, γ=13.01,
                 ρ_λ=0.981, s_λ=0.12 * 0.0015,
                 μ_c=0.0016, φ_c=0.0015,
                 ρ=0.983, ρ_π=-0.0075, φ_z=0.13 * 0.0015,
                 ρ_c=0.992, s_c=0.104, ρ_z=0.980, s_z=0.09,
                 ρ_ππ=0.985, φ_zπ=0.08 * 0.0015, ρ_zπ=0.970, s_zπ=0.271):
        self.β, self.ψ, self.γ = β, ψ, γ
        self.ρ_λ, self.s_λ, self.μ_c, self.φ_c, self.ρ = ρ_λ, s_λ, μ_c, φ_c, ρ
        self.ρ_π, self.φ_z, self.ρ_c = ρ_π, φ_z, ρ_c
        self.s_c, self.ρ_z, self.s_z = s_c, ρ_z, s_z
        self.ρ_ππ, self.φ_zπ, self.ρ_zπ, self.s_zπ = ρ_ππ, φ_zπ, ρ_zπ, s_zπ
        self.params = (β, ψ, γ, ρ_λ, s_λ, μ_c, φ_c, ρ, ρ_π, φ_z, ρ_c, s_c, ρ_z, s_z,
                       ρ_ππ, φ_zπ, ρ_zπ, s_zπ)


def wc_loglinear_factory(gcy):
    """Constant terms of the log-linear approximation of the W/C ratio and a function that
    evaluates it (log w) at a state (h_λ, h_c, h_z, h_zπ, z, z_π) -- mirror of gcy_model.py:80-159."""
    from .loglinear import gcy_factory
    return gcy_factory(gcy)
