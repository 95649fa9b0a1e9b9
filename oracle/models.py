"""Oracle restatement of the model parameter objects (test infrastructure only).

Follows /root/reference/code/ssy/ssy_model.py:50-81 (SSY) and
/root/reference/code/gcy/gcy_model.py:43-75 (GCY): same keyword names, same
defaults, same ``.params`` ordering.  Pinned by tests/golden/model_defaults.json,
which is generated from the importable reference files.
"""
import math


class SSY:
    """Schorfheide-Song-Yaron parameters; state (h_lam, h_c, h_z, z)."""

    def __init__(self, **kw):
        d = dict(β=0.999, γ=8.89, ψ=1.97, ρ=0.987, ρ_z=0.992, ρ_c=0.991,
                 ρ_λ=0.959, s_z=math.sqrt(0.0039), s_c=math.sqrt(0.0096),
                 s_λ=0.0004, μ_c=0.0016,
                 ϕ_z=0.215 * 0.0035 * math.sqrt(1 - 0.987 ** 2),
                 ϕ_c=1.00 * 0.0035)
        for k, v in kw.items():
            import unicodedata
            k = unicodedata.normalize("NFKC", k)
            if k not in d:
                raise TypeError(f"unexpected SSY parameter {k!r}")
            d[k] = v
        self.__dict__.update(d)
        self.θ = (1 - self.γ) / (1 - 1 / self.ψ)
        self.params = (self.β, self.γ, self.ψ, self.μ_c, self.ρ, self.ϕ_z,
                       self.ϕ_c, self.ρ_z, self.ρ_c, self.ρ_λ, self.s_z,
                       self.s_c, self.s_λ)


class GCY:
    """Gomez-Cram-Yaron parameters; state (z, z_pi, h_z, h_c, h_zpi, h_lam)."""

    def __init__(self, **kw):
        d = dict(β=0.9987, ψ=1.5, γ=13.01, ρ_λ=0.981, s_λ=0.12 * 0.0015,
                 μ_c=0.0016, φ_c=0.0015, ρ=0.983, ρ_π=-0.0075,
                 φ_z=0.13 * 0.0015, ρ_c=0.992, s_c=0.104, ρ_z=0.980, s_z=0.09,
                 ρ_ππ=0.985, φ_zπ=0.08 * 0.0015, ρ_zπ=0.970, s_zπ=0.271)
        for k, v in kw.items():
            if k not in d:
                raise TypeError(f"unexpected GCY parameter {k!r}")
            d[k] = v
        self.__dict__.update(d)
        self.params = (self.β, self.ψ, self.γ, self.ρ_λ, self.s_λ, self.μ_c,
                       self.φ_c, self.ρ, self.ρ_π, self.φ_z, self.ρ_c,
                       self.s_c, self.ρ_z, self.s_z, self.ρ_ππ, self.φ_zπ,
                       self.ρ_zπ, self.s_zπ)

    @property
    def θ(self):      # the reference GCY has no .θ; provided for convenience
        return (1 - self.γ) / (1 - 1 / self.ψ)
