"""SSY (Schorfheide-Song-Yaron) parameter object: host mirror of the reference's ``SSY``
(/root/reference/code/ssy/ssy_model.py:50-81) -- same keyword names, defaults (Table VII of the
paper), ``.θ`` and ``.params`` order.  State x = (h_λ, h_c, h_z, z), indexed (l, k, i, j).
"""
import math

from ._params import ParameterSet

_SIGMA_BAR = 0.0035


class SSY(ParameterSet):
    _TABLE = (
        ("β", 0.999), ("γ", 8.89), ("ψ", 1.97),
        ("ρ", 0.987), ("ρ_z", 0.992), ("ρ_c", 0.991), ("ρ_λ", 0.959),
        ("s_z", math.sqrt(0.0039)), ("s_c", math.sqrt(0.0096)), ("s_λ", 0.0004),
        ("μ_c", 0.0016),
        ("φ_z", 0.215 * _SIGMA_BAR * math.sqrt(1 - 0.987 ** 2)),
        ("φ_c", 1.00 * _SIGMA_BAR),
    )
    _PARAMS_ORDER = ("β", "γ", "ψ", "μ_c", "ρ", "φ_z", "φ_c", "ρ_z", "ρ_c", "ρ_λ", "s_z", "s_c", "s_λ")

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.θ = (1 - self.γ) / (1 - 1 / self.ψ)


def wc_loglinear_factory(ssy):
    """Constants of the log-linear approximation of the W/C ratio and a function evaluating it
    (log w) at a state (h_λ, h_c, h_z, z): mirror of ssy_model.py:86-156."""
    from .loglinear import ssy_factory
    return ssy_factory(ssy)
