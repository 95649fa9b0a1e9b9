"""GCY (Gomez-Cram-Yaron) parameter object: host mirror of the reference's ``GCY``
(/root/reference/code/gcy/gcy_model.py:43-75) -- same keyword names, defaults and ``.params``
order (ψ before γ, unlike SSY); like the reference it has no ``.θ`` attribute.
State x = (z, z_π, h_z, h_c, h_zπ, h_λ).
"""
from ._params import ParameterSet

_SIGMA = 0.0015


class GCY(ParameterSet):
    _TABLE = (
        ("β", 0.9987), ("ψ", 1.5), ("γ", 13.01),
        ("ρ_λ", 0.981), ("s_λ", 0.12 * _SIGMA),
        ("μ_c", 0.0016), ("φ_c", _SIGMA),
        ("ρ", 0.983), ("ρ_π", -0.0075), ("φ_z", 0.13 * _SIGMA),
        ("ρ_c", 0.992), ("s_c", 0.104), ("ρ_z", 0.980), ("s_z", 0.09),
        ("ρ_ππ", 0.985), ("φ_zπ", 0.08 * _SIGMA), ("ρ_zπ", 0.970), ("s_zπ", 0.271),
    )
    _PARAMS_ORDER = tuple(name for name, _ in _TABLE)


def wc_loglinear_factory(gcy):
    """Constants of the log-linear approximation of the W/C ratio and a function evaluating it
    (log w) at a state (h_λ, h_c, h_z, h_zπ, z, z_π): mirror of gcy_model.py:80-159."""
    from .loglinear import gcy_factory
    return gcy_factory(gcy)
