"""Rouwenhorst AR(1) discretisation (oracle; test infrastructure only).

The reference calls ``quantecon.rouwenhorst(n, rho, sigma, mu)``
(/root/reference/code/ssy/discrete/ssy_wc_ratio.py:48-50,63 and
/root/reference/code/gcy/discrete/gcy_wc_ratio.py:65-68,97,115).  quantecon is a
third-party dependency that is neither vendored under /root/reference nor
version-pinned there (the sandpit notebook's warning text implies quantecon
>= 0.7, the ``(n, rho, sigma, mu=0.)`` signature) and it is not installable in
this image, so its published algorithm is restated here:

  y_sd = sqrt(sigma^2 / (1 - rho^2));  p = q = (1 + rho)/2;  psi = y_sd*sqrt(n-1)
  states = linspace(-psi, psi, n) + mu/(1 - rho)
  Theta_2 = [[p, 1-p], [1-q, q]];  Theta_n from Theta_{n-1} by the four shifted
  embeddings (p top-left, 1-p top-right, 1-q bottom-left, q bottom-right),
  summed, interior rows halved.

Pin: only the reference's recorded Newton trace (sandpit.ipynb), see
tests/test_oracle_solvers.py::test_sandpit_trace.
"""
import numpy as np


def rouwenhorst(n, rho, sigma, mu=0.0):
    if n < 2:
        raise ValueError("The number of states must be >= 2")
    y_sd = np.sqrt(sigma ** 2 / (1 - rho ** 2))
    p = (1 + rho) / 2
    q = p
    psi = y_sd * np.sqrt(n - 1)
    states = np.linspace(-psi, psi, n)
    theta = np.array([[p, 1 - p], [1 - q, q]])
    for m in range(3, n + 1):
        nxt = np.zeros((m, m))
        nxt[:m - 1, :m - 1] += p * theta
        nxt[:m - 1, 1:] += (1 - p) * theta
        nxt[1:, :m - 1] += (1 - q) * theta
        nxt[1:, 1:] += q * theta
        nxt[1:m - 1, :] /= 2
        theta = nxt
    states = states + mu / (1 - rho)
    return states, theta
