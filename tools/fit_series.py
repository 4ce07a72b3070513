"""Derives the constants of svgf_device.cuh economised_series3 (three-term normal-weight series).

The reference's normal weight is pow(sat(n.n'), phiN) (src/Filter.cuh:407-427) = 2^-E, E = -phiN log2(1-u), u = 1-d.
Truncating the series of E after u^3 leaves c u^4/4 (c = phiN log2 e).  In s = c ln2 u the weight is ~e^-s, so the
weighted minimax problem  min_{a,b} max_s e^-s |s^4 - a s^3 - b s^2|  is independent of phiN; its solution is folded
into the u^3 and u^2 coefficients.  Prints (alpha, beta) fitted per phiN and the resulting max weight error.
"""
import numpy as np
from scipy.optimize import linprog


def fit3(phiN):
    c = phiN * np.log2(np.e); L = np.log(2)
    s = np.linspace(0, 40, 8001)[1:]
    u = s / (c * L)
    E = -phiN * np.log2(1 - u); w = 2.0 ** (-E)
    base = c * u + c * u ** 2 / 2 + c * u ** 3 / 3
    r0 = w * L * (E - base) * 1e6
    A = np.stack([w * L * s ** 2, w * L * s ** 3], 1) * 1e6
    m = len(s)
    Aub = np.block([[A, -np.ones((m, 1))], [-A, -np.ones((m, 1))]])
    r = linprog([0, 0, 1], A_ub=Aub, b_ub=np.concatenate([r0, -r0]), bounds=[(None, None)] * 3, method="highs-ipm")
    K = 4 * c ** 3 * L ** 4
    return r.x[0] * K, r.x[1] * K, r.x[2] * 1e-6


def max_weight_error(phiN, k2_corr=-3.3930, k3_corr=2.0594):
    """max over u of |2^-(u P(u)) - (1-u)^phiN| for the shipped constants (float64 evaluation)."""
    c = phiN * np.log2(np.e)
    u = np.linspace(0, 1, 2_000_001)[:-1]
    exact = (1 - u) ** phiN
    approx = 2.0 ** (-(u * (c + u * ((c / 2 + k2_corr / c) + u * (c / 3 + k3_corr)))))
    return float(np.max(np.abs(approx - exact)))


if __name__ == "__main__":
    for phiN in (100, 128, 256, 1024):
        beta, alpha, t = fit3(phiN)
        print(f"phiN {phiN}: beta {beta:.4f} alpha {alpha:.4f} minimax weight error {t:.3e}; shipped constants: {max_weight_error(phiN):.3e}")
