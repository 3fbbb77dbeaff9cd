"""Analytic test function behind tests/golden/sobol_reference.npz (shared by the generator script and the tests)."""
import numpy as np


def analytic_function(x):
    """Smooth vector-valued function on [0,1]^3 -> 4 outputs in which every input has a first-order effect on every
    output (a zero effect makes the clamped bootstrap statistic degenerate and scipy's BCa limits NaN)."""
    z = 2.0 * np.pi * (x - 0.5)
    y0 = np.sin(z[:, 0]) + 0.7 * np.sin(z[:, 1]) ** 2 + 0.4 * z[:, 2] + 0.05 * z[:, 2] ** 2 * np.sin(z[:, 0]) + 0.8 * x[:, 1]
    y1 = x[:, 0] + 0.5 * x[:, 1] * x[:, 2] + 0.3 * x[:, 2] + 0.2 * x[:, 1]
    y2 = np.exp(-x[:, 0]) * (1.0 + 0.5 * x[:, 2]) + x[:, 1] ** 2
    y3 = 0.3 * y0 - y1 + 0.4 * x[:, 2] ** 2
    return np.stack([y0, y1, y2, y3], axis=1)
