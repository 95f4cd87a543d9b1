"""Fits the odd polynomial P of csrc/conv_tc.cu:gelu_tanh2: erf(x / sqrt 2) ~= tanh(x * P(x^2)), minimising the maximum
error of 0.5 x (1 + tanh(x P(x^2))) against the erf GELU (iteratively re-weighted least squares).  CPU only."""
import numpy as np
from scipy.optimize import least_squares
from scipy.special import erf

x = np.linspace(1e-3, 8, 40001)
gelu = lambda v: 0.5 * v * (1 + erf(v / np.sqrt(2)))


def model(c, v):
    v2 = v * v
    return 0.5 * v * (1 + np.tanh(v * (c[0] + c[1] * v2 + c[2] * v2 * v2)))


f = lambda c: np.concatenate([model(c, x) - gelu(x), model(c, -x) - gelu(-x)])
c = least_squares(f, [0.7978845608, 0.0356774, 0.0], xtol=1e-15, ftol=1e-15).x
for _ in range(30):
    e = np.abs(f(c))
    w = 1 + 50 * e / e.max()
    c = least_squares(lambda cc: f(cc) * w, c, xtol=1e-15, ftol=1e-15).x
print('P coefficients (x^0, x^2, x^4):', c, ' max |error| =', np.abs(f(c)).max())
