"""Test infrastructure (not product code): derives the coefficients of the tanh-form fit to the exact
erf GELU (nn.GELU() at mirage/utils.py:143,156) used by the GEMM epilogue (csrc/common.cuh gelu_fast)
and reports its maximum absolute error.  Run: python oracle/fit_gelu.py"""
import numpy as np
from scipy.optimize import least_squares
from scipy.special import erf


def gelu_exact(x):
    return 0.5 * x * (1.0 + erf(x / np.sqrt(2.0)))


def gelu_fit(p, x):
    x2 = np.minimum(x * x, 64.0)
    return 0.5 * x * (1.0 + np.tanh(x * (p[0] + p[1] * x2 + p[2] * x2 * x2)))


if __name__ == "__main__":
    x = np.linspace(-8, 8, 200001)
    r = least_squares(lambda p: gelu_fit(p, x) - gelu_exact(x), [0.7978845608, 0.0356774, 0.0])
    xx = np.linspace(-30, 30, 600001)
    print("a, b, c =", r.x)
    print("max |err| on [-30, 30]:", np.abs(gelu_fit(r.x, xx) - gelu_exact(xx)).max())
