"""Natural cubic-spline interpolation.

The reference interpolates the 1-D Fraunhofer pattern of a cylinder with
``cora.util.cubicspline.Interpolater`` (drift/telescope/cylbeam.py:95), an external
helper; this is the textbook natural spline (second derivative zero at both ends).
"""

import numpy as np


class Interpolater:
    def __init__(self, x, y):
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        order = np.argsort(x)
        self.x, self.y = x[order], y[order]
        n = self.x.size
        h = np.diff(self.x)
        # tridiagonal system for the second derivatives m_1 .. m_{n-2}; m_0 = m_{n-1} = 0
        m = np.zeros(n)
        if n > 2:
            lower = h[:-1].copy()
            diag = 2.0 * (h[:-1] + h[1:])
            upper = h[1:].copy()
            rhs = 6.0 * (np.diff(self.y[1:]) / h[1:] - np.diff(self.y[:-1]) / h[:-1])
            # Thomas algorithm
            for i in range(1, n - 2):
                w = lower[i] / diag[i - 1]
                diag[i] -= w * upper[i - 1]
                rhs[i] -= w * rhs[i - 1]
            sol = np.zeros(n - 2)
            sol[-1] = rhs[-1] / diag[-1]
            for i in range(n - 4, -1, -1):
                sol[i] = (rhs[i] - upper[i] * sol[i + 1]) / diag[i]
            m[1:-1] = sol
        self.m = m
        self.h = h

    def __call__(self, xq):
        xq = np.asarray(xq, dtype=np.float64)
        i = np.clip(np.searchsorted(self.x, xq, side="right") - 1, 0, self.x.size - 2)
        h = self.h[i]
        a = (self.x[i + 1] - xq) / h
        b = (xq - self.x[i]) / h
        return (
            a * self.y[i] + b * self.y[i + 1]
            + ((a**3 - a) * self.m[i] + (b**3 - b) * self.m[i + 1]) * h * h / 6.0
        )
