"""Host mirror of the field-solver choice (el_mag_fld{ solver }, emf_solver{ ... }): the Yee solver and the
customized-stencil solver of emf/solver/os-emf-solver-fei.f03.  The coefficients are host-side set-up work
(read_input_fei); the device only ever sees coef_e / coef_b (osgpu_emf_set_solver)."""
from math import factorial

import numpy as np


def std_coef(order, dtype=np.float32):
    """std_coef (os-emf-solver-fei.f03:1079-1098): coefficients of the standard FDTD stencil of even order
    `order` for the first derivative on a staggered grid, n = order / 2 values."""
    n = order // 2
    out = np.zeros(n, dtype)
    for k in range(1, n + 1):
        out[k - 1] = dtype((-1.0) ** (k + 1) * 16.0 ** (1 - n) * factorial(2 * n - 1) ** 2 /
                           ((2 * k - 1) ** 2 * factorial(n + k - 1) * factorial(n - k) * factorial(n - 1) ** 2))
    return out


class Solver:
    """t_emf_solver: name "yee" or "custom".  For "custom", type "standard" (solver_ord) or
    "customized-coef" (n_coef, coef_e, coef_b from the deck); the other types of read_input_fei (bump, xu,
    dual) produce coefficients by numerical integration on the host and arrive here as "customized-coef"."""

    def __init__(self, name="yee", type="standard", solver_ord=2, n_coef=None, coef_e=None, coef_b=None):
        self.name, self.type = name, type
        if name == "yee":
            self.n_coef, self.coef_e, self.coef_b = 1, np.ones(1), np.ones(1)
        elif name == "custom":
            if type == "standard":
                self.coef_e = std_coef(solver_ord, np.float64)
                self.coef_b = self.coef_e.copy()
                self.n_coef = len(self.coef_e)
            elif type == "customized-coef":
                self.n_coef = int(n_coef)
                self.coef_e = np.asarray(coef_e, np.float64)[:self.n_coef].copy()
                self.coef_b = np.asarray(coef_b, np.float64)[:self.n_coef].copy()
            else:
                raise ValueError(f'emf_solver type "{type}": hand its coefficients over as "customized-coef"')
        else:
            raise ValueError(f'solver "{name}" is not available on the device')

    def min_gc(self, dim):
        """min_gc_yee (os-emf-solver-yee.f03:458-468) / min_gc_fei (os-emf-solver-fei.f03:1995-2026)"""
        lo, up = [1] * dim, [2] * dim
        if self.name == "custom":
            lo[0], up[0] = 2 * self.n_coef - 1, 2 * self.n_coef
        return lo, up

    def courant(self, dx):
        """largest stable dt: test_stability_yee / test_stability_fei (std, bump, coef branch :1370-1384)"""
        dx = [float(v) for v in dx]
        if self.name == "yee":
            return float(np.sqrt(1.0 / sum(1.0 / d ** 2 for d in dx)))
        c = float(np.sum(np.abs(self.coef_e)) * np.sum(np.abs(self.coef_b))) / dx[0] ** 2
        c += sum(1.0 / d ** 2 for d in dx[1:])
        return float(np.sqrt(1.0 / c))
