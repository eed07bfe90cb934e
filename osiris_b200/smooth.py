"""Host mirror of t_smooth (source/vdf/os-vdf-smooth.f90): the `smooth{}` section of a deck and the
filter kernels `setup_smooth` derives from it.  The filtering itself runs on the GPU
(osgpu_current_smooth, csrc/fields.cu); this module only does what the Fortran host keeps doing
in a drop-in build -- parse type/order and build scoef -- and hands the result to
osgpu_current_set_smooth.

    smooth { type(1:2) = "5pass", "binomial", order(2) = 2 }   ->   Smooth(2, type=["5pass", "binomial"], order=[0, 2])
"""
import math

import numpy as np

P_ORDER_MAX = 15  # os-vdf-smooth.f90:34
TYPES = ("none", "custom", "binomial", "compensated", "5pass", "tri6", "tri8", "digital", "laserkdx")


def combine_kernel3(kern3, n, nmax):
    """n 3-point kernels -> one (2 nmax + 1)-point kernel (os-vdf-smooth.f90:155-179), double."""
    out = [0.0] * (2 * nmax + 1)
    out[nmax] = 1.0
    for k in range(1, n + 1):
        f1 = 0.0
        for i in range((nmax + 1) - (k - 1), (nmax + 1) + k + 1):  # Fortran index i <-> out[i - 1]
            ci = out[i - 1]
            out[i - 2] = kern3[k - 1][2] * ci + out[i - 2]
            f2 = kern3[k - 1][0] * ci
            out[i - 1] = kern3[k - 1][1] * ci + f1
            f1 = f2
    return out


def besseli0(x):
    """besseli0_rdoub (os-math.f90:450-475); literals without a `d` exponent are single precision."""
    f = lambda v: float(np.float32(v))
    ax = abs(x)
    if ax < 3.75:
        y = (x / 3.75) ** 2
        return 1.0 + y * (f(3.5156229) + y * (f(3.0899424) + y * (f(1.2067492) + y * (f(0.2659732) + y * (
            0.360768e-1 + y * 0.45813e-2)))))
    y = 3.75 / ax
    return (math.exp(ax) / math.sqrt(ax)) * (f(0.39894228) + y * (0.1328592e-1 + y * (0.225319e-2 + y * (
        -0.157565e-2 + y * (0.916281e-2 + y * (-0.2057706e-1 + y * (0.2635537e-1 + y * (
            -0.1647633e-1 + y * 0.392377e-2))))))))


def digital_filter(fc, A, n):
    """Kaiser-window low-pass (os-vdf-smooth.f90:117-148)."""
    if A <= 21.0:
        alpha = 0.0
    elif A >= 50.0:
        alpha = 0.1102 * (A - 8.7)
    else:
        alpha = 0.5842 * ((A - 21.0) ** 0.4) + 0.07886 * (A - 21.0)
    out = [0.0] * (2 * n + 1)
    out[n] = fc
    for i in range(1, n + 1):
        r = np.float32(i) / np.float32(n)  # real(i)/n: default real
        v = besseli0(alpha * math.sqrt(1.0 - float(r * r))) / besseli0(alpha) * math.sin(i * math.pi * fc) / (
            i * math.pi)
        out[n + i] = v
        out[n - i] = v
    return out


class Smooth:
    """t_smooth after read_nml_smooth + setup_smooth.

    type[d], order[d] per axis as in the deck; swfj[d][k] = (w-, w0, w+) 3-point kernels for
    type "custom"; digital_fc / digital_A for "digital"; type[0] == "laserkdx" with
    laserkdx_omega0, laserkdx_dir (1-based) for the laser-compensated filter."""

    def __init__(self, x_dim, type="none", order=0, swfj=None, digital_fc=0.5, digital_A=60.0,
                 laserkdx_omega0=10.0, laserkdx_dir=1):
        as_list = lambda v: list(v) if isinstance(v, (list, tuple)) else [v] * x_dim
        self.x_dim = x_dim
        types, orders = [str(t) for t in as_list(type)], [int(o) for o in as_list(order)]
        self.swfj = swfj
        self.digital_fc, self.digital_A = as_list(digital_fc), as_list(digital_A)
        self.laserkdx_omega0, self.laserkdx_dir = float(laserkdx_omega0), int(laserkdx_dir)
        self.type, self.order = [], []
        if types[0] == "laserkdx":  # sets the filter in all directions (:247-262)
            if not 1 <= self.laserkdx_dir <= x_dim:
                raise ValueError("laserkdx_dir must be a valid direction")
            for d in range(x_dim):
                self.type.append("laserkdx")
                self.order.append(2 if d == self.laserkdx_dir - 1 else 0)
        else:
            for t, o in zip(types, orders):  # (:276-341)
                if t == "none":
                    o = 0
                elif t in ("custom", "binomial"):
                    if o < 1 or o > P_ORDER_MAX:
                        raise ValueError("order must be >= 1 and <= p_order_max when using custom smoothing.")
                elif t == "compensated":
                    o = max(o, 2)
                elif t == "5pass":
                    o = 5
                elif t == "tri6":
                    o = 3
                elif t == "tri8":
                    o = 4
                elif t == "digital":
                    if o < 1:
                        raise ValueError("order must be >= 1 when using digital smoothing.")
                else:
                    raise ValueError(f'Invalid smooth type : "{t}"')
                if o > P_ORDER_MAX:
                    raise ValueError("order exceeds p_order_max")
                self.type.append(t)
                self.order.append(o)
        self.max_order = max(self.order)
        self.scoef = None

    def if_smooth(self):
        return self.max_order > 0

    def smooth_order(self):  # os-vdf-smooth.f90:474-485
        return self.max_order

    def setup(self, dx, gamma=1.0, dtype=np.float32):
        """scoef[x_dim][2 max_order + 1] in field precision, normalised (os-vdf-smooth.f90:373-467)."""
        m, W = self.max_order, 2 * self.max_order + 1
        sc = np.zeros((self.x_dim, W), dtype)
        B = [1.0, 2.0, 1.0]
        for d in range(self.x_dim):
            n, t = self.order[d], self.type[d]
            if n == 0:
                continue
            k3 = None
            if t == "laserkdx":
                kdx = self.laserkdx_omega0 * dx[self.laserkdx_dir - 1] / (gamma + math.sqrt(gamma ** 2 - 1))
                k3 = [B, [-1.0, 4 + 2 * math.cos(kdx), -1.0]]
            elif t == "custom":
                k3 = [list(self.swfj[d][j]) for j in range(n)]
            elif t == "binomial":
                k3 = [B] * n
            elif t == "compensated":
                comp = float((np.float32(4.0) + np.float32(2.0) * np.float32(n - 1)) / np.float32(n - 1))
                k3 = [B] * (n - 1) + [[-1.0, comp, -1.0]]
            elif t == "5pass":
                k3 = [B] * 4 + [[-5.0, 15.0, -5.0]]
            if k3 is not None:
                sc[d, :] = np.asarray(combine_kernel3(k3, n, m), np.float64).astype(dtype)
            elif t == "tri6":
                sc[d, m - 3:m + 4] = [0.015625, -0.09375, 0.234375, 0.6875, 0.234375, -0.09375, 0.015625]
            elif t == "tri8":
                sc[d, m - 4:m + 5] = [-0.00390625, 0.03125, -0.109375, 0.21875, 0.7265625, 0.21875, -0.109375,
                                      0.03125, -0.00390625]
            elif t == "digital":
                sc[d, m - n:m + n + 1] = np.asarray(digital_filter(self.digital_fc[d], self.digital_A[d], n),
                                                    np.float64).astype(dtype)
            total = 0.0
            for j in range(m - n, m + n + 1):
                total = total + float(sc[d, j])
            for j in range(m - n, m + n + 1):
                sc[d, j] = dtype(float(sc[d, j]) / total)
        self.scoef = sc
        return sc

    def gc_num(self, gc_std):
        """Current guard cells with smoothing: max(std, (order, order + 1)) per axis
        (current/os-current-define.f03:199-205; smooth_order() is the largest order)."""
        m = self.max_order
        lo = [max(int(v), m) for v in gc_std[0]]
        up = [max(int(v), m + 1) for v in gc_std[1]]
        return lo, up
