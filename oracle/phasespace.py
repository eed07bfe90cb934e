"""TEST INFRASTRUCTURE -- CPU restatement (numpy) of the reference's phasespace deposit,
spec/diagnostics/os-spec-phasespace.f03, statement by statement and in the reference's precisions
(p_diag_prec = single, dutil/os-dutil.f03:26; the particle kind for the products the reference forms in it):

  * get_position        spec/os-spec-define.f03:1098-1123  (position_idx_comp_4 / position_comp_4)
  * get_l               os-spec-phasespace.f03:314-409
  * axis quantity       get_phasespace_axis_spec :3196-3238
  * deposited quantity  get_phasespace_charge :2627-2737
  * deposit_1D :1633-1689, deposit_2D :1693-1775, deposit_3D :1779-1902

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module.  Parity unpinned by a
reference binary (Fortran, not compilable here) and the reference holds no golden vector for it: pinned by the
invariants of tests/test_phasespace_cpu.py (total deposited quantity, first moment, hand-computed cases).

Particles are ix[np, dim] (1-based local cell), x[np, dim], p[np, 3], q[np]; the phasespace is returned as a
float32 array [(n3,) n2, n1] (first axis fastest, the reference's buffer order) or [n1].
"""
import numpy as np

F = np.float32
AXIS = {"x": 1, "p": 2, "g": 3, "gl": 4, "l": 5, "kel": 6}
TYPE = {"charge": 0, "m": 1, "ene": 2, "q1": 3, "q2": 4, "q3": 5, "|charge|": 6, "j1": 7, "j2": 8, "j3": 9}


def nint(v):
    """Fortran nint: nearest integer, halves away from zero"""
    return np.where(v >= 0, np.floor(v + 0.5), -np.floor(-v + 0.5)).astype(np.int64)


def position(comp, ix, x, g_box_lo, dx, my_nx_p_lo):
    """pos = xmin + dx * real((ix + ixmin) + x, p_single), ixmin = my_nx_p(1, comp) - 2"""
    T = x.dtype.type
    ixmin = int(my_nx_p_lo[comp]) - 2
    inner = (ix[:, comp] + ixmin).astype(T) + x[:, comp]
    return F(g_box_lo[comp]) + F(dx[comp]) * inner.astype(F)


def ang_momentum(i, dim, ix, x, p, g_box, dx, my_nx_p_lo):
    """get_l: component i (0-based) of (x - centre) x p with the absent coordinates dropped"""
    center = [F((g_box[0][d] + g_box[1][d]) / 2) for d in range(dim)]
    arm = lambda cx, cp: (position(cx, ix, x, g_box[0], dx, my_nx_p_lo) - center[cx]) * p[:, cp].astype(F)
    if dim == 2:
        if i == 0:
            return arm(1, 2)
        if i == 1:
            return -arm(0, 2)
        return arm(0, 1) - arm(1, 0)
    if i == 0:
        return arm(1, 2) - arm(2, 1)
    if i == 1:
        return -arm(0, 2) + arm(2, 0)
    return arm(0, 1) - arm(1, 0)


def axis_value(kind, comp, dim, ix, x, p, g_box, dx, my_nx_p_lo):
    T = p.dtype.type
    k = AXIS[kind]
    if k == 1:
        return position(comp - 1, ix, x, g_box[0], dx, my_nx_p_lo)
    if k == 2:
        return p[:, comp - 1].astype(F)
    if k in (3, 4):
        g = np.sqrt((p[:, 0] ** 2 + p[:, 1] ** 2 + p[:, 2] ** 2 + T(1)).astype(F))
        return g if k == 3 else np.log10(g)
    if k == 5:
        return ang_momentum(comp - 1, dim, ix, x, p, g_box, dx, my_nx_p_lo)
    u2 = (p[:, 0] ** 2 + p[:, 1] ** 2) + p[:, 2] ** 2
    ke = (u2 / (np.sqrt(T(1) + u2) + T(1))).astype(F)
    return np.log10(np.maximum(ke, np.finfo(F).tiny))


def quantity(ps_type, p, q, rqm, ene_bin=None):
    T = p.dtype.type
    t = TYPE[ps_type]
    u2 = (p[:, 0] ** 2 + p[:, 1] ** 2 + p[:, 2] ** 2).astype(F)
    gamma = np.sqrt(u2 + F(1))
    if t == 0:
        c = q.astype(F)
    elif t == 1:
        c = (T(rqm) * q).astype(F)
    elif t == 2:
        c = (T(rqm) * q).astype(F) * u2 / (gamma + F(1))
    elif t <= 5:
        kin = (T(rqm) * q).astype(F) * u2 / (gamma + F(1))
        c = kin * p[:, t - 3].astype(F) / gamma
    elif t == 6:
        c = np.abs(q.astype(F))
    else:
        c = (q * p[:, t - 7]).astype(F) / gamma
    if ene_bin is not None:
        kin = u2 / (gamma + F(1))
        c = np.where((kin <= F(ene_bin[0])) | (kin > F(ene_bin[1])), F(0), c)
    return c.astype(F)


def deposit(axes_values, lq, n, lo, hi, T, buffer=None):
    """deposit_1D / _2D / _3D with ixmin = 1, ixmax = n (the whole phasespace on this rank).  The sums run
    in particle order like the reference's loop (float32 accumulation), so this is O(np) python: small cases."""
    nd = len(axes_values)
    buf = np.zeros(tuple(reversed(n[:nd])), F) if buffer is None else buffer
    idx, eps = [], []
    for k in range(nd):
        rdx = T(n[k]) / (T(hi[k]) - T(lo[k]))
        nx1 = (axes_values[k].astype(T) - T(lo[k])) * rdx
        i1 = nint(nx1)
        idx.append(i1)
        eps.append((nx1 - i1.astype(T)).astype(F) + F(0.5))
    one = F(1)
    for l in range(len(lq)):
        if nd == 1:
            i = int(idx[0][l])
            if 1 <= i <= n[0]:
                buf[i - 1] += (one - eps[0][l]) * lq[l]
            if 1 <= i + 1 <= n[0]:
                buf[i] += eps[0][l] * lq[l]
        elif nd == 2:
            i, j = int(idx[0][l]), int(idx[1][l])
            e1, e2 = eps[0][l], eps[1][l]
            if 1 <= j <= n[1]:
                if 1 <= i <= n[0]:
                    buf[j - 1, i - 1] += (one - e1) * (one - e2) * lq[l]
                if 1 <= i + 1 <= n[0]:
                    buf[j - 1, i] += e1 * (one - e2) * lq[l]
            if 1 <= j + 1 <= n[1]:
                if 1 <= i <= n[0]:
                    buf[j, i - 1] += (one - e1) * e2 * lq[l]
                if 1 <= i + 1 <= n[0]:
                    buf[j, i] += e1 * e2 * lq[l]
        else:  # deposit_3D (:1779-1902)
            i, j, k = int(idx[0][l]), int(idx[1][l]), int(idx[2][l])
            e1, e2, e3 = eps[0][l], eps[1][l], eps[2][l]
            for kk, w3 in ((k, one - e3), (k + 1, e3)):
                if not 1 <= kk <= n[2]:
                    continue
                for jj, w2 in ((j, one - e2), (j + 1, e2)):
                    if not 1 <= jj <= n[1]:
                        continue
                    if 1 <= i <= n[0]:
                        buf[kk - 1, jj - 1, i - 1] += (one - e1) * w2 * w3 * lq[l]
                    if 1 <= i + 1 <= n[0]:
                        buf[kk - 1, jj - 1, i] += e1 * w2 * w3 * lq[l]
    return buf


def phasespace(axes, n, lo, hi, dim, ix, x, p, q, rqm, g_box, dx, my_nx_p_lo, ps_type="charge", ene_bin=None,
               buffer=None):
    """The particle loop of save_phasespace_1D / _2D for one rank's particles"""
    vals = [axis_value(kind, comp, dim, ix, x, p, g_box, dx, my_nx_p_lo) for kind, comp in axes]
    lq = quantity(ps_type, p, q, rqm, ene_bin)
    return deposit(vals, lq, n, lo, hi, x.dtype.type, buffer)
