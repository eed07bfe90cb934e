"""TEST INFRASTRUCTURE: the oracle time step on a domain-decomposed grid, one rank per process over
torch.distributed (gloo on CPU).  The exchange protocol is the product's: neighbours and slab
ranges come from the library's host-side helpers (osgpu_cart_neighbors, osgpu_exchange_ranges) and
the migration re-indexing convention is the one csrc/particles.cu uses (upper message re-indexed by
the sender, lower message by the receiver).  Reference: vdf/os-vdf-comm.f90:665-722,
spec/os-spec-boundary.f03:1134-1225."""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from oracle_sim import OracleSim
from osiris_b200 import lib
from osiris_b200.simulation import NodeConf


class DecompOracleSim(OracleSim):
    def __init__(self, dim, order, nx_local, dx, dt, precision, node_dims, rank, periodic=None, part_bc=None, **kw):
        super().__init__(dim, order, nx_local, dx, dt, precision, periodic=periodic, part_bc=part_bc, **kw)
        self.no_co = NodeConf(dim, node_dims, self.periodic, rank)
        self.rank = rank
        self.L = lib.load()
        self.move_pending = [0] * dim  # cells the window moved since the last E/B guard exchange

    # ---- grids ---------------------------------------------------------------------------
    def _slab(self, f, gc, axis, c0, c1):
        """view of cells c0..c1 (1-based, inclusive) along `axis`, full padded extent elsewhere"""
        ax = self.dim - 1 - axis
        sl = [slice(None)] * f.ndim
        sl[ax] = slice(c0 - 1 + gc[0][axis], c1 + gc[0][axis])
        return tuple(sl)

    def _exchange(self, fields, gc, mode, move=None):
        for d in range(self.dim):
            lo, up = self.no_co.neighbor[d]
            if lo == self.rank and up == self.rank:
                for f in fields:
                    self.orc.update_periodic(self.dim, f, self.nx, gc, "replace" if mode == 0 else "add", axes=[d])
                continue
            r = (C.c_int32 * 8)()
            self.L.osgpu_exchange_ranges_move(self.nx[d], gc[0][d], gc[1][d], mode, move[d] if move else 0, r)
            s_lo = [f[self._slab(f, gc, d, r[0], r[1])].copy() for f in fields]  # copies: packed before any unpack
            s_up = [f[self._slab(f, gc, d, r[2], r[3])].copy() for f in fields]
            if lo < 0 and up < 0:
                continue  # physical walls on both sides (p_bc_move_c and friends: nothing done here)
            reqs, hold = [], []
            for k, (a, b) in enumerate(zip(s_lo, s_up)):
                hold += [torch.from_numpy(a), torch.from_numpy(b)]
                if lo >= 0:
                    reqs.append(dist.isend(hold[-2], lo, tag=10 + k))
                if up >= 0:
                    reqs.append(dist.isend(hold[-1], up, tag=20 + k))
            for k, f in enumerate(fields):
                if up >= 0:
                    from_up = np.empty_like(f[self._slab(f, gc, d, r[6], r[7])])
                    dist.recv(torch.from_numpy(from_up), up, tag=10 + k)   # its "to lower" message
                    if mode == 0:
                        f[self._slab(f, gc, d, r[6], r[7])] = from_up
                    else:
                        f[self._slab(f, gc, d, r[6], r[7])] += from_up
                if lo >= 0:
                    from_lo = np.empty_like(f[self._slab(f, gc, d, r[4], r[5])])
                    dist.recv(torch.from_numpy(from_lo), lo, tag=20 + k)   # its "to upper" message
                    if mode == 0:
                        f[self._slab(f, gc, d, r[4], r[5])] = from_lo
                    else:
                        f[self._slab(f, gc, d, r[4], r[5])] += from_lo
            for q in reqs:
                q.wait()

    def emf_update_boundary(self):
        self._exchange([self.e, self.b], self.gc_emf, 0, move=self.move_pending)
        self.move_pending = [0] * self.dim

    def move_window(self, nmove):
        super().move_window(nmove)
        for d in range(self.dim):
            self.move_pending[d] += int(nmove[d])

    def current_update_boundary(self):
        self._exchange([self.j], self.gc_jay, 1)

    # ---- particles ------------------------------------------------------------------------
    def _wall(self, sp, d, side):
        """physical wall on `side` of axis d (process_phys_boundary, spec/os-spec-boundary.f03:1238-1325)"""
        nxd = self.nx[d]
        hit = (sp["ix"][:, d] < 1) if side == 0 else (sp["ix"][:, d] > nxd)
        if self.part_bc[side][d] == "specular":
            sp["p"][hit, d] = -sp["p"][hit, d]
            sp["x"][hit, d] = -sp["x"][hit, d]
            sp["ix"][hit, d] = nxd if side else 1
        else:  # absorbing
            for k in ("ix", "x", "p", "q", "tag"):
                sp[k] = np.ascontiguousarray(sp[k][~hit])

    def particles_update_boundary(self):
        for sp in self.species:
            for d in range(self.dim):
                lo, up = self.no_co.neighbor[d]
                nxd = self.nx[d]
                if lo == self.rank and up == self.rank:
                    i = sp["ix"][:, d]
                    i[i < 1] += nxd
                    i[i > nxd] -= nxd
                    continue
                if lo < 0:
                    self._wall(sp, d, 0)
                if up < 0:
                    self._wall(sp, d, 1)
                if lo < 0 and up < 0:
                    continue
                go_lo = (sp["ix"][:, d] < 1) if lo >= 0 else np.zeros(len(sp["q"]), bool)
                go_up = (sp["ix"][:, d] > nxd) if up >= 0 else np.zeros(len(sp["q"]), bool)
                keys = ("ix", "x", "p", "q", "tag")
                msg_lo = {k: sp[k][go_lo].copy() for k in keys}
                msg_up = {k: sp[k][go_up].copy() for k in keys}
                msg_up["ix"][:, d] -= nxd                     # re-indexed by the sender
                keep = ~(go_lo | go_up)
                for k in keys:
                    sp[k] = sp[k][keep]
                # every tensor handed to isend is kept alive until its request completes
                hold = [torch.tensor([len(msg_lo["q"])]), torch.tensor([len(msg_up["q"])])]
                reqs = []
                if lo >= 0:
                    reqs.append(dist.isend(hold[0], lo, tag=30))
                if up >= 0:
                    reqs.append(dist.isend(hold[1], up, tag=31))
                n_up, n_lo = torch.zeros(1, dtype=torch.int64), torch.zeros(1, dtype=torch.int64)
                if up >= 0:
                    dist.recv(n_up, up, tag=30)
                if lo >= 0:
                    dist.recv(n_lo, lo, tag=31)
                for q in reqs:
                    q.wait()
                reqs = []
                for kk, k in enumerate(keys):
                    hold += [torch.from_numpy(np.ascontiguousarray(msg_lo[k])),
                             torch.from_numpy(np.ascontiguousarray(msg_up[k]))]
                    if lo >= 0:
                        reqs.append(dist.isend(hold[-2], lo, tag=40 + kk))
                    if up >= 0:
                        reqs.append(dist.isend(hold[-1], up, tag=50 + kk))
                for kk, k in enumerate(keys):
                    shp = lambda n: (n,) + sp[k].shape[1:]
                    a = np.empty(shp(int(n_up)), sp[k].dtype)
                    b = np.empty(shp(int(n_lo)), sp[k].dtype)
                    if up >= 0:
                        dist.recv(torch.from_numpy(a), up, tag=40 + kk)
                    if lo >= 0:
                        dist.recv(torch.from_numpy(b), lo, tag=50 + kk)
                    if k == "ix":
                        a[:, d] += nxd                        # lower message re-indexed by the receiver
                    sp[k] = np.concatenate([sp[k], a, b])
                for q in reqs:
                    q.wait()
