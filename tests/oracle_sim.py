"""TEST INFRASTRUCTURE: the reference time step composed from the oracle's pieces, with the same
method names as osiris_b200.simulation.Simulation so a test can drive both side by side.
Single rank, periodic (the reference's update_periodic_vdf / periodic particle wrap)."""
import numpy as np

from oracle.bindings import Oracle, aligned


class OracleSim:
    def __init__(self, dim, order, nx, dx, dt, precision="f32", gc_emf=None, gc_jay=None, smooth=None, emf_smooth=None,
                 periodic=None, part_bc=None, window_order="device", solver=None, emf_bc=None):
        self.orc = Oracle(precision)
        self.dtype = self.orc.dtype
        self.dim, self.order, self.nx, self.dx, self.dt = dim, order, list(nx), list(dx), float(dt)
        lo, up = max(order, 1), max(order + 1, 2)
        self.gc_emf = gc_emf if gc_emf is not None else ([lo] * dim, [up] * dim)
        self.gc_jay = gc_jay if gc_jay is not None else ([lo] * dim, [up] * dim)
        sh = lambda gc: tuple(self.nx[d] + gc[0][d] + gc[1][d] for d in reversed(range(dim))) + (3,)
        self.e = np.zeros(sh(self.gc_emf), self.dtype)
        self.b = np.zeros(sh(self.gc_emf), self.dtype)
        self.j = np.zeros(sh(self.gc_jay), self.dtype)
        self.species = []
        # periodic[d]; part_bc = ([lower per axis], [upper per axis]) of "absorbing" | "specular" on the others
        self.periodic = list(periodic) if periodic is not None else [True] * dim
        self.part_bc = part_bc if part_bc is not None else (["absorbing"] * dim, ["absorbing"] * dim)
        # "device": the particles a window move pushes out meet the wall / migration at once (what
        # osgpu_move_window does); "reference": they stay in cell 0 for one more push and are removed by the
        # next update_boundary (move_window_species only lowers ix, spec/os-species.f03:171-229) -- needs one
        # more lower guard cell on the moving axis, as the reference allocates (emf/os-emf.f03:684-688)
        self.window_order = window_order
        self.smooth = smooth  # (order[dim], max_order, scoef[dim, 2 max_order + 1]) from Oracle.smooth_setup
        self.emf_smooth = emf_smooth  # the same for el_mag_fld{ smooth_type = "stand" }
        self.e_part, self.b_part = self.e, self.b
        # particles use dt rounded to their precision (spec/os-spec-vpush.f03:374)
        self.dtp = float(self.dtype(self.dt))
        self.n = 0  # n(tstep)
        self.solver = solver  # osiris_b200.solver.Solver or None (Yee)
        # ([lower per axis], [upper per axis]) of "none" | "conducting" | "reflecting" on the non-periodic axes
        self.emf_bc = emf_bc if emf_bc is not None else (["none"] * dim, ["none"] * dim)

    def add_species(self, rqm, ix, x, p, q, tag=None, udist=None):
        sp = dict(rqm=float(rqm), udist=udist, ix=np.ascontiguousarray(ix, np.int32).copy(),
                  x=np.ascontiguousarray(x, self.dtype).copy(), p=np.ascontiguousarray(p, self.dtype).copy(),
                  q=np.ascontiguousarray(q, self.dtype).copy())
        if tag is not None:
            sp["tag"] = np.asarray(tag, np.int64).copy()
        self.species.append(sp)
        return sp

    def _paxes(self):
        return [d for d in range(self.dim) if self.periodic[d]]

    def emf_update_boundary(self):
        # physical walls of a moving window (p_bc_move_c) do nothing to the fields
        self.orc.update_periodic(self.dim, self.e, self.nx, self.gc_emf, "replace", axes=self._paxes())
        self.orc.update_periodic(self.dim, self.b, self.nx, self.gc_emf, "replace", axes=self._paxes())

    def emf_update_particle_fld(self):
        """update_particle_fld (emf/os-emf.f03:1065-1150) with smooth_type "stand" """
        if self.emf_smooth is None:
            self.e_part, self.b_part = self.e, self.b
            return
        order, mo, sc = self.emf_smooth
        self.e_part, self.b_part = self.e.copy(), self.b.copy()
        self.orc.smooth(self.dim, self.e_part, self.nx, self.gc_emf, order, mo, sc)
        self.orc.smooth(self.dim, self.b_part, self.nx, self.gc_emf, order, mo, sc)

    def advance_deposit(self):
        self.j[...] = 0
        ene = []
        for sp in self.species:
            if sp["q"].shape[0] == 0:
                ene.append(0.0)
                continue
            ud = sp.get("udist")
            if ud is not None and (ud.use_accelerate(self.n) or ud.use_q_incr(self.n)):  # push_species dispatch
                u_frac, q_frac = ud.fractions(self.n)
                ene.append(self.orc.accelerate_deposit(self.dim, self.order, sp["ix"], sp["x"], sp["p"], sp["q"],
                                                       u_frac, q_frac, ud.ufl[0], ud.use_particle_uacc, self.j,
                                                       self.gc_jay, self.dx, self.dtp))
                continue
            ene.append(self.orc.advance_deposit(self.dim, self.order, sp["ix"], sp["x"], sp["p"], sp["q"],
                                                sp["rqm"], self.e_part, self.b_part, self.gc_emf, self.j, self.gc_jay,
                                                self.dx, self.dtp))
        return np.array(ene)

    def particles_update_boundary(self):
        bc = tuple(["periodic" if self.periodic[d] else self.part_bc[sd][d] for d in range(self.dim)] for sd in range(2))
        for sp in self.species:
            if all(self.periodic):
                self.orc.particles_periodic(self.dim, sp["ix"], self.nx)
                continue
            keep = self.orc.particles_bc(self.dim, sp["ix"], sp["x"], sp["p"], self.nx, bc)
            for k in ("ix", "x", "p", "q", "tag"):
                if k in sp:
                    sp[k] = np.ascontiguousarray(sp[k][keep])

    def _walls(self):
        """(axis, side, kind) of the PEC / PMC walls in the reference's loop order"""
        kinds = {"conducting": "pec", "pec": "pec", "reflecting": "pmc", "pmc": "pmc"}
        return [(d, sd, kinds[self.emf_bc[sd][d]]) for d in range(self.dim) for sd in range(2)
                if self.emf_bc[sd][d] != "none" and not self.periodic[d]]

    def current_update_boundary(self):
        from oracle import walls
        self.orc.update_periodic(self.dim, self.j, self.nx, self.gc_jay, "add", axes=self._paxes())
        for d, sd, _ in self._walls():   # update_boundary_current (current/os-current-define.f03:708-723)
            walls.current_wall(self.j, self.dim, d, sd, self.order, self.nx, self.gc_jay[0])

    def move_window(self, nmove):
        """emf%move_window + part%move_window (os-simulation.f03:529-548)"""
        self.orc.move_window_vdf(self.dim, self.b, self.nx, self.gc_emf, nmove)
        self.orc.move_window_vdf(self.dim, self.e, self.nx, self.gc_emf, nmove)
        for sp in self.species:
            sp["ix"] -= np.asarray(nmove, np.int32)[None, :self.dim]
        if self.window_order == "device":
            self.particles_update_boundary()

    def inject(self, isp, ix, x, p, q, tag=None):
        sp = self.species[isp]
        sp["ix"] = np.ascontiguousarray(np.concatenate([sp["ix"], np.asarray(ix, np.int32)]))
        for k, v in (("x", x), ("p", p), ("q", q)):
            sp[k] = np.ascontiguousarray(np.concatenate([sp[k], np.asarray(v, self.dtype)]))
        if "tag" in sp:
            sp["tag"] = np.concatenate([sp["tag"], np.asarray(tag, np.int64)])

    def current_smooth(self):
        if self.smooth is not None:
            order, mo, sc = self.smooth
            self.orc.smooth(self.dim, self.j, self.nx, self.gc_jay, order, mo, sc)

    def emf_advance(self):
        assert self.gc_emf == self.gc_jay
        if self._walls():
            # advance_yee with the wall hooks (emf/solver/os-emf-solver-yee.f03:542-568)
            from oracle import walls
            assert self.solver is None or self.solver.name == "yee"
            parity = "odd" if self.order % 2 == 1 else "even"   # init_emf_bound, os-emf-boundary.f90:490-511
            def hook(f, which):
                for d, sd, kind in self._walls():
                    walls.field_wall(f, which, kind, parity, self.dim, d, sd, self.nx, self.gc_emf[0])
            self.orc.dbdt(self.dim, self.b, self.e, self.nx, self.gc_emf, self.dx, self.dt / 2.0); hook(self.b, "b")
            self.orc.dedt(self.dim, self.e, self.b, self.j, self.nx, self.gc_emf, self.dx, self.dt); hook(self.e, "e")
            self.orc.dbdt(self.dim, self.b, self.e, self.nx, self.gc_emf, self.dx, self.dt / 2.0); hook(self.b, "b")
            return
        if self.solver is not None and self.solver.name == "custom":
            self.orc.fei_advance(self.dim, self.e, self.b, self.j, self.nx, self.gc_emf, self.dx, self.dt,
                                 self.solver.coef_e, self.solver.coef_b)
            return
        self.orc.yee_advance(self.dim, self.e, self.b, self.j, self.nx, self.gc_emf, self.dx, self.dt)

    def iter(self, nsteps=1):
        ene = None
        for _ in range(nsteps):
            self.emf_update_boundary()
            self.emf_update_particle_fld()
            ene = self.advance_deposit()
            self.particles_update_boundary()
            self.current_update_boundary()
            self.current_smooth()
            self.emf_advance()
            self.n += 1
        return ene

    def field_energy(self):
        return np.concatenate([self.orc.field_energy(self.dim, f, self.nx, self.gc_emf, self.dx)
                               for f in (self.e, self.b)])


def random_case(dim, order, prec, nx, ppc, seed=0, umax=0.6, gc=None, efield=0.5):
    """random particles + random fields in the reference layouts"""
    dt_ = np.float32 if prec == "f32" else np.float64
    rng = np.random.default_rng(seed)
    lo, up = max(order, 1), max(order + 1, 2)
    gc = gc if gc is not None else ([lo] * dim, [up] * dim)
    shape = tuple(nx[d] + gc[0][d] + gc[1][d] for d in reversed(range(dim))) + (3,)
    e = (rng.standard_normal(shape) * efield).astype(dt_)
    b = (rng.standard_normal(shape) * efield).astype(dt_)
    n = int(np.prod(nx)) * ppc
    ix = np.stack([rng.integers(1, nx[d] + 1, n) for d in range(dim)], -1).astype(np.int32)
    x = (rng.random((n, dim)) - 0.5).astype(dt_)
    x[x >= 0.5] = dt_(0.49)
    u = ((rng.random((n, 3)) * 2 - 1) * umax * 3).astype(dt_)
    q = (-(rng.random(n) + 0.5) / ppc).astype(dt_)
    return dict(nx=list(nx), gc=gc, e=e, b=b, ix=ix, x=x, u=u, q=q)
