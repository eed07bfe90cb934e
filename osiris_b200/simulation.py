"""Host-side mirror of the OSIRIS objects that bound the hot path, driving the C ABI.

Names follow the reference so the parity tests read like its time step
(source/os-simulation.f03:498-666):

    sim.emf.update_boundary(); sim.part.advance_deposit(); sim.part.update_boundary();
    sim.jay.update_boundary(); sim.emf.advance()            # == sim.iter()

and the device-residency hooks keep their name (`fill_data`, spec/os-species.f03:1867-1894,
emf/os-emf.f03:1158-1184, current/os-current-define.f03:734-757).  Host arrays use the Fortran
layouts: particles x(D,np) -> numpy (np, D); grids f(3,S1,S2,S3) -> numpy (S3, S2, S1, 3).
All compute goes through libosiris_b200.so; nothing here touches particle or grid values.
"""
import ctypes as C

import numpy as np

from . import lib as _lib
from .lib import Config, check, ptr


def min_gc(order, smooth_order=0):
    """Guard cells (lower[D]-like scalar pair) the reference would allocate for E/B and J:
    particles need (order, order+1) (particles/os-particles.f03:329-346), Yee (1,2)
    (emf/solver/os-emf-solver-yee.f03:458-468), + smoothing order on both sides
    (emf/os-emf.f03:663-724, current/os-current-define.f03:192-205)."""
    out = (C.c_int32 * 2)()
    check(_lib.load().osgpu_min_gc(order, out))
    return int(out[0]) + smooth_order, int(out[1]) + smooth_order


def min_gc_axes(dim, order, smooth_order=0, if_move=None):
    """The same per axis, ([lower], [upper]): an axis the window moves along gets one more lower guard
    cell (emf/os-emf.f03:684-688, current/os-current-define.f03:194-197)."""
    lo, up = min_gc(order, smooth_order)
    mv = [bool(v) for v in (if_move if if_move is not None else [False] * dim)]
    return [lo + (1 if mv[d] else 0) for d in range(dim)], [up] * dim


class NodeConf:
    """t_node_conf (source/os-nconf.f03:45-107).  topology = "mpi" (the reference's default: ranks laid
    out as MPI_Cart_create does, last axis fastest, create_topology_mpi :506-568), "old" (x1 fastest,
    create_topology_old :574-625) or "host": the caller supplies `neighbor` (and `coords`), the way a
    Fortran host hands over t_node_conf%neighbor - 1 whatever mpi_cart_create decided."""

    def __init__(self, dim, node_number=None, if_periodic=None, rank=0, topology="mpi", neighbor=None, coords=None):
        self.dim = dim
        self.topology = {"mpi": _lib.TOPO_MPI, "old": _lib.TOPO_OLD, "host": _lib.TOPO_HOST}[topology]
        self.node_number = list(node_number) if node_number is not None else [1] * dim
        self.if_periodic = [bool(p) for p in (if_periodic if if_periodic is not None else [True] * dim)]
        self.rank = rank
        self.nranks = int(np.prod(self.node_number))
        L = _lib.load()
        dims = (C.c_int32 * 3)(*(self.node_number + [1] * (3 - dim)))
        if self.topology == _lib.TOPO_HOST:
            if neighbor is None or coords is None:
                raise ValueError('topology "host" needs the neighbor table and the coordinates')
            self.coords = [int(v) for v in coords]
            self.neighbor = [[int(v) for v in pair] for pair in neighbor]
            return
        co = (C.c_int32 * 3)()
        L.osgpu_cart_coords(self.topology, dim, dims, rank, co)
        self.coords = [int(co[d]) for d in range(dim)]
        nb = (C.c_int32 * 6)()
        per = (C.c_int32 * 3)(*([int(p) for p in self.if_periodic] + [0] * (3 - dim)))
        L.osgpu_cart_neighbors(self.topology, dim, dims, per, rank, nb)
        # neighbor(2, D): [d][0] lower, [d][1] upper
        self.neighbor = [[int(nb[2 * d]), int(nb[2 * d + 1])] for d in range(dim)]


class Space:
    """t_space (source/os-space.f90): the part the moving window needs -- which axes move, how far
    the window has travelled, and how many cells it moves this step.  The host keeps this object
    (deck section `space{ xmin, xmax, if_move }`); the device is told nx_move through
    Simulation.move_window."""

    def __init__(self, xmin, xmax, if_move=None, move_vel=0.0):
        self.x_dim = len(xmin)
        self.x_bnd_initial = [list(map(float, xmin)), list(map(float, xmax))]
        self.x_bnd = [list(map(float, xmin)), list(map(float, xmax))]
        self.if_move = [bool(v) for v in (if_move if if_move is not None else [False] * self.x_dim)]
        self.move_vel = float(move_vel)
        self.xmoved = [0.0] * self.x_dim
        self.total_i_moved = [0] * self.x_dim
        self.move_num = [0] * self.x_dim

    def move_boundary(self, dx, dt):
        """move_boundary_space_global (os-space.f90:508-562): accumulate the distance travelled at c (or
        move_vel) and move by ONE cell when it reaches dx; returns nx_move per axis."""
        for d in range(self.x_dim):
            self.move_num[d] = 0
            if not self.if_move[d]:
                continue
            self.xmoved[d] += self.move_vel * dt if self.move_vel > 0.0 else dt
            if self.xmoved[d] >= dx[d]:
                self.move_num[d] = 1
                self.xmoved[d] -= dx[d]
                self.total_i_moved[d] += 1
                for s_ in range(2):
                    self.x_bnd[s_][d] = self.x_bnd_initial[s_][d] + self.total_i_moved[d] * dx[d]
        return list(self.move_num)

    def nx_move(self):
        return list(self.move_num)


def grad_linear(n, nmax):
    """gradient_linear (os-math.f90:743-753): real(n) / real(nmax) in default (single) real"""
    return float(np.float32(n) / np.float32(nmax))


def grad_regressive(n, nmax):
    """gradient_regressive (os-math.f90:761-775): st x / sqrt((st x)^2 + 1) scaled to reach 1 at n = nmax"""
    st = 3.0
    scaling = float(np.sqrt(np.float64(np.float32(st) ** 2 + np.float32(1.0)))) / st
    x = float(np.float32(n) / np.float32(nmax))
    return st * x / np.sqrt((st * x) ** 2 + 1.0) * scaling


class UDist:
    """The part of t_udist (spec/os-spec-udist.f03) the pusher dispatch reads: beam initialisation by
    free streaming (n_accelerate) and by growing charge (n_q_incr); *_type is "linear" | "regressive"."""

    def __init__(self, ufl=(0.0, 0.0, 0.0), n_accelerate=-1, n_accelerate_type="linear", n_q_incr=-1,
                 n_q_incr_type="linear", use_particle_uacc=False):
        self.ufl = [float(v) for v in ufl]
        self.n_accelerate, self.n_accelerate_type = int(n_accelerate), n_accelerate_type
        self.n_q_incr, self.n_q_incr_type = int(n_q_incr), n_q_incr_type
        self.use_particle_uacc = bool(use_particle_uacc)

    def use_accelerate(self, n):   # use_accelerate_udist (:1233-1243)
        return n < self.n_accelerate

    def use_q_incr(self, n):       # use_q_incr_udist (:1249-1259)
        return n < self.n_q_incr

    def fractions(self, n):
        """(u_frac, q_frac) of step n as accelerate_deposit forms them (spec/os-spec-accel.f90:76-113)"""
        g = {"linear": grad_linear, "regressive": grad_regressive}
        u_frac = g[self.n_accelerate_type](n + 1, self.n_accelerate) if self.n_accelerate > 0 else 1.0
        q_frac = g[self.n_q_incr_type](n + 1, self.n_q_incr) if self.n_q_incr > 0 else 1.0
        return u_frac, q_frac


class Species:
    """t_species (spec/os-spec-define.f03:327-549) device mirror."""

    def __init__(self, sim, sp_id, name, rqm, num_par_max, udist=None):
        self.sim, self.sp_id, self.name, self.rqm = sim, sp_id, name, float(rqm)
        self.udist = udist if udist is not None else UDist()
        check(sim._L.osgpu_species_create(sim._h, sp_id, float(rqm), int(num_par_max)))

    def upload(self, ix, x, p, q, tag=None):
        dt = self.sim.dtype
        ix = np.ascontiguousarray(ix, np.int32)
        x, p, q = (np.ascontiguousarray(a, dt) for a in (x, p, q))
        tag = None if tag is None else np.ascontiguousarray(tag, np.int64)
        check(self.sim._L.osgpu_species_upload(self.sim._h, self.sp_id, q.shape[0], ptr(ix), ptr(x), ptr(p),
                                               ptr(q), ptr(tag)))

    def upload_begin(self, ix, x, p, q, tag=None):
        """Queue the host -> device copies and return; the arrays (contiguous, of the simulation's
        precision, preferably page-locked) must stay untouched until upload_end()."""
        dt = self.sim.dtype
        assert ix.dtype == np.int32 and x.dtype == dt and p.dtype == dt and q.dtype == dt
        assert all(a.flags.c_contiguous for a in (ix, x, p, q))
        check(self.sim._L.osgpu_species_upload_begin(self.sim._h, self.sp_id, q.shape[0], ptr(ix), ptr(x), ptr(p),
                                                     ptr(q), ptr(tag)))

    def upload_end(self):
        check(self.sim._L.osgpu_species_upload_end(self.sim._h, self.sp_id))

    @property
    def num_par(self):
        n = self.sim._L.osgpu_species_count(self.sim._h, self.sp_id)
        if n < 0:
            raise _lib.OsgpuError(-1, "species_count failed")
        return int(n)

    def fill_data(self, sort_by_tag=True):
        """Device -> host copy of the particle buffers: returns dict(ix, x, p, q, tag)."""
        n, D, dt = self.num_par, self.sim.dim, self.sim.dtype
        out = dict(ix=np.zeros((n, D), np.int32), x=np.zeros((n, D), dt), p=np.zeros((n, 3), dt),
                   q=np.zeros(n, dt), tag=np.zeros(n, np.int64))
        check(self.sim._L.osgpu_species_download(self.sim._h, self.sp_id, ptr(out["ix"]), ptr(out["x"]),
                                                 ptr(out["p"]), ptr(out["q"]), ptr(out["tag"])))
        if sort_by_tag:
            o = np.argsort(out["tag"], kind="stable")
            out = {k: v[o] for k, v in out.items()}
        return out

    def inject(self, ix, x, p, q, tag=None):
        """New particles (reference layout, ix local to this rank) join the species"""
        dt = self.sim.dtype
        ix = np.ascontiguousarray(ix, np.int32)
        x, p, q = (np.ascontiguousarray(a, dt) for a in (x, p, q))
        tag = None if tag is None else np.ascontiguousarray(tag, np.int64)
        check(self.sim._L.osgpu_species_inject(self.sim._h, self.sp_id, q.shape[0], ptr(ix), ptr(x), ptr(p), ptr(q),
                                               ptr(tag)))

    def push(self, n=None):
        """push_species (spec/os-species.f03:1461-1520): advance + deposit this species alone; its
        current is added to J (zero it first with sim.jay.zero()).  n = n(tstep) selects the beam
        initialisation pusher while n < n_accelerate or n < n_q_incr, as the reference's dispatch does."""
        if n is not None and (self.udist.use_accelerate(n) or self.udist.use_q_incr(n)):
            return self.accelerate(*self.udist.fractions(n))
        check(self.sim._L.osgpu_species_push(self.sim._h, self.sp_id))

    def accelerate(self, u_frac, q_frac):
        """accelerate_deposit (spec/os-spec-accel.f90:41-121)"""
        check(self.sim._L.osgpu_species_accelerate(self.sim._h, self.sp_id, float(u_frac), float(q_frac),
                                                   self.udist.ufl[0], int(self.udist.use_particle_uacc)))

    def deposit_charge(self, accumulate=False):
        """deposit_rho + guard-cell sum (spec/os-spec-charge.f90:139-197): the charge density on the grid nodes,
        array [(S3,) S2, S1] with the J array's guard cells"""
        shape = self.sim.jay_shape[:-1]
        rho = np.zeros(shape, self.sim.dtype)
        check(self.sim._L.osgpu_species_deposit_charge(self.sim._h, self.sp_id, ptr(rho), int(accumulate)))
        return rho

    # axis quantities and deposited quantities of a phasespace, by the reference's deck names
    # (spec/diagnostics/os-spec-phasespace.f03:28-46; "p1x1" = x1 on the first axis, p1 on the second)
    PHASE_AXIS = {"x": 1, "p": 2, "g": 3, "gl": 4, "l": 5, "kel": 6}
    PHASE_TYPE = {"charge": 0, "m": 1, "ene": 2, "q1": 3, "q2": 4, "q3": 5, "|charge|": 6, "j1": 7, "j2": 8, "j3": 9}

    def phasespace(self, axes, n, lo, hi, ps_type="charge", ene_bin=None, g_box=None, my_nx_p_lo=None,
                   buffer=None):
        """The particle loop of save_phasespace_1D / _2D / _3D (os-spec-phasespace.f03:970-1276): deposit this
        rank's particles of the species on the phasespace grid.  axes: one to three of ("x", 1), ("p", 2),
        ("g", 0), ("gl", 0), ("l", 3), ("kel", 0), FIRST axis first; n, lo, hi: points and range per axis.
        g_box ([lower], [upper]) and my_nx_p_lo (global index of this rank's first cell, 1-based) default to
        a box starting at 0 split evenly over the nodes.  Returns float32 [(n3,) n2, n1] (or [n1]); buffer, when
        given, is added to.  The sum over ranks and the normalisation stay with the caller."""
        sim = self.sim
        ps = _lib.Phasespace()
        ps.ndims = len(axes)
        for k, (kind, comp) in enumerate(axes):
            ps.x_or_p[k], ps.xp_dim[k] = self.PHASE_AXIS[kind], int(comp)
            ps.n[k], ps.min[k], ps.max[k] = int(n[k]), float(lo[k]), float(hi[k])
        ps.ps_type = self.PHASE_TYPE[ps_type]
        if ene_bin is not None:
            ps.use_ene_bin, ps.ene_bin[0], ps.ene_bin[1] = 1, float(ene_bin[0]), float(ene_bin[1])
        for d in range(sim.dim):
            glo = g_box[0][d] if g_box is not None else 0.0
            gup = g_box[1][d] if g_box is not None else sim.nx[d] * sim.dx[d] * sim.no_co.node_number[d]
            first = my_nx_p_lo[d] if my_nx_p_lo is not None else sim.no_co.coords[d] * sim.nx[d] + 1
            ps.xmin[d], ps.center[d], ps.ixmin[d] = glo, (glo + gup) / 2, int(first) - 2
        shape = tuple(int(v) for v in reversed(n[:len(axes)]))
        out = np.zeros(shape, np.float32) if buffer is None else buffer
        assert out.dtype == np.float32 and out.shape == shape and out.flags.c_contiguous
        check(sim._L.osgpu_species_phasespace(sim._h, self.sp_id, C.byref(ps), ptr(out), int(buffer is not None)))
        return out

    @property
    def energy(self):
        """this%energy(1) of the last push (sum q u^2/(gamma+1); get_energy multiplies by rqm * cell volume)"""
        e = C.c_double(0.0)
        check(self.sim._L.osgpu_species_energy(self.sim._h, self.sp_id, C.byref(e)))
        return e.value

    def update_boundary(self):
        """update_boundary_species (spec/os-spec-boundary.f03:1099-1232)"""
        check(self.sim._L.osgpu_species_update_boundary(self.sim._h, self.sp_id))

    def fill_data_async(self, ix, x, p, q, tag=None):
        """Queue the device -> host copy into the given (preferably page-locked) arrays and return;
        they are valid after sim.sync().  Arrays must hold num_par particles."""
        check(self.sim._L.osgpu_species_download_async(self.sim._h, self.sp_id, ptr(ix), ptr(x), ptr(p), ptr(q),
                                                       ptr(tag)))


class _Emf:
    """t_emf (emf/os-emf-define.f03:197-281)"""

    def __init__(self, sim):
        self.sim = sim

    def upload(self, e, b):
        dt = self.sim.dtype
        e, b = np.ascontiguousarray(e, dt), np.ascontiguousarray(b, dt)
        assert e.shape == self.sim.emf_shape and b.shape == self.sim.emf_shape, (e.shape, self.sim.emf_shape)
        check(self.sim._L.osgpu_fields_upload(self.sim._h, ptr(e), ptr(b)))

    def fill_data(self):
        e = np.zeros(self.sim.emf_shape, self.sim.dtype)
        b = np.zeros(self.sim.emf_shape, self.sim.dtype)
        check(self.sim._L.osgpu_fields_download(self.sim._h, ptr(e), ptr(b)))
        return e, b

    def update_boundary(self):
        check(self.sim._L.osgpu_emf_update_boundary(self.sim._h))

    def advance(self):
        check(self.sim._L.osgpu_emf_advance(self.sim._h))

    def set_smooth(self, smooth):
        """el_mag_fld{ smooth_type = "stand" } with the given osiris_b200.smooth.Smooth; None = "none"."""
        order = np.zeros(3, np.int32)
        if smooth is None or not smooth.if_smooth():
            check(self.sim._L.osgpu_emf_set_smooth(self.sim._h, ptr(order), 0, None))
            return
        sc = smooth.scoef if smooth.scoef is not None else smooth.setup(self.sim.dx, dtype=self.sim.dtype)
        sc = np.ascontiguousarray(sc, self.sim.dtype)
        order[:self.sim.dim] = smooth.order
        check(self.sim._L.osgpu_emf_set_smooth(self.sim._h, ptr(order), int(smooth.max_order), ptr(sc)))

    def set_solver(self, solver):
        """el_mag_fld{ solver = "yee" | "custom" } + emf_solver{...}: an osiris_b200.solver.Solver"""
        dt = self.sim.dtype
        if solver is None or solver.name == "yee":
            check(self.sim._L.osgpu_emf_set_solver(self.sim._h, 0, 0, None, None))
            return
        ce, cb = np.ascontiguousarray(solver.coef_e, dt), np.ascontiguousarray(solver.coef_b, dt)
        check(self.sim._L.osgpu_emf_set_solver(self.sim._h, 1, int(solver.n_coef), ptr(ce), ptr(cb)))

    def update_particle_fld(self):
        """update_particle_fld (emf/os-emf.f03:1065-1150): refresh e_part / b_part"""
        check(self.sim._L.osgpu_emf_update_particle_fld(self.sim._h))

    def energy(self):
        out = np.zeros(6, np.float64)
        check(self.sim._L.osgpu_field_energy(self.sim._h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out


class _Current:
    """t_current (current/os-current-define.f03:42-95); pf(1) only: one J grid per GPU"""

    def __init__(self, sim):
        self.sim = sim

    def upload(self, j):
        j = np.ascontiguousarray(j, self.sim.dtype)
        assert j.shape == self.sim.jay_shape
        check(self.sim._L.osgpu_current_upload(self.sim._h, ptr(j)))

    def fill_data(self):
        j = np.zeros(self.sim.jay_shape, self.sim.dtype)
        check(self.sim._L.osgpu_current_download(self.sim._h, ptr(j)))
        return j

    def update_boundary(self):
        check(self.sim._L.osgpu_current_update_boundary(self.sim._h))

    def zero(self):
        check(self.sim._L.osgpu_current_zero(self.sim._h))

    def set_smooth(self, smooth):
        """Hand setup_smooth's result (osiris_b200.smooth.Smooth) to the device; None switches it off."""
        if smooth is None or not smooth.if_smooth():
            order = np.zeros(3, np.int32)
            check(self.sim._L.osgpu_current_set_smooth(self.sim._h, ptr(order), 0, None))
            return
        sc = smooth.scoef if smooth.scoef is not None else smooth.setup(self.sim.dx, dtype=self.sim.dtype)
        sc = np.ascontiguousarray(sc, self.sim.dtype)
        order = np.zeros(3, np.int32)
        order[:self.sim.dim] = smooth.order
        check(self.sim._L.osgpu_current_set_smooth(self.sim._h, ptr(order), int(smooth.max_order), ptr(sc)))

    def smooth(self):
        """jay%smooth() (current/os-current-define.f03:298-316)"""
        check(self.sim._L.osgpu_current_smooth(self.sim._h))


class _Particles:
    """t_particles (particles/os-particles.f03): the container of all species"""

    def __init__(self, sim):
        self.sim = sim
        self.species = []

    def advance_deposit(self):
        """Returns the time-centred kinetic energy sum(q u^2/(gamma+1)) per species."""
        ene = np.zeros(max(1, len(self.species)), np.float64)
        check(self.sim._L.osgpu_advance_deposit(self.sim._h, ene.ctypes.data_as(C.POINTER(C.c_double))))
        return ene

    def update_boundary(self):
        check(self.sim._L.osgpu_particles_update_boundary(self.sim._h))

    def sort(self):
        check(self.sim._L.osgpu_sort(self.sim._h))


class Simulation:
    """The hot part of t_simulation (os-simulation.f03).  One instance per rank / GPU."""

    def __init__(self, dim, interpolation, nx, dx, dt, precision="f32", nspecies=1, gc_emf=None, gc_jay=None,
                 node_conf=None, nccl_id=None, device=-1, smooth_order=0, add_tag=True, smooth=None,
                 emf_smooth=None, part_bc=None, if_move=None, solver=None, emf_bc=None):
        self._L = _lib.load()
        self.dim, self.order = int(dim), int(interpolation)
        self.dtype = np.float32 if precision in ("f32", 4, np.float32) else np.float64
        self.nx = [int(v) for v in nx]
        self.dx = [float(v) for v in dx]
        self.dt = float(dt)
        if emf_smooth is not None:
            smooth_order = max(smooth_order, emf_smooth.smooth_order())
        glo, gup = min_gc_axes(dim, self.order, smooth_order, if_move)
        if solver is not None:  # the solver's stencil may need more (os-simulation.f03:360-365)
            slo, sup = solver.min_gc(dim)
            glo, gup = [max(a, b) for a, b in zip(glo, slo)], [max(a, b) for a, b in zip(gup, sup)]
        self.gc_emf = gc_emf if gc_emf is not None else (list(glo), list(gup))
        self.gc_jay = gc_jay if gc_jay is not None else (list(glo), list(gup))
        if smooth is not None and gc_jay is None:  # setup of t_current (current/os-current-define.f03:192-205)
            self.gc_jay = smooth.gc_num(self.gc_jay)
        self.no_co = node_conf if node_conf is not None else NodeConf(dim)
        cfg = Config()
        cfg.dim, cfg.order, cfg.real_bytes, cfg.nspecies = self.dim, self.order, np.dtype(self.dtype).itemsize, nspecies
        for d in range(dim):
            cfg.nx[d] = self.nx[d]
            cfg.dx[d] = self.dx[d]
            for s in range(2):
                cfg.gc_emf[s][d] = int(self.gc_emf[s][d])
                cfg.gc_jay[s][d] = int(self.gc_jay[s][d])
            cfg.node_dims[d] = self.no_co.node_number[d]
            cfg.node_coords[d] = self.no_co.coords[d]
            cfg.periodic[d] = int(self.no_co.if_periodic[d])
            for s in range(2):
                cfg.neighbor[s][d] = self.no_co.neighbor[d][s]
        cfg.topology = self.no_co.topology
        cfg.dt = self.dt
        cfg.rank, cfg.nranks, cfg.device, cfg.smooth_order = self.no_co.rank, self.no_co.nranks, device, smooth_order
        cfg.no_tags = 0 if add_tag else 1  # species{add_tag} (spec/os-spec-define.f03:431)
        # species boundary condition at physical walls, ([lower per axis], [upper per axis]) of
        # "absorbing" | "specular"; a moving window along x is (["absorbing", ..], ["specular", ..])
        if part_bc is not None:
            for s_ in range(2):
                for d in range(dim):
                    cfg.part_bc[s_][d] = {"absorbing": 0, "specular": 1}[part_bc[s_][d]]
        # field boundary condition at physical walls, ([lower per axis], [upper per axis]) of emf_bound{ type }:
        # "none" | "conducting" ("pec") | "reflecting" ("pmc"); the current is folded at the same walls
        if emf_bc is not None:
            names = {"none": 0, "conducting": 1, "pec": 1, "reflecting": 2, "pmc": 2}
            for s_ in range(2):
                for d in range(dim):
                    cfg.emf_bc[s_][d] = names[emf_bc[s_][d]]
        self.add_tag = bool(add_tag)
        self._cfg = cfg
        self._h = C.c_void_p()
        idbuf = None
        if nccl_id is not None:
            idbuf = (C.c_char * _lib.NCCL_ID_BYTES).from_buffer_copy(bytes(nccl_id))
        check(self._L.osgpu_init(C.byref(self._h), C.byref(cfg), idbuf))
        self.emf, self.jay, self.part = _Emf(self), _Current(self), _Particles(self)
        self.nspecies = nspecies
        self.n = 0  # n(tstep): steps taken by iter()
        if smooth is not None:
            self.jay.set_smooth(smooth)
        if emf_smooth is not None:
            self.emf.set_smooth(emf_smooth)
        if solver is not None:
            self.emf.set_solver(solver)

    # padded host array shapes (numpy order: slowest first, component last)
    def _shape(self, gc):
        return tuple(self.nx[d] + gc[0][d] + gc[1][d] for d in reversed(range(self.dim))) + (3,)

    @property
    def emf_shape(self):
        return self._shape(self.gc_emf)

    @property
    def jay_shape(self):
        return self._shape(self.gc_jay)

    def add_species(self, name, rqm, num_par_max, udist=None):
        sp = Species(self, len(self.part.species), name, rqm, num_par_max, udist)
        self.part.species.append(sp)
        return sp

    def iter(self, nsteps=1):
        """nsteps time steps entirely on the GPU; returns per-species kinetic energy sums of the
        last step (iter_sim, os-simulation.f03:561-661).  Steps in which a species still runs its beam
        initialisation pusher (n < n_accelerate / n_q_incr) go phase by phase through the per-species
        plug points, the others as one osgpu_step."""
        ene = np.zeros(max(1, self.nspecies), np.float64)
        done = 0
        while done < nsteps:
            n = self.n
            if any(sp.udist.use_accelerate(n) or sp.udist.use_q_incr(n) for sp in self.part.species):
                self.emf.update_boundary()
                self.emf.update_particle_fld()
                self.jay.zero()
                for sp in self.part.species:
                    sp.push(n)
                for i, sp in enumerate(self.part.species):
                    ene[i] = sp.energy
                    sp.update_boundary()
                self.jay.update_boundary()
                self.jay.smooth()
                self.emf.advance()
                self.n += 1
                done += 1
                continue
            last = max((max(sp.udist.n_accelerate, sp.udist.n_q_incr) for sp in self.part.species), default=0)
            k = nsteps - done  # (no species switches pusher inside these steps: n >= every threshold)
            assert n >= last
            check(self._L.osgpu_step(self._h, int(k), ene.ctypes.data_as(C.POINTER(C.c_double))))
            self.n += k
            done += k
        return ene

    def move_window(self, nmove):
        """emf%move_window + part%move_window (os-simulation.f03:529-548) by nmove[dim] cells"""
        nm = np.zeros(3, np.int32)
        nm[:self.dim] = nmove
        check(self._L.osgpu_move_window(self._h, ptr(nm)))

    def phase_times_ms(self):
        out = np.zeros(8, np.float64)
        check(self._L.osgpu_phase_times_ms(self._h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return dict(zip(("update emf boundary", "advance deposit", "update particle boundary",
                         "update current boundary", "field solver", "loop", "push kernel", "current smooth"),
                        [float(v) for v in out]))

    @property
    def launch_count(self):
        return int(self._L.osgpu_launch_count(self._h))

    @property
    def stream(self):
        return self._L.osgpu_stream(self._h)

    def sync(self):
        check(self._L.osgpu_sync(self._h))

    def close(self):
        if self._h:
            self._L.osgpu_finalize(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nccl_unique_id():
    buf = (C.c_char * _lib.NCCL_ID_BYTES)()
    check(_lib.load().osgpu_nccl_unique_id(buf))
    return bytes(buf)


def vadvance_deposit(dim, order, ix, x, u, q, rqm, e, b, emf_gc, j, j_gc, dx, dt, i0=1, i1=None):
    """The stateless pusher with the reference's C signature (host arrays in and out, J and the
    energy accumulated): osgpu_vadvance_deposit_{2d,3d}_s{order}[_d].  Returns the energy added."""
    L = _lib.load()
    dtp = x.dtype
    n = q.shape[0] if i1 is None else i1
    esz = np.ascontiguousarray(e.shape[::-1][1:1 + dim], np.int32)
    jsz = np.ascontiguousarray(j.shape[::-1][1:1 + dim], np.int32)
    eoff = np.ascontiguousarray(emf_gc[0], np.int32)
    joff = np.ascontiguousarray(j_gc[0], np.int32)
    dxa = np.ascontiguousarray(dx, np.float64)
    ene = C.c_double(0.0)
    check(L.osgpu_vadvance_deposit(dim, order, dtp.itemsize, ptr(ix), ptr(x), ptr(u), ptr(q), i0, n, float(rqm),
                                   ptr(e), ptr(b), ptr(esz), ptr(eoff), ptr(j), ptr(jsz), ptr(joff), ptr(dxa),
                                   float(dt), C.byref(ene)))
    return ene.value
