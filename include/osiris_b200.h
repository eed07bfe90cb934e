/*
 * osiris_b200.h -- C ABI of the B200-native PIC hot path (drop-in boundary for OSIRIS).
 *
 * Plain C: pointers, sizes and scalars only.  This is what OSIRIS' Fortran host binds with
 * ISO_C_BINDING (see INTEGRATION.md for the `bind(c)` interface module and the three
 * call-site patches).  Every entry point names the reference interface it replaces;
 * paths are relative to /root/reference/source.
 *
 * Conventions (the reference's own, so host arrays can be passed unchanged):
 *   - particle arrays are the Fortran buffers of t_species (spec/os-spec-define.f03:391-393):
 *       x(D,np) reals, ix(D,np) int32, p(3,np) reals, q(np) reals; component fastest;
 *       ix is the 1-based local cell, x in [-0.5,0.5) relative to it;
 *   - grids are t_vdf buffers (vdf/os-vdf-define.f03:345-351): f(3, S1, S2[, S3]),
 *       S_d = nx_d + gc[0][d] + gc[1][d], cell 1 at padded position gc[0][d];
 *   - "real" is float when real_bytes == 4 and double when 8 (p_k_part == p_k_fld);
 *   - all functions return 0 on success or a negative OSGPU_ERR_* code; the text of the last
 *     error is available from osgpu_last_error().  The Fortran shim turns a non-zero status
 *     into ERROR()/abort_program(p_err_invalid) (system/os-sys-multi.f03:1203-1240).
 *   - one host thread drives one GPU (n_threads must be 1 for push_type = "cuda").
 */
#ifndef OSIRIS_B200_H
#define OSIRIS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OSGPU_OK 0
#define OSGPU_ERR_INVALID -1     /* bad argument / unsupported configuration          */
#define OSGPU_ERR_CUDA -2        /* CUDA runtime or driver error                      */
#define OSGPU_ERR_NCCL -3        /* NCCL error                                        */
#define OSGPU_ERR_CAPACITY -4    /* particle / mover buffer too small                 */
#define OSGPU_ERR_STABILITY -5   /* Courant condition violated (test_stability_yee)   */
#define OSGPU_ERR_NO_DEVICE -6   /* no CUDA device: there is no CPU fallback          */

#define OSGPU_MAX_DIM 3
#define OSGPU_NCCL_ID_BYTES 128

typedef struct osgpu_sim osgpu_sim; /* opaque: owns every device buffer */

/* Everything init needs; mirrors what init_sim hands to emf/jay/part
 * (os-simulation.f03:232-400) and t_node_conf (os-nconf.f03:45-107). */
typedef struct osgpu_config {
  int32_t dim;        /* p_x_dim: 2 or 3 (os-param.f90:20)                               */
  int32_t order;      /* particles{interpolation}: 1 linear .. 4 quartic                 */
  int32_t real_bytes; /* 4 | 8 (PRECISION_SINGLE|DOUBLE, os-param.f90:27-55)             */
  int32_t nspecies;   /* number of t_species                                             */
  int32_t nx[OSGPU_MAX_DIM];        /* local cells my_nx(3,:) (grid/os-grid-define.f03)  */
  int32_t gc_emf[2][OSGPU_MAX_DIM]; /* guard cells of the host E/B arrays [lower|upper]  */
  int32_t gc_jay[2][OSGPU_MAX_DIM]; /* guard cells of the host J array                   */
  double dx[OSGPU_MAX_DIM];         /* cell size                                         */
  double dt;                        /* time step (full double; particles round it)       */
  /* domain decomposition: node_conf{node_number, if_periodic} (os-nconf.f03:222)        */
  int32_t node_dims[OSGPU_MAX_DIM];   /* ranks per axis (1,1,1 on a single GPU)          */
  int32_t node_coords[OSGPU_MAX_DIM]; /* this rank's cartesian coordinates               */
  int32_t periodic[OSGPU_MAX_DIM];    /* periodic axes                                   */
  int32_t rank;       /* NCCL rank of this process = its rank in the host's communicator  */
  int32_t nranks;     /* (t_node_conf%my_aid - 1, os-nconf.f03:540); see `topology`       */
  int32_t device;     /* CUDA device ordinal, -1 = rank % device_count                   */
  int32_t smooth_order; /* current smoothing passes folded into gc (0 = none)            */
  int32_t no_tags;    /* 1: particles carry no tag (species{add_tag} off, the reference's default,
                       * spec/os-spec-define.f03:431); by-tag downloads are then unavailable      */
  /* species boundary condition at a physical wall (a side with no neighbour: non-periodic axis, rank
   * on the edge), [lower/upper][axis]: spec/os-spec-boundary.f03:1293-1325.  A moving window makes the
   * lower wall absorbing and the upper one specular on its axis (:306-318, :1276).  Field boundary
   * conditions at such walls are NOT applied (that is what a moving window needs, p_bc_move_c;
   * conducting / open walls are SURVEY.md 8f row 3). */
  int32_t part_bc[2][OSGPU_MAX_DIM];
  /* How ranks map to node_coords (node_conf{topology}, os-nconf.f03:461-482):
   *   OSGPU_TOPO_MPI  (0, the reference's default "mpi"): row-major, LAST axis fastest -- the order
   *                   MPI_Cart_create / MPI_Cart_coords use (create_topology_mpi, :506-568);
   *   OSGPU_TOPO_OLD  (1, "old"): x1 fastest (create_topology_old, :574-625);
   *   OSGPU_TOPO_HOST (2): no assumption -- the neighbours are `neighbor` below, which the host copies
   *                   from t_node_conf%neighbor(2,D) - 1 (mpi_cart_create may reorder ranks).
   * For MPI / OLD the neighbours are derived from node_coords, and osgpu_init fails when `rank` is not
   * the rank that topology gives those coordinates. */
  int32_t topology;
  int32_t neighbor[2][OSGPU_MAX_DIM]; /* [lower|upper][axis] NCCL rank, -1 none (OSGPU_TOPO_HOST) */
  /* field boundary condition at a physical wall, [lower|upper][axis]: emf_bound{ type } (emf/boundary/
   * os-emf-boundary.f90:204-384).  OSGPU_EBC_NONE: nothing is done (periodic / other node / moving window,
   * p_bc_move_c); OSGPU_EBC_PEC ("pec", "conducting") and OSGPU_EBC_PMC ("pmc", "reflecting"): the mirror
   * walls of os-emf-pec.f90 / os-emf-pmc.f90 after every B and E (sub)step of the solver, in the variant
   * the particle interpolation selects (odd orders: wall on the cell edge, even: through the cell centre,
   * :490-511), and the specular fold of the current at the same wall (current/os-current-boundary.f03;
   * t_current takes its boundary type from emf_bound, current/os-current-define.f03:240-251).  Pair it with
   * part_bc = OSGPU_PBC_SPECULAR for the particles.  Open walls (lindman, vpml) are not built. */
  int32_t emf_bc[2][OSGPU_MAX_DIM];
} osgpu_config;
#define OSGPU_EBC_NONE 0
#define OSGPU_EBC_PEC 1
#define OSGPU_EBC_PMC 2
#define OSGPU_PBC_ABSORB 0
#define OSGPU_PBC_SPECULAR 1
#define OSGPU_TOPO_MPI 0
#define OSGPU_TOPO_OLD 1
#define OSGPU_TOPO_HOST 2

/* ---- library / topology helpers (no GPU needed) -------------------------------------- */
const char *osgpu_version(void);
const char *osgpu_last_error(void);
int osgpu_device_count(void);
/* cartesian rank <-> coords for topology OSGPU_TOPO_MPI (last axis fastest, MPI_Cart_create) or
 * OSGPU_TOPO_OLD (x1 fastest); neighbour(2,D) of t_node_conf (os-nconf.f03:547-559, 596-625):
 * neighbors[2*d+0] = lower, [2*d+1] = upper, -1 where there is none.  osgpu_cart_rank returns -1
 * for OSGPU_TOPO_HOST (the mapping is the host's). */
int osgpu_cart_rank(int topology, int dim, const int32_t *dims, const int32_t *coords);
int osgpu_cart_coords(int topology, int dim, const int32_t *dims, int rank, int32_t *coords);
int osgpu_cart_neighbors(int topology, int dim, const int32_t *dims, const int32_t *periodic, int rank,
                         int32_t *neighbors);
/* guard-cell exchange ranges along `axis` (vdf/os-vdf-comm.f90:756-784, 866-894), cell indices
 * inclusive.  mode 0 = replace (E,B), 1 = add (J).  out = {send_lo0, send_lo1, send_up0, send_up1,
 * recv_lo0, recv_lo1, recv_up0, recv_up1} */
int osgpu_exchange_ranges(int nx, int gcl, int gcu, int mode, int32_t *out8);
/* the same for the exchange that follows a window move of nx_move cells along the axis: the replace
 * ranges slide down so the upper neighbour also refreshes the cells the shift pulled out of stale
 * guard cells (nx_move argument of isend_vdf / irecv_vdf, vdf/os-vdf-comm.f90:770-776, 880-887) */
int osgpu_exchange_ranges_move(int nx, int gcl, int gcu, int mode, int nx_move, int32_t *out8);
/* the mirror rule the device applies at a PEC / PMC wall of `axis` (0-based), side 0 lower / 1 upper, to the
 * three components of E (is_b = 0) or B, and to J for an interpolation order: per component
 * {m, s, lo, hi, plane, plane_op} -- guard cells i = lo..hi take s * f(m - i) (fields) or are added with sign s
 * onto cell m - i (current); `plane` (INT32_MIN: none) is set to zero (plane_op 1) or doubled (2).  Restates
 * pec_bc_* / pmc_bc_* (emf/boundary/os-emf-pec.f90, os-emf-pmc.f90) and spec_bc_* (current/
 * os-current-boundary.f03); exported so the rule can be checked against those statements without a GPU. */
int osgpu_field_wall_op(int nx, int axis, int side, int is_b, int pmc, int even, int gcl, int gcu, int32_t *out18);
int osgpu_current_wall_op(int nx, int axis, int side, int order, int32_t *out18);
/* minimum guard cells the particles need for an interpolation order
 * (particles/os-particles.f03:329-346) merged with the Yee solver's (1,2)
 * (emf/solver/os-emf-solver-yee.f03:458-468): out = {lower, upper} */
int osgpu_min_gc(int order, int32_t *out2);
/* 16-byte NCCL unique id for osgpu_init; rank 0 creates it, the host broadcasts it (MPI_Bcast
 * in OSIRIS, torch.distributed in bench.py) */
int osgpu_nccl_unique_id(void *id_bytes /* OSGPU_NCCL_ID_BYTES */);

/* ---- lifetime ------------------------------------------------------------------------- */
/* replaces: init of t_emf/t_current/t_particles device mirrors (os-simulation.f03:232-400) */
int osgpu_init(osgpu_sim **sim, const osgpu_config *cfg, const void *nccl_id /* or NULL */);
int osgpu_finalize(osgpu_sim *sim);

/* ---- species: t_species buffers -------------------------------------------------------- */
/* replaces: alloc of t_species buffers (spec/os-spec-define.f03:1420-1468).  np_max bounds the
 * particles this rank can hold (num_par_max). */
int osgpu_species_create(osgpu_sim *sim, int sp_id, double rqm, int64_t np_max);
/* host AoS -> device tile-binned SoA.  tag may be NULL (then tag = 0-based index).
 * replaces nothing in the reference: it is the H2D half of the fill_data contract
 * (spec/os-species.f03:1867-1894). */
int osgpu_species_upload(osgpu_sim *sim, int sp_id, int64_t np, const int32_t *ix, const void *x,
                         const void *p, const void *q, const int64_t *tag);
/* the same in two halves: _begin queues the host -> device copies on a copy stream and returns (the
 * host arrays must stay untouched until _end returns), _end sorts the staged particles into the
 * species' buffer.  osgpu_species_upload == _begin + _end + wait.  At most two uploads in flight. */
int osgpu_species_upload_begin(osgpu_sim *sim, int sp_id, int64_t np, const int32_t *ix, const void *x,
                               const void *p, const void *q, const int64_t *tag);
int osgpu_species_upload_end(osgpu_sim *sim, int sp_id);
int64_t osgpu_species_count(osgpu_sim *sim, int sp_id);
/* device -> host AoS (fill_data target).  Arrays hold osgpu_species_count() particles, in device
 * (tile) order; tag (may be NULL) identifies them. */
int osgpu_species_download(osgpu_sim *sim, int sp_id, int32_t *ix, void *x, void *p, void *q,
                           int64_t *tag);

/* The same without the wait: returns once the gather and the device -> host copies are queued
 * (the copies run on a second stream); the arrays are valid after osgpu_sync().  Together with the
 * per-species phases below this lets a host stream one species -- or one chunk of a species, the way
 * the reference calls its pusher once per thread chunk (spec/os-spec-vpush.f03:380-397) -- after the
 * other: the download of one overlaps the upload of the next. */
int osgpu_species_download_async(osgpu_sim *sim, int sp_id, int32_t *ix, void *x, void *p, void *q,
                                 int64_t *tag);

/* ---- moving window ------------------------------------------------ */
/* One call = emf%move_window + part%move_window of iter_sim (os-simulation.f03:529-548) for
 * nmove[dim] = nx_move(g_space) cells (the host keeps t_space and decides when the window moves):
 * E, B slide down and are zero-filled (vdf/os-vdf.f03:56-205), every particle's ix drops by nmove
 * (spec/os-species.f03:171-229) and leavers meet migration / part_bc.  Call it before
 * osgpu_step; follow it with osgpu_species_inject for the plasma the window uncovered.
 * Deviation: the reference leaves the particles that dropped to cell 0 in the buffer for one more
 * push and removes them at the next update_boundary; here they meet the wall at once. */
int osgpu_move_window(osgpu_sim *sim, const int32_t *nmove);
/* np new particles (host arrays in the reference layout, ix local to this rank, inside 1..nx) join
 * the species: the inject() of move_window_species (spec/os-species.f03:218-226) with the particles
 * created by the host's psource.  tag may be NULL. */
int osgpu_species_inject(osgpu_sim *sim, int sp_id, int64_t np, const int32_t *ix, const void *x,
                         const void *p, const void *q, const int64_t *tag);

/* ---- fields: t_emf%e, t_emf%b, t_current%pf(1) ------------------------------------------ */
int osgpu_fields_upload(osgpu_sim *sim, const void *e, const void *b);
int osgpu_fields_download(osgpu_sim *sim, void *e, void *b);     /* emf/os-emf.f03:1158-1184 */
int osgpu_current_upload(osgpu_sim *sim, const void *j);
int osgpu_current_download(osgpu_sim *sim, void *j); /* current/os-current-define.f03:734-757 */

/* ---- the time step and its phases ------------------------------------------------------- */
/* One call = the hot part of iter_sim (os-simulation.f03:561-661), nsteps times:
 *   emf%update_boundary -> part%advance_deposit -> part%update_boundary -> (sort) ->
 *   jay%update_boundary -> emf%advance.
 * ene_kin[nspecies] (may be NULL) receives, for the LAST step, sum(q*u^2/(gamma+1)) per species
 * (time-centred, spec/os-spec-push.f03:500); multiply by rqm*prod(dx) for HIST/par_ene. */
int osgpu_step(osgpu_sim *sim, int nsteps, double *ene_kin);
/* the phases, individually (parity tests call these one at a time) */
int osgpu_emf_update_boundary(osgpu_sim *sim);     /* emf/boundary/os-emf-boundary.f90:723-830 */
int osgpu_advance_deposit(osgpu_sim *sim, double *ene_kin); /* particles/os-particles.f03:1362 */
int osgpu_particles_update_boundary(osgpu_sim *sim); /* spec/os-spec-boundary.f03:1099-1232   */
/* the two particle phases for ONE species, the granularity of the reference's own plug points:
 * osgpu_current_zero = jay%pf%zero() (particles/os-particles.f03:1396-1410), osgpu_species_push =
 * push_species (spec/os-species.f03:1461-1520; the species' current is ADDED to J),
 * osgpu_species_update_boundary = update_boundary_species (spec/os-spec-boundary.f03:1099-1232).
 * osgpu_advance_deposit == current_zero + species_push for every species. */
int osgpu_current_zero(osgpu_sim *sim);
int osgpu_species_push(osgpu_sim *sim, int sp_id);
int osgpu_species_update_boundary(osgpu_sim *sim, int sp_id);
/* replaces: accelerate_deposit (spec/os-spec-accel.f90:41-397), the pusher push_species selects while
 * n(tstep) < udist%n_accelerate or < udist%n_q_incr (spec/os-species.f03:1499-1503; beam initialisation,
 * decks/PWFAs).  The particles free-stream along x1 with the velocity of u_accel = u_frac * ufl1
 * (use_particle_uacc: u_frac * their own p1), momenta stay, the current is deposited with q * q_frac
 * and ADDED to J.  u_frac / q_frac are the host's grad_linear / grad_regressive values of the step
 * (os-math.f90:743-775).  The species' energy is left as the plain sum q u^2/(gamma+1)
 * (p_ene_recalc, spec/os-spec-define.f03:1040-1052).  Followed by osgpu_species_update_boundary. */
int osgpu_species_accelerate(osgpu_sim *sim, int sp_id, double u_frac, double q_frac, double ufl1,
                             int use_particle_uacc);
/* energy sum of the species' last push (this%energy(1), spec/os-spec-define.f03:1030-1086; multiply by
 * rqm * cell volume for get_energy) */
int osgpu_species_energy(osgpu_sim *sim, int sp_id, double *ene);
/* replaces: deposit_rho (spec/os-spec-charge.f90:139-197, templates :463-589) followed by the guard-cell
 * sum update_boundary(rho, p_vdf_add) of the charge diagnostics: the species' charge density on the grid
 * nodes, rho(i + k) += q prod S(k) with the species' interpolation order.  rho is the host array
 * rho(1, S1, S2[, S3]) with the guard cells of the J array (or NULL: keep the sum on the device);
 * accumulate != 0 adds this species to the density the previous call left (total charge).  The species
 * must not have unprocessed leavers (call it after osgpu_species_update_boundary / osgpu_step). */
int osgpu_species_deposit_charge(osgpu_sim *sim, int sp_id, void *rho, int accumulate);

/* One phasespace of the species' diagnostics (spec/diagnostics/os-spec-phasespace.f03), deposited on the
 * device: the axes and the deposited quantity are the reference's own codes. */
typedef struct osgpu_phasespace {
  int32_t ndims;       /* 1 | 2 | 3 (save_phasespace_1D :970-1067, _2D :1071-1169, _3D :1173-1276)             */
  int32_t x_or_p[3];   /* axis quantity: 1 x, 2 p, 3 gamma, 4 log10 gamma, 5 l (angular momentum), 6 log10 ke  */
                       /* (p_phase_x .. p_phase_kel :28-33; get_phasespace_axis_spec :3196-3238, get_l :314)   */
  int32_t xp_dim[3];   /* component 1..3 of x, p or l (ignored for gamma, log_gamma, log_ke)                    */
  int32_t ps_type;     /* deposited quantity: 0 charge, 1 mass, 2 ene, 3-5 q1..q3, 6 |charge|, 7-9 j1..j3       */
                       /* (p_pstype_normal .. p_pstype_j3 :37-46; get_phasespace_charge :2627-2737)            */
  int32_t n[3];        /* grid points per axis                                                                   */
  int32_t use_ene_bin; /* != 0: only particles with ene_bin[0] < u^2/(gamma+1) <= ene_bin[1] count (:2726-2735) */
  double min[3], max[3]; /* axis range (params%min / %max)                                                       */
  double ene_bin[2];
  /* get_position (spec/os-spec-define.f03:1098-1150): pos = xmin + dx * ((ix + ixmin) + x) with
   * xmin = g_box(p_lower, comp), ixmin = my_nx_p(1, comp) - 2; center = middle of g_box (get_l :335) */
  double xmin[OSGPU_MAX_DIM], center[OSGPU_MAX_DIM];
  int32_t ixmin[OSGPU_MAX_DIM];
} osgpu_phasespace;
/* replaces: the particle loop of save_phasespace_1D / _2D -- get_phasespace_axis, get_phasespace_charge,
 * deposit_1D / _2D / _3D (:1633-1902: linear weighting onto the two nearest grid points per axis, points
 * outside 1..n dropped) -- over the species on this rank.  buffer: n[0] (* n[1] (* n[2])) values of p_diag_prec (single,
 * dutil/os-dutil.f03:26), ADDED to when accumulate != 0, otherwise overwritten.  The sum over ranks
 * (reduce_array_size) and normalize_1D / _2D stay with the host. */
int osgpu_species_phasespace(osgpu_sim *sim, int sp_id, const osgpu_phasespace *ps, float *buffer, int accumulate);
int osgpu_sort(osgpu_sim *sim);                      /* spec/os-species.f03:269-332           */
int osgpu_current_update_boundary(osgpu_sim *sim); /* current/os-current-define.f03:683-727 */
int osgpu_emf_advance(osgpu_sim *sim);
/* replaces: the choice of the field solver, el_mag_fld{ solver = "yee" | "custom" } (new_solver_emf,
 * emf/solver/os-emf-solver.f03:36-52), and for "custom" the stencil of t_emf_solver_fei
 * (emf/solver/os-emf-solver-fei.f03): coef_e(1:n_coef) in dB/dt, coef_b(1:n_coef) in dE/dt, reals of the
 * handle's precision, as read_input_fei leaves them for any emf_solver{ type } ("customized-coef" takes
 * them from the deck, decks/PWFAs/beam-focus-3D).  Needs x1 guard cells (2 n_coef - 1, 2 n_coef) on E, B
 * (min_gc_fei, :1995-2026); returns OSGPU_ERR_STABILITY when dt exceeds test_stability_fei's limit for
 * these coefficients (:1370-1384).  Boundary tapering (taper_bnd) and the spectral current filter /
 * correction (filter_current, correct_current: FFTW) are not built. */
#define OSGPU_SOLVER_YEE 0
#define OSGPU_SOLVER_CUSTOM 1
int osgpu_emf_set_solver(osgpu_sim *sim, int type, int n_coef, const void *coef_e, const void *coef_b);             /* emf/os-emf-define.f03:467-482 (Yee)   */
/* Current smoothing.  set_smooth takes what setup_smooth leaves in
 * t_smooth (vdf/os-vdf-smooth.f90:373-467): order[dim] and scoef[dim][2*max_order+1] in FIELD
 * precision (real(p_k_fld)), centre tap at index max_order; all orders 0 switches it off.
 * Needs gc_jay >= (order, order+1) per axis (current/os-current-define.f03:199-205).  Once set,
 * osgpu_step filters J between jay%update_boundary and emf%advance (os-simulation.f03:654-657);
 * osgpu_current_smooth is that phase alone = smooth_current (current/os-current-define.f03:298-316)
 * -> filter_2d/filter_3d (vdf/os-vdf-smooth.f90:639-1105). */
int osgpu_current_set_smooth(osgpu_sim *sim, const int32_t *order, int max_order, const void *scoef);
int osgpu_current_smooth(osgpu_sim *sim);
/* Field smoothing for the particles, el_mag_fld{ smooth_type = "stand" } (emf/os-emf.f03:425-443,
 * 774-792): the particles interpolate e_part / b_part = smooth(e), smooth(b) with the t_smooth of
 * t_emf (same order / scoef convention as above); the solver keeps the unsmoothed fields.  Needs
 * gc_emf >= interpolation guard cells + order (emf/os-emf.f03:663-724).  Once set, osgpu_step
 * refreshes e_part, b_part after emf%update_boundary; osgpu_emf_update_particle_fld is that phase
 * alone (update_particle_fld_emf, emf/os-emf.f03:1065-1150). */
int osgpu_emf_set_smooth(osgpu_sim *sim, const int32_t *order, int max_order, const void *scoef);
int osgpu_emf_update_particle_fld(osgpu_sim *sim);
/* field energy per component, interior cells: out6 = {E1,E2,E3,B1,B2,B3}
 * (emf/diagnostics/os-emf-diagnostics.f03:43-58); summed over this rank only */
int osgpu_field_energy(osgpu_sim *sim, double *out6);
/* device-timed duration (ms, CUDA events on the handle's stream) of the last osgpu_step call,
 * summed over its steps: out8 = {emf_bnd, push_deposit, part_bnd, jay_bnd, solver, whole call
 * (first launch to last completion, gaps included), push kernel alone (single species),
 * current smooth (also counted inside jay_bnd)};
 * feeds the m_logprof events
 * 'update emf boundary', 'advance deposit', 'update particle boundary', 'update current
 * boundary', 'field solver' (os-logprof.f90:146-260). */
int osgpu_phase_times_ms(osgpu_sim *sim, double *out8);
/* number of kernel launches issued by this handle so far */
int64_t osgpu_launch_count(osgpu_sim *sim);
/* cudaStream_t the handle launches on (for CUDA-event timing by the caller) */
void *osgpu_stream(osgpu_sim *sim);
int osgpu_sync(osgpu_sim *sim);
/* page-locked host memory for arrays that cross the boundary every step (the Fortran side can
 * c_f_pointer these); plain pageable arrays work too, at pageable-copy speed */
void *osgpu_host_alloc(size_t bytes);
void osgpu_host_free(void *p);

/* ---- stateless pushers with EXACTLY the reference C signature --------------------------- */
/* replace: vadvance_deposit_{2d,3d}_s{1..4}_ (simd/os-spec-push-avx2.c:471-476, 1141-1145;
 * call sites spec/os-spec-vpush.f03:152-155, 294-297).  Host arrays in, host arrays out:
 * x, ix, u are advanced in place for particles i0..i1 (1-based, inclusive), the current is
 * ACCUMULATED into j and the time-centred energy into *ene, as the reference does.
 * Single precision: osgpu_vadvance_deposit_*; double: osgpu_vadvance_deposit_*_d.
 * On failure they print the error and exit(1), like the reference (:1213-1216). */
#define OSGPU_DECL_VADV(name, real)                                                            \
  void name(int *ix, real *x, real *u, real *q, int *i0, int *i1, real *rqm, real *efield,     \
            real *bfield, int *emf_size, int *emf_offset, real *j, int *j_size, int *j_offset, \
            double *dx, double *dt, double *ene)
OSGPU_DECL_VADV(osgpu_vadvance_deposit_2d_s1, float);
OSGPU_DECL_VADV(osgpu_vadvance_deposit_2d_s2, float);
OSGPU_DECL_VADV(osgpu_vadvance_deposit_2d_s3, float);
OSGPU_DECL_VADV(osgpu_vadvance_deposit_2d_s4, float);
OSGPU_DECL_VADV(osgpu_vadvance_deposit_3d_s1, float);
OSGPU_DECL_VADV(osgpu_vadvance_deposit_3d_s2, float);
OSGPU_DECL_VADV(osgpu_vadvance_deposit_3d_s3, float);
OSGPU_DECL_VADV(osgpu_vadvance_deposit_3d_s4, float);
OSGPU_DECL_VADV(osgpu_vadvance_deposit_2d_s1_d, double);
OSGPU_DECL_VADV(osgpu_vadvance_deposit_2d_s2_d, double);
OSGPU_DECL_VADV(osgpu_vadvance_deposit_2d_s3_d, double);
OSGPU_DECL_VADV(osgpu_vadvance_deposit_2d_s4_d, double);
OSGPU_DECL_VADV(osgpu_vadvance_deposit_3d_s1_d, double);
OSGPU_DECL_VADV(osgpu_vadvance_deposit_3d_s2_d, double);
OSGPU_DECL_VADV(osgpu_vadvance_deposit_3d_s3_d, double);
OSGPU_DECL_VADV(osgpu_vadvance_deposit_3d_s4_d, double);
/* generic form of the above (dim, order, precision at run time); returns a status instead of
 * exiting.  i0, i1 1-based inclusive. */
int osgpu_vadvance_deposit(int dim, int order, int real_bytes, int *ix, void *x, void *u, void *q,
                           int i0, int i1, double rqm, const void *efield, const void *bfield,
                           const int *emf_size, const int *emf_offset, void *j, const int *j_size,
                           const int *j_offset, const double *dx, double dt, double *ene);
/* The stateless entry points keep the device state of the last call (grids, particle buffers,
 * streams) per host thread and reuse it while dim/order/sizes/rqm are unchanged -- the reference
 * calls its pusher once per thread chunk per step (spec/os-spec-vpush.f03:380-397).  This drops it. */
void osgpu_vadvance_deposit_release(void);

#ifdef __cplusplus
}
#endif
#endif /* OSIRIS_B200_H */
