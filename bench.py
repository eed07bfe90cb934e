#!/usr/bin/env python
"""bench.py -- particle-steps/s of the PIC hot path (push + deposit + sort/binning + Yee solve).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload uniform3d|weibel2d|thermal2d|lwfa3d|pwfa3d] [--prec f32|f64]

One process per GPU (N > 1: launched by torch.distributed.run; RANK / LOCAL_RANK / WORLD_SIZE from
the environment).  A step is one pass of the hot path (E/B guards -> push+deposit -> particle
boundary + re-binning -> J guard sum -> Yee) over synthetic particles resident in HBM.  Rank 0
prints ONE JSON line (contract in the round prompt; keys `roofline`, `cpu_baseline`, `e2e`,
`clocks`, `gpu_launches`).

The particles are loaded on the reference's lattice and advanced --settle steps (default 20, named in
`config.load`, the same on both arms) before the W warm-up steps: lattice-loaded particles reach their first cell
faces together in steps 3-8, which cost up to 20 % more than any later step (profiles/r2_steptimes.txt).

Weak scaling: every GPU owns the workload's per-GPU box (3D uniform: 256^3 cells x 8 ppc); ranks
form a cartesian grid (8 -> 2x2x2) with NCCL guard-cell / migration exchanges inside the library.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from osiris_b200.loads import WORKLOADS, lattice_species  # noqa: E402

METRIC = "particle-steps/s (push+deposit+solve)"
UNIT = "particle-steps/s"
SETTLE_STEPS = 20
# algorithmic bytes (SURVEY.md 8d / BASELINE.md 2): per particle-step and per cell-step
B_P = {(2, 4): 60, (3, 4): 76, (2, 8): 104, (3, 8): 128}
B_C = {4: 96, 8: 192}
# share of B_C that belongs to the push+deposit kernel: stage E,B (6 words) + write J (3 words)
B_C_PUSH_WORDS = 9


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def node_dims(n, dim):
    table = {1: [1, 1, 1], 2: [2, 1, 1], 4: [2, 2, 1], 8: [2, 2, 2]} if dim == 3 else \
            {1: [1, 1], 2: [2, 1], 4: [2, 2], 8: [4, 2]}
    return table[n][:dim]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")] + [time.perf_counter()])

    def stop(self, t0=None, t1=None):
        """Samples read between t0 and t1 (perf_counter): the timed region.  nvidia-smi needs a few hundred ms to
        produce its first line, so the sampler is started before the warm-up and the rows are filtered here; a timed
        region shorter than the sampling period takes the rows of the second before it as well (GPU under the same
        load: the warm-up steps) and says so in `window`."""
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        rows, window = list(self.rows), "all"
        if t0 is not None:
            inside = [r for r in rows if t0 <= r[-1] <= t1]
            window = "timed region"
            if not inside:
                inside = [r for r in rows if t0 - 1.0 <= r[-1] <= t1 + 0.05]
                window = "timed region and the second of warm-up before it"
            rows = inside
        sm = [float(r[1]) for r in rows if len(r) >= 10 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 10 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) >= 10:
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def cpu_reference_run(w, prec, steps, warmup, sample_nx, threads=None):
    """The reference's own CPU implementation of the path on this box's host cores: its compiled
    SIMD push+deposit kernels (oracle/_ref, built from /root/reference/source/simd) under its
    threading model (oracle/cpu_driver.c), plus the oracle's restatement of the Fortran-only parts
    (guard cells, particle wrap, Yee).  Returns (particle-steps/s, info)."""
    import ctypes as C
    from oracle.bindings import Oracle, aligned, HERE as ORA, REFDIR
    dim, order = w["dim"], w["order"]
    dt_ = np.float32 if prec == "f32" else np.float64
    tag = "f32" if prec == "f32" else "f64"
    isa = "avx" if (prec == "f32" and order >= 3 and dim == 3) else "avx2"
    path = os.path.join(REFDIR, f"libref_{isa}_{tag}_perf.so")
    kind = "reference"
    if not os.path.exists(path):
        path = os.path.join(REFDIR, f"libref_{isa}_{tag}.so")
    orc = Oracle(prec)
    drv = C.CDLL(os.path.join(ORA, "libcpu_driver.so"))
    drv.cpu_driver_advance_deposit.restype = C.c_double
    T = threads or drv.cpu_driver_max_threads()
    if os.path.exists(path):
        ref = C.CDLL(path)
        fn = C.cast(getattr(ref, f"vadvance_deposit_{dim}d_s{order}_"), C.c_void_p)
    else:  # the reference could not be built: time the port through the same ABI
        kind, fn = "port", None
    nx = list(sample_nx)
    lo, up = max(order, 1), max(order + 1, 2)
    gc = ([lo] * dim, [up] * dim)
    shape = tuple(nx[d] + lo + up for d in reversed(range(dim))) + (3,)
    e, b = aligned(shape, dt_), aligned(shape, dt_)
    jcount = int(np.prod(shape))
    jth = aligned((T * jcount,), dt_)
    sps = []
    for i, s in enumerate(w["species"]):
        part = lattice_species(nx, w["ppc"], dt_, rqm=s["rqm"], uth=s["uth"], ufl=s["ufl"], seed=42 + i)
        al = {}
        for k, v in part.items():
            a = aligned(v.shape, v.dtype)
            a[...] = v
            al[k] = a
        al["rqm"] = s["rqm"]
        sps.append(al)
    npart = sum(s["q"].shape[0] for s in sps)
    esz = np.ascontiguousarray(shape[::-1][1:1 + dim], np.int32)
    off = np.ascontiguousarray(gc[0], np.int32)
    dxa = np.ascontiguousarray(w["dx"], np.float64)
    dtp = float(dt_(w["dt"]))
    pv = lambda a: a.ctypes.data_as(C.c_void_p)
    j0 = jth[:jcount].reshape(shape)

    def step():
        orc.update_periodic(dim, e, nx, gc, "replace")
        orc.update_periodic(dim, b, nx, gc, "replace")
        first = True
        for s in sps:
            if fn is not None:
                # species after the first accumulate into the reduced grid: run them single-pass
                # on grid 0 of a fresh set, then add (the reference loops species inside a thread)
                if not first:
                    keep = j0.copy()
                drv.cpu_driver_advance_deposit(fn, T, np.dtype(dt_).itemsize, dim, pv(s["ix"]), pv(s["x"]), pv(s["p"]),
                                               pv(s["q"]), C.c_long(s["q"].shape[0]), C.c_double(s["rqm"]), pv(e), pv(b),
                                               pv(esz), pv(off), pv(jth), C.c_long(jcount), pv(esz), pv(off), pv(dxa),
                                               C.c_double(dtp))
                if not first:
                    j0[...] += keep
            else:
                if first:
                    j0[...] = 0
                orc.advance_deposit(dim, order, s["ix"], s["x"], s["p"], s["q"], s["rqm"], e, b, gc, j0, gc, w["dx"], dtp)
            first = False
            orc.particles_periodic(dim, s["ix"], nx)
        orc.update_periodic(dim, j0, nx, gc, "add")
        orc.yee_advance(dim, e, b, j0, nx, gc, w["dx"], w["dt"])

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    el = time.perf_counter() - t0
    info = {"kind": kind, "cores": int(T), "npart": int(npart), "nx": nx, "isa": isa, "seconds": el,
            "sample": f"{'x'.join(map(str, nx))} cells x {int(np.prod(w['ppc']))} ppc ({npart} particles), "
                      f"{steps} steps, {isa.upper()} kernels, {T} threads"}
    return npart * steps / el, info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5, help="untimed steps (>= 3)")
    ap.add_argument("--settle", type=int, default=SETTLE_STEPS,
                    help="steps run while the workload is set up, before the warm-up, on both arms: the particles are "
                    "loaded on the reference's lattice (inject_std) and cross their first cell faces together in "
                    "steps 3-8, which run up to 20 %% slower than every later step (profiles/r2_steptimes.txt); the "
                    "metric is quoted on the thermalised plasma")
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="uniform3d")
    ap.add_argument("--prec", default="f32")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the per-GPU box (debug only)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-chunks", type=int, default=16, help="chunks per species the host hands over in the e2e leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    N = args.gpus
    steps, warmup = args.steps, max(args.warmup, 3 if args.impl == "ours" else 0)
    w = dict(WORKLOADS[args.workload])
    dim, order = w["dim"], w["order"]
    nx = [max(16, int(round(n * args.scale))) for n in w["nx"]]
    es = 4 if args.prec == "f32" else 8
    # the same dictionary on both arms (the driver compares them): everything that names the workload,
    # nothing that was measured
    cfg_common = {"workload": args.workload, "dim": dim, "interpolation": order, "cells_per_gpu": nx,
                  "ppc": int(np.prod(w["ppc"])), "species": len(w["species"]), "precision": args.prec,
                  "particles_per_gpu": int(np.prod(nx)) * int(np.prod(w["ppc"])) * len(w["species"]),
                  "node_dims": node_dims(N, dim),
                  "load": f"lattice (inject_std) + thermal momenta, advanced {args.settle} steps before the warm-up",
                  "cache": "inputs larger than L2 (particle buffers >> 126 MB)"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        # bounded sample of the same workload: 1/8 of the per-GPU volume
        sample = [max(16, n // 2) for n in nx]
        val, info = cpu_reference_run(w, args.prec, steps, warmup + args.settle, sample)
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": N, "steps": steps,
                "warmup": warmup, "ms_per_step": 1e3 * info["seconds"] / steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": args.prec, "data": "synthetic",
                "config": cfg_common,
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
                                 "sample": info["sample"]},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ---- our arm --------------------------------------------------------------------------
    from osiris_b200 import lib
    from osiris_b200.simulation import NodeConf, Simulation, nccl_unique_id
    L = lib.load()
    if L.osgpu_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    dist = None
    nccl_id = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo", rank=rank, world_size=world)
        ids = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        nccl_id = ids[0]
    nd = node_dims(world, dim)
    assert world == N or world == 1, f"--gpus {N} but WORLD_SIZE {world}"
    no_co = NodeConf(dim, nd, [True] * dim, rank)
    sim = Simulation(dim, order, nx, w["dx"], w["dt"], precision=args.prec, nspecies=len(w["species"]),
                     node_conf=no_co, nccl_id=nccl_id, device=local, add_tag=False)  # species{add_tag} default
    gnx = [nx[d] * nd[d] for d in range(dim)]
    off = [no_co.coords[d] * nx[d] for d in range(dim)]
    dt_ = np.float32 if args.prec == "f32" else np.float64
    npart = 0
    host_parts, sps = [], []
    for i, s in enumerate(w["species"]):
        part = lattice_species(nx, w["ppc"], dt_, rqm=s["rqm"], uth=s["uth"], ufl=s["ufl"], seed=42 + i,
                               global_nx=gnx, offset=off)
        n = part["q"].shape[0]
        sp = sim.add_species(f"species{i}", s["rqm"], int(n * 1.25) + 65536)
        sp.upload(part["ix"], part["x"], part["p"], part["q"])
        sps.append(sp)
        npart += n
        host_parts.append(part if (world == 1 and args.e2e_steps > 0) else None)

    def barrier():
        if dist is not None:
            dist.barrier()
        sim.sync()

    sampler = ClockSampler(local)
    if rank == 0:  # (rank 0 prints the line; seven more nvidia-smi loops would only load the host)
        sampler.start()
    sim.iter(args.settle)
    sim.iter(warmup)
    barrier()
    l0 = sim.launch_count
    t0 = time.perf_counter()
    sim.iter(steps)
    barrier()
    t1 = time.perf_counter()
    wall = t1 - t0
    clocks = sampler.stop(t0, t1)
    launches = sim.launch_count - l0
    pt = sim.phase_times_ms()
    dev_ms = pt["loop"]  # CUDA events on the launching stream, first launch to last completion
    if dist is not None:
        import torch
        t = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms = float(t[0]), float(t[1])
    else:
        wall_ms = wall * 1e3
    assert npart == cfg_common["particles_per_gpu"]
    total_part = npart * world
    ms_per_step = dev_ms / steps
    value = total_part / (ms_per_step * 1e-3)

    # roofline of the dominant kernel (push+deposit), rank 0's launches
    peak, peak_src = peaks()
    ncell = int(np.prod(nx))
    kern_ms = (pt["push kernel"] if len(w["species"]) == 1 else pt["advance deposit"]) / steps
    kern_bytes = npart * B_P[(dim, es)] + ncell * B_C_PUSH_WORDS * es
    achieved = kern_bytes / (kern_ms * 1e-3) / 1e9
    step_bytes = npart * B_P[(dim, es)] + ncell * B_C[es]
    roof = {"bound": "hbm", "kernel": "push_deposit_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": kern_bytes, "kernel_ms": kern_ms,
            "step_frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
            "phase_ms_per_step": {k: v / steps for k, v in pt.items()}}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            roof["traffic"] = json.load(open(tr)).get(f"{args.workload}_{args.prec}")
        except Exception:
            pass

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": N, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.prec, "data": "synthetic",
            "config": cfg_common, "wall_ms_per_step": wall_ms / steps,
            "roofline": roof, "clocks": clocks, "gpu_launches": int(launches)}

    # ---- end to end through the C ABI with HOST buffers (N = 1) -----------------------------
    emf_shape = sim.emf_shape
    if world == 1 and args.e2e_steps > 0:  # the same thermalised particles, back on the host
        for i, sp in enumerate(sps):
            d = sp.fill_data(sort_by_tag=False)
            host_parts[i] = {k: d[k] for k in ("ix", "x", "p", "q")}
    sim.close()
    if world == 1 and args.e2e_steps > 0 and rank == 0:
        line["e2e"] = e2e_run(lib, w, nx, host_parts, args.e2e_steps, dt_, args.prec, emf_shape, args.e2e_chunks)

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            sample = [max(16, n // 2) for n in nx]
            cs = max(2, min(10, steps))
            val, info = cpu_reference_run(w, args.prec, cs, 1 + args.settle, sample)
            line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
                                    "sample": info["sample"]}
        except Exception as ex:  # the checker is optional for the product line, never the reverse
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port",
                                    "sample": f"unavailable: {ex}"}
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def e2e_run(lib, w, nx, host_parts, nsteps, dt_, prec, emf_shape, nchunk):
    """Same metric through the reference-facing call sequence with HOST buffers: every step copies
    the particle and field arrays host -> device (pinned memory), runs the step, and copies them
    back -- what a drop-in at the pusher/solver boundary costs when nothing stays resident.

    The host hands the particles over in `nchunk` chunks per species, as the reference hands them
    to its pusher one thread chunk at a time (spec/os-spec-vpush.f03:380-397): chunk k is uploaded,
    pushed (its current added to J), wrapped and queued for download while chunk k+1 uploads, so
    both PCIe directions are busy.  Then J guard cells, Yee, and the fields come back.  The phases
    are the ones osgpu_step runs, per species (SURVEY.md 8b plug points)."""
    from osiris_b200.simulation import Simulation
    L = lib.load()
    nsp = len(host_parts)
    sim = Simulation(w["dim"], w["order"], nx, w["dx"], w["dt"], precision=prec, nspecies=nsp * nchunk, add_tag=False)
    pins, chunks = [], []
    for si, part in enumerate(host_parts):
        d = {}
        for k, v in part.items():
            a = lib.pinned_empty(v.shape, v.dtype)
            a[...] = v
            d[k] = a
        pins.append(d)
        n = d["q"].shape[0]
        edge = (16, 16) if w["dim"] == 2 else (8, 8, 4)  # cell tiles (csrc/tile.cuh); every tile keeps 32 spare slots
        ntile = int(np.prod([-(-nx[i] // edge[i]) for i in range(w["dim"])]))
        for k in range(nchunk):
            lo, hi = n * k // nchunk, n * (k + 1) // nchunk
            sp = sim.add_species(f"s{si}c{k}", w["species"][si]["rqm"], int((hi - lo) * 1.25) + 48 * ntile + 65536)
            chunks.append((sp, d, lo, hi))
    e = lib.pinned_empty(emf_shape, dt_)
    b = lib.pinned_empty(emf_shape, dt_)
    e[...] = 0
    b[...] = 0
    h2d = sum(sum(a.nbytes for a in d.values()) for d in pins) + e.nbytes + b.nbytes
    d2h = h2d

    def begin(k):
        sp, d, lo, hi = chunks[k]
        sp.upload_begin(d["ix"][lo:hi], d["x"][lo:hi], d["p"][lo:hi], d["q"][lo:hi])

    def one():
        begin(0)
        sim.emf.upload(e, b)
        sim.emf.update_boundary()
        sim.jay.zero()
        for k, (sp, d, lo, hi) in enumerate(chunks):
            sp.upload_end()
            if k + 1 < len(chunks):
                begin(k + 1)  # copies under the push of chunk k
            sp.push()
            sp.update_boundary()
            sp.fill_data_async(d["ix"][lo:hi], d["x"][lo:hi], d["p"][lo:hi], d["q"][lo:hi])
        sim.jay.update_boundary()
        sim.emf.advance()
        lib.check(L.osgpu_fields_download(sim._h, lib.ptr(e), lib.ptr(b)))
        sim.sync()

    one()  # warm-up
    t0 = time.perf_counter()
    for _ in range(nsteps):
        one()
    el = time.perf_counter() - t0
    npart = sum(d["q"].shape[0] for d in pins)
    assert sum(sp.num_par for sp, _, _, _ in chunks) == npart
    sim.close()
    for d in pins:
        for a in d.values():
            lib.pinned_free(a)
    lib.pinned_free(e)
    lib.pinned_free(b)
    return {"value": npart * nsteps / el, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "steps": nsteps, "ms_per_step": 1e3 * el / nsteps, "chunks_per_species": nchunk,
            "path": "per step: osgpu_fields_upload, osgpu_emf_update_boundary, osgpu_current_zero, per chunk "
                    "[osgpu_species_upload_begin (one chunk ahead), osgpu_species_upload_end, osgpu_species_push, osgpu_species_update_boundary, "
                    "osgpu_species_download_async], osgpu_current_update_boundary, osgpu_emf_advance, "
                    "osgpu_fields_download, osgpu_sync; pinned host arrays"}


if __name__ == "__main__":
    sys.exit(main())
