#!/usr/bin/env python
"""Benchmark of the demagnetization hot path (BASELINE.json metric: apply_demag seconds and
demag_tensor entries/s at 8k / 64k cells, 1/2/4/8 B200 next to the CPU path).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME] [--solver lu|gmres]

One "step" = one full pass of the hot path on the workload: assemble Q = I - chi*T (9 n^2 FP64
entries) and solve Q x = J0 + chi*H_ext.  Headline unit: T-entries per second of the whole step
(9 n^2 / step seconds); ms_per_step is the apply_demag device time itself.

  value  : inputs (cell geometry, chi, rhs) already resident in HBM, CUDA events around K steps.
  e2e    : the same step through the reference-signature object API -- apply_demag(Collection) on
           mesh_Cuboid objects (host objects in, solved host objects out; copy(), H2D of the geometry and
           D2H of the solution inside the timed region), K steps, wall clock between synchronizes.
  e2e_arrays   : the same through the array-level API apply_demag_arrays (host numpy in/out).
  parity_check : before timing, a 1728-cell two-magnet system solved with LU and GMRES through the public
                 API on ALL ranks and compared with the CPU oracle on rank 0 (non-zero exit above 1e-8).
  config4      : (N >= 2) after the timed region, one step of BASELINE configs[3] (64 000 cells, 295 GB).
  roofline     : dominant kernel (LU trailing DMMA update, or the matvec for --solver gmres), timed live
                 with CUDA events inside the library (mmr_ctx_profile); FP64 peaks measured in this run
                 (mmr_fp64_peaks) right before the timed region.
  cpu_baseline : the numpy oracle (restated reference path) on a bounded sample, rank 0, N=1.

--impl reference times the reference's own CPU algorithm (oracle port: magpylib is not installable here,
see DESIGN.md) on the host cores for the same metric and config: ONE full-size step is always taken
(recorded in cpu_baseline.full_size with its stage times); the W + K steps run at full size when that fits
the time budget (MMR_REF_BUDGET_S, default 420 s), otherwise on the largest sample of the workload that does.
"""

from __future__ import annotations

import argparse
import json
import os
import pathlib
import statistics
import subprocess
import sys
import tempfile
import time

# Host threads for the CPU arm: torchrun exports OMP_NUM_THREADS=1 to its workers, which would run the
# oracle's BLAS / LAPACK on one core.  Rank 0 -- the only rank that ever runs the oracle --
# claims every core; this has to happen before numpy loads its BLAS.
CPU_THREADS = os.cpu_count() or 1
if int(os.environ.get("RANK", "0")) == 0:
    for _var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_var] = str(CPU_THREADS)

import numpy as np  # noqa: E402

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
# stdout carries exactly one JSON line.  Native libraries write to file descriptor 1 behind Python's back
# (NCCL prints its version banner there): main() points fd 1 at stderr for the whole process and keeps a
# private handle on the real stdout for the result line.  (Not done at import: tests import make_workload.)
_RESULT_OUT = None


def _claim_stdout():
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _RESULT_OUT if _RESULT_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


FLOP_PER_PAIR_ALIGNED = 2162.0  # SURVEY 8d convention (sqrt=20, log=50, atan2=65, add/mul=1)
FLOP_PER_PAIR_ROTATED = 2288.0


# ----------------------------------------------------------------------------------------
# workloads (synthetic geometry of the BASELINE configs; arrays only, no objects)
# ----------------------------------------------------------------------------------------
def mesh_arrays(dimension, nnn, position=(0, 0, 0), quat=(0, 0, 0, 1)):
    from scipy.spatial.transform import Rotation as R

    dim0 = np.asarray(dimension, dtype=float)
    nnn = np.asarray(nnn, dtype=int)
    axes = [np.linspace(d / 2 * (1 / n - 1), d / 2 * (1 - 1 / n), n) for d, n in zip(dim0, nnn, strict=True)]
    grid = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, 3)
    rot = R.from_quat(quat)
    n = len(grid)
    return rot.apply(grid) + np.asarray(position, float), np.tile(np.asarray(quat, float), (n, 1)), np.tile(dim0 / nnn, (n, 1))


def make_workload(name):
    """Returns dict(pos, quat, dim, pol, chi, H_ext, desc).  pol in the LOCAL frame."""
    from scipy.spatial.transform import Rotation as R

    if name.startswith("two_magnets_"):
        # BASELINE configs[1]: hard magnet (chi 0.5) + soft-iron cube (chi 1000) rotated 45 deg
        # about y, docs/examples/soft_magnets.md:47-57; each meshed a x a x a/2
        n = int(name.split("_")[-1])
        a = round((n / 2 * 2) ** (1 / 3))
        nnn = (a, a, a // 2)
        p1, q1, d1 = mesh_arrays((1e-3, 1e-3, 2e-3), nnn, position=(0, 0, 0.5e-3))
        p2, q2, d2 = mesh_arrays((1e-3, 1e-3, 1e-3), nnn, position=(1.5e-3, 0, 0),
                                 quat=R.from_euler("y", 45, degrees=True).as_quat())
        m = len(p1)
        return dict(
            pos=np.vstack([p1, p2]), quat=np.vstack([q1, q2]), dim=np.vstack([d1, d2]),
            pol=np.vstack([np.tile((0, 0, 1.0), (m, 1)), np.zeros((m, 3))]),
            chi=np.vstack([np.full((m, 3), 0.5), np.full((m, 3), 1000.0)]), H_ext=np.zeros((2 * m, 3)), groups=[m, m],
            desc=f"hard magnet chi=0.5 + soft-iron cube chi=1000 rotated 45deg, 2 x {nnn[0]}x{nnn[1]}x{nnn[2]} = {2 * m} cells",
        )
    if name.startswith("soft_cube_"):
        # north-star target / configs[3]: single soft-iron cube, chi=1000, H_ext=(0,0,1)
        n = int(name.split("_")[-1])
        a = round(n ** (1 / 3))
        p, q, d = mesh_arrays((1e-3,) * 3, (a, a, a))
        m = len(p)
        return dict(pos=p, quat=q, dim=d, pol=np.zeros((m, 3)), chi=np.full((m, 3), 1000.0),
                    H_ext=np.tile((0, 0, 1.0), (m, 1)),
                    desc=f"soft-iron cube chi=1000, H_ext=(0,0,1), {a}x{a}x{a} = {m} cells")
    if name.startswith("lattice_"):
        # BASELINE configs[2]: k x k x k lattice of 1 mm cubes, pitch 1.5 mm, each meshed
        # a x a x a identical cells, J = (0,0,1) T, anisotropic chi = (0.3, 0.2, 0.1); run with pairs_matching.
        # lattice_27000 = 3^3 magnets x 10^3 cells; lattice_<k>x<a> builds reduced sizes for the parity tests.
        spec = name.split("_")[1]
        k, a = (int(v) for v in spec.split("x")) if "x" in spec else (3, round((int(spec) / 27) ** (1 / 3)))
        parts = [mesh_arrays((1e-3,) * 3, (a, a, a), position=(1.5e-3 * i, 1.5e-3 * j, 1.5e-3 * l))
                 for i in range(k) for j in range(k) for l in range(k)]
        p_, q_, d_ = (np.vstack([pt[c] for pt in parts]) for c in range(3))
        m = len(p_)
        return dict(pos=p_, quat=q_, dim=d_, pol=np.tile((0, 0, 1.0), (m, 1)), chi=np.tile((0.3, 0.2, 0.1), (m, 1)),
                    H_ext=np.zeros((m, 3)), pairs_matching=True,
                    desc=f"{k}x{k}x{k} lattice of 1 mm cubes, pitch 1.5 mm, each {a}x{a}x{a} cells = {m} cells, "
                         "chi=(0.3,0.2,0.1), pairs_matching")
    if name.startswith("array_"):
        # BASELINE configs[4]: k x k planar array of 1 mm cubes, pitch 2 mm, each a^3 cells,
        # J = (0,0,1) T, chi = 0.5, max_dist truncation (pairs farther than max_dist cell edges are dropped) and
        # split chunking.  array_51200 = 10x10 magnets x 8^3 cells; array_<k>x<a> for reduced sizes.
        spec = name.split("_")[1]
        k, a = (int(v) for v in spec.split("x")) if "x" in spec else (10, round((int(spec) / 100) ** (1 / 3)))
        parts = [mesh_arrays((1e-3,) * 3, (a, a, a), position=(2e-3 * i, 2e-3 * j, 0.0)) for i in range(k) for j in range(k)]
        p_, q_, d_ = (np.vstack([pt[c] for pt in parts]) for c in range(3))
        m = len(p_)
        return dict(pos=p_, quat=q_, dim=d_, pol=np.tile((0, 0, 1.0), (m, 1)), chi=np.full((m, 3), 0.5),
                    H_ext=np.zeros((m, 3)), max_dist=12.0, split=8,
                    desc=f"{k}x{k} planar array of 1 mm cubes, pitch 2 mm, each {a}x{a}x{a} cells = {m} cells, "
                         "chi=0.5, max_dist=12, split=8")
    msg = f"unknown workload {name}"
    raise ValueError(msg)


# ----------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.split(",") for r in pathlib.Path(self.tmp.name).read_text().strip().splitlines() if r.count(",") >= 6]
        os.unlink(self.tmp.name)
        if not rows:
            return out
        sm = [float(r[0]) for r in rows]
        pw = [float(r[2]) for r in rows]
        loaded = [s for s, p in zip(sm, pw, strict=True) if p > 0.5 * max(pw)] or sm
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names) if any("Active" in r[3 + k] and "Not" not in r[3 + k] for r in rows)]
        out.update(sm_mhz=statistics.median(loaded), sm_max_mhz=float(rows[0][1]), reasons=reasons, samples=len(rows),
                   power_w_max=max(pw))
        return out


# ----------------------------------------------------------------------------------------
# shared description of a workload: BOTH arms print this dict as `config`
# ----------------------------------------------------------------------------------------
def bench_config(name, w, solver):
    n = len(w["pos"])
    N = 3 * n
    q_gb = 8.0 * N * N / 1e9
    return {"workload": name, "desc": w["desc"], "n_cells": n, "matrix_order": N,
            "solver": "lu (dgesv-equivalent: LU with partial pivoting, 1 rhs)" if solver == "lu" else solver,
            "l2": "inputs larger than L2 (Q is %.1f GB)" % q_gb if q_gb > 0.5 else "256 MB flush buffer written before every step"}


def make_collection(name):
    """The workload as the objects a user of the reference API holds: Cuboids meshed with mesh_Cuboid
    (docs/examples/soft_magnets.md:47-57 for two_magnets_*), susceptibilities / H_ext as attributes."""
    from magpylib_material_response_b200 import Collection, Cuboid, mesh_Cuboid

    n = int(name.split("_")[-1])
    if name.startswith("two_magnets_"):
        a = round(n ** (1 / 3))
        nnn = (a, a, a // 2)
        hard = Cuboid(polarization=(0, 0, 1), dimension=(1e-3, 1e-3, 2e-3)).move((0, 0, 0.5e-3))
        hard.susceptibility = 0.5
        soft = Cuboid(polarization=(0, 0, 0), dimension=(1e-3, 1e-3, 1e-3))
        soft.rotate_from_angax(45, "y").move((1.5e-3, 0, 0))
        soft.susceptibility = 1000.0
        return Collection(mesh_Cuboid(hard, nnn), mesh_Cuboid(soft, nnn))
    if name.startswith("soft_cube_"):
        a = round(n ** (1 / 3))
        cube = Cuboid(polarization=(0, 0, 0), dimension=(1e-3, 1e-3, 1e-3))
        cube.susceptibility = 1000.0
        coll = mesh_Cuboid(cube, (a, a, a))
        coll.H_ext = (0.0, 0.0, 1.0)
        return coll
    return None


def collection_polarizations(coll):
    """Local-frame polarizations of all cells, (n,3), without creating cell objects."""
    parts = []
    for g in coll._source_groups():
        blk = getattr(g, "_block", None)
        parts.append(blk.pol if blk is not None else np.asarray(g.polarization, dtype=float)[None, :])
    return np.vstack(parts)


# ----------------------------------------------------------------------------------------
# reference arm / cpu baseline (oracle port on host cores)
# ----------------------------------------------------------------------------------------
def cpu_sample_workload(name, n_sample):
    kind = name.rsplit("_", 1)[0]
    return make_workload(f"{kind}_{n_sample}")


def time_oracle_step(w, split=None, stages=None):
    """One full reference step on the CPU: 3 getH passes, dense S@T, np.linalg.solve.  The getH passes run
    on a thread pool over source chunks (bit-identical to the single-threaded order; the reference itself
    runs them on one thread), BLAS/LAPACK on all cores: every host thread the path can use."""
    from threadpoolctl import threadpool_limits

    from oracle import demag_ref as oref

    n = len(w["pos"])
    if split is None:  # keep the ~60 live n_src x n temporaries of a getH chunk below ~1 GB per thread
        split = max(4, 2 * CPU_THREADS, int(np.ceil(n * n * 60 * 8 * CPU_THREADS / 16e9)))
    t0 = time.perf_counter()
    with threadpool_limits(limits=CPU_THREADS):
        oref.apply_demag_ref(w["pos"], w["quat"], w["dim"], w["pol"], w["chi"], H_ext=w["H_ext"], split=split,
                             timings=stages, threads=CPU_THREADS)
    return time.perf_counter() - t0


def host_mem_available():
    try:
        import psutil

        return float(psutil.virtual_memory().available)
    except Exception:  # pragma: no cover
        return 0.0


def oracle_fits(n):
    """The reference path keeps ~7 dense N x N FP64 arrays alive (T4, T, S, S@T, eye, Q, LAPACK copy)."""
    return 7.5 * 8.0 * (3 * n) ** 2 < 0.8 * host_mem_available()


def sample_sizes(name):
    """Cell counts of the workload family, largest first (cube edges in steps of 2)."""
    n_full = int(name.split("_")[-1])
    a_full = round(n_full ** (1 / 3))
    return [a ** 3 for a in range(a_full, 5, -2)]


def cpu_baseline(name, budget_s=30.0):
    """Oracle timed on a bounded sample of the same workload geometry (about 10-30 s of CPU work)."""
    if name.startswith(("two_magnets_", "soft_cube_")):
        n_sample = next((m for m in sample_sizes(name) if m <= 4096), 1000)
    else:
        n_sample = 1728
    w = cpu_sample_workload(name, n_sample)
    n = len(w["pos"])
    stages = {}
    t = time_oracle_step(w, stages=stages)
    return {
        "value": 9.0 * n * n / t, "unit": "T-entries/s", "cores": CPU_THREADS, "kind": "port",
        "stages_s": {k: round(v, 4) for k, v in stages.items()},
        "assembly_entries_per_s": 9.0 * n * n / stages["assembly_s"],
        "sample": f"{w['desc']} (reduced from the full workload), one full reference step "
                  f"(3 getH passes + dense S@T + LAPACK dgesv) in {t:.2f} s; getH passes on a {CPU_THREADS}-thread pool over "
                  "source chunks (the reference runs them on one thread), BLAS/LAPACK on all cores; the full-size step is "
                  "timed by `bench.py --impl reference` (cpu_baseline.full_size)",
        "seconds": t, "n_cells": n,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload or default_workload(args.gpus)
    w_full = make_workload(name)
    n_full = len(w_full["pos"])
    budget = float(os.environ.get("MMR_REF_BUDGET_S", "420"))
    total = args.steps + args.warmup
    t_begin = time.perf_counter()

    # (1) one full-size step, always (when the host can hold it)
    full = None
    if oracle_fits(n_full):
        st = {}
        t_full = time_oracle_step(w_full, stages=st)
        full = {"n_cells": n_full, "seconds": round(t_full, 3), "stages_s": {k: round(v, 3) for k, v in st.items()},
                "value": 9.0 * n_full * n_full / t_full, "unit": "T-entries/s"}
    # (2) size of the W + K steps: full size if the budget allows, else the largest sample that fits.
    # Model from the full-size stage times: assembly ~ n^2, dense S@T and dgesv ~ n^3.
    used = time.perf_counter() - t_begin
    steps_left = total - (1 if full else 0)
    if full:
        asm, cub = full["stages_s"]["assembly_s"], full["stages_s"]["build_Q_s"] + full["stages_s"]["solve_s"]
        nf = n_full
    else:  # calibrate on a small sample
        wc = cpu_sample_workload(name, 1000)
        st = {}
        time_oracle_step(wc, stages=st)
        asm, cub, nf = st["assembly_s"], st["build_Q_s"] + st["solve_s"], len(wc["pos"])
        used = time.perf_counter() - t_begin

    def predict(m):
        return 1.15 * (asm * (m / nf) ** 2 + cub * (m / nf) ** 3)

    n_run = None
    for m in sample_sizes(name):
        if (m == n_full and not full) or not oracle_fits(m):
            continue
        if steps_left * predict(m) <= budget - used:
            n_run = m
            break
    if n_run is None:
        n_run = sample_sizes(name)[-1]
    w = w_full if n_run == n_full else cpu_sample_workload(name, n_run)
    n = len(w["pos"])
    times = []
    done_warm = 1 if (full and n_run == n_full) else 0  # the full-size step above is the first warm-up
    for _ in range(max(0, args.warmup - done_warm)):
        time_oracle_step(w)
    times = [time_oracle_step(w) for _ in range(args.steps)]
    t = sum(times) / len(times)
    value = 9.0 * n * n / t
    if n_run == n_full:
        sample = f"full size: {w['desc']}; every step is the whole workload, {t:.2f} s per step"
    else:
        sample = (f"{w['desc']} -- the largest member of the workload family whose {total} steps fit the {budget:.0f} s "
                  f"budget ({t:.2f} s per step); the full-size step ({n_full} cells) is in full_size")
    sample += (f"; one step = the reference's 3 getH passes + dense S@T + LAPACK dgesv; getH passes on a {CPU_THREADS}-thread "
               "pool over source chunks (the reference runs them on one thread), BLAS/LAPACK on all cores")
    line = {
        "impl": "reference", "metric": "apply_demag_T_entries_per_s", "value": value, "unit": "T-entries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(name, w_full, "lu"),
        "cpu_baseline": {"value": value, "unit": "T-entries/s", "cores": CPU_THREADS, "kind": "port",
                         "sample": sample, "sample_n_cells": n, "full_size": full,
                         "note": "restated oracle of the reference CPU path (magpylib is not installable here, DESIGN.md 2)"},
        "e2e": {"value": value, "unit": "T-entries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "host_threads_env": {v: os.environ.get(v) for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS")},
    }
    emit(line)


# ----------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------
def default_workload(ngpus):
    return "two_magnets_8000"


def load_peaks(ctx):
    """Roofline denominators: HBM from MEASURED_PEAKS.json (driver-written), FP64 measured NOW on this
    device through the C ABI (mmr_fp64_peaks); the tracked round-1 file is only the cross-check."""
    peaks = {"hbm_gbs": 6650.0, "hbm_src": "fallback (B200_PROFILING.md)"}
    mp = ROOT / "MEASURED_PEAKS.json"
    if mp.exists():
        d = json.loads(mp.read_text())
        peaks["hbm_gbs"], peaks["hbm_src"] = d.get("hbm_gbs", peaks["hbm_gbs"]), "MEASURED_PEAKS.json"
    fma, dmma = ctx.fp64_peaks()
    peaks["fp64_fma_tflops"], peaks["fp64_dmma_tflops"] = fma, dmma
    peaks["fp64_src"] = ("mmr_fp64_peaks, measured in this run right before the timed region "
                         f"(DMMA {dmma:.2f} TF, FMA {fma:.2f} TF; burst figure of independent instruction chains on all SMs)")
    fp = ROOT / "profiles" / "fp64_peaks_r1.json"
    if fp.exists():
        d = json.loads(fp.read_text())
        peaks["fp64_src"] += f"; round-1 file: DMMA {d['fp64_dmma_tflops']:.2f}, FMA {d['fp64_fma_tflops']:.2f}"
    return peaks


PARITY_TOL = 1e-8  # north star: solved polarizations within rtol 1e-8


def parity_check(world, rank):
    """Driver-visible correctness of THIS run's configuration (N ranks): a 1728-cell two-magnet system
    (rotated chi = 1000 cube included) through the public array API on all ranks -- LU, GMRES, and the
    pairs_matching / split / demag_store variants -- against the CPU oracle on rank 0."""
    from magpylib_material_response_b200.demag import apply_demag_arrays

    w = make_workload("two_magnets_1728")
    n = len(w["pos"])
    geo = (w["pos"], w["quat"], w["dim"], w["pol"], w["chi"], w["H_ext"])
    runs = [("lu", {}), ("gmres", {}), ("lu", {"split": 8}), ("lu", {"pairs_matching": True})]
    xs = []
    for solver, kw in runs:
        xs.append(apply_demag_arrays(*geo, solver=solver, tol=1e-12, **kw))
    store = {}
    apply_demag_arrays(*geo, solver="lu", demag_store=store)
    xs.append(apply_demag_arrays(*geo, solver="lu", demag_store=store))  # second call: factors reused
    runs.append(("lu", {"demag_store": "hit"}))
    out = []
    if rank == 0:
        from threadpoolctl import threadpool_limits

        from oracle import demag_ref as oref

        with threadpool_limits(limits=CPU_THREADS):
            ref = oref.apply_demag_ref(*geo[:5], H_ext=w["H_ext"], split=2 * CPU_THREADS, threads=CPU_THREADS)
        scale = np.abs(ref).max()
        for (solver, kw), x in zip(runs, xs, strict=True):
            out.append({"solver": solver, **kw, "n": n, "max_rel_err": float(np.abs(x - ref).max() / scale)})
    return out


def run_config4(ctx, world, rank, dist):
    """One step of BASELINE configs[3] (64 000-cell soft-iron cube, N = 192 000, 295 GB of Q) on the N
    ranks of this run: distributed assembly + LU + solve, then the true residual on a fresh Q."""
    import torch

    from magpylib_material_response_b200.demag import _padded_ld
    from magpylib_material_response_b200.distributed import DEFAULT_TGT_BLOCK, owned_cells

    w = make_workload("soft_cube_64000")
    n = len(w["pos"])
    N = 3 * n
    ld = _padded_ld(N)
    tgt_block = DEFAULT_TGT_BLOCK
    n_local = len(owned_cells(n, tgt_block, world, rank))
    need = 8.0 * 3 * n_local * ld + 8.0 * N * (2 * 256 + 128 + 16) + 2e9
    torch.cuda.empty_cache()
    free = torch.cuda.mem_get_info(ctx.device)[0]
    ok = torch.tensor([1 if need < free else 0], device=ctx.device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) == 0:
        return {"skipped": f"needs {need / 1e9:.0f} GB per GPU at {world} GPUs, {free / 1e9:.0f} GB free"}
    rhs = (w["pol"] + w["chi"] * w["H_ext"]).reshape(N)  # identity orientation
    d_pos, d_quat, d_dim = ctx.to_device(w["pos"]), ctx.to_device(w["quat"]), ctx.to_device(w["dim"])
    d_chi, d_rhs = ctx.to_device(w["chi"].reshape(N)), ctx.to_device(rhs)
    Q = ctx.empty(3 * n_local, ld)
    d_x = ctx.empty(N)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]

    def assemble():
        ctx.assemble(d_pos, d_quat, d_dim, Q, ld, chi=d_chi, n_tgt_local=n_local, tgt_block=tgt_block,
                     tgt_nparts=world, tgt_part=rank)

    dist.barrier()
    torch.cuda.synchronize()
    ctx.set_profiling(True)
    ev[0].record()
    assemble()
    ev[1].record()
    d_x.copy_(d_rhs)
    info = ctx.lu_solve_dist(N, 3 * n_local, tgt_block, Q, ld, d_x)
    ev[2].record()
    torch.cuda.synchronize()
    prof = ctx.profile()
    ctx.set_profiling(False)
    tt = torch.tensor([ev[0].elapsed_time(ev[2]), ev[0].elapsed_time(ev[1]), prof["gemm"][0]], dtype=torch.float64,
                      device=ctx.device)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_all, t_asm, t_gemm = (float(v) * 1e-3 for v in tt)
    assemble()
    _, relres = ctx.gmres_dist(N, 3 * n_local, tgt_block, Q, ld, d_rhs, d_x, tol=1.0, restart=1, maxit=1)
    x = d_x.cpu().numpy().reshape(n, 3)
    del Q
    torch.cuda.empty_cache()
    lu_s = t_all - t_asm
    return {"workload": "soft_cube_64000", "n_cells": n, "matrix_order": N, "q_gb_total": 8.0 * N * N / 1e9,
            "apply_demag_s": t_all, "assembly_s": t_asm, "assembly_entries_per_s": 9.0 * n * n / t_asm,
            "lu_factor_tflops": (2.0 / 3.0) * N ** 3 / lu_s / 1e12, "gemm_frac": t_gemm / t_all,
            "relres": relres, "lu_info": int(info), "mean_Jz": float(x[:, 2].mean()),
            "note": "one un-warmed step (max over ranks, CUDA events); relres = ||b - Qx||/||b|| on a re-assembled Q"}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from magpylib_material_response_b200 import _native
    from magpylib_material_response_b200.demag import _padded_ld, apply_demag, apply_demag_arrays

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the demag path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    name = args.workload or default_workload(world)
    w = make_workload(name)
    n = len(w["pos"])
    N = 3 * n
    rotated = bool(np.any(w["quat"][:, :3] != 0))
    ctx = _native.default_context(local_rank)
    if world > 1:
        ctx.comm_init_from_torch()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness of this configuration first (not timed) --------------------------------------
    parity = None
    if not args.no_parity:
        parity = parity_check(world, rank)
        bad = torch.tensor([1 if any(not (r["max_rel_err"] <= PARITY_TOL) for r in parity) else 0], device=ctx.device)
        if world > 1:
            dist.broadcast(bad, src=0)
        if int(bad.item()):
            if rank == 0:
                emit({"metric": "apply_demag_T_entries_per_s", "value": None, "n_gpus": world, "parity_check": parity,
                      "error": f"parity check failed (tolerance {PARITY_TOL})"})
            if world > 1:
                dist.destroy_process_group()
            raise SystemExit(3)

    from scipy.spatial.transform import Rotation as R

    rot = R.from_quat(w["quat"])
    rhs_host = (rot.apply(w["pol"]) + w["chi"] * w["H_ext"]).reshape(N)
    d_pos, d_quat, d_dim = ctx.to_device(w["pos"]), ctx.to_device(w["quat"]), ctx.to_device(w["dim"])
    d_chi, d_rhs = ctx.to_device(w["chi"].reshape(N)), ctx.to_device(rhs_host)
    ld = _padded_ld(N)

    # row-block ownership (block-cyclic over cells); single GPU owns everything
    from magpylib_material_response_b200.distributed import DEFAULT_TGT_BLOCK

    tgt_block = DEFAULT_TGT_BLOCK if world > 1 else n
    nblk = -(-n // tgt_block)
    my_blocks = list(range(rank, nblk, world))
    n_local = sum(min(tgt_block, n - b * tgt_block) for b in my_blocks)
    Q = ctx.empty(3 * n_local, ld)
    ipiv = ctx.empty(N, dtype=torch.int32)
    d_x = ctx.empty(N)
    solver = args.solver
    from magpylib_material_response_b200.demag import _precond_blocks

    bj_blocks = _precond_blocks(w.get("groups"), w["chi"]) if (solver == "gmres" and world == 1) else None
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=ctx.device) if 8.0 * N * N < 0.5e9 else None

    def step():
        if flush is not None:
            flush.zero_()  # inputs smaller than L2 would otherwise stay cached
        ctx.assemble(d_pos, d_quat, d_dim, Q, ld, chi=d_chi, n_tgt_local=n_local, tgt_block=tgt_block,
                     tgt_nparts=world, tgt_part=rank)
        d_x.copy_(d_rhs)
        if world == 1:
            if solver == "lu":
                info = ctx.lu_factor(Q, ld, ipiv)
                assert info == 0
                ctx.lu_solve(Q, ld, ipiv, d_x)
                return None
            d_x.zero_()
            if bj_blocks:
                return ctx.gmres_bj(Q, ld, d_rhs, d_x, bj_blocks, tol=args.tol, restart=args.restart, maxit=5000)
            return ctx.gmres(Q, ld, d_rhs, d_x, tol=args.tol, restart=args.restart, maxit=5000)
        if solver == "lu":
            info = ctx.lu_solve_dist(N, 3 * n_local, tgt_block, Q, ld, d_x)
            assert info == 0
            return None
        d_x.zero_()
        return ctx.gmres_dist(N, 3 * n_local, tgt_block, Q, ld, d_rhs, d_x, tol=args.tol, restart=args.restart, maxit=5000)

    # The timed region records event pairs only for the kernel classes the line reports a roofline for (the LU
    # trailing GEMM, the assembly, the matvec, the triangular sweeps): every recorded launch costs two event
    # records of host time, which shows in the multi-GPU step where the panel chain exposes the host (5 ms of a
    # 68 ms step at 8 GPUs with everything recorded).  The full per-stage table comes from one extra, untimed
    # step after the timed region.  The last warm-up step runs with the same records so that the event pool
    # exists before the clock starts.
    timed_kinds = ["assemble", "gemm", "matvec", "trisolve"]
    for it in range(args.warmup):
        if it == args.warmup - 1:
            ctx.set_profiling(True, kinds=timed_kinds)
        step()
    barrier()
    peaks = load_peaks(ctx)  # FP64 peaks measured now, clocks already up from the warm-up steps
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ctx.set_profiling(True, kinds=timed_kinds)
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gm = None
    for _ in range(args.steps):
        gm = step()
    e1.record()
    barrier()
    t_dev = e0.elapsed_time(e1) * 1e-3
    prof = ctx.profile()
    launches = ctx.launch_count - launches0
    # one more step, untimed, with every class recorded: the per-stage table
    ctx.set_profiling(True)
    step()
    barrier()
    prof_all = ctx.profile()
    ctx.set_profiling(False)
    clocks = sampler.stop() if sampler else None
    if world > 1:
        tt = torch.tensor([t_dev], dtype=torch.float64, device=ctx.device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev = float(tt.item())
    t_step = t_dev / args.steps
    x_dev = d_x.cpu().numpy().reshape(n, 3)

    # residual of the last solution (not timed): ||Q x - b|| / ||b|| on a fresh Q
    ctx.assemble(d_pos, d_quat, d_dim, Q, ld, chi=d_chi, n_tgt_local=n_local, tgt_block=tgt_block,
                 tgt_nparts=world, tgt_part=rank)
    relres = None
    if world == 1:
        y = ctx.empty(N)
        ctx.matvec(Q, ld, d_x, y)
        relres = float(torch.linalg.norm(y - d_rhs) / torch.linalg.norm(d_rhs))
    else:
        # distributed true residual ||b - Q x|| / ||b|| (tol=1: report the residual of x, no iterations)
        _, relres = ctx.gmres_dist(N, 3 * n_local, tgt_block, Q, ld, d_rhs, d_x, tol=1.0, restart=1, maxit=1)
    del Q
    torch.cuda.empty_cache()

    # ---- end to end through the public APIs with HOST inputs/outputs -------------------------------
    # every rank makes the same call (SPMD) and gets the full solution back; wall clock between
    # synchronizes, max over ranks, K steps
    def time_calls(fn):
        fn()  # warm (allocations)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = fn()
        torch.cuda.synchronize()
        t = (time.perf_counter() - t0) / args.steps
        if world > 1:
            tt = torch.tensor([t], dtype=torch.float64, device=ctx.device)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt.item())
        return t, res

    e2e = e2e_arrays = None
    h2d = 8 * (n * 3 + n * 4 + n * 3 + N + N) * world  # pos, quat, dim, chi, rhs on every rank
    d2h = 8 * N * world
    if not args.no_e2e:
        kw = dict(solver=solver, tol=args.tol)
        t_arr, xs = time_calls(lambda: apply_demag_arrays(w["pos"], w["quat"], w["dim"], w["pol"], w["chi"], w["H_ext"],
                                                          groups=w.get("groups"), **kw))
        e2e_arrays = {"value": 9.0 * n * n / t_arr, "unit": "T-entries/s", "seconds_per_step": t_arr, "steps": args.steps,
                      "api": "magpylib_material_response_b200.demag.apply_demag_arrays (host numpy in/out)",
                      "max_abs_diff_vs_device_path": float(np.abs(rot.apply(xs) - x_dev).max())}
        t0 = time.perf_counter()
        coll = make_collection(name)
        t_mesh = time.perf_counter() - t0
        if coll is not None:
            t_obj, solved = time_calls(lambda: apply_demag(coll, **kw))
            pol_obj = collection_polarizations(solved)
            e2e = {"value": 9.0 * n * n / t_obj, "unit": "T-entries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "seconds_per_step": t_obj, "steps": args.steps, "timer": "time.perf_counter between synchronizes, max over ranks",
                   "api": "apply_demag(collection) -- reference signature (demag.py:324-333) on mesh_Cuboid objects; returns the "
                          "solved copy (inplace=False); Collection.copy(), gather, H2D, assembly, LU, D2H and write-back are inside",
                   "mesh_Cuboid_s": t_mesh, "object_api_overhead_s": t_obj - t_step,
                   "max_abs_diff_vs_device_path": float(np.abs(rot.apply(pol_obj) - x_dev).max())}
        else:
            e2e = dict(e2e_arrays, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h)

    # ---- BASELINE configs[3] on the ranks of this run (N >= 2) -------------------------------------
    config4 = None
    if world > 1 and not args.no_config4 and solver == "lu":
        del d_x, ipiv
        config4 = run_config4(ctx, world, rank, dist)
    elif world == 1:
        config4 = {"skipped": "64 000 cells = 295 GB of Q: needs >= 2 GPUs (run by bench.py --gpus N, N >= 2)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    tpath = ROOT / "profiles" / "ncu_traffic_r2.json"
    traffic = json.loads(tpath.read_text()) if tpath.exists() else {}
    asm_ms, asm_pairs, asm_cnt = prof["assemble"]
    gemm_ms, gemm_flops, gemm_cnt = prof["gemm"]
    mv_ms, mv_bytes, mv_cnt = prof["matvec"]
    fpp = FLOP_PER_PAIR_ROTATED if rotated else FLOP_PER_PAIR_ALIGNED
    ncu_pipe = traffic.get("assemble_cell_major", {}).get("fp64_pipe_frac_rotated" if rotated else "fp64_pipe_frac")
    roof_asm = {"bound": "fp64", "achieved": asm_pairs * fpp / (asm_ms * 1e-3) / 1e12 if asm_ms else None,
                "peak": peaks["fp64_fma_tflops"], "unit": "TFLOP/s (flop-equivalents of the reference algorithm, SURVEY 8d convention)",
                "frac": ncu_pipe, "frac_src": "ncu sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active of this kernel "
                                              "(profiles/, see ncu_traffic_r2.json); NOT achieved/peak: the far-field path evaluates "
                                              "3 atan2 + 3 log per pair where the reference algorithm counts 24 + 6",
                "traffic": traffic.get("assemble_cell_major", {}).get("ratio", 0) * 72.0 * asm_pairs / max(1, asm_cnt) or None,
                "traffic_src": "72 B/pair x ncu dram/algorithmic ratio (profiles/ncu_traffic_r2.json)",
                "kernel": "assemble_cell_major_kernel", "ms_per_launch": asm_ms / max(1, asm_cnt),
                "entries_per_s": 9.0 * asm_pairs / (asm_ms * 1e-3) if asm_ms else None, "peak_src": peaks["fp64_src"]}
    k_update = 256 if world == 1 else 3 * tgt_block  # rank of the trailing updates (two 128-column panels / one ownership block)
    if solver == "lu" and gemm_ms:
        ach = gemm_flops / (gemm_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": ach, "peak": peaks["fp64_dmma_tflops"], "unit": "TFLOP/s",
                    "frac": ach / peaks["fp64_dmma_tflops"],
                    "traffic": traffic.get("dgemm_nn", {}).get("ratio", 0) * (gemm_flops * 8.0 / k_update) / gemm_cnt or None,
                    "traffic_src": f"C read+write bytes (16 B per C element = flops*8/k, rank-{k_update} updates) x ncu dram/algorithmic ratio (profiles/ncu_traffic_r2.json)",
                    "kernel": "LU trailing update (DMMA.8x8x4 GEMM)",
                    "launches": gemm_cnt, "ms_per_launch": gemm_ms / gemm_cnt,
                    "peak_src": "FP64 DMMA peak: " + peaks["fp64_src"] + "; MEASURED_PEAKS.json has no FP64 entry"}
    elif mv_ms:
        ach = mv_bytes / (mv_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                    "traffic": None, "kernel": "matvec_rows_kernel (GMRES)", "launches": mv_cnt, "ms_per_launch": mv_ms / mv_cnt,
                    "peak_src": peaks["hbm_src"]}
    else:
        roofline = roof_asm

    stages = {k: {"ms": v[0], "launches": v[2]} for k, v in prof_all.items() if v[2]}
    stages["_note"] = "one untimed step after the timed region with every kernel class recorded (event pairs per launch)"
    lu_flops = (2.0 / 3.0) * N**3
    # stages overlap (look-ahead) / are per rank: use the step time minus assembly and the solve
    lu_ms = t_step * 1e3 - (prof["assemble"][0] + prof["trisolve"][0]) / args.steps
    worst_parity = max((r["max_rel_err"] for r in parity), default=None) if parity else None
    line = {
        "metric": "apply_demag_T_entries_per_s", "value": 9.0 * n * n / t_step, "unit": "T-entries/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "parity_check": parity, "relres_check": relres, "config4": config4,
        "apply_demag_s": t_step,
        "assembly_entries_per_s": roof_asm["entries_per_s"],
        "lu_factor_tflops": lu_flops / (lu_ms * 1e-3) / 1e12 if (solver == "lu" and lu_ms) else None,
        "e2e": e2e, "e2e_arrays": e2e_arrays,
        "gpu_launches": launches,
        "config": bench_config(name, w, solver),
        "sharding": {"row_block_cells": tgt_block, "ranks": world, "scheme": "block-cyclic target-row blocks" if world > 1 else "single GPU"},
        "gmres": {"iters": gm[0], "relres": gm[1], "block_jacobi_blocks": bj_blocks} if gm else None,
        "roofline": roofline,
        "clocks": clocks,
        "stages": stages,
        "roofline_assembly": roof_asm,
    }
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(name)
    # the facts a truncated tail must still show
    line["summary"] = {"n_gpus": world, "ms_per_step": t_step * 1e3, "parity_max_rel_err": worst_parity, "relres_check": relres,
                       "config4_apply_demag_s": (config4 or {}).get("apply_demag_s"), "config4_relres": (config4 or {}).get("relres"),
                       "e2e_object_api_s": (e2e or {}).get("seconds_per_step"), "roofline_frac": roofline.get("frac")}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, help="two_magnets_<n> | soft_cube_<n>")
    ap.add_argument("--solver", default="lu", choices=["lu", "gmres"])
    ap.add_argument("--dist-lu", action="store_true", help="(kept for compatibility; the distributed LU is the default)")
    ap.add_argument("--tol", type=float, default=1e-10)
    ap.add_argument("--restart", type=int, default=300, help="GMRES restart length")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiler runs only)")
    ap.add_argument("--no-parity", action="store_true", help="skip the pre-timing parity check against the oracle")
    ap.add_argument("--no-config4", action="store_true", help="skip the 64 000-cell step that follows the timed region at N >= 2")
    args = ap.parse_args()
    _claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
