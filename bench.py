#!/usr/bin/env python
"""bench.py -- cell-updates/s per RK stage of the per-iteration solve on B200, and the reference CPU arm.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 the CUDA path (one rank per GPU)
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W] the reference's own CPU solver (oracle/_ref)
    torchrun --nproc-per-node N ... bench.py --gpus N ...               multi-GPU (weak scaling, NCCL halo exchange)

Workload (BASELINE.json configs[4], the one the metric is quoted on): synthetic 3D hex polyMesh, 256^3 = 16.8 M
cells PER GPU (blocks 1x1x1, 2x1x1, 2x2x1, 2x2x2 for N = 1,2,4,8), M2 scheme + laminar viscous terms + sponge,
RK5, fp64.  One "step" is one time step = 5 RK stages over every cell of the job; value = cells * 5 * K / time.
The working set (about 7 GB per GPU) is far larger than the 126 MB L2, so no L2 flush is needed between steps.

Printed JSON keys follow the driver contract; see DESIGN.md section "Measurement" for roofline / cpu_baseline.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from lfm_public_b200 import host_api  # noqa: E402
from lfm_public_b200.tools import casegen, meshgen  # noqa: E402

BLOCKS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
MU = 7.17948717948718e-05          # examples/3D_Cylinder_Re3900/S thermophysicalProperties
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "lfm_solve_ref")
REF_BIN_SP = os.path.join(ROOT, "oracle", "_ref", "lfm_solve_ref_sp")


def b_alg(D, s, fbar):
    """Algorithmic bytes per cell per RK stage: gradient pass G + flux/update pass F."""
    G = s * ((D + 3) + (D * D + D)) + fbar * ((D + 1) * s + 8)
    F = s * (2 * (D + 2) + (D * D + D) + 2 + 2 * (D + 2)) + fbar * ((2 * D + 1) * s + 8)
    return G, F


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------------------
# synthetic block case (fields are analytic in the GLOBAL coordinates so that every rank of a job agrees)
# ----------------------------------------------------------------------------------------------------------
def block_fields(m, n, blocks, rank):
    nx, ny, nz = n
    bx, by, bz = blocks
    bi, bj, bk = rank % bx, (rank // bx) % by, rank // (bx * by)
    h = 1.0 / (nx * bx)
    L = (1.0, h * ny * by, h * nz * bz)
    xc = (np.arange(nx) + 0.5 + bi * nx) * h
    yc = (np.arange(ny) + 0.5 + bj * ny) * h
    zc = (np.arange(nz) + 0.5 + bk * nz) * h
    Z, Y, X = np.meshgrid(zc, yc, xc, indexing="ij")      # cell id = i + nx*(j + ny*k)
    X, Y, Z = X.ravel(), Y.ravel(), Z.ravel()
    sx, sy, sz = 2 * np.pi * X / L[0], 2 * np.pi * Y / L[1], 2 * np.pi * Z / L[2]
    p = 1.0 + 0.01 * np.sin(sx) * np.cos(sy)
    T = np.ones_like(p)
    U = np.stack([0.2 * (1.0 + 0.05 * np.sin(sy)), 0.01 * np.sin(sx), 0.01 * np.sin(sz)], axis=1)
    alpha = np.minimum(X, L[0] - X)                       # distance to the inlet / outlet planes
    return dict(p=p, T=T, U=U, alpha=alpha), h


def build_rank_case(n, blocks, rank, n_ranks, precision, scheme, tile=None, brick_order="morton", z_cyclic=None):
    # z_cyclic=None: z periodic while there is one block in z, walls otherwise; True: periodic for every block layout (the
    # cut periodic plane becomes processorCyclic patches)
    m = meshgen.hex_block((n, n, n), blocks, rank, tile=tile, brick_order=brick_order, z_cyclic=z_cyclic)
    fields, h = block_fields(m, (n, n, n), blocks, rank)
    if "new_of_old" in m:                                  # fields were built in lexicographic order: renumber them too
        perm = m["new_of_old"]
        for k, v in fields.items():
            out = np.empty_like(v)
            out[perm] = v
            fields[k] = out
    dt = 0.5 * h / 1.2                                     # acoustic CFL ~ 0.5 (c = 1, |U| = 0.2)
    o = host_api.default_opts(solver=scheme, dimension=3, delta_t=dt, Ls=0.15, mu0=MU, mach=0.2, comm_type=2,
                              double_precision=1 if precision == 8 else 0)
    o.U_inf[0] = 0.2
    case = host_api.Case.from_mesh(m, o, fields, rank=rank, n_ranks=n_ranks)
    return case, dt


# ----------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        exe = shutil.which("nvidia-smi")
        if not exe:
            return
        try:
            self.proc = subprocess.Popen([exe, f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return None
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))


# ----------------------------------------------------------------------------------------------------------
# reference CPU arm: the reference's own solver binary on a bounded sample of the workload
# ----------------------------------------------------------------------------------------------------------
REF_BLOCKS = {1: None, 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2), 16: (4, 2, 2), 32: (4, 4, 2), 64: (4, 4, 4)}


def reference_case(tmp, n_sample, n_ranks, scheme, precision):
    """Writes the reference's input (an OpenFOAM case: n_sample^3 hex box in the mesh generator's lexicographic numbering,
    decomposed into n_ranks blocks, the bench workload's dictionaries and fields) into `tmp`; returns (dt, block layout)."""
    blocks = REF_BLOCKS[n_ranks]
    m = meshgen.hex_box(n_sample, n_sample, n_sample, lengths=(1.0, 1.0, 1.0), z_cyclic=(blocks is None or blocks[2] == 1))
    h = 1.0 / n_sample
    dt = 0.5 * h / 1.2
    cr = meshgen.block_assignment(m, blocks) if blocks else None
    fields, _ = block_fields(m, (n_sample, n_sample, n_sample), (1, 1, 1), 0)
    casegen.write_case(tmp, m, fields=fields, cell_rank=cr, solver=scheme, dimension=3, deltaT=dt, endTime=dt,
                       writeInterval=10 ** 7, Ls=0.15, mu=MU, haveResiduals=False, printInfoFreq=10 ** 6,
                       doublePrecision=(precision == 8), commType=2)
    return dt, blocks


def reference_exec(tmp, dt, n_sample, n_ranks, steps, precision):
    """Runs oracle/_ref/lfm_solve_ref[_sp] (the reference's unmodified src/*.cpp over the mini-MPI shim: one forked process per
    rank) for `steps` time steps on the case in `tmp`; cell-stages/s from the reference's own `Rk Loop` timer
    (src/mesh_solver.cpp:906-919)."""
    exe = REF_BIN if precision == 8 else REF_BIN_SP
    if not os.path.exists(exe):
        raise RuntimeError(f"{exe} missing (built by `make -C oracle ref` where /root/reference exists)")
    casegen.set_end_time(tmp, dt * steps)
    env = dict(os.environ)
    args = [exe]
    if REF_BLOCKS[n_ranks] is not None:
        env["LFM_MPI_NP"] = str(n_ranks)
        args.append("-p")
    out = subprocess.run(args, cwd=tmp, env=env, capture_output=True, text=True, timeout=1800)
    if "Simulation finished successfully" not in out.stdout:
        raise RuntimeError("reference run failed: " + out.stdout[-1500:] + out.stderr[-1500:])
    rk = float(re.search(r"\[\s*([0-9.eE+-]+)\]: Rk Loop", out.stdout).group(1))
    return dict(cells=n_sample ** 3, steps=steps, rk_loop_s=rk, value=n_sample ** 3 * 5 * steps / rk)


def reference_run(n_sample, n_ranks, steps, scheme, precision, warm_steps=0):
    """One bounded sample of the workload through the reference binary: case written once, optional untimed warm-up run
    (page cache, binary), then the timed run."""
    tmp = tempfile.mkdtemp(prefix="lfm_ref_")
    try:
        dt, _ = reference_case(tmp, n_sample, n_ranks, scheme, precision)
        if warm_steps > 0:
            reference_exec(tmp, dt, n_sample, n_ranks, warm_steps, precision)
        return reference_exec(tmp, dt, n_sample, n_ranks, steps, precision)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def pick_cpu_ranks():
    cores = os.cpu_count() or 1
    p = 1
    while p * 2 <= min(cores, 64):
        p *= 2
    return p, cores


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ranks, cores = pick_cpu_ranks()
    n_sample = args.ref_arm_n
    steps = max(1, args.steps)
    r = reference_run(n_sample, ranks, steps, args.scheme, args.precision, warm_steps=max(0, min(args.warmup, 2)))
    ms = 1e3 * r["rk_loop_s"] / steps
    bl = REF_BLOCKS[ranks] or (1, 1, 1)
    sample = (f"{n_sample}^3 hex box ({r['cells']} cells) in {ranks} blocks, {steps} time steps x 5 RK stages, reference binary "
              f"oracle/_ref/lfm_solve_ref over the mini-MPI shim (one process per rank), reference's own 'Rk Loop' timer")
    # a second, smaller sample shows how the CPU figure depends on the mesh size (cache residency vs halo share)
    small = None
    if args.ref_n != n_sample:
        try:
            rs = reference_run(args.ref_n, ranks, max(steps, args.ref_steps), args.scheme, args.precision)
            small = {"value": rs["value"], "sample": f"{args.ref_n}^3 hex box ({rs['cells']} cells) in {ranks} blocks, {rs['steps']} steps"}
        except Exception as e:
            small = {"value": None, "sample": f"failed: {e}"[:200]}
    cfg = workload_config(args, args.gpus)
    # this line's workload is the mesh the CPU arm actually advanced: a bounded sample of the GPU arm's workload (same
    # scheme, dictionaries, fields and RK5; smaller box, lexicographic numbering, one block per host core)
    cfg["workload"] = (f"bounded CPU sample of the GPU arm's workload: synthetic 3D hex polyMesh {n_sample}^3 = {n_sample ** 3} cells in "
                       f"{bl[0]}x{bl[1]}x{bl[2]} blocks (one per host core used), {'M2' if args.scheme == 1 else 'M1'} + laminar viscous + "
                       f"sponge, RK5, commType 2; GPU arm: {args.n}^3 cells per GPU")
    cfg["cell_numbering"] = "lexicographic (mesh generator order), decomposed into blocks"
    cfg["cells_per_gpu"] = None
    cfg["total_cells"] = n_sample ** 3
    cfg["l2_policy"] = "n/a (CPU)"
    line = {
        "impl": "reference", "metric": "cell-updates/s per RK stage", "value": r["value"], "unit": "cell-updates/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64" if args.precision == 8 else "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": r["value"], "unit": "cell-updates/s", "cores": ranks, "kind": "reference", "sample": sample,
                         "host_cores": cores, "smaller_sample": small,
                         "caveat": "the ranks talk through the in-tree mini-MPI shim (forked processes, AF_UNIX sockets), not a tuned MPI; "
                                   "the sample is smaller than one GPU's block"},
        "e2e": {"value": r["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def numbering_of(args):
    """(tile argument of meshgen.hex_block, description) of the cell numbering the bench mesh is generated in."""
    if args.numbering == "lex":
        return None, "lexicographic (the mesh generator's order, no renumbering)"
    if args.numbering == "morton":
        return "morton", "Z-order curve through the cell centres == `python -m lfm_public_b200.tools.renumber --method morton` on this block"
    tile = tuple(int(x) for x in args.tile.split(","))
    return tile, f"bricks {args.tile} aligned with the interior submesh, {args.brick_order} brick order (hand-fitted to the 128-cell tiles)"


def workload_config(args, n_gpus):
    n = args.n
    bl = BLOCKS[n_gpus]
    return {"workload": f"synthetic 3D hex polyMesh weak scaling, {n}^3 = {n ** 3} cells per GPU, blocks {bl[0]}x{bl[1]}x{bl[2]} "
                        f"({n ** 3 * n_gpus} cells), {'M2' if args.scheme == 1 else 'M1'} + laminar viscous + sponge, RK5, commType 2"
                        + (", z periodic (processorCyclic)" if not getattr(args, "z_walls", False) and bl[2] > 1 else ""),
            "cell_numbering": numbering_of(args)[1], "cells_per_gpu": n ** 3, "total_cells": n ** 3 * n_gpus, "rk_stages_per_step": 5,
            "l2_policy": "inputs larger than L2 (about 7 GB of state per GPU at 256^3), no flush"}


# ----------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------
def quick_measure(n, precision, scheme, tile, brick_order, steps, device=0):
    """Device-resident throughput of one single-rank block (sub-records of the N = 1 line): ms per step, cell-updates/s,
    per-kernel times and the two roofline fractions, on a freshly built case."""
    from lfm_public_b200 import gpu_api
    case, dt = build_rank_case(n, (1, 1, 1), 0, 1, precision, scheme, tile, brick_order)
    case.finish()
    g = gpu_api.GpuSolver(case, device)
    try:
        g.warmup()
        g.step(scheme, dt, 3)
        g.sync()
        g.event_record(0)
        g.step(scheme, dt, steps)
        g.event_record(1)
        g.sync()
        ms = g.event_elapsed_ms(0, 1) / steps
        g.enable_kernel_timing(True)
        g.step(scheme, dt, 2)
        g.sync()
        kt = {}
        for name in ("tile_stage", "tile_grad"):
            t, nl = g.kernel_time(name)
            if nl:
                kt[name] = t / 2
        g.enable_kernel_timing(False)
        peak, _ = measured_peak()
        Gb, Fb = b_alg(g.D, precision, 3.0)
        nc = g.n_cells
        out = {"cells": nc, "ms_per_step": ms, "value": nc * 5 / (ms * 1e-3), "kernel_ms_per_step": kt,
               "stage_frac": (Gb + Fb) * nc * 5 / (ms * 1e-3) / 1e9 / peak, "tiles": g.tile_info()}
        if "tile_stage" in kt:
            out["frac"] = Fb * nc * 5 / (kt["tile_stage"] * 1e-3) / 1e9 / peak
        return out
    finally:
        g.close()
        case.close()


def bind_to_gpu_cores(local_rank, world):
    """Pins this rank to its own share of the cores NVML lists as close to its GPU (pinned host buffers are allocated after
    this, so they land on that NUMA node): eight ranks sharing every core of the host was one suspect of the round-1 e2e
    collapse at N = 8.  Returns what was done (goes into the e2e record)."""
    if world == 1:
        return "unchanged (one rank: the CPU baseline that follows uses every core)"
    try:
        import pynvml
        pynvml.nvmlInit()
        hdl = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(hdl, (os.cpu_count() + 63) // 64)
        cores = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if not allowed:
            return "nvml affinity empty: unchanged"
        per = max(1, len(allowed) // max(1, world))
        mine = allowed[(local_rank * per) % len(allowed):][:per] or allowed
        os.sched_setaffinity(0, mine)
        return f"cores {mine[0]}-{mine[-1]} of the {len(allowed)} NVML lists for GPU {local_rank}"
    except Exception as e:
        return f"unchanged ({type(e).__name__})"


def nccl_parity_check(args, dist, rank, world, local_rank, n=24, steps=3):
    """Outside the timed region, N > 1 only: a small block case (n^3 cells per rank, the bench workload's layout and
    dictionaries) advanced `steps` time steps (a) on this rank's GPU with the halo exchange over NCCL and (b) by the CPU
    restatement of the reference (oracle/, the checker) for ALL ranks in lockstep inside this process.  Conservatives and
    ghost states of this rank must be bit-identical (fp64; fp32: relative 1e-5).  Returns (ok, sha256 of q on the GPU)."""
    import hashlib
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from lfm_public_b200 import gpu_api
    blocks = BLOCKS[world]
    z_cyc = None if args.z_walls else True
    cases = []
    for r in range(world):
        c, dt = build_rank_case(n, blocks, r, world, args.precision, args.scheme, "morton", args.brick_order, z_cyclic=z_cyc)
        cases.append(c)
    host_api.exchange_in_process(cases)
    oracles = [oracle_lib.Oracle(c) for c in cases]
    oracle_lib.run(oracles, args.scheme, dt, steps, first=True)
    want_q, want_g = oracles[rank].download(0), oracles[rank].download(7)
    g = gpu_api.GpuSolver(cases[rank], local_rank)
    ids = [gpu_api.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    g.comm_init_nccl(ids[0], rank, world)
    g.warmup()
    g.step(args.scheme, dt, steps)
    g.sync()
    got_q, got_g = g.download(0), g.download(7)
    if args.precision == 8:
        ok = bool(np.array_equal(got_q, want_q) and np.array_equal(got_g, want_g))
    else:
        ok = bool(np.abs(got_q - want_q).max() <= 1e-5 * np.abs(want_q).max())
    digest = hashlib.sha256(np.ascontiguousarray(got_q).tobytes()).hexdigest()
    g.close()
    for o in oracles:
        o.close()
    for c in cases:
        c.close()
    return ok, digest


def run_gpu_arm(args):
    from lfm_public_b200 import gpu_api
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one process per GPU)")
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_
        dist = dist_
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if gpu_api.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    blocks = BLOCKS[world]
    affinity = bind_to_gpu_cores(local_rank, world)
    parity_nccl = None
    if world > 1 and not args.no_parity:
        ok, digest = nccl_parity_check(args, dist, rank, world, local_rank)
        import torch
        t = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        parity_nccl = bool(t.item() == 1.0)
        if not ok:
            print(f"bench.py: rank {rank}: NCCL-exchanged run differs from the oracle (sha256 of q {digest})", file=sys.stderr, flush=True)
    t0 = time.time()
    tile = numbering_of(args)[0]
    z_cyc = None if args.z_walls else True          # z stays periodic in every block layout (processorCyclic at 2x2x2)
    case, dt = build_rank_case(args.n, blocks, rank, world, args.precision, args.scheme, tile, args.brick_order, z_cyclic=z_cyc)
    if world > 1:
        host_api.exchange_distributed(case, rank)
    else:
        case.finish()
    t_setup = time.time() - t0
    g = gpu_api.GpuSolver(case, local_rank)
    g.set_option("use_tiles", args.use_tiles)
    if world > 1:
        import torch
        ids = [gpu_api.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        g.comm_init_nccl(ids[0], rank, world)
    n_cells = g.n_cells
    NQ, D, s = g.NQ, g.D, args.precision
    K, W = args.steps, args.warmup

    def barrier():
        g.sync()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput -------------------------------------------------------------------------
    g.warmup()
    g.step(args.scheme, dt, W)
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = g.launch_count
    g.event_record(0)
    g.step(args.scheme, dt, K)
    g.event_record(1)
    barrier()
    ms_total = max_over_ranks(g.event_elapsed_ms(0, 1))
    launches = g.launch_count - l0
    clk = clocks.stop() if rank == 0 else None
    q_check = g.download(0)
    finite = bool(np.isfinite(q_check).all())
    del q_check

    # ---- per-kernel device times (separate pass, events around every launch) ---------------------------------
    g.enable_kernel_timing(True)
    g.step(args.scheme, dt, max(1, min(K, 3)))
    g.sync()
    ktimes = {}
    for name in ("tile_stage", "tile_grad", "k_flux_face", "k_update_cell", "k_grad_cell", "k_set_bc", "k_pack", "k_unpack"):
        ms, nl = g.kernel_time(name)
        if nl:
            ktimes[name] = (ms, nl)
    g.enable_kernel_timing(False)

    # ---- end-to-end through the C ABI with host buffers ----------------------------------------------------------
    # Every step (= one batch) uploads its conservatives from pinned host memory, advances one time step and downloads the
    # result to pinned host memory.  Batches are pipelined (lfmgpu_pipe_*): the upload of batch k+1 and the download of
    # batch k-1 run on copy streams while batch k computes; the timed region holds exactly Ke uploads, Ke steps, Ke downloads.
    real = np.float64 if s == 8 else np.float32
    pin_in = [gpu_api.PinnedArray((NQ, n_cells), real) for _ in range(2)]
    pin_out = [gpu_api.PinnedArray((NQ, n_cells), real) for _ in range(2)]
    g.download_q_soa_async(pin_in[0].ptr, pin_in[0].nbytes)
    g.sync()
    pin_in[1].array[...] = pin_in[0].array
    Ke = max(1, min(K, 5))
    g.pipe_in_start(pin_in[0].ptr, pin_in[0].nbytes)
    for it in range(1 + Ke):
        if it == 1:
            barrier()
            g.event_record(2)
        g.pipe_in_commit()
        g.pipe_in_start(pin_in[(it + 1) % 2].ptr, pin_in[(it + 1) % 2].nbytes)     # next batch, overlaps this batch's compute
        g.step(args.scheme, dt, 1)
        g.pipe_out_start()
        g.pipe_out_fetch(pin_out[it % 2].ptr, pin_out[it % 2].nbytes)
    g.event_record(3)
    barrier()
    ms_e2e = max_over_ranks(g.event_elapsed_ms(2, 3)) / Ke
    # the same loop without the time step: what the host <-> device path alone allows at this rank count (the ceiling of e2e)
    g.pipe_in_start(pin_in[0].ptr, pin_in[0].nbytes)
    for it in range(1 + Ke):
        if it == 1:
            barrier()
            g.event_record(4)
        g.pipe_in_commit()
        g.pipe_in_start(pin_in[(it + 1) % 2].ptr, pin_in[(it + 1) % 2].nbytes)
        g.pipe_out_start()
        g.pipe_out_fetch(pin_out[it % 2].ptr, pin_out[it % 2].nbytes)
    g.event_record(5)
    barrier()
    ms_copy = max_over_ranks(g.event_elapsed_ms(4, 5)) / Ke
    e2e_ok = bool(np.isfinite(pin_out[0].array).all() and np.isfinite(pin_out[1].array).all() and pin_out[Ke % 2].array.std() > 0)
    for p_ in pin_in + pin_out:
        p_.free()

    tile_info = g.tile_info()
    total_cells = n_cells * world if dist is None else int(max_over_ranks(float(n_cells))) * world
    ms_step = ms_total / K
    value = total_cells * 5 / (ms_step * 1e-3)
    e2e_value = total_cells * 5 / (ms_e2e * 1e-3)

    if rank != 0:
        g.close()
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        if parity_nccl is False:
            raise SystemExit(3)
        return

    # ---- roofline of the dominant kernel ------------------------------------------------------------------------------
    peak, peak_src = measured_peak()
    fbar = 3.0
    Gb, Fb = b_alg(D, s, fbar)
    alg = {"tile_stage": Fb, "tile_grad": Gb, "k_grad_cell": Gb, "k_flux_face": Fb, "k_update_cell": Fb}
    dom = max((k for k in ktimes if k in alg), key=lambda k: ktimes[k][0], default=None)
    roofline = None
    if dom:
        ms, nl = ktimes[dom]
        per_launch_cells = n_cells * (max(1, min(K, 3)) * 5) / nl      # cells one launch covers
        achieved = alg[dom] * per_launch_cells / (ms / nl * 1e-3) / 1e9
        traffic, traffic_note = None, None
        try:
            # measured DRAM bytes per launch from the committed ncu capture -- valid only for the very binary that was profiled:
            # the capture is stamped with the SHA-256 of the library's device code (.nv_fatbin; scripts/ncu_traffic.py,
            # tools/libstamp.py) and dropped when the kernels differ
            from lfm_public_b200.tools.libstamp import device_code_sha256
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[dom]["f64" if s == 8 else "f32"]
            so = device_code_sha256(os.path.join(ROOT, "lfm_public_b200", "liblfmgpu.so"))
            if abs(per_launch_cells - n_cells) > 0.5:
                # (a rank with neighbours launches the kernel once per submesh: `achieved` is per average launch there, the capture
                # is of a launch over the whole rank)
                traffic_note = "the committed ncu capture is of a launch over the whole rank; this run launches per submesh: not reported"
            elif tr["n"] == args.n and tr.get("numbering") == args.numbering and tr.get("fatbin_sha256") == so:
                traffic = tr["bytes_per_launch"]
            else:
                traffic_note = "the committed ncu capture is of another binary or workload: not reported"
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src, "alg_bytes_per_cell": alg[dom],
                    "avg_launch_ms": ms / nl, "kernel_ms_per_step": {k: v[0] / max(1, min(K, 3)) for k, v in ktimes.items()},
                    "stage_frac": (Gb + Fb) * n_cells * 5 / (ms_step * 1e-3) / 1e9 / peak}

    # ---- CPU baseline: the reference binary on a bounded sample (rank 0, N == 1 only) ----------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            ranks, cores = pick_cpu_ranks()
            r = reference_run(args.ref_n, ranks, args.ref_steps, args.scheme, args.precision)
            cpu = {"value": r["value"], "unit": "cell-updates/s", "cores": ranks, "kind": "reference", "host_cores": cores,
                   "sample": f"{args.ref_n}^3 hex box in {ranks} blocks, {args.ref_steps} steps x 5 RK stages, oracle/_ref/lfm_solve_ref "
                             f"(reference src/*.cpp unchanged, mini-MPI shim), reference 'Rk Loop' timer = {r['rk_loop_s']:.3f} s"}
        except Exception as e:      # the baseline is reported, never required for the GPU number
            cpu = {"value": None, "unit": "cell-updates/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"[:300]}

    # ---- sub-records of the N = 1 line: the fp32 arm of the sweep and the same block in the other cell numberings ----------
    extras = None
    if world == 1 and not args.no_extras:
        g.close()
        case.close()
        g = None
        extras = {}
        try:
            if args.precision == 8:
                r = quick_measure(args.n, 4, args.scheme, tile, args.brick_order, max(3, min(K, 5)), local_rank)
                extras["f32"] = dict(r, workload=f"the headline workload in fp32 ({args.n}^3 cells, {numbering_of(args)[0]!r} numbering)")
            variants = {"morton": "morton", "bricks": tuple(int(x) for x in args.tile.split(",")), "lex": None}
            ne = args.extras_n
            extras["numbering"] = {"n": ne, "precision": args.precision,
                                   "note": f"{ne}^3 cells, same scheme; morton = tools/renumber.py --method morton, bricks = {args.tile} bricks "
                                           "aligned with the interior submesh (hand-fitted), lex = the generator's order"}
            for name, tl in variants.items():
                extras["numbering"][name] = quick_measure(ne, args.precision, args.scheme, tl, args.brick_order, 5, local_rank)
        except Exception as e:      # sub-records never take the headline down
            extras["error"] = str(e)[:300]

    line = {
        "metric": "cell-updates/s per RK stage", "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64" if s == 8 else "f32", "data": "synthetic", "config": workload_config(args, world),
        "roofline": roofline, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "cell-updates/s", "h2d_bytes_per_step": int(NQ * n_cells * s) * world,
                "d2h_bytes_per_step": int(NQ * n_cells * s) * world, "ms_per_step": ms_e2e, "finite": e2e_ok,
                "copy_only_ms_per_step": ms_copy, "copy_only_ceiling": total_cells * 5 / (ms_copy * 1e-3),
                "frac_of_copy_ceiling": ms_copy / ms_e2e, "cpu_affinity": affinity,
                "mode": "pipelined batches: H2D of batch k+1 and D2H of batch k-1 overlap the step of batch k (pinned host buffers)"},
        "gpu_launches": int(launches), "clocks": clk, "finite": finite, "setup_s": t_setup,
        "tiles": tile_info, "use_tiles": args.use_tiles, "extra": extras,
    }
    if parity_nccl is not None:
        line["parity_nccl"] = parity_nccl
        line["parity_nccl_case"] = "24^3 cells per rank, the bench layout, 3 time steps over NCCL vs the lockstep CPU oracle, q + ghost states bit-exact"
    print(json.dumps(line), flush=True)
    if g is not None:
        g.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if parity_nccl is False:
        raise SystemExit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gpu", choices=["gpu", "reference"])
    ap.add_argument("--n", type=int, default=int(os.environ.get("LFM_BENCH_N", "256")), help="cells per side of each GPU's block")
    ap.add_argument("--precision", type=int, default=8, choices=[4, 8])
    ap.add_argument("--scheme", type=int, default=1, choices=[0, 1], help="0 = M1, 1 = M2")
    ap.add_argument("--use-tiles", type=int, default=1)
    ap.add_argument("--tile", default="8,4,4", help="brick numbering of the synthetic mesh (renumberMesh analogue), or 'none'")
    ap.add_argument("--brick-order", default="morton", choices=["lex", "morton"], help="order of the bricks in the numbering")
    ap.add_argument("--ref-n", type=int, default=64, help="cells per side of the CPU sample timed next to the GPU arm (cpu_baseline)")
    ap.add_argument("--ref-arm-n", type=int, default=int(os.environ.get("LFM_REF_ARM_N", "128")), help="cells per side of the --impl reference workload")
    ap.add_argument("--ref-steps", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--z-walls", action="store_true", help="close z with walls when the block layout cuts it (8 GPUs); default: z stays periodic through processorCyclic patches")
    ap.add_argument("--numbering", default=os.environ.get("LFM_BENCH_NUMBERING", "morton"), choices=["morton", "bricks", "lex"],
                    help="cell numbering of the synthetic mesh: morton = what tools/renumber.py --method morton produces (default), "
                         "bricks = --tile bricks aligned with the interior submesh (hand-fitted), lex = none")
    ap.add_argument("--no-parity", action="store_true", help="skip the NCCL-transport parity check of the multi-GPU runs")
    ap.add_argument("--no-extras", action="store_true", help="skip the sub-records (fp32 arm, other numberings) of the N = 1 line")
    ap.add_argument("--extras-n", type=int, default=128, help="cells per side of the block the numbering variants are timed on")
    args = ap.parse_args()
    if args.gpus not in BLOCKS:
        raise SystemExit("--gpus must be 1, 2, 4 or 8")
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
