#!/usr/bin/env python
"""bench.py -- throughput of the 2-D FDTD time-stepping path on B200.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W`
prints ONE JSON line on rank 0.  For N > 1 it is launched under torchrun, one
rank per GPU.

Workload (BASELINE.json configs[4], the one the 1/2/4/8-GPU metric is quoted on):
zigzagModel, TM UPML (solver id 2), 16384 x 16384 cells PER GPU, stacked along y
(global grid 16384 x 16384*N), h_u = 10 nm, pml = 10, lambda = 500 nm, angle 0.
A "step" is one update(): H phase, E phase + source, NTFF surface sample; the
deferred NTFF projection of the K timed steps is inside the timed region too.
9 complex fields x 4 GiB per GPU -- far larger than the 126 MB L2, so no flush is
needed between iterations.

  value      whole-job Gcell-updates/s, state resident in HBM, CUDA-event timed on
             the engine's stream, max over ranks.
  e2e        the same metric through the host-facing C API with HOST buffers in
             the timed region: eps map H2D from pinned memory, K x update(), Ez
             D2H into the pinned mirror the getter hands out.
  roofline   H-phase kernel alone (the dominant kernel): algorithmic bytes
             (144 B/cell: reads Ez,Mx,Bx,My,By, writes Mx,Bx,My,By; Hx/Hy are not stored,
             the E phase forms them as B/mu0) over its mean duration, against
             MEASURED_PEAKS.json hbm_gbs.  `e_phase` is the other kernel (120 B/cell);
             `step` carries the contract figure of SURVEY 8(d): 264 B/cell-update x rate,
             which is exactly what the two kernels move.
  lean_interior
             the same K steps with B200FDTD_OPT_LEAN_INTERIOR (opt-in tolerance form: cells
             outside the absorbing frame advance B / D directly, 168 instead of 264 B per TM
             cell-update; fields within 1e-12 of the reference instead of bit-identical).
             Reported BESIDE `value`, which stays the reference's arithmetic in every cell;
             `--lean` makes it the measured form of the whole line instead.
  cpu_baseline / --impl reference
             the UNMODIFIED reference (oracle/_ref/libref.so, built from
             /root/reference) on the host cores, one serial solver instance per
             core (how main.c uses its MPI ranks), on a bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

N_PER_GPU = 16384
BYTES_STEP_TM = 264          # SURVEY 8(d) contract figure = what the step moves: 144 + 120
BYTES_H_TM = 144             # H-phase kernel: reads Ez,Mx,Bx,My,By (80) + writes Mx,Bx,My,By (64)
BYTES_E_TM = 120             # E-phase kernel: reads Bx,By,Jz,Dz (64) + eps (8) + writes Jz,Dz,Ez (48)
FALLBACK_HBM_GBS = 6650.0    # B200_PROFILING.md fallback


def ncu_traffic(kernel_substr, cells):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    of this same command (profiles/*_full.json, written by scripts/ncu_summary.py); None when no
    capture at this launch size is on file."""
    import glob
    for path in sorted(glob.glob(os.path.join(HERE, "profiles", "*_full.json")), reverse=True):
        try:
            for rec in json.load(open(path))["launches"]:
                name = rec["kernel"].replace("double, ", "").replace(" ", "")      # "<double, 0>" == "<0>"
                if kernel_substr in name and abs(rec.get("cells", 0) - cells) < 0.01 * cells:
                    return rec["traffic_bytes"], os.path.relpath(path, HERE)
        except Exception:
            continue
    return None, None


def measured_hbm_peak():
    try:
        with open(os.path.join(HERE, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


# ----------------------------------------------------------------------------
# CPU arm: the reference itself, one serial instance per host core
# ----------------------------------------------------------------------------
CPU_WORKER = r"""
import os, sys, time
sys.path.insert(0, %(here)r)
core, n, steps, warm, with_ntff = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
try:
    os.sched_setaffinity(0, {core})
except Exception:
    pass
devnull = os.open(os.devnull, os.O_WRONLY)
saved = os.dup(1); os.dup2(devnull, 1)          # the reference printf()s a lot
from oracle import reflib
sim = reflib.RefSim("ZIGZAG", "TM_UPML_2D", n, steps=steps + warm)
if with_ntff:
    sim.step(warm)
    t0 = time.perf_counter(); sim.step(steps); dt = time.perf_counter() - t0
else:
    sim.step_fn("refhook_tm_upml_update_no_ntff", warm)
    t0 = time.perf_counter(); sim.step_fn("refhook_tm_upml_update_no_ntff", steps); dt = time.perf_counter() - t0
os.dup2(saved, 1)
print("RATE %%.6e %%.6f" %% (n * n * steps / dt, dt))
"""


def run_cpu_reference(n, steps, warm, with_ntff=True, max_workers=None):
    """Aggregate cell-updates/s of R concurrent reference instances (R = host cores)."""
    from oracle import reflib
    if not reflib.available():
        return None
    try:
        cores = sorted(os.sched_getaffinity(0))
    except Exception:
        cores = list(range(os.cpu_count() or 1))
    if max_workers:
        cores = cores[:max_workers]
    code = CPU_WORKER % {"here": HERE}
    procs = [subprocess.Popen([sys.executable, "-c", code, str(c), str(n), str(steps), str(warm),
                               "1" if with_ntff else "0"],
                              stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
                              cwd=tempfile.gettempdir())
             for c in cores]
    rates, times = [], []
    for p in procs:
        out, _ = p.communicate()
        for line in out.splitlines():
            if line.startswith("RATE"):
                rates.append(float(line.split()[1]))
                times.append(float(line.split()[2]))
    if not rates:
        return None
    return {"rate": sum(rates), "per_core": statistics.mean(rates), "cores": len(rates),
            "max_loop_s": max(times)}


def reference_arm(args):
    """bench.py --impl reference: the reference's own CPU path, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = args.cpu_n
    t0 = time.perf_counter()
    res = run_cpu_reference(n, args.steps, args.warmup, with_ntff=True)
    wall = time.perf_counter() - t0
    if res is None:
        print(json.dumps({"impl": "reference", "unavailable":
                          "oracle/_ref/libref.so missing (built only where /root/reference exists)"}))
        return 0
    value = res["rate"] / 1e9
    sample = ("%d concurrent serial TM_UPML instances (one per host core, as main.c uses MPI "
              "ranks), each zigzagModel %dx%d, %d timed steps after %d warm-up, NTFF included"
              % (res["cores"], n, n, args.steps, args.warmup))
    line = {
        "impl": "reference", "metric": "Gcell-updates/s", "value": value, "unit": "Gcell-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * res["max_loop_s"] / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus, sample_note="CPU arm runs a bounded sample: " + sample, n=args.n),
        "cpu_baseline": {"value": value, "unit": "Gcell-updates/s", "cores": res["cores"],
                         "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "Gcell-updates/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall,
    }
    print(json.dumps(line))
    return 0


MODEL_NAMES = {"ZIGZAG": "zigzagModel", "LAYER": "multiLayerModel", "MORPHO_SCALE": "morphoScaleModel",
               "MIE_CYLINDER": "MieCylinderModel", "NO_MODEL": "noModel"}


def workload_config(n_gpus, sample_note=None, n=N_PER_GPU, solver="TM_UPML_2D", halo="peer", model="ZIGZAG",
                    strong=False):
    n_py = n if strong else n * n_gpus
    per_gpu = n * n_py // n_gpus
    cfg = {"workload": "%s %s %s scaling, %d x %d cells per GPU, global %d x %d, y-slabs"
                       % (MODEL_NAMES[model], solver, "strong" if strong else "weak", n, n_py // n_gpus, n, n_py),
           "baseline_config": ("BASELINE.json configs[2]" if strong else "BASELINE.json configs[4]"),
           "h_u_nm": 10, "pml": 10, "lambda_nm": 500, "angle_deg": 0,
           "cells_per_gpu": per_gpu,
           "l2_policy": "working set %.1f GiB per GPU >> 126 MB L2, no flush needed" % (per_gpu * 152.0 / 2**30),
           "parallelism": "y-slab x%d" % n_gpus}
    if n_gpus > 1:
        cfg["halo"] = halo
    if sample_note:
        cfg["note"] = sample_note
    return cfg


# ----------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = os.path.join(tempfile.gettempdir(), "bench_clocks_%d.csv" % os.getpid())
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                sm.append(float(f[1]))
                smax = float(f[2])
                for name, val in zip(names, f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        except Exception:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------
def gpu_arm(args):
    import numpy as np
    import torch
    from mpifdtd_b200 import binding as B
    from mpifdtd_b200.slab import SlabRun, TorchHaloComm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if B.device_count() < 1:
        raise SystemExit("no CUDA device: this benchmark has no CPU fallback for the GPU arm")

    # CPU baseline first (before this process holds a CUDA context), rank 0 at N=1 only
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res = run_cpu_reference(args.cpu_n, args.cpu_steps, 2, with_ntff=True)
        res_stencil = run_cpu_reference(args.cpu_n, args.cpu_steps, 2, with_ntff=False)
        if res is not None:
            cpu = {"value": res["rate"] / 1e9, "unit": "Gcell-updates/s", "cores": res["cores"],
                   "kind": "reference",
                   "sample": "%d concurrent serial TM_UPML instances of oracle/_ref/libref.so (one per "
                             "host core), zigzagModel %dx%d, %d timed steps, NTFF included"
                             % (res["cores"], args.cpu_n, args.cpu_n, args.cpu_steps),
                   "per_core_mcells": res["per_core"] / 1e6,
                   "stencil_only_value": (res_stencil["rate"] / 1e9) if res_stencil else None}
        else:
            cpu = {"value": None, "unit": "Gcell-updates/s", "cores": 0, "kind": "reference",
                   "sample": "unavailable: oracle/_ref/libref.so not shipped to this box"}

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # (NCCL prints its version banner on stdout at NCCL_DEBUG >= VERSION: main() has pointed fd 1
        # at stderr for the whole run, the JSON line goes out through the saved descriptor)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n_px, n_py = args.n, (args.n if args.strong else args.n * world)
    K, W = args.steps, args.warmup
    total_steps = W + K
    stream = torch.cuda.Stream()
    comm = None
    with torch.cuda.stream(stream):
        if world > 1:
            comm = TorchHaloComm(n_px, torch.device("cuda", local_rank))
        run = SlabRun(args.model, args.solver, n_px, n_py, total_steps, rank=rank, world=world,
                      device=local_rank, comm=comm, precision=args.precision)
        run.engine.set_stream(stream.cuda_stream)
        if args.lean:
            run.engine.set_option(B.OPT_LEAN_INTERIOR, 1)
        if comm is not None:
            run.attach_halo_buffers(*comm.pointers())
            if args.halo == "peer":
                run.enable_peer_halos(comm.gather_blobs)

        def barrier():
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()

        # 0 full kernels, 1 unit-coefficient interior + frame, 2 lean interior + frame, 3 one-pass step
        form = run.engine.step_form()
        two_kernel_form = form
        if form == 3 and world > 1 and args.halo != "peer":
            form = -1       # NCCL send/recv halos run between the two phase kernels: no one-pass step there
        if form in (3, -1):       # phase_h / phase_e (timed separately below, for reference) use this form:
            run.engine.set_option(B.OPT_FUSED, 0)
            two_kernel_form = run.engine.step_form()
            run.engine.set_option(B.OPT_FUSED, 2)
            if form == -1:
                form = two_kernel_form

        # ---- value: device-resident K steps + deferred projection ----------------
        for _ in range(W):
            run.step()
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        launches0 = run.engine.launches()
        run.engine.timer_start()
        for _ in range(K):
            run.step()
        run.project()
        ms = run.engine.timer_stop()
        launches = run.engine.launches() - launches0
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())
        cells = float(n_px) * float(n_py)
        value = cells * K / (ms_max * 1e-3) / 1e9

        # ---- per-kernel timing for the roofline (rank-local, same state) ----------
        reps = max(3, min(K, 20))
        run.engine.phase_h(run.args); run.engine.sync()
        run.engine.timer_start()
        for _ in range(reps):
            run.engine.phase_h(run.args)
        ms_h = run.engine.timer_stop() / reps
        run.engine.phase_e(run.args); run.engine.sync()
        run.engine.timer_start()
        for _ in range(reps):
            run.engine.phase_e(run.args)
        ms_e = run.engine.timer_stop() / reps
        ms_fused = None
        if form == 3:        # the step is ONE pass: edge pre-pass + TMA-staged marching kernel
            run.engine.phase_fused(run.args); run.engine.sync()
            run.engine.timer_start()
            for _ in range(reps):
                run.engine.phase_fused(run.args)
            ms_fused = run.engine.timer_stop() / reps
        barrier()

        # ---- the opt-in lean-interior form, same state, same K steps (reported beside `value`)
        lean = None
        if not args.lean and not args.no_lean_leg:
            run.engine.zero()
            run.L.field_reset()
            run.engine.set_option(B.OPT_LEAN_INTERIOR, 1)
            barrier()
            for _ in range(W):
                run.step()
            barrier()
            run.engine.timer_start()
            for _ in range(K):
                run.step()
            run.project()
            t = torch.tensor([run.engine.timer_stop()], dtype=torch.float64, device="cuda")
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_lean = float(t.item())
            run.engine.phase_h(run.args); run.engine.sync()
            run.engine.timer_start()
            for _ in range(reps):
                run.engine.phase_h(run.args)
            ms_h_lean = run.engine.timer_stop() / reps
            run.engine.phase_e(run.args); run.engine.sync()
            run.engine.timer_start()
            for _ in range(reps):
                run.engine.phase_e(run.args)
            ms_e_lean = run.engine.timer_stop() / reps
            run.engine.sync()
            run.engine.set_option(B.OPT_LEAN_INTERIOR, 0)
            barrier()
            lean = (ms_lean, ms_h_lean, ms_e_lean)

        # ---- e2e: host buffers inside the timed region ------------------------------
        run.engine.zero()
        run.L.field_reset()
        barrier()            # every rank has zeroed (peer-halo flags included) before anyone steps
        eps_pinned = torch.from_numpy(run.eps_host[0]).pin_memory()
        ez_pinned = torch.empty((n_px, run.nj, 2), dtype=torch.float64).pin_memory()
        for _ in range(min(W, 3)):
            run.step()
        barrier()
        t0 = time.perf_counter()
        B.check(run.L.b200fdtd_set_eps_slab(run.engine.h, 0, eps_pinned.data_ptr()), "set_eps_slab")
        for _ in range(K):
            run.step()
        B.check(run.L.b200fdtd_get_field_slab(run.engine.h, 0, ez_pinned.data_ptr()), "get_field_slab")
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_value = cells * K / float(t.item()) / 1e9
        h2d = run.eps_host[0].nbytes / K
        d2h = n_px * run.nj * 16 / K

    if rank == 0:
        peak, peak_kind = measured_hbm_peak()
        cells_rank = float(n_px) * float(run.nj)
        tm = args.solver == "TM_UPML_2D"
        # TE: H phase reads Ex,Ey,Mz,Bz (64) + writes Mz,Bz (32); E phase reads Bz,Jx,Dx,Jy,Dy (80) +
        # eps x2 (16) + writes Jx,Dx,Jy,Dy,Ex,Ey (96); step = SURVEY 8(d)'s 288 B
        bytes_h, bytes_e, bytes_step = (BYTES_H_TM, BYTES_E_TM, BYTES_STEP_TM) if tm else (96, 192, 288)
        # lean interior: H reads Ez,Bx,By + writes Bx,By (TE: Ex,Ey,Bz + Bz); E reads Bx,By,Dz,eps +
        # writes Dz,Ez (TE: Bz,Dx,Dy,2 eps + Dx,Dy,Ex,Ey); the 10-cell frame adds < 0.3 % at 16384^2
        lean_h, lean_e = (80, 88) if tm else (64, 128)
        if args.lean:
            bytes_h, bytes_e = lean_h, lean_e
        if args.precision == "f32":     # complex64 fields, f32 eps: every array element is half as wide
            bytes_h, bytes_e, bytes_step = bytes_h // 2, bytes_e // 2, bytes_step // 2
            lean_h, lean_e = lean_h // 2, lean_e // 2
        kname = "tm" if tm else "te"
        ach_h = bytes_h * cells_rank / (ms_h * 1e-3) / 1e9
        ach_e = bytes_e * cells_rank / (ms_e * 1e-3) / 1e9
        step_gbs = bytes_step * (value / world) * 1e9 / 1e9
        # ncu prints the kernel as <double, STORE_H, LEAN>: "<0,0>" default form, "<0,1>" lean
        # kernel names as ncu prints them, minus "double, " and blanks: the full kernels are
        # <STORE_H, RECTS> ("<0>" in captures older than the RECTS parameter), the lean ones <STORE_H>
        traffic, traffic_src = None, None
        if args.precision == "f64":
            tags = {0: [kname + "_upml_h_kernel<0,0>", kname + "_upml_h_kernel<0>"],
                    1: [kname + "_unit_h_kernel<0>"], 2: [kname + "_lean_h_kernel<0>"],
                    3: [kname + "_upml_fused_tma_kernel<0"]}[form]
            for tag in tags:
                traffic, traffic_src = ncu_traffic(tag, cells_rank)
                if traffic is not None:
                    break
        stem = kname + {0: "_upml", 1: "_unit", 2: "_lean"}[two_kernel_form]
        h_name, e_name = stem + "_h_kernel<STORE_H=false>", stem + "_e_kernel<FROM_B=true>"
        line = {
            "metric": "Gcell-updates/s", "value": value, "unit": "Gcell-updates/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_max / K,
            "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": workload_config(world, n=args.n, solver=args.solver, model=args.model, strong=args.strong,
                                      halo={"peer": "direct NVLink peer stores + device flags",
                                            "nccl": "NCCL send/recv"}[args.halo]),
            "roofline": {"bound": "hbm", "kernel": h_name, "achieved": ach_h,
                         "peak": peak, "unit": "GB/s", "frac": ach_h / peak, "peak_source": peak_kind,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": bytes_h * cells_rank, "ms_per_launch": ms_h,
                         "algorithmic_bytes_per_cell": bytes_h,
                         "e_phase": {"kernel": e_name, "achieved": ach_e,
                                     "frac": ach_e / peak, "ms_per_launch": ms_e,
                                     "algorithmic_bytes_per_cell": bytes_e},
                         "step": {"algorithmic_bytes_per_cell_update": bytes_step,
                                  "achieved": step_gbs, "frac": step_gbs / peak,
                                  "frac_of_nominal_8TBs": step_gbs / 8000.0}},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "Gcell-updates/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "path": "b200fdtd_set_eps_slab(pinned host eps) + K x [mpifdtd_upml_step_args + "
                            "b200fdtd_step + field_nextStep] + b200fdtd_get_field_slab(Ez -> pinned host)"},
            "gpu_launches": int(launches),
            "step_form": {3: "one pass: edge pre-pass + TMA-staged marching kernel (H and E fused, 232 B per "
                             "cell-update, bit-identical to the two-kernel forms; B200FDTD_OPT_FUSED, default on "
                             "large single-slab TM grids)",
                          0: "one full kernel per phase",
                          1: "unit-coefficient interior kernel + frame kernel per phase (bit-identical to the "
                             "one-kernel form; B200FDTD_OPT_UNIT_SPLIT, default on large grids)",
                          2: "lean interior kernel + frame kernel per phase (tolerance form)"}[form],
            "clocks": clocks,
            "device_bytes": run.engine.device_bytes(),
        }
        if form == 3:
            bytes_fused = 232 if args.precision == "f64" else 116
            ach_f = bytes_fused * cells_rank / (ms_fused * 1e-3) / 1e9
            two = dict(line["roofline"])
            line["roofline"] = {
                "bound": "hbm", "kernel": kname + "_upml_fused_tma_kernel<STORE_H=false, 8 warps, 4 stages> (+ its "
                                          "edge pre-pass, timed together)",
                "achieved": ach_f, "peak": peak, "unit": "GB/s", "frac": ach_f / peak, "peak_source": peak_kind,
                "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": bytes_fused * cells_rank, "ms_per_launch": ms_fused,
                "algorithmic_bytes_per_cell": bytes_fused,
                "note": "one pass reads Ez,Mx,Bx,My,By,Jz,Dz + eps (120 B) and writes Mx,Bx,My,By,Jz,Dz,Ez (112 B): "
                        "232 B per cell-update against SURVEY 8(d)'s 264 B contract figure for two passes",
                "step": {"algorithmic_bytes_per_cell_update": bytes_step,
                         "achieved": step_gbs, "frac": step_gbs / peak, "frac_of_nominal_8TBs": step_gbs / 8000.0,
                         "moved_bytes_per_cell_update": bytes_fused,
                         "moved": bytes_fused * (value / world), "moved_frac": bytes_fused * (value / world) / peak},
                "two_kernel_form": {"h_phase": {k: two[k] for k in ("kernel", "achieved", "frac", "ms_per_launch",
                                                                    "algorithmic_bytes_per_cell")},
                                    "e_phase": two["e_phase"]}}
        if args.lean:
            line["config"]["form"] = ("lean interior (B200FDTD_OPT_LEAN_INTERIOR): tolerance form, fields within "
                                      "1e-12 of the reference; moves %d B per cell-update" % (bytes_h + bytes_e))
            line["roofline"]["step"]["moved_bytes_per_cell_update"] = bytes_h + bytes_e
            line["roofline"]["step"]["moved"] = (bytes_h + bytes_e) * (value / world)
            line["roofline"]["step"]["moved_frac"] = (bytes_h + bytes_e) * (value / world) / peak
        if lean is not None:
            ms_lean, ms_h_lean, ms_e_lean = lean
            v_lean = cells * K / (ms_lean * 1e-3) / 1e9
            moved = (lean_h + lean_e) * (v_lean / world)
            line["lean_interior"] = {
                "value": v_lean, "unit": "Gcell-updates/s", "ms_per_step": ms_lean / K,
                "speedup_vs_value": v_lean / value,
                "note": "opt-in B200FDTD_OPT_LEAN_INTERIOR: cells outside the absorbing frame skip the M / J "
                        "recurrences (all coefficients exactly 1 there); tolerance form, fields within 1e-12 of "
                        "the reference (tests/test_gpu_lean.py); `value` above is the bit-exact default",
                "moved_bytes_per_cell_update": lean_h + lean_e,
                "moved": moved, "moved_frac": moved / peak,
                "h_phase": {"ms_per_launch": ms_h_lean, "bytes_per_cell": lean_h,
                            "achieved": lean_h * cells_rank / (ms_h_lean * 1e-3) / 1e9,
                            "frac": lean_h * cells_rank / (ms_h_lean * 1e-3) / 1e9 / peak},
                "e_phase": {"ms_per_launch": ms_e_lean, "bytes_per_cell": lean_e,
                            "achieved": lean_e * cells_rank / (ms_e_lean * 1e-3) / 1e9,
                            "frac": lean_e * cells_rank / (ms_e_lean * 1e-3) / 1e9 / peak}}
        emit(json.dumps(line))
    run.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


_json_out = None


def emit(text):
    """The one JSON line, on the process's ORIGINAL stdout."""
    if _json_out is None:
        print(text, flush=True)
    else:
        os.write(_json_out, (text + "\n").encode())


def quiet_stdout():
    """stdout carries the JSON line only: whatever libraries print to fd 1 during the run (the NCCL
    version banner, ...) goes to stderr instead."""
    global _json_out
    sys.stdout.flush()
    _json_out = os.dup(1)
    os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--size", dest="n", type=int, default=N_PER_GPU,
                    help="cells per side per GPU (use --size under torchrun, whose own parser trips over --n)")
    ap.add_argument("--cpu-n", type=int, default=1024, help="grid side of the CPU sample")
    ap.add_argument("--cpu-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU halo transport: direct NVLink peer stores (default) or NCCL send/recv")
    ap.add_argument("--model", default="ZIGZAG", choices=sorted(MODEL_NAMES),
                    help="material model of the synthetic structure (ZIGZAG = BASELINE configs[4])")
    ap.add_argument("--strong", action="store_true",
                    help="fixed global grid n x n split over the GPUs (BASELINE configs[2]: --model LAYER "
                         "--n 8192 --strong at 2/4 GPUs) instead of n x n per GPU")
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"],
                    help="f64 is the reference's arithmetic and the BASELINE metric; f32 is the optional "
                         "single-precision path (own tolerance), reported for information only")
    ap.add_argument("--lean", action="store_true",
                    help="measure the whole line in the opt-in lean-interior form (tolerance form, not the "
                         "bit-exact default)")
    ap.add_argument("--no-lean-leg", action="store_true", help="skip the extra lean_interior measurement")
    ap.add_argument("--solver", default="TM_UPML_2D", choices=["TM_UPML_2D", "TE_UPML_2D"],
                    help="TM_UPML_2D is the BASELINE workload; TE_UPML_2D is reported for information")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return reference_arm(args)
    quiet_stdout()
    return gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
