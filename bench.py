#!/usr/bin/env python
"""bench.py -- throughput of the 2-D FDTD time-stepping path on B200.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W`
prints ONE JSON line on rank 0.  For N > 1 it is launched under torchrun, one
rank per GPU.

Workload (BASELINE.json configs[4], the one the 1/2/4/8-GPU metric is quoted on):
zigzagModel, TM UPML (solver id 2), 16384 x 16384 cells PER GPU, stacked along y
(global grid 16384 x 16384*N), h_u = 10 nm, pml = 10, lambda = 500 nm, angle 0.
A "step" is one update(): H phase + E phase + source (ONE pass of the one-pass kernel,
fused_kernels.cu) and the NTFF surface sample; the deferred NTFF projection of the K
timed steps is inside the timed region too.  9 complex fields x 4 GiB per GPU -- far
larger than the 126 MB L2, so no flush is needed between iterations.

  value      whole-job Gcell-updates/s, state resident in HBM, CUDA-event timed on
             the engine's stream, max over ranks.
  roofline   the one-pass kernel (+ its edge pre-pass, timed together, live, CUDA
             events): algorithmic bytes per launch (192 B per cell-update in vacuum row-strips -- rows of a tile
             whose cells all have eps == 1, where E holds the bits of D and does not move; the engine reports
             how many cells those are -- and elsewhere TM 232 B per cell-update: reads
             Ez,Mx,Bx,My,By,Jz,Dz + eps, writes Mx,Bx,My,By,Jz,Dz,Ez; TE 272) over its
             mean duration, against MEASURED_PEAKS.json hbm_gbs.  `step` carries SURVEY
             8(d)'s contract figure (264 / 288 B) for comparison; `two_kernel_form` the two
             phase kernels the step would otherwise launch.
  e2e        the same metric with HOST buffers inside the timed region -- what simulator_init /
             K x simulator_calc / simulator_finish move for a run: the permittivity map H2D from
             pinned memory (as the plugin ships it: 16-bit indices into the table of its distinct
             values), K x update(), NTFF projection (+ reduce over ranks), spectrum, the 321 x 360
             far-field table D2H -- wall clock, max over ranks.  `with_field_snapshot` adds the
             getter's D2H of the whole Ez plane (4 GiB per rank: round 1's definition of this key).
             `e2e_plugin` (N = 1) is the reference-facing call sequence itself -- simulator_init /
             K x simulator_calc / fdtdTM_upml_getEz / simulator_finish -- at a size whose host-side
             permittivity build fits the bench budget, with `init_s` reported beside it.
  lean_interior
             the same K steps with B200FDTD_OPT_LEAN_INTERIOR (opt-in tolerance form: cells
             outside the absorbing frame advance B / D directly; one pass, 136 B per TM
             cell-update; fields within 1e-12 of the reference instead of bit-identical).
             Reported BESIDE `value`, which stays the reference's arithmetic in every cell.
  dense      `value` again on a material-dense structure (--fill, default 0.35 of the
             cells with eps != 1, every one of them firing the pulse's exp / sincos): the
             regime the vacuum-dominated zigzag workload does not show.
  parity_check
             N > 1: before anything is timed, a 512 x (512 N) random-state, random-eps case
             runs on the SAME ranks, peer halos and step form, and every one of the nine
             arrays is compared bit for bit (position-mixed 64-bit digests summed over the
             slabs) with a single-slab run on rank 0; U/W after the reduce within 1e-12.
  cpu_baseline / --impl reference
             the UNMODIFIED reference (oracle/_ref/libref.so, built from
             /root/reference) on the host cores, one serial solver instance per
             core (how main.c uses its MPI ranks), on a bounded sample whose size is
             stated in config.workload.
"""
import argparse
import ctypes
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
FALLBACK_HBM_GBS = 6650.0    # B200_PROFILING.md fallback
# algorithmic bytes per cell-update: one-pass forms, two-kernel phases, SURVEY 8(d)
BYTES = {"TM_UPML_2D": {"one_pass": 232, "one_pass_lean": 136, "h": 144, "e": 120, "contract": 264,
                        "lean_h": 80, "lean_e": 88},
         "TE_UPML_2D": {"one_pass": 272, "one_pass_lean": 176, "h": 96, "e": 192, "contract": 288,
                        "lean_h": 64, "lean_e": 128}}


def ncu_traffic(kernel_substr, cells):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    of this same command (profiles/*_full.json, written by scripts/ncu_summary.py); None when no
    capture at this launch size is on file."""
    import glob
    for path in sorted(glob.glob(os.path.join(HERE, "profiles", "*_full.json")), reverse=True):
        try:
            for rec in json.load(open(path))["launches"]:
                name = rec["kernel"].replace("double, ", "").replace(" ", "").replace("(bool)", "").replace("(int)", "")
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


def cpu_sample_text(res, n, steps, warm):
    return ("%d concurrent serial TM_UPML instances of the unmodified reference (oracle/_ref/libref.so, one per "
            "host core, as main.c uses its MPI ranks), each zigzagModel %dx%d, %d timed steps after %d warm-up, "
            "NTFF included" % (res["cores"], n, n, steps, warm))


def reference_arm(args):
    """bench.py --impl reference: the reference's own CPU path, same metric; the workload string
    states the bounded sample that was actually run."""
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
    sample = cpu_sample_text(res, n, args.steps, args.warmup)
    cfg = workload_config(args.gpus, n=args.n)
    cfg["workload"] = ("CPU SAMPLE of that workload: %d x zigzagModel TM_UPML_2D %d x %d (one reference instance per "
                       "host core); the GPU arm runs %s" % (res["cores"], n, n, cfg["workload"]))
    cfg["cpu_sample_cells_per_instance"] = n * n
    cfg["note"] = ("bounded sample (BASELINE.md 4.5): the reference's layout needs 78 GiB per 16384^2 instance; its "
                   "per-core rate falls with grid size (survey: -35 % from 1024^2 to 4096^2), so the sample flatters "
                   "the CPU")
    line = {
        "impl": "reference", "metric": "Gcell-updates/s", "value": value, "unit": "Gcell-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * res["max_loop_s"] / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
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


def workload_config(n_gpus, n=N_PER_GPU, solver="TM_UPML_2D", halo="peer", model="ZIGZAG", strong=False):
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
    return cfg


# ----------------------------------------------------------------------------
# host placement: this rank's threads and pinned buffers next to its GPU
# ----------------------------------------------------------------------------
def bind_near_gpu(local_rank):
    """CPU affinity + memory policy (MPOL_PREFERRED) of this process on the NUMA node its GPU hangs
    off, so the pinned eps / Ez buffers of the e2e leg are node-local: at 8 ranks the host copies
    otherwise cross the socket interconnect (round 1: 50 -> 13 GB/s per GPU).  Best effort."""
    info = {"numa_node": None, "bound": False}
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip()
        bdf = out.lower()
        if bdf.startswith("0000"):
            bdf = bdf[4:]                  # nvidia-smi prints an 8-digit domain, sysfs a 4-digit one
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        info["numa_node"] = node
        if node < 0:
            return info
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        mask = ctypes.c_ulong(1 << node)
        libc = ctypes.CDLL(None, use_errno=True)
        # set_mempolicy(MPOL_PREFERRED = 1, &mask, maxnode): x86-64 syscall 238
        rc = libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(8 * ctypes.sizeof(mask)))
        info["bound"] = bool(allowed) and rc == 0
        info["cpus"] = len(allowed)
    except Exception as err:            # no sysfs entry, container without the syscall, ...
        info["error"] = str(err)[:80]
    return info


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
# multi-GPU parity check on the communicator the benchmark uses
# ----------------------------------------------------------------------------
def parity_check(B, SlabRun, solver, world, rank, local_rank, comm, stream, dist, torch, steps=48, n=512):
    """A 512 x (512 * world) case from a random state with a random permittivity map (one cell in
    three a material cell in the middle third of the rows, vacuum elsewhere) and the pulse on, stepped on all ranks with peer halos in the one-pass
    form the benchmark runs; then the same global case as ONE slab on rank 0.  Every one of the
    nine arrays must agree bit for bit: the position-mixed digests of the slabs add up (mod 2^64) to
    the single-slab digest.  U/W after the NCCL reduce against the single slab's: <= 1e-12."""
    import numpy as np
    kind = 2 if solver == "TM_UPML_2D" else 3
    npx, npy = n, n * world
    rng = np.random.default_rng(20261017)
    n_eps = 1 if kind == 2 else 2
    eps = [np.where(rng.random((npx, npy), dtype=np.float32) < 0.67, 1.0, 1.5 + rng.random((npx, npy), dtype=np.float32))
           for _ in range(n_eps)]
    state = [rng.standard_normal((npx, npy), dtype=np.float32).astype(np.float64) +
             1j * rng.standard_normal((npx, npy), dtype=np.float32).astype(np.float64) for _ in range(9)]
    for e in eps:                       # material in the middle third of the rows only: above and below it whole tile
        e[:npx // 3, :] = 1.0           # rows are vacuum, where the one-pass step keeps no E arrays (vacuum row-strips)
        e[2 * npx // 3:, :] = 1.0
    for h, b in (((3, 5), (6, 8)) if kind == 2 else ((6, 8),)):      # H == B/mu0, the solver's invariant
        state[h] = (state[b].real / B.MU_0_S) + 1j * (state[b].imag / B.MU_0_S)
    for arr in state:                   # a slab's ghost columns start at zero: so do the columns they mirror
        for k in range(1, world):
            arr[:, k * n - 2:k * n + 2] = 0

    def run_case(r, w, communicator):
        run = SlabRun("NO_MODEL", solver, npx, npy, steps, rank=r, world=w, device=local_rank, comm=communicator,
                      angle_deg=30, n_bins="full")
        run.engine.set_stream(stream.cuda_stream)
        run.engine.set_option(B.OPT_FUSED, 1)                 # the benchmark's form, forced at this size
        for slot, e in enumerate(eps):
            run.engine.set_eps(slot, e)
        for slot in range(9):
            run.engine.set_field(slot, state[slot])
        if communicator is not None and w > 1:
            run.enable_peer_halos(communicator.gather_blobs)
        form = run.engine.step_form()
        for _ in range(steps):
            run.step()
        run.engine.sync()
        digests = [run.engine.digest(s) for s in range(9)]
        run.project()
        uw = None
        ptr, count = run.engine.uw_device()
        if communicator is not None and w > 1:
            communicator.reduce_sum_to_root(ptr, count)
        if r == 0:
            uw = np.stack([run.engine.uw(s) for s in range(3)])
        torch.cuda.synchronize()
        if communicator is not None and w > 1:
            dist.barrier()                 # nobody frees a slab a neighbour may still store into
        run.close()
        return digests, uw, form

    mine, uw_multi, form = run_case(rank, world, comm)
    t = torch.tensor([d - (1 << 64) if d >= (1 << 63) else d for d in mine], dtype=torch.int64, device="cuda")
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    out = None
    if rank == 0:
        total = [sum(int(g[s].item()) for g in gathered) & ((1 << 64) - 1) for s in range(9)]
        want, uw_single, _ = run_case(0, 1, None)
        bad = [s for s in range(9) if total[s] != want[s]]
        scale = float(np.abs(uw_single).max())
        err = float(np.abs(uw_multi - uw_single).max() / scale) if scale > 0 else float("inf")
        ok = not bad and err <= 1e-12
        out = {"result": "bit-identical" if ok else "MISMATCH",
               "grid": "%d x %d, %d y-slabs of %d columns" % (npx, npy, world, n), "steps": steps,
               "arrays_compared": 9, "mismatching_arrays": bad, "step_form": form,
               "ntff_uw_rel_err_after_reduce": err,
               "how": "random state + random eps (1/3 material cells in the middle third of the rows, vacuum "
                      "row-strips elsewhere) + pulse; digests of the slabs summed mod 2^64 vs a single-slab run of the "
                      "same global case on rank 0"}
    dist.barrier()
    return out


def dense_eps(np, n_px, nj, fill, seed):
    """Material-dense permittivity map: square blocks of 8 x 8 cells, a fraction `fill` of them
    dielectric (eps 2.56, the reference's n = 1.6) -- every such cell divides by eps and fires the
    pulse's exp / sincos each step."""
    rng = np.random.default_rng(seed)
    blocks = rng.random(((n_px + 7) // 8, (nj + 7) // 8), dtype=np.float32) < fill
    eps = np.where(np.repeat(np.repeat(blocks, 8, axis=0), 8, axis=1)[:n_px, :nj], 2.56, 1.0)
    eps[:12, :] = 1.0; eps[-12:, :] = 1.0                      # keep the absorbing frame in vacuum
    return np.ascontiguousarray(eps)


# ----------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------
def gpu_arm(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    placement = bind_near_gpu(local_rank)          # before anything allocates host memory

    import numpy as np
    import torch
    from mpifdtd_b200 import binding as B
    from mpifdtd_b200.slab import SlabRun, TorchHaloComm

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
                   "kind": "reference", "sample": cpu_sample_text(res, args.cpu_n, args.cpu_steps, 2),
                   "per_core_mcells": res["per_core"] / 1e6,
                   "stencil_only_value": (res_stencil["rate"] / 1e9) if res_stencil else None,
                   "note": "the GPU's NTFF cost in a run this short is the surface sample only (the deferred "
                           "projection finds few taps inside K steps), so `stencil_only_value` is the like-for-like "
                           "CPU figure for `value`"}
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
    tm = args.solver == "TM_UPML_2D"
    by = BYTES[args.solver]
    stream = torch.cuda.Stream()
    comm = None
    with torch.cuda.stream(stream):
        if world > 1:
            comm = TorchHaloComm(n_px, torch.device("cuda", local_rank))

        def barrier():
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()

        def max_over_ranks(x):
            t = torch.tensor([x], dtype=torch.float64, device="cuda")
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        # ---- correctness of the multi-GPU path, on this communicator, before anything is timed
        parity = None
        if world > 1 and args.halo == "peer" and args.precision == "f64" and not args.no_parity_check:
            parity = parity_check(B, SlabRun, args.solver, world, rank, local_rank, comm, stream, dist, torch)

        run = SlabRun(args.model, args.solver, n_px, n_py, total_steps, rank=rank, world=world,
                      device=local_rank, comm=comm, precision=args.precision)
        run.engine.set_stream(stream.cuda_stream)
        if args.lean:
            run.engine.set_option(B.OPT_LEAN_INTERIOR, 1)
        if comm is not None:
            run.attach_halo_buffers(*comm.pointers())
            if args.halo == "peer":
                run.enable_peer_halos(comm.gather_blobs)

        # 0 full kernels, 1 unit-coefficient interior + frame, 2 lean interior + frame, 3 one pass, 4 one pass lean
        form = run.engine.step_form()
        one_pass = form in (3, 4) and not (world > 1 and args.halo != "peer")
        cells = float(n_px) * float(n_py)
        cells_rank = float(n_px) * float(run.nj)

        def timed_steps():
            for _ in range(W):
                run.step()
            barrier()
            launches0 = run.engine.launches()
            run.engine.timer_start()
            for _ in range(K):
                run.step()
            run.project()
            ms = run.engine.timer_stop()
            launches = run.engine.launches() - launches0
            barrier()
            return max_over_ranks(ms), launches

        def restart():
            run.engine.zero()
            run.L.field_reset()
            barrier()

        # ---- value: device-resident K steps + deferred projection ----------------
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        ms_max, launches = timed_steps()
        clocks = sampler.stop() if rank == 0 else None
        value = cells * K / (ms_max * 1e-3) / 1e9
        # cells of this rank in vacuum row-strips, where the one-pass step keeps no E arrays (B200FDTD_OPT_DERIVED_E)
        vac_cells = float(run.engine.vacuum_cells()) if form in (3, 4) else 0.0

        # ---- per-kernel timing for the roofline (rank-local, same state, CUDA events) -------
        def time_phase(fn, reps):
            fn(run.args); run.engine.sync()
            run.engine.timer_start()
            for _ in range(reps):
                fn(run.args)
            return run.engine.timer_stop() / reps

        reps = max(3, min(K, 20))
        ms_one = time_phase(run.engine.phase_fused, reps) if form in (3, 4) else None
        if form in (3, 4):             # phase_h / phase_e time the two-kernel form the step would otherwise take
            run.engine.set_option(B.OPT_FUSED, 0)
        two_form = run.engine.step_form()
        ms_h = time_phase(run.engine.phase_h, reps)
        ms_e = time_phase(run.engine.phase_e, reps)
        if form in (3, 4):
            run.engine.set_option(B.OPT_FUSED, 2)
        barrier()

        # ---- the opt-in lean-interior form, same K steps (reported beside `value`) --------------
        lean = None
        if not args.lean and not args.no_lean_leg:
            restart()
            run.engine.set_option(B.OPT_LEAN_INTERIOR, 1)
            barrier()
            lean_form = run.engine.step_form()
            ms_lean, _ = timed_steps()
            ms_lean_kernel = time_phase(run.engine.phase_fused, reps) if lean_form == 4 else None
            run.engine.sync()
            run.engine.set_option(B.OPT_LEAN_INTERIOR, 0)
            barrier()
            lean = (ms_lean, ms_lean_kernel, lean_form)

        # ---- material-dense structure: same engine, another permittivity map --------------------
        dense = None
        if args.fill > 0 and tm and args.precision == "f64":
            restart()
            d_eps = dense_eps(np, n_px, run.nj, args.fill, 7 + rank)
            frac = float((d_eps != 1.0).mean())
            B.check(run.L.b200fdtd_set_eps_slab(run.engine.h, 0, d_eps.ctypes.data), "set_eps_slab")
            del d_eps
            ms_dense, _ = timed_steps()
            ms_dense_kernel = time_phase(run.engine.phase_fused, reps) if form in (3, 4) else None
            vac_dense = float(run.engine.vacuum_cells()) if form in (3, 4) else 0.0
            B.check(run.L.b200fdtd_set_eps_slab(run.engine.h, 0, run.eps_host[0].ctypes.data), "set_eps_slab")
            dense = (ms_dense, ms_dense_kernel, frac, vac_dense)

        # ---- e2e: host buffers inside the timed region ------------------------------------------
        # What simulator_init / K x simulator_calc / simulator_finish move between host and device for a
        # run: the permittivity map in (as the plugin ships it: 16-bit indices into the table of its
        # distinct values, b200fdtd_set_eps_palette), the step arguments each step, the far-field table
        # out.  The second figure adds what a caller who ALSO wants a field snapshot pays (the getter's
        # D2H of the whole Ez plane: round 1's definition of this key).
        restart()
        pinned = []                        # pinned by the library (cudaHostAlloc), node-local after bind_near_gpu()

        def host_alloc(nbytes):
            p = ctypes.c_void_p()
            B.check(run.L.b200fdtd_host_alloc(ctypes.byref(p), nbytes), "host_alloc")
            pinned.append(p)
            return p

        n_eps = len(run.eps_host)
        palettes = []
        for e in run.eps_host:
            idx = host_alloc(e.size * 2)
            tab = host_alloc(65536 * 8)
            count = run.L.mpifdtd_eps_palette(e.ctypes.data, e.size, idx, tab)
            if count < 1:                  # a map too rich for a palette goes up as doubles
                idx = host_alloc(e.nbytes)
                ctypes.memmove(idx, e.ctypes.data, e.nbytes)
            palettes.append((idx, tab, count))
        field_bytes = n_px * run.nj * 16
        ez_pinned = host_alloc(field_bytes)
        far_table = np.zeros((321, 360))
        for _ in range(min(W, 3)):
            run.step()
        barrier()
        t0 = time.perf_counter()
        for slot, (idx, tab, count) in enumerate(palettes):
            if count > 0:
                B.check(run.L.b200fdtd_set_eps_palette(run.engine.h, slot, idx, run.nj, tab, count), "set_eps_palette")
            else:
                B.check(run.L.b200fdtd_set_eps_slab(run.engine.h, slot, idx), "set_eps_slab")
        t1 = time.perf_counter()
        for _ in range(K):
            run.step()
        far = run.far_field()              # projection, reduce over ranks (N > 1), spectrum, table D2H on rank 0
        run.engine.sync()
        t2 = time.perf_counter()
        dt_run = t2 - t0
        B.check(run.L.b200fdtd_get_field_slab(run.engine.h, 0, ez_pinned), "get_field_slab")
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        e2e_value = cells * K / max_over_ranks(dt_run) / 1e9
        e2e_snapshot_value = cells * K / max_over_ranks(t3 - t0) / 1e9
        h2d_total = sum((e.size * 2 + c * 8) if c > 0 else e.nbytes for e, (_, _, c) in zip(run.eps_host, palettes))
        h2d = h2d_total / K
        d2h = 321 * 360 * 8 / K
        e2e_parts = {"h2d_s": max_over_ranks(t1 - t0), "steps_and_far_field_s": max_over_ranks(t2 - t1),
                     "h2d_gbs_per_rank": h2d_total / max_over_ranks(t1 - t0) / 1e9,
                     "field_snapshot_d2h_s": max_over_ranks(t3 - t2),
                     "field_snapshot_d2h_gbs_per_rank": field_bytes / max_over_ranks(t3 - t2) / 1e9,
                     "palette_values": [c for _, _, c in palettes]}
        for p in pinned:
            run.L.b200fdtd_host_free(p)

    # ---- e2e through the plugin surface itself (N = 1) ----------------------------------------------
    device_bytes = run.engine.device_bytes()
    e2e_plugin = None
    if rank == 0 and world == 1 and not args.no_plugin_leg and args.precision == "f64" and not args.lean:
        if run is not None:
            run.close()
            run = None
        e2e_plugin = plugin_leg(B, args, K)

    # ---- the deferred NTFF projection over a history of realistic length (N = 1) -----------------------
    ntff_leg = None
    if rank == 0 and world == 1 and not args.no_ntff_leg:
        if run is not None:
            run.close()
            run = None
        ntff_leg = ntff_projection_leg(B, args, local_rank)

    if rank == 0:
        peak, peak_kind = measured_hbm_peak()
        kname = "tm" if tm else "te"
        scale = 1 if args.precision == "f64" else 0.5          # complex64 / f32 eps: every element half as wide
        step_gbs = by["contract"] * scale * (value / world)
        stems = {0: "_upml", 1: "_unit", 2: "_lean"}
        two = {"h_phase": {"kernel": kname + stems[two_form] + "_h_kernel<STORE_H=false>", "ms_per_launch": ms_h,
                           "algorithmic_bytes_per_cell": by["h"] * scale,
                           "achieved": by["h"] * scale * cells_rank / (ms_h * 1e-3) / 1e9},
               "e_phase": {"kernel": kname + stems[two_form] + "_e_kernel<FROM_B=true>", "ms_per_launch": ms_e,
                           "algorithmic_bytes_per_cell": by["e"] * scale,
                           "achieved": by["e"] * scale * cells_rank / (ms_e * 1e-3) / 1e9}}
        for ph in two.values():
            ph["frac"] = ph["achieved"] / peak
        # a vacuum row-strip moves no E (read + write) and no eps: TM 40 B, TE 80 B per cell less, exact and lean alike
        vac_saved = 40 if tm else 80
        if one_pass:
            b_one = by["one_pass_lean" if form == 4 else "one_pass"]
            launch_bytes = b_one * cells_rank - vac_saved * vac_cells
            ach = launch_bytes / (ms_one * 1e-3) / 1e9
            traffic, traffic_src = ncu_traffic(kname + "_onepass_kernel<%d,0" % (1 if form == 4 else 0), cells_rank)
            roof = {"bound": "hbm",
                    "kernel": "%s_onepass_kernel<LEAN=%s, STORE_H=false, 8 warps, 4 row buffers> (+ its edge pre-pass, "
                              "timed together)" % (kname, "true" if form == 4 else "false"),
                    "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": peak_kind,
                    "traffic": traffic, "traffic_source": traffic_src,
                    "algorithmic_bytes_per_launch": launch_bytes, "ms_per_launch": ms_one,
                    "algorithmic_bytes_per_cell": launch_bytes / cells_rank,
                    "algorithmic_bytes_per_cell_by_region": {"vacuum_row_strips": b_one - vac_saved, "elsewhere": b_one,
                                                             "vacuum_row_strip_cells": vac_cells, "cells": cells_rank},
                    "note": ("one pass reads Ez,Mx,Bx,My,By,Jz,Dz + eps (120 B) and writes Mx,Bx,My,By,Jz,Dz,Ez (112 B): "
                             "232 B per cell-update against SURVEY 8(d)'s 264 B contract figure for two passes; in "
                             "vacuum row-strips (every cell of a tile row has eps == 1, so Ez holds the bits of Dz) "
                             "neither Ez nor eps moves: 192 B"
                             if tm and form == 3 else "see DESIGN.md section 4 for the byte count of this form"),
                    "step": {"algorithmic_bytes_per_cell_update": by["contract"], "achieved": step_gbs,
                             "frac": step_gbs / peak, "frac_of_nominal_8TBs": step_gbs / 8000.0,
                             "moved_bytes_per_cell_update": launch_bytes / cells_rank,
                             "moved": launch_bytes / cells_rank * (value / world),
                             "moved_frac": launch_bytes / cells_rank * (value / world) / peak},
                    "two_kernel_form": two}
        else:
            bh = (by["lean_h"] if form == 2 else by["h"]) * scale
            be = (by["lean_e"] if form == 2 else by["e"]) * scale
            ach = bh * cells_rank / (ms_h * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": two["h_phase"]["kernel"], "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "peak_source": peak_kind, "traffic": None, "traffic_source": None,
                    "algorithmic_bytes_per_launch": bh * cells_rank, "ms_per_launch": ms_h,
                    "algorithmic_bytes_per_cell": bh,
                    "e_phase": dict(two["e_phase"], algorithmic_bytes_per_cell=be,
                                    achieved=be * cells_rank / (ms_e * 1e-3) / 1e9),
                    "step": {"algorithmic_bytes_per_cell_update": by["contract"] * scale, "achieved": step_gbs,
                             "frac": step_gbs / peak, "frac_of_nominal_8TBs": step_gbs / 8000.0,
                             "moved_bytes_per_cell_update": bh + be, "moved": (bh + be) * (value / world),
                             "moved_frac": (bh + be) * (value / world) / peak}}
        forms = {3: "one pass: edge pre-pass + TMA-staged marching kernel (H and E fused; bit-identical to the "
                    "two-kernel forms; B200FDTD_OPT_FUSED, default on grids of >= 2^22 cells)",
                 4: "one pass, lean interior (tolerance form)",
                 0: "one full kernel per phase",
                 1: "unit-coefficient interior kernel + frame kernel per phase (bit-identical to the one-kernel form)",
                 2: "lean interior kernel + frame kernel per phase (tolerance form)"}
        line = {
            "metric": "Gcell-updates/s", "value": value, "unit": "Gcell-updates/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_max / K,
            "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": workload_config(world, n=args.n, solver=args.solver, model=args.model, strong=args.strong,
                                      halo={"peer": "direct NVLink peer stores + device flags",
                                            "nccl": "NCCL send/recv"}[args.halo]),
            "roofline": roof,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "Gcell-updates/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "path": "what simulator_init / K x simulator_calc / simulator_finish move: b200fdtd_set_eps_palette("
                            "pinned host index map + table) + K x [mpifdtd_upml_step_args + b200fdtd_step + "
                            "field_nextStep] + b200fdtd_ntff_project%s + b200fdtd_ntff_spectrum (321 x 360 far-field "
                            "table -> host)" % (" + NCCL reduce of U/W" if world > 1 else ""),
                    "definition_note": "round 1 read back the whole %s plane (4 GiB per rank) instead of the far-field "
                                       "table and uploaded eps as doubles; that figure is `with_field_snapshot` "
                                       "(eps as palette)" % ("Ez" if tm else "Ex"),
                    "with_field_snapshot": {"value": e2e_snapshot_value, "d2h_bytes_per_step": d2h + field_bytes / K},
                    "host_placement": placement, "where_the_time_goes": e2e_parts},
            "gpu_launches": int(launches),
            "step_form": forms[form if one_pass or form < 3 else two_form],
            "clocks": clocks,
            "device_bytes": device_bytes,
        }
        if e2e_plugin is not None:
            line["e2e_plugin"] = e2e_plugin
        if ntff_leg is not None:
            ntff_leg["value_with_amortised_projection"] = (
                cells / ((ms_max / K + ntff_leg["ms_per_step_amortised"]) * 1e-3) / 1e9)
            line["ntff_projection"] = ntff_leg
        if parity is not None:
            line["parity_check"] = parity["result"]
            line["parity_detail"] = parity
        if args.lean:
            line["config"]["form"] = "lean interior (B200FDTD_OPT_LEAN_INTERIOR): tolerance form, fields within 1e-12"
        if lean is not None:
            ms_lean, ms_lean_kernel, lean_form = lean
            v_lean = cells * K / (ms_lean * 1e-3) / 1e9
            b_lean = by["one_pass_lean"] if lean_form == 4 else by["lean_h"] + by["lean_e"]
            if lean_form == 4:
                b_lean -= vac_saved * vac_cells / cells_rank
            line["lean_interior"] = {
                "value": v_lean, "unit": "Gcell-updates/s", "ms_per_step": ms_lean / K,
                "speedup_vs_value": v_lean / value, "step_form": forms[lean_form],
                "note": "opt-in B200FDTD_OPT_LEAN_INTERIOR: cells outside the absorbing frame skip the M / J "
                        "recurrences (all coefficients exactly 1 there); tolerance form, fields within 1e-12 of "
                        "the reference (tests/test_gpu_lean.py); `value` above is the bit-exact default",
                "moved_bytes_per_cell_update": b_lean, "moved": b_lean * (v_lean / world),
                "moved_frac": b_lean * (v_lean / world) / peak}
            if ms_lean_kernel:
                ach = b_lean * cells_rank / (ms_lean_kernel * 1e-3) / 1e9
                line["lean_interior"]["kernel"] = {"ms_per_launch": ms_lean_kernel, "achieved": ach, "frac": ach / peak}
        if dense is not None:
            ms_dense, ms_dense_kernel, frac, vac_dense = dense
            v_dense = cells * K / (ms_dense * 1e-3) / 1e9
            line["dense"] = {"value": v_dense, "unit": "Gcell-updates/s", "ms_per_step": ms_dense / K,
                             "material_cell_fraction": frac, "ratio_to_value": v_dense / value,
                             "structure": "8 x 8-cell dielectric blocks (eps 2.56) on %.0f %% of the grid; every "
                                          "material cell divides by eps and evaluates the pulse's exp / sincos each "
                                          "step (field.c:224-256)" % (100 * frac)}
            if ms_dense_kernel:
                b_dense = by["one_pass"] - vac_saved * vac_dense / cells_rank
                ach = b_dense * cells_rank / (ms_dense_kernel * 1e-3) / 1e9
                line["dense"]["roofline"] = {"ms_per_launch": ms_dense_kernel, "achieved": ach, "frac": ach / peak,
                                             "algorithmic_bytes_per_cell": b_dense, "vacuum_row_strip_cells": vac_dense}
        emit(json.dumps(line))
    if run is not None:
        run.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def ntff_projection_leg(B, args, device, history=2048):
    """The K timed steps of `value` record K surface samples, and the deferred projection over so
    short a history finds few taps -- the reference, by contrast, pays its 360-direction binning every
    step.  This leg prices the GPU's share honestly: the projection kernel over a FULL history of
    `history` steps on the benchmark's surface (65 k points at 16384^2), on an engine that holds only
    the NTFF machinery (B200FDTD_GRID_NTFF_ONLY), amortised per step."""
    import numpy as np
    L = B.lib()
    n = args.n
    kind = 2 if args.solver == "TM_UPML_2D" else 3
    L.models_setModel(B.MODELS["NO_MODEL"])
    L.field_init(B.FieldInfo(n * 10, n * 10, 10, 10, 500, 0, history))
    grid = B.Grid(kind, n, n, 10, 0, n, 1, n - 2, 1, n - 2, device, 0, B.MU_0_S, 1, 1)      # flags = NTFF only
    h = ctypes.c_void_p()
    B.check(L.b200fdtd_create(ctypes.byref(grid), ctypes.byref(h)), "create")
    try:
        box = L.field_getNTFFInfo()
        n_points = L.mpifdtd_ntff_point_count(ctypes.byref(box))
        ptr = L.mpifdtd_ntff_time_shift(ctypes.byref(box), 360, 0.0 if kind == 2 else 0.5, 0, n)
        plan = B.NtffPlan(box.top, box.bottom, box.left, box.right, n_points, n_points, history, history, 360,
                          box.arraySize, ptr)
        B.check(L.b200fdtd_set_ntff_plan(h, ctypes.byref(plan)), "set_ntff_plan")
        L.free(ptr)
        rng = np.random.default_rng(3)
        samples = rng.standard_normal(2 * n_points) + 0.0
        L.b200fdtd_ntff_push_samples.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]
        for t in (0, history // 2, history - 1):          # the last one marks the whole history as recorded
            B.check(L.b200fdtd_ntff_push_samples(h, t, samples.ctypes.data, samples.ctypes.data), "push_samples")
        B.check(L.b200fdtd_ntff_project(h), "project")      # warm-up
        B.check(L.b200fdtd_sync(h), "sync")
        ms = ctypes.c_float(0)
        B.check(L.b200fdtd_timer_start(h), "timer_start")
        B.check(L.b200fdtd_ntff_project(h), "project")
        B.check(L.b200fdtd_timer_stop(h, ctypes.byref(ms)), "timer_stop")
        return {"history_steps": history, "surface_points": n_points, "directions": 360,
                "project_ms": ms.value, "ms_per_step_amortised": ms.value / history,
                "note": "ntff_project_kernel over a full %d-step history of the benchmark's surface; the reference "
                        "scatters 360 x points x 6 taps EVERY step (ntffTM.c:279-371), ~85 %% of its step" % history}
    finally:
        L.b200fdtd_destroy(h)


def plugin_leg(B, args, K):
    """The reference-facing call sequence itself, timed with the host clock: simulator_init (host
    permittivity build + uploads + NTFF plan: `init_s`), K x simulator_calc (deferred CUDA-graph
    replay), the borrowed-pointer getter (D2H into the pinned mirror) and simulator_finish (NTFF
    projection, spectrum, the two far-field files).  Size: the largest square grid whose init stays
    within a few seconds of host work."""
    import numpy as np
    n = args.plugin_n
    cwd = os.getcwd()
    work = tempfile.mkdtemp(prefix="bench_plugin_")
    os.chdir(work)
    try:
        L = B.lib()
        t0 = time.perf_counter()
        gpu = B.Plugin(args.model, args.solver, n, n, steps=K + 3, h_u_nm=10)
        gpu.sync()
        init_s = time.perf_counter() - t0
        gpu.step(3)
        gpu.sync()
        t1 = time.perf_counter()
        gpu.step(K)
        getter = "fdtdTM_upml_getEz" if args.solver == "TM_UPML_2D" else "fdtdTE_upml_getEx"
        ptr = getattr(L, getter)()
        t2 = time.perf_counter()
        peak = float(np.abs(np.frombuffer((ctypes.c_double * 64).from_address(ptr), dtype=np.float64)).max())
        gpu.finish()
        t3 = time.perf_counter()
        return {"value": n * n * K / (t2 - t1) / 1e9, "unit": "Gcell-updates/s",
                "grid": "%d x %d" % (n, n), "steps": K,
                "init_s": init_s, "calc_plus_getter_s": t2 - t1, "finish_s": t3 - t2,
                "value_including_finish": n * n * K / (t3 - t1) / 1e9,
                "d2h_bytes_per_step": n * n * 16 / K,
                "path": "simulator_init | timed: K x simulator_calc + %s() (pinned mirror) | simulator_finish "
                        "(projection, spectrum, .txt + .dat)" % getter,
                "first_values_abs_max": peak}
    finally:
        os.chdir(cwd)


_json_out = None


def emit(text):
    """The one JSON line, on the process's ORIGINAL stdout."""
    if _json_out is None:
        print(text, flush=True)
    else:
        os.write(_json_out, (text + "\n").encode())


def quiet_stdout():
    """stdout carries the JSON line only: whatever libraries print to fd 1 during the run (the NCCL
    version banner, the plugin's own printf()s, ...) goes to stderr instead."""
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
    ap.add_argument("--cpu-n", type=int, default=2048, help="grid side of the CPU sample (one instance per core)")
    ap.add_argument("--cpu-steps", type=int, default=6)
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
    ap.add_argument("--fill", type=float, default=0.35,
                    help="material-cell fraction of the extra `dense` measurement (0 = skip)")
    ap.add_argument("--no-parity-check", action="store_true", help="N > 1: skip the bit-identity check")
    ap.add_argument("--no-plugin-leg", action="store_true", help="N = 1: skip the e2e_plugin measurement")
    ap.add_argument("--no-ntff-leg", action="store_true", help="N = 1: skip the NTFF projection measurement")
    ap.add_argument("--plugin-n", type=int, default=4096, help="grid side of the e2e_plugin leg")
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
