"""y-slab runs of the UPML solvers across GPUs (one process per GPU).

Replaces the reference's 2-D Cartesian MPI decomposition with per-step
Sendrecv (mpiTM_UPML.c:196-334, 718-748) by a 1-D split along y: rank r owns
global columns j in [r*N_PY/G, (r+1)*N_PY/G) plus one ghost column each side.
Per step and interface two halo columns move: after the H phase the top owned
column of Hx (TM) / Hz (TE) goes up, after the E phase the bottom owned column
of Ez / Ex goes down.  NTFF partial sums (each rank projects only the surface
points it owns) are reduced to rank 0 at the end -- legal because the transform
is linear in the fields.

Host-side preparation (grid, eps maps, coefficient tables, NTFF plan, per-step
source parameters) comes from the plugin's C helpers, i.e. the same code the
serial shim runs.  torch.distributed is plumbing only: rendezvous, the NCCL
send/recv of the two halo buffers and the final reduce.
"""
import ctypes as C

import numpy as np

from . import binding as B

KIND = {"TM_UPML_2D": 2, "TE_UPML_2D": 3}


def split_columns(n_py, world, rank):
    """Contiguous, near-equal column ranges; first ranks take the remainder."""
    base, extra = divmod(n_py, world)
    j0 = rank * base + min(rank, extra)
    return j0, base + (1 if rank < extra else 0)


def halo_peers(which, rank, world):
    """(send_to, recv_from) for one halo phase; None at the ends of the stack.
    which = 0: the H-phase column (TM Hx / TE Hz) goes UP, the ghost comes from below;
    which = 1: the E-phase column (TM Ez / TE Ex) goes DOWN, the ghost comes from above."""
    up = rank + 1 if rank + 1 < world else None
    down = rank - 1 if rank - 1 >= 0 else None
    return (up, down) if which == 0 else (down, up)


def exchange_halo(engine, comm, which, rank, world, n_px, halo):
    """pack -> send/recv -> unpack for one phase (engine: halo_pack/halo_unpack)."""
    send_ptr, recv_ptr = halo
    send_to, recv_from = halo_peers(which, rank, world)
    if send_to is not None:
        engine.halo_pack(which, send_ptr)
    comm.exchange(send_ptr, recv_ptr, n_px, send_to, recv_from)
    if recv_from is not None:
        engine.halo_unpack(which, recv_ptr)


class SlabRun:
    """One rank's slab.  `comm` is None for a single slab, else an object with
    exchange(which, send_ptr, recv_ptr, n_complex, send_to, recv_from) and
    reduce_sum_to_root(dev_ptr, n_doubles) (see TorchHaloComm)."""

    def __init__(self, model, solver, n_px, n_py, steps, rank=0, world=1, device=-1,
                 h_u_nm=10, pml=10, lambda_nm=500, angle_deg=0, comm=None, n_bins=None,
                 with_ntff=True, precision=0):
        self.L = B.lib()
        self.kind = KIND[solver] if isinstance(solver, str) else int(solver)
        self.rank, self.world, self.comm = rank, world, comm
        self.n_px, self.n_py, self.steps = n_px, n_py, steps
        model_id = B.MODELS[model] if isinstance(model, str) else int(model)
        self.L.models_setModel(model_id)
        self.L.field_init(B.FieldInfo(n_px * h_u_nm, n_py * h_u_nm, h_u_nm, pml, lambda_nm,
                                      angle_deg, steps))
        self.L.models_initModel()
        self.j0, self.nj = split_columns(n_py, world, rank)
        self.engine = B.Engine(self.kind, n_px, n_py, pml, self.j0, self.nj, device, precision=precision)
        if comm is not None and hasattr(comm, "torch") and str(getattr(comm, "device", "")).startswith("cuda"):
            # NCCL send/recv runs on torch's current stream: the engine's pack / unpack kernels must be
            # ordered with it, so the engine launches there too
            self.engine.set_stream(comm.torch.cuda.current_stream().cuda_stream)

        ti = np.empty((B.UPML_TABS, n_px))
        tj = np.empty((B.UPML_TABS, n_py))
        self.L.mpifdtd_upml_tables(self.kind, ti.ctypes.data, tj.ctypes.data)
        self.engine.set_tables(ti, tj)

        # eps maps for this slab only: TM EPS_EZ (i, j) D_XY; TE EPS_EX (i+1/2, j) D_Y,
        # EPS_EY (i, j+1/2) D_X  (fdtdTM_upml.c:237, fdtdTE_upml.c:374-375)
        modes = [(0.0, 0.0, B.D_XY)] if self.kind == 2 else [(0.5, 0.0, B.D_Y), (0.0, 0.5, B.D_X)]
        self.eps_host = []
        for slot, (xo, yo, mode) in enumerate(modes):
            eps = np.empty((n_px, self.nj))
            self.L.mpifdtd_fill_eps_slab(eps.ctypes.data, xo, yo, mode, self.j0, self.nj)
            B.check(self.L.b200fdtd_set_eps_slab(self.engine.h, slot, eps.ctypes.data), "set_eps_slab")
            self.eps_host.append(eps)

        self.box = self.L.field_getNTFFInfo()
        self.with_ntff = with_ntff
        if with_ntff:
            n_local = self.L.mpifdtd_ntff_local_count(C.byref(self.box), self.j0, self.nj)
            ptr = self.L.mpifdtd_ntff_time_shift(C.byref(self.box), 360,
                                                 0.0 if self.kind == 2 else 0.5, self.j0, self.nj)
            n_points = self.L.mpifdtd_ntff_point_count(C.byref(self.box))
            # bins kept: the first `steps` (all the far field reads, ntffTM.c:181), a number, or "full" =
            # the reference's whole arraySize (every tap of a short run lands somewhere)
            self.engine.n_bins = steps if n_bins is None else (self.box.arraySize if n_bins == "full" else n_bins)
            plan = B.NtffPlan(self.box.top, self.box.bottom, self.box.left, self.box.right,
                              n_points, n_local, steps, self.engine.n_bins, 360,
                              self.box.arraySize, ptr)
            B.check(self.L.b200fdtd_set_ntff_plan(self.engine.h, C.byref(plan)), "set_ntff_plan")
            self.L.free(ptr)
            self.n_local = n_local
        self.args = B.StepArgs()
        self.halo = None
        self.peer_halos = False

    def enable_peer_halos(self, gather_blobs):
        """Switch from NCCL send/recv to direct NVLink stores into the neighbours' ghost
        columns.  `gather_blobs(my_blob) -> [blob of rank 0, ..., blob of rank world-1]` is the
        only communication needed (once); afterwards a step issues no collective at all."""
        if self.world == 1:
            return
        blobs = gather_blobs(self.engine.peer_export())
        if self.rank > 0:
            self.engine.peer_attach(0, blobs[self.rank - 1])
        if self.rank + 1 < self.world:
            self.engine.peer_attach(1, blobs[self.rank + 1])
        self.peer_halos = True

    # -- halo buffers are allocated by the communicator (torch tensors) ------
    def attach_halo_buffers(self, send_ptr, recv_ptr):
        self.halo = (send_ptr, recv_ptr)

    def _exchange(self, which):
        if self.comm is None or self.world == 1:
            return
        exchange_halo(self.engine, self.comm, which, self.rank, self.world, self.n_px, self.halo)

    def step(self):
        """One update() of the serial solver, slab-wise: H, [halo], E + source,
        [halo], NTFF sample; then the host clock advances (simulator_calc)."""
        self.L.mpifdtd_upml_step_args(self.kind, 0, C.byref(self.args))
        e = self.engine
        if self.world == 1:
            e.step(self.args)
        elif self.peer_halos:
            # halo columns travel as peer stores inside the kernels; b200fdtd_step runs the flag
            # protocol around whatever form the step takes (one pass, or H phase / E phase)
            e.step(self.args)
        else:
            e.phase_h(self.args)
            self._exchange(0)
            e.phase_e(self.args)
            self._exchange(1)
            if self.with_ntff:
                e.phase_sample(self.args)
        self.L.field_nextStep()

    def project(self):
        self.engine.project()

    def far_field(self):
        """Reduce U/W over ranks, then translate/FFT/interpolate on rank 0."""
        self.engine.project()
        if self.comm is not None and self.world > 1:
            ptr, n = self.engine.uw_device()
            self.comm.reduce_sum_to_root(ptr, n)
        if self.rank != 0:
            return None
        out = np.zeros((321, 360))
        self.L.mpifdtd_upml_far_field(self.engine.h, self.kind, 0, out.ctypes.data)
        return out

    def gather_field(self, slot):
        """This rank's columns of a field, as [n_px, nj]."""
        return self.engine.get_field_slab(slot)

    def close(self):
        self.engine.close()


class TorchHaloComm:
    """NCCL plumbing through torch.distributed: two pinned-size device buffers of
    n_px complex doubles, batched isend/irecv on the current stream."""

    def __init__(self, n_px, device):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.device = device
        self.send = torch.zeros(2 * n_px, dtype=torch.float64, device=device)
        self.recv = torch.zeros(2 * n_px, dtype=torch.float64, device=device)

    def pointers(self):
        return self.send.data_ptr(), self.recv.data_ptr()

    def gather_blobs(self, blob):
        """all_gather of the 256-byte peer blobs (CUDA IPC handles)."""
        out = [None] * self.dist.get_world_size()
        self.dist.all_gather_object(out, blob)
        return out

    def exchange(self, send_ptr, recv_ptr, n_complex, send_to, recv_from):
        """send_ptr / recv_ptr are this object's own buffers (pointers()); kept in the signature so other
        communicators (the gloo protocol test) can take raw pointers"""
        dist = self.dist
        ops = []
        if send_to is not None:
            ops.append(dist.P2POp(dist.isend, self.send, send_to))
        if recv_from is not None:
            ops.append(dist.P2POp(dist.irecv, self.recv, recv_from))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def reduce_sum_to_root(self, dev_ptr, n_doubles):
        torch, dist = self.torch, self.dist
        # wrap the engine's U/W block without copying
        if str(self.device).startswith("cuda"):
            class _Raw:
                pass
            raw = _Raw()
            raw.__cuda_array_interface__ = {"shape": (int(n_doubles),), "typestr": "<f8",
                                            "data": (int(dev_ptr), False), "version": 3}
            t = torch.as_tensor(raw, device=self.device)
        else:       # host pointer (CPU protocol tests over gloo)
            buf = (C.c_double * int(n_doubles)).from_address(int(dev_ptr))
            t = torch.frombuffer(buf, dtype=torch.float64)
        dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
