"""Host-side logic of the y-slab (multi-GPU) path on CPU: column partition, per-slab
NTFF plans and eps maps, and the halo protocol over a 2-rank gloo group with a fake
engine (the real engine needs a GPU).  Replaces Connection_ISend_IRecvH/E of
mpiTM_UPML.c:252-296."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from mpifdtd_b200 import binding as B
from mpifdtd_b200.slab import halo_peers, split_columns

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n_py,world", [(16384, 8), (1000, 3), (7, 7), (131072, 8), (129, 2)])
def test_split_columns_partitions_exactly(n_py, world):
    covered = 0
    for r in range(world):
        j0, nj = split_columns(n_py, world, r)
        assert j0 == covered and nj >= n_py // world
        covered += nj
    assert covered == n_py


def test_halo_peers():
    assert halo_peers(0, 0, 4) == (1, None) and halo_peers(0, 3, 4) == (None, 2)
    assert halo_peers(1, 0, 4) == (None, 1) and halo_peers(1, 2, 4) == (1, 3)
    assert halo_peers(0, 0, 1) == (None, None)


@pytest.mark.parametrize("stagger", [0.0, 0.5])
def test_slab_time_shift_tables_tile_the_global_table(plugin_lib, stagger):
    """Concatenating every rank's local NTFF table, in surface order, gives the
    single-slab table bit for bit (ntffTM.c:326-369 recurrence is walked from the
    edge start on every rank)."""
    L = plugin_lib
    L.field_init(B.FieldInfo(900, 2600, 10, 10, 500, 0, 50))
    box = L.field_getNTFFInfo()
    P = L.mpifdtd_ntff_point_count(C.byref(box))
    n_ang = 360

    def table(j0, nj):
        n = L.mpifdtd_ntff_local_count(C.byref(box), j0, nj)
        ptr = L.mpifdtd_ntff_time_shift(C.byref(box), n_ang, stagger, j0, nj)
        arr = np.ctypeslib.as_array((C.c_double * (n_ang * max(n, 1))).from_address(ptr)).copy()
        L.free(ptr)
        return arr[:n_ang * n].reshape(n_ang, n), n

    full, n_full = table(0, 260)
    assert n_full == P == 2 * (box.right - box.left) + 2 * (box.top - box.bottom)
    world = 3
    # global order: bottom | right | top | left ; owner of each point by its column j
    nx, ny = box.right - box.left, box.top - box.bottom
    j_of = np.concatenate([np.full(nx, box.bottom), np.arange(box.bottom, box.top),
                           np.full(nx, box.top), np.arange(box.bottom, box.top)])
    total = 0
    for r in range(world):
        j0, nj = split_columns(260, world, r)
        local, n = table(j0, nj)
        mask = (j_of >= j0) & (j_of < j0 + nj)
        assert n == int(mask.sum())
        assert np.array_equal(local.view(np.uint64), full[:, mask].view(np.uint64))
        total += n
    assert total == P


@pytest.mark.parametrize("stagger", [0.0, 0.5])
def test_direct_time_shift_tables_of_the_mpi_variant_solvers_tile_too(plugin_lib, stagger):
    """Ids 4/5 over several slabs: the direct-formula table (mpiTM_UPML.c:849-1037) of every slab keeps the
    points whose SAMPLED cell -- one column below the nominal one -- it owns; together they are the
    single-engine table."""
    L = plugin_lib
    L.field_init(B.FieldInfo(900, 2600, 10, 10, 500, 0, 50))
    box = L.field_getNTFFInfo()
    P = L.mpifdtd_ntff_point_count(C.byref(box))
    n_ang, dj = 360, -1
    L.mpifdtd_ntff_time_shift_direct.restype = C.c_void_p
    L.mpifdtd_ntff_time_shift_direct.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int]
    L.mpifdtd_ntff_local_count_shifted.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]

    def table(j0, nj):
        n = L.mpifdtd_ntff_local_count_shifted(C.byref(box), dj, j0, nj)
        ptr = L.mpifdtd_ntff_time_shift_direct(C.byref(box), n_ang, stagger, dj, j0, nj)
        arr = np.ctypeslib.as_array((C.c_double * (n_ang * max(n, 1))).from_address(ptr)).copy()
        L.free(C.c_void_p(ptr))
        return arr[:n_ang * n].reshape(n_ang, n), n

    full, n_full = table(0, 260)
    assert n_full == P
    nx = box.right - box.left
    j_of = np.concatenate([np.full(nx, box.bottom), np.arange(box.bottom, box.top),
                           np.full(nx, box.top), np.arange(box.bottom, box.top)]) + dj
    total = 0
    for world in (2, 5):
        total = 0
        for r in range(world):
            j0, nj = split_columns(260, world, r)
            local, n = table(j0, nj)
            mask = (j_of >= j0) & (j_of < j0 + nj)
            assert n == int(mask.sum())
            assert np.array_equal(local.view(np.uint64), full[:, mask].view(np.uint64))
            total += n
        assert total == P


WORKER = r"""
import os, sys, ctypes as C
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from mpifdtd_b200.slab import TorchHaloComm, exchange_halo, split_columns

rank, world, n_px, n_py = int(sys.argv[1]), int(sys.argv[2]), 12, 10
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=sys.argv[3], RANK=str(rank), WORLD_SIZE=str(world))
dist.init_process_group("gloo", rank=rank, world_size=world)

class FakeEngine:
    '''columns of two "fields" with ghosts, numpy-backed; same pack/unpack contract as
    b200fdtd_halo_pack/unpack (which 0: last owned -> low ghost, 1: first owned -> high ghost)'''
    def __init__(self, j0, nj):
        self.j0, self.nj = j0, nj
        self.f = {0: np.zeros((n_px, nj + 2), complex), 1: np.zeros((n_px, nj + 2), complex)}
    def _buf(self, ptr):
        return np.ctypeslib.as_array((C.c_double * (2 * n_px)).from_address(ptr)).view(complex)
    def halo_pack(self, which, ptr):
        col = self.nj if which == 0 else 1
        self._buf(ptr)[:] = self.f[which][:, col]
    def halo_unpack(self, which, ptr):
        col = 0 if which == 0 else self.nj + 1
        self.f[which][:, col] = self._buf(ptr)

j0, nj = split_columns(n_py, world, rank)
eng = FakeEngine(j0, nj)
comm = TorchHaloComm(n_px, "cpu")
halo = comm.pointers()
ok = True
for step in range(3):
    for which in (0, 1):
        # owned value encodes (field, step, global column, row)
        for c in range(nj):
            eng.f[which][:, c + 1] = (which + 1) * 1000 + step * 100 + (j0 + c) + 1j * np.arange(n_px)
        exchange_halo(eng, comm, which, rank, world, n_px, halo)
        if which == 0 and rank > 0:
            want = 1000 + step * 100 + (j0 - 1) + 1j * np.arange(n_px)
            ok &= np.array_equal(eng.f[0][:, 0], want)
        if which == 1 and rank < world - 1:
            want = 2000 + step * 100 + (j0 + nj) + 1j * np.arange(n_px)
            ok &= np.array_equal(eng.f[1][:, nj + 1], want)
    if rank == 0:
        ok &= np.all(eng.f[0][:, 0] == 0)          # bottom of the stack: ghost stays zero
    if rank == world - 1:
        ok &= np.all(eng.f[1][:, nj + 1] == 0)
# NTFF partial-sum reduce to rank 0
part = np.full(6, float(rank + 1))
comm.reduce_sum_to_root(part.ctypes.data, part.size)
if rank == 0:
    ok &= np.all(part == sum(range(1, world + 1)))
dist.barrier()
dist.destroy_process_group()
print("OK" if ok else "FAIL")
"""


@pytest.mark.parametrize("world", [2, 3])
def test_halo_protocol_over_gloo(world, tmp_path):
    import socket
    import subprocess
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    code = WORKER % {"root": ROOT}
    procs = [subprocess.Popen([sys.executable, "-c", code, str(r), str(world), str(port)],
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(world)]
    for p in procs:
        out, err = p.communicate(timeout=180)
        assert p.returncode == 0, err[-2000:]
        assert out.strip().endswith("OK"), (out, err[-2000:])
