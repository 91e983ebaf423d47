"""config.txt reader (parser.c:3-20, configSample.txt:6-22, main.c:319-366)."""
import ctypes as C
import subprocess
import sys
import os

from mpifdtd_b200 import binding as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SAMPLE = """# simulation parameters
# '#' starts a comment, units are nm

2560  #width
2560  #height
10    #h_u
10    #pml
500   #lambda
2000  #step

#incidence angle, degrees
0     #start
90    #end
5     #delta

#Model
3

#Solver
2
"""


class Config(C.Structure):
    _fields_ = [("field_info", B.FieldInfo), ("startAngle", C.c_int), ("endAngle", C.c_int),
                ("deltaAngle", C.c_int), ("ModelType", C.c_int), ("SolverType", C.c_int)]


def test_read_config_eleven_values(plugin_lib, tmp_path):
    path = tmp_path / "config.txt"
    path.write_text(SAMPLE)
    cfg = Config()
    assert plugin_lib.mpifdtd_readConfig(str(path).encode(), C.byref(cfg)) == 0
    fi = cfg.field_info
    assert (fi.width_nm, fi.height_nm, fi.h_u_nm, fi.pml, fi.lambda_nm, fi.stepNum) == (2560, 2560, 10, 10, 500, 2000)
    assert (cfg.startAngle, cfg.endAngle, cfg.deltaAngle, cfg.ModelType, cfg.SolverType) == (0, 90, 5, 3, 2)
    assert fi.angle_deg == 0


def test_next_line_skips_comments_and_blanks(plugin_lib, tmp_path):
    path = tmp_path / "c.txt"
    path.write_text("# a\n\n   12 #x\n#b\n7\n")
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    plugin_lib.parser_nextLine.argtypes = [C.c_void_p, C.c_char_p]
    fp = libc.fopen(str(path).encode(), b"r")
    buf = C.create_string_buffer(1024)
    assert plugin_lib.parser_nextLine(fp, buf) and int(buf.value.split(b"#")[0]) == 12
    assert plugin_lib.parser_nextLine(fp, buf) and int(buf.value) == 7
    assert not plugin_lib.parser_nextLine(fp, buf)
    libc.fclose(fp)


def test_short_config_exits_2(tmp_path):
    path = tmp_path / "short.txt"
    path.write_text("100\n200\n")
    code = ("import sys, ctypes as C; sys.path.insert(0, %r)\n"
            "from mpifdtd_b200 import binding as B\n"
            "buf = (C.c_int * 16)()\n"
            "B.lib().mpifdtd_readConfig(%r, buf)\n" % (ROOT, str(path).encode()))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert p.returncode == 2 and "needs 11 values" in p.stdout


# ---- initConfigFromText (main.c:368-394): rank 0 reads, everybody else receives ----------------
BCAST_WORKER = r"""
import ctypes as C, os, sys
sys.path.insert(0, %(root)r)
sys.path.insert(0, os.path.join(%(root)r, "tests"))
from mpifdtd_b200 import binding as B
from test_parser_cpu import Config
rank, world, path = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
L = B.lib()
cfg = Config()
L.mpifdtd_initConfigFromText.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
L.mpifdtd_initConfigFromText(path.encode(), rank, world, C.byref(cfg), None, None, None)
print("CFG", rank, cfg.field_info.width_nm, cfg.field_info.height_nm, cfg.field_info.h_u_nm, cfg.field_info.pml,
      cfg.field_info.lambda_nm, cfg.field_info.angle_deg, cfg.field_info.stepNum, cfg.startAngle, cfg.endAngle,
      cfg.deltaAngle, cfg.ModelType, cfg.SolverType, flush=True)
"""


def test_config_broadcast_built_in_transport(plugin_lib, tmp_path):
    """Three ranks started by one launcher, no message layer: only rank 0 can see config.txt (the
    others get a path that does not exist) and all end up with the same struct; the banner is
    printed once."""
    path = tmp_path / "config.txt"
    path.write_text(SAMPLE)
    env = dict(os.environ, MPIFDTD_JOB_ID="pytest%d" % os.getpid(), MPIFDTD_BCAST_TIMEOUT_S="30")
    code = BCAST_WORKER % {"root": ROOT}
    procs = [subprocess.Popen([sys.executable, "-c", code, str(r), "3", str(path) if r == 0 else "/nonexistent/config.txt"],
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
             for r in (2, 1, 0)]                       # the readers start first and have to wait
    outs = [p.communicate(timeout=120) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    lines = sorted(l for o, _ in outs for l in o.splitlines() if l.startswith("CFG"))
    assert len(lines) == 3 and len({l.split(None, 2)[2] for l in lines}) == 1
    assert lines[0].split()[2:] == ["2560", "2560", "10", "10", "500", "0", "2000", "0", "90", "5", "3", "2"]
    assert sum(o.count("FieldSetting") for o, _ in outs) == 1
    B.lib().mpifdtd_configBroadcastDone()


GLOO_WORKER = r"""
import ctypes as C, os, sys
sys.path.insert(0, %(root)r)
sys.path.insert(0, os.path.join(%(root)r, "tests"))
import torch, torch.distributed as dist
from mpifdtd_b200 import binding as B
from test_parser_cpu import Config
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
SEND = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_void_p)
RECV = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_void_p)
def send(buf, count, dest, ctx):
    dist.send(torch.tensor([buf[k] for k in range(count)], dtype=torch.int32), dst=dest, tag=1)   # MPI tag 1 upstream
    return 0
def recv(buf, count, src, ctx):
    t = torch.zeros(count, dtype=torch.int32)
    dist.recv(t, src=src, tag=1)
    for k in range(count):
        buf[k] = int(t[k])
    return 0
L = B.lib()
cfg = Config()
L.mpifdtd_initConfigFromText.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, SEND, RECV, C.c_void_p]
L.mpifdtd_initConfigFromText((sys.argv[1] if rank == 0 else "/nonexistent").encode(), rank, world, C.byref(cfg),
                             SEND(send), RECV(recv), None)
print("CFG", rank, cfg.field_info.width_nm, cfg.field_info.stepNum, cfg.endAngle, cfg.ModelType, cfg.SolverType, flush=True)
dist.barrier()
dist.destroy_process_group()
"""


def test_config_broadcast_over_a_message_layer_gloo(plugin_lib, tmp_path):
    """The same with the caller's transport: torch.distributed send / recv over gloo, 2 ranks."""
    path = tmp_path / "config.txt"
    path.write_text(SAMPLE)
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER % {"root": ROOT})
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script), str(path)],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    lines = sorted(l for l in p.stdout.splitlines() if l.startswith("CFG"))
    assert [l.split()[2:] for l in lines] == [["2560", "2000", "90", "3", "2"]] * 2
