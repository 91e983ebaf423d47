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
