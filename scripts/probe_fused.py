import os, sys
sys.path.insert(0, os.getcwd())
os.chdir("/tmp")
from mpifdtd_b200 import binding as B
os.environ["B200FDTD_FUSED"] = "2"
gpu = B.Plugin("MIE_CYLINDER", "TM_UPML_2D", 2048, 2100, steps=24, h_u_nm=10, angle_deg=20)
print("created", flush=True)
gpu.run()
gpu.sync()
print("ran", gpu.launches(), flush=True)
