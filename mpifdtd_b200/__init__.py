"""mpifdtd_b200 -- B200-native time-stepping path behind rennone/mpiFDTD's C plugin
surface.  The product is libmpifdtd_b200.so (csrc/); this package is the ctypes
harness used by tests/ and bench.py."""
from . import binding  # noqa: F401
