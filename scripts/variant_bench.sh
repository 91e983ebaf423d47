#!/bin/bash
# A/B of engine builds on one box: for every build/variants/lib_<name>.so, swap it in and run
# bench.py; prints value and the lean-interior figures.  usage: scripts/variant_bench.sh [bench args]
cp mpifdtd_b200/libmpifdtd_b200.so /tmp/lib_default.so
for lib in /tmp/lib_default.so build/variants/lib_*.so; do
  cp "$lib" mpifdtd_b200/libmpifdtd_b200.so
  python bench.py --no-cpu-baseline --steps 10 "$@" > /tmp/vb.json 2>/tmp/vb.err || { echo "$lib FAILED"; tail -3 /tmp/vb.err; continue; }
  python - "$lib" <<'PY'
import json, sys
d = json.loads([l for l in open('/tmp/vb.json') if l.startswith('{')][0])
li = d.get('lean_interior') or {}
print("%-36s value %.2f (h %.3f e %.3f ms)  lean %.2f (h %.3f e %.3f ms)" % (
    sys.argv[1].split('/')[-1], d['value'], d['roofline']['ms_per_launch'], d['roofline']['e_phase']['ms_per_launch'],
    li.get('value', 0), li.get('h_phase', {}).get('ms_per_launch', 0), li.get('e_phase', {}).get('ms_per_launch', 0)))
PY
done
cp /tmp/lib_default.so mpifdtd_b200/libmpifdtd_b200.so
