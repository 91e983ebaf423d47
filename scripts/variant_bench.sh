#!/bin/bash
# A/B of engine builds on one box: for every build/variants/lib_<name>.so, swap it in and run
# bench.py (TM and TE); prints value and the per-phase times of the default and the lean form.
# usage: scripts/variant_bench.sh [bench args]
cp mpifdtd_b200/libmpifdtd_b200.so /tmp/lib_default.so
for lib in /tmp/lib_default.so build/variants/lib_*.so; do
  cp "$lib" mpifdtd_b200/libmpifdtd_b200.so
  for solver in TM_UPML_2D TE_UPML_2D; do
  python bench.py --no-cpu-baseline --steps 10 --solver $solver "$@" > /tmp/vb.json 2>/tmp/vb.err || { echo "$lib FAILED"; tail -3 /tmp/vb.err; continue; }
  python - "$lib" $solver <<'PY'
import json, sys
d = json.loads([l for l in open('/tmp/vb.json') if l.startswith('{')][0])
li = d.get('lean_interior') or {}
print("%-22s %s value %.2f (h %.3f e %.3f ms)  lean %.2f (h %.3f e %.3f ms)" % (
    sys.argv[1].split('/')[-1], sys.argv[2][:2], d['value'], d['roofline']['ms_per_launch'], d['roofline']['e_phase']['ms_per_launch'],
    li.get('value', 0), li.get('h_phase', {}).get('ms_per_launch', 0), li.get('e_phase', {}).get('ms_per_launch', 0)))
PY
  done
done
cp /tmp/lib_default.so mpifdtd_b200/libmpifdtd_b200.so
