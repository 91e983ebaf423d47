#!/bin/bash
# Round 2, job 27: the whole GPU suite + smoke with the final build, the benchmark lines, the launch list,
# BASELINE configs[1] / configs[3] end to end, the sweep measurements.
mkdir -p gpurun_out
O=gpurun_out/r02_j27
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > $O.pytest.log 2>&1
tail -n 6 $O.pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; print(g.smoke())" ) > $O.smoke.log 2>&1
tail -n 4 $O.smoke.log
( timeout 900 python bench.py --steps 20 --warmup 5 ) > $O.bench_n1_tm.json 2> $O.bench_n1_tm.err
( timeout 900 python bench.py --steps 20 --warmup 5 --solver TE_UPML_2D --no-cpu-baseline --no-plugin-leg --no-ntff-leg ) > $O.bench_n1_te.json 2> $O.bench_n1_te.err
( timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > $O.bench_reference_arm.json 2> $O.bench_reference_arm.err
python - <<'PY'
import json
for f in ('bench_n1_tm','bench_n1_te','bench_reference_arm'):
    try:
        d=json.loads(open('gpurun_out/r02_j27.%s.json'%f).read().strip().splitlines()[-1])
        print(f,'value',d.get('value'),'e2e',(d.get('e2e') or {}).get('value'),'frac',(d.get('roofline') or {}).get('frac'),'clocks',d.get('clocks'))
    except Exception as ex:
        print(f,'FAILED',ex)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O.launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-lean-leg --no-plugin-leg --no-ntff-leg > $O.ncu_bench.log 2>&1
tail -n 2 $O.ncu_bench.log
( CONFIG2_SAVE=/tmp/c2_default CONFIG2_OUT=r02_j27.config2_full.json timeout 900 python scripts/config2_full.py ) > $O.config2.log 2>&1
( B200FDTD_FUSED=1 CONFIG2_SAVE=/tmp/c2_onepass CONFIG2_OUT=r02_j27.config2_full_onepass.json timeout 600 python scripts/config2_full.py --no-ref ) > $O.config2_onepass.log 2>&1
python - <<'PY' > gpurun_out/r02_j27.config2_forms.log 2>&1
import numpy as np
for solver in ("TM_UPML_2D", "TE_UPML_2D"):
    for what in ("field", "table"):
        a = np.load("/tmp/c2_default_%s_%s.npy" % (solver, what)); b = np.load("/tmp/c2_onepass_%s_%s.npy" % (solver, what))
        print(solver, what, "one-pass vs default form: bit-identical" if a.tobytes() == b.tobytes() else "DIFFER max %g" % np.abs(a - b).max())
PY
cat $O.config2.log $O.config2_onepass.log gpurun_out/r02_j27.config2_forms.log | grep -v "^saved\|^output\|mode\|simulator_finish\|^time" | tail -n 12
( timeout 900 python scripts/config4_ns_sweep.py ) > $O.config4.log 2>&1
tail -n 3 $O.config4.log
( timeout 600 python scripts/sweep_bench.py 256 1024 ) > $O.sweep_tm_upml.json 2> $O.sweep_tm_upml.err
( timeout 600 python scripts/sweep_bench.py --solver 7 256 1024 ) > $O.sweep_ns_te.json 2> $O.sweep_ns_te.err
cat $O.sweep_tm_upml.json $O.sweep_ns_te.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['workload'], 'batched %.2f s, one at a time %.2f s, x%.2f' % (d['batched']['seconds'], d['one_angle_at_a_time']['seconds'], d['speedup_batched']))"
( timeout 300 python scripts/split_bench.py 4096 ) > $O.split_4096.log 2>&1
grep solver_id $O.split_4096.log | cut -c1-160
