#!/bin/bash
# Round 2, job 33: final build with the vacuum row-strips: the whole GPU suite + smoke, bench lines (TM, TE), ncu --set full of
# the four one-pass kernels at 16384^2 (DRAM traffic per cell), launch list of the bench command.
mkdir -p gpurun_out
O=gpurun_out/r02_j33
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > $O.pytest.log 2>&1
tail -n 6 $O.pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; print(g.smoke())" ) > $O.smoke.log 2>&1
grep "smoke\[" $O.smoke.log | cut -c1-200
( timeout 900 python bench.py --steps 20 --warmup 5 ) > $O.bench_n1_tm.json 2> $O.bench_n1_tm.err
( timeout 900 python bench.py --steps 20 --warmup 5 --solver TE_UPML_2D --no-cpu-baseline --no-plugin-leg --no-ntff-leg ) > $O.bench_n1_te.json 2> $O.bench_n1_te.err
python - <<'PY'
import json
for f in ('bench_n1_tm','bench_n1_te'):
    try:
        d=json.loads(open('gpurun_out/r02_j33.%s.json'%f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f,'value',d['value'],'e2e',d['e2e']['value'],'frac',r['frac'],'B/cell',r['algorithmic_bytes_per_cell'],'lean',d['lean_interior']['value'],d['lean_interior'].get('kernel'),'dense',(d.get('dense') or {}).get('value'),'clocks',d['clocks'])
    except Exception as ex:
        print(f,'FAILED',ex); print(open('gpurun_out/r02_j33.%s.err'%f).read()[-2000:])
PY
for cfg in "TM_UPML_2D 0 tm_onepass" "TM_UPML_2D 1 tm_onepass_lean" "TE_UPML_2D 0 te_onepass" "TE_UPML_2D 1 te_onepass_lean"; do
  set -- $cfg
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:onepass_kernel -s 2 -c 1 \
      -o $O.$3 -f python scripts/onepass_one.py $1 $2 20 32 16384 4 > $O.$3.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O.launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-lean-leg --no-plugin-leg --no-ntff-leg > $O.ncu_bench.log 2>&1
ls -la $O.*ncu-rep
