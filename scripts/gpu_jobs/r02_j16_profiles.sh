#!/bin/bash
# Round 2, job 16: the evidence files for profiles/: launch list of the bench command, ncu --set full of the one-pass
# kernels (TM exact / lean, TE exact / lean) at 16384^2 and of the NS TE split kernels at 4096^2, default bench line.
mkdir -p gpurun_out
O=gpurun_out/r02_j16
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O.launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-plugin-leg --no-lean-leg --fill 0 > $O.launches.log 2>&1
for cfg in "TM_UPML_2D 0 tm_onepass" "TM_UPML_2D 1 tm_onepass_lean" "TE_UPML_2D 0 te_onepass" "TE_UPML_2D 1 te_onepass_lean"; do
  set -- $cfg
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:onepass_kernel -s 2 -c 1 \
      -o $O.$3 -f python scripts/onepass_one.py $1 $2 20 32 16384 4 > $O.$3.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:split_te -s 40 -c 2 -o $O.split_ns_te -f \
    python scripts/split_bench.py 4096 7 > $O.split.log 2>&1
( timeout 900 python bench.py --steps 20 --warmup 5 ) > $O.bench_n1.json 2> $O.bench_n1.err
( timeout 900 python bench.py --steps 20 --warmup 5 --impl reference ) > $O.bench_ref.json 2> $O.bench_ref.err
( timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-plugin-leg --solver TE_UPML_2D ) > $O.bench_te.json 2> $O.bench_te.err
ls -la gpurun_out/r02_j16*; head -c 600 $O.bench_n1.json; echo; head -c 400 $O.bench_ref.json
