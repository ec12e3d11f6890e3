#!/bin/bash
# Round-end evidence run (one GPU): bench line, ncu launch list of the SAME bench command, in-graph timeline,
# ncu --set full of the roofline kernel.  Outputs under gpurun_out/ with the given tag.
TAG=${1:-final}
O=gpurun_out
mkdir -p $O
python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_short_$TAG.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/bench_launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_bench_$TAG.log 2>&1
python scripts/ncu_summary.py $O/bench_launches_$TAG.csv > $O/bench_launches_summary_$TAG.txt 2>&1; head -12 $O/bench_launches_summary_$TAG.txt
python scripts/trace_step.py > $O/trace_$TAG.txt 2>/dev/null; grep "step span" $O/trace_$TAG.txt
python scripts/roofline_kernel.py > $O/plain_roof_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:spmm_v4 -c 2 -f -o $O/roof_$TAG python scripts/roofline_kernel.py > $O/ncu_roof_$TAG.log 2>&1
ncu -i $O/roof_$TAG.ncu-rep --page raw --csv > $O/roof_$TAG.raw.csv 2>/dev/null
tail -1 $O/plain_roof_$TAG.log | cut -c1-300
