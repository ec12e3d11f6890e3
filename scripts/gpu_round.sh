#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench, ncu launch list of one training step.
#   gpurun --timeout 900 -- 'bash scripts/gpu_round.sh TAG [full]'
# Every ncu command runs only after the same command has exited 0 without ncu.  Outputs: gpurun_out/*_TAG.*
TAG=${1:-x}
FULL=${2:-}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke_$TAG.log
python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"; cat $O/bench_$TAG.json
python scripts/profile_step.py > $O/plain_$TAG.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/launches_$TAG.csv python scripts/profile_step.py > $O/ncu_$TAG.log 2>&1
python scripts/ncu_summary.py $O/launches_$TAG.csv $O/seq_$TAG.txt > $O/sum_$TAG.txt 2>&1; head -30 $O/sum_$TAG.txt
if [ -n "$FULL" ]; then
  python scripts/roofline_kernel.py > $O/plain_roof_$TAG.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"$FULL" -c 2 -f -o $O/roof_$TAG \
      python scripts/roofline_kernel.py > $O/ncu_roof_$TAG.log 2>&1
  ncu -i $O/roof_$TAG.ncu-rep --page raw --csv > $O/roof_$TAG.raw.csv 2>/dev/null
  tail -5 $O/plain_roof_$TAG.log
fi
