#!/bin/bash
# One GPU call that produces the round's ncu artefacts under gpurun_out/ (summarised into profiles/ by
# scripts/summarize_ncu.py <tag>):  launch list of the bench command + one full capture of yee_E / yee_H.
# usage (on the GPU box, from the repo root): bash scripts/profile_round.sh r02
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --c5-n 0 > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:yee_ -s 10 -c 2 -o gpurun_out/${TAG}_coupler \
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --c5-n 0 > gpurun_out/${TAG}_full.log 2>&1
python bench.py --steps 1000 --warmup 5 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
tail -c 600 gpurun_out/${TAG}_bench_n1.json
