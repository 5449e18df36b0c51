#!/bin/bash
# Everything the round-end evidence comes from, in one GPU-box call (run under gpurun from the repo root):
#   bash tools/round_measure.sh <tag>
# bench lines are taken WITHOUT a profiler; the ncu passes re-run small commands separately.
set -u
TAG=${1:-rX}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 600 gpurun_out/bench_${TAG}.err
python bench.py --impl reference > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err
# launch list of the bench command (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_${TAG}_bench.csv \
    python bench.py --steps 1 --warmup 3 --premises-per-step 2048 --skip-e2e --skip-cpu-baseline > gpurun_out/bench_under_ncu.json 2>/dev/null
# full captures: layer 0 of the first 262144-token chunk, and one retrieve (1024 x 200k, k = 100)
ncu --set full --clock-control none --import-source on -k 'regex:gemm_tc2_kernel|t5_attention' -s 0 -c 5 -f \
    -o gpurun_out/prof_${TAG}_encode python tools/profile_step.py --mode encode --layers 2 --premises 1024 --warm 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:gemm_tc2?_kernel|select_rescore|sample_threshold' -s 0 -c 4 -f \
    -o gpurun_out/prof_${TAG}_retrieve python tools/profile_step.py --mode retrieve --warm 0 > /dev/null 2>&1
# one state on the latency path: in-kernel timeline (no profiler), then the launch list of the same command
python tools/timeline_latency.py 225 > gpurun_out/timeline_${TAG}.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}_latency.csv \
    python tools/timeline_latency.py 225 > /dev/null 2>&1
# launch lists of one retrieve call at 1024 queries and at 1 query (4 calls each: the last is warm)
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${TAG}_retrieve.csv \
    python tools/profile_step.py --mode retrieve --warm 3 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${TAG}_retrieve_q1.csv \
    python tools/profile_step.py --mode retrieve --nq 1 --warm 3 > /dev/null 2>&1
python tools/latency_sweep.py > gpurun_out/latency_sweep_${TAG}.txt 2>&1
ls -la gpurun_out/*${TAG}*
head -c 1500 gpurun_out/bench_${TAG}.json
