#!/bin/bash
# Round-2 profile capture (run on the GPU box through gpurun from the repo root):
#   1. launch list of one short bench run (per-launch durations, cold-cache and serialised under ncu)
#   2. one `--set full` capture of the three stage kernels of a steady-state step (source-level, -lineinfo)
# A number printed by a run under ncu is never a bench value.
TAG=${1:-r02}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:kuq:: -c 400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_${TAG}.log 2>&1
# matching launches: warm-up step (0-2), two timed steps (3-8), the instrumented stats step (9-11), then plain steps
ncu --set full --clock-control none --import-source on -k regex:'k_scan|k_lookup|k_resolve' --launch-skip 12 --launch-count 3 \
    -o gpurun_out/prof_${TAG} -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/prof_${TAG}.log 2>&1
ls -la gpurun_out/prof_${TAG}.ncu-rep
