#!/bin/bash
# One GPU-box call that produces everything profiles/ holds for a round (run through gpurun):
#   bench line, ncu launch list of the same command, one ncu --set full capture of a whole step, clocks.
set -u
mkdir -p gpurun_out
python bench.py --steps 200 --warmup 5 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
# full capture of the 9 kernels of the fourth step of tools/one_step.py
ncu --set full --clock-control none --import-source on -s 27 -c 9 -o gpurun_out/step_full python tools/one_step.py > gpurun_out/one_step_under_ncu.log 2>&1
