#!/bin/bash
# One GPU-box call that produces everything profiles/ holds for a round (run through gpurun):
#   bench line, reference arm, ncu launch list of the same bench command, one ncu --set full capture of a whole step.
# usage: bash tools/collect_profiles.sh r2
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
python bench.py --steps 200 --warmup 5 > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
# launch list of the bench command (kernel nodes of the replayed step graph are listed like direct launches)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-subrecords > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
# full capture of one whole step of tools/one_step.py (direct launches: COVO_GRAPH=0, so -s counts kernels only).
# default optimize_sigma path of a single environment (dense, D1-D3): hessian(3) + lanczos + inverses + combine + cholesky + rollout = 8 kernels per step
COVO_GRAPH=0 ncu --set full --clock-control none --import-source on -s 24 -c 8 -o gpurun_out/${TAG}_step_full python tools/one_step.py > gpurun_out/${TAG}_one_step_under_ncu.log 2>&1
# the tridiagonal path (E1-E3, the default of environment batches): 9 kernels per step
COVO_GRAPH=0 COVO_SIGMA=tridiag ncu --set full --clock-control none --import-source on -s 27 -c 9 -o gpurun_out/${TAG}_step_full_tridiag python tools/one_step.py > gpurun_out/${TAG}_one_step_tridiag_under_ncu.log 2>&1
# launch list of one batched step of BASELINE config 5 (512 environments behind one handle)
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${TAG}_envbatch_launches.csv python tools/bench_env_batch.py 512 3 > gpurun_out/${TAG}_envbatch_under_ncu.log 2>&1
