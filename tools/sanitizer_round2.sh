#!/bin/bash
# compute-sanitizer over the round-2 kernels (run through gpurun): memcheck + synccheck on the tridiagonal step, the dense optimize_sigma
# step (adaptive cluster Lanczos with its checker warp, blocked cluster Gauss-Jordan), the fused peer exchange, the MPPI covariance
# update and the auto-reset; racecheck on the fast step (informational: its hand-overs are ordered by mbarriers / fences, which the
# tool does not model).
set -u
OUT=gpurun_out/r2_sanitizer.log
: > $OUT
run() { echo "=== $*" >> $OUT; "$@" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|error:|hazard|ok \[|passed|failed" | head -40 >> $OUT; }
run env COVO_SIGMA=tridiag compute-sanitizer --tool memcheck python tools/one_step.py 512 12
run env COVO_SIGMA=tridiag compute-sanitizer --tool synccheck python tools/one_step.py 512 12
run env COVO_SIGMA=dense compute-sanitizer --tool memcheck python tools/one_step.py 512 50
run env COVO_SIGMA=dense compute-sanitizer --tool synccheck python tools/one_step.py 512 50
run env COVO_SIGMA=dense compute-sanitizer --tool racecheck python tools/one_step.py 256 12
# the DSMEM bulk copy by itself (static shared memory, 8-CTA cluster): does memcheck accept cp.async.bulk shared::cta -> shared::cluster at all?
run compute-sanitizer --tool memcheck ./tools/microbench/dsmem_latency
run compute-sanitizer --tool memcheck python -m pytest tests/test_step_gpu.py -q -m gpu -k "fused_peer or sample_sharded"
run compute-sanitizer --tool memcheck python -m pytest tests/test_rollout_gpu.py -q -m gpu -k "covariance_update"
run compute-sanitizer --tool memcheck python -m pytest tests/test_env_gpu.py -q -m gpu -k "auto_reset"
cat $OUT
