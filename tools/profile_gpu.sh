#!/bin/bash
# Run on the GPU box (under gpurun): launch list + full captures of the top kernels.
# Usage: tools/profile_gpu.sh <round-tag> [log_steps]
TAG=${1:-r01}
LOG=${2:-22}
mkdir -p gpurun_out
BENCH="python bench.py --log-steps $LOG --steps 1 --warmup 0 --no-cpu-baseline"
# every launch of the 2nd proof of the run (the first one is the cold start)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2300 -c 2300 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/ncu_launches_${TAG}.log 2>&1
for K in air_program_kernel fft_pass_kernel merkle_layer_kernel quotients_kernel eap_stage1_kernel; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 40 -c 12 \
        -f -o gpurun_out/full_${K}_${TAG} $BENCH > gpurun_out/ncu_full_${K}_${TAG}.log 2>&1
done
ls -la gpurun_out
