#!/bin/bash
# Run on the GPU box (under gpurun): launch list + full captures of the top kernels.
# The .ncu-rep files are summarised ON the box (tools/ncu_summary.py) and deleted: gpurun only
# copies back 64 MiB.   Usage: tools/profile_gpu.sh <round-tag> [log_steps] [kernel regexes...]
TAG=${1:-r01}
LOG=${2:-22}
shift 2
KERNELS=${@:-fft4_pass_kernel merkle_layer_kernel merkle_multi_kernel quotients_fast_kernel k_store_fp_imm_constraints k_store_fp_imm_logup}
mkdir -p gpurun_out
BENCH="python bench.py --log-steps $LOG --steps 1 --warmup 0 --no-cpu-baseline --no-adapter"
SKIP=${SKIP:-12}    # launches of the kernel skipped before the full capture (warm-up proofs); low-launch kernels: SKIP=1
COUNT=${COUNT:-8}
if [ -z "$NOLIST" ]; then
# every launch of the run's 3 proofs (value proof, staging proof, e2e proof): shares are per kernel over identical proofs
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/ncu_launches_${TAG}.log 2>&1
python tools/ncu_summary.py launches gpurun_out/launches_${TAG}.csv gpurun_out/launches_${TAG}.md > /dev/null
fi
for K in $KERNELS; do
    N=$(echo $K | tr -cd 'a-z0-9_')
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c $COUNT \
        -f -o /tmp/full_${N} $BENCH > gpurun_out/ncu_full_${N}_${TAG}.log 2>&1
    python tools/ncu_summary.py full /tmp/full_${N}.ncu-rep gpurun_out/full_${N}_${TAG}.md > /dev/null
    ncu -i /tmp/full_${N}.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/source_${N}_${TAG}.csv.gz
    rm -f /tmp/full_${N}.ncu-rep
done
ls -la gpurun_out | tail -20
