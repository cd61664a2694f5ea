#!/bin/bash
# Round-2 final-state profile (run under gpurun): launch list + full captures.  Large launches of a kernel are picked by
# skipping its warm-up / small instances (SKIP per kernel), see tools/profile_gpu.sh for the general form.
TAG=${1:-r02b}
mkdir -p gpurun_out
BENCH="python bench.py --log-steps 22 --steps 1 --warmup 0 --no-cpu-baseline --no-adapter --sync-proofs"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/ncu_launches_${TAG}.log 2>&1
python tools/ncu_summary.py launches gpurun_out/launches_${TAG}.csv gpurun_out/launches_${TAG}.md > /dev/null
cap() {  # kernel-regex skip count
    N=$(echo $1 | tr -cd 'a-z0-9_')
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 \
        -f -o /tmp/full_${N} $BENCH > gpurun_out/ncu_full_${N}_${TAG}.log 2>&1
    python tools/ncu_summary.py full /tmp/full_${N}.ncu-rep gpurun_out/full_${N}_${TAG}.md > /dev/null
    ncu -i /tmp/full_${N}.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/source_${N}_${TAG}.csv.gz
    python tools/ncu_stalls.py gpurun_out/source_${N}_${TAG}.csv.gz > gpurun_out/stalls_${N}_${TAG}.txt
    rm -f /tmp/full_${N}.ncu-rep
}
cap merkle_warp_kernel 4 4
cap "fft4_pass_kernel.*13" 0 6
cap air_program_batch_kernel 0 3
cap fri_tail_kernel 0 2
cap k_store_fp_imm_logup 1 2
ls -la gpurun_out | tail -12
