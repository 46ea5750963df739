#!/bin/bash
# The single-seed part of tools/run_profiles.sh alone (bench line, ncu launch list, full capture of the last step).
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/run_profiles_single.sh r02d'
set -u
O=gpurun_out
T="${1:-r02d}"
python bench.py > $O/${T}_bench_single.json 2> $O/${T}_bench_single.err
python tools/profile_step.py --steps 3 > $O/${T}_plain_single.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches_single_fp32.csv \
      python tools/profile_step.py --steps 3 > $O/${T}_ncu_single.log 2>&1
L1=$(( $(grep -o '[0-9]* launches/step' $O/${T}_plain_single.log | grep -o '^[0-9]*') + 1 ))
echo "single $L1" > $O/${T}_launches_per_step.txt
ncu --set full --clock-control none -k regex:"gemm_sk_kernel|gemm_fwd2_kernel|critic_head|policy_head_kernel|step_tail_kernel|policy_grad_kernel|replay_gather_kernel" --launch-skip $((2 * L1)) --launch-count $L1 -o $O/${T}_single_fp32 -f \
    python tools/profile_step.py --steps 3 > $O/${T}_full_single.log 2>&1
ncu -i $O/${T}_single_fp32.ncu-rep --page raw --csv > $O/${T}_single_fp32_raw.csv 2>/dev/null; rm -f $O/${T}_single_fp32.ncu-rep
ls -la $O/${T}_*
