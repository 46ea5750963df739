#!/bin/bash
# Round measurements on the GPU box: benches (no profiler attached) first, then the ncu launch lists and full captures of
# the same programs.  Everything lands in gpurun_out/; tools/summarize_profiles.py turns it into profiles/*.txt.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/run_profiles.sh'
set -u
O=gpurun_out
T="${1:-r02b}"
python bench.py > $O/${T}_bench_single.json 2> $O/${T}_bench_single.err
python bench.py --impl reference --steps 20 --warmup 5 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err
python bench.py --seeds-per-gpu 64 --steps 100 --warmup 5 --total-seeds 0 > $O/${T}_bench_64seeds.json 2> $O/${T}_bench_64seeds.err
python bench.py --seeds-per-gpu 8 --steps 200 --warmup 10 --total-seeds 0 > $O/${T}_bench_8seeds.json 2> $O/${T}_bench_8seeds.err
python tools/stage_profile.py --seeds 8 16 32 64 > $O/${T}_stages.txt 2>&1
# launch lists (per-launch gpu__time_duration; cold-cache, serialised)
python tools/profile_step.py --steps 3 > $O/${T}_plain_single.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches_single_fp32.csv \
      python tools/profile_step.py --steps 3 > $O/${T}_ncu_single.log 2>&1
python tools/profile_step.py --steps 2 --seeds 64 --gemm-path tf32 > $O/${T}_plain_64.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches_64seeds_tf32.csv \
      python tools/profile_step.py --steps 2 --seeds 64 --gemm-path tf32 > $O/${T}_ncu_64.log 2>&1
python tools/profile_step.py --steps 2 --seeds 8 --gemm-path tf32 > $O/${T}_plain_8.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches_8seeds_tf32.csv \
      python tools/profile_step.py --steps 2 --seeds 8 --gemm-path tf32 > $O/${T}_ncu_8.log 2>&1
# launches per step of each program (+ 1: the gather), read off the plain runs: the full captures take the LAST step
per_step() { echo $(( $(grep -o '[0-9]* launches/step' $1 | grep -o '^[0-9]*') + 1 )); }
L1=$(per_step $O/${T}_plain_single.log); L64=$(per_step $O/${T}_plain_64.log); L8=$(per_step $O/${T}_plain_8.log)
echo "single $L1 64seeds $L64 8seeds $L8" > $O/${T}_launches_per_step.txt
# full captures: the last step's launches of each configuration
ncu --set full --clock-control none -k regex:"gemm_sk_kernel|gemm_fwd2_kernel|critic_head|policy_head_kernel|step_tail_kernel|policy_grad_kernel|replay_gather_kernel" --launch-skip $((2 * L1)) --launch-count $L1 -o $O/${T}_single_fp32 -f \
    python tools/profile_step.py --steps 3 > $O/${T}_full_single.log 2>&1
ncu --set full --clock-control none -k regex:"gemm_chain|gemm_ws|adam_stream|critic_head|policy_head|policy_grad|rank1|step_tail|replay_gather" --launch-skip $L64 --launch-count $L64 -o $O/${T}_64seeds_tf32 -f \
    python tools/profile_step.py --steps 2 --seeds 64 --gemm-path tf32 > $O/${T}_full_64.log 2>&1
ncu --set full --clock-control none -k regex:"gemm_chain|gemm_ws|adam_stream|critic_head|policy_head|policy_grad|rank1|step_tail|replay_gather" --launch-skip $L8 --launch-count $L8 -o $O/${T}_8seeds_tf32 -f \
    python tools/profile_step.py --steps 2 --seeds 8 --gemm-path tf32 > $O/${T}_full_8.log 2>&1
tail -2 $O/${T}_full_64.log
# gpurun brings back at most 64 MiB: keep the raw metric pages (what tools/summarize_profiles.py reads), drop the reports
for r in single_fp32 64seeds_tf32 8seeds_tf32; do
  if [ -f $O/${T}_$r.ncu-rep ]; then ncu -i $O/${T}_$r.ncu-rep --page raw --csv > $O/${T}_${r}_raw.csv 2>/dev/null; rm -f $O/${T}_$r.ncu-rep; fi
done
ls -la $O/${T}_*
