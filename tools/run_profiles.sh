#!/bin/bash
# Round measurements on the GPU box: benches (no profiler attached), then the ncu launch lists and full captures of the
# same commands.  Everything lands in gpurun_out/; tools/summarize_profiles.py turns it into profiles/*.txt.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/run_profiles.sh'
set -u
PHASE=${1:-all}        # bench | ncu | single | all  (gpurun copies back at most 64 MiB per call: the two full captures go alone)
O=gpurun_out
T="r01n"
if [ "$PHASE" == "single" ]; then      # only the single-seed configuration (bench line, launch list, full capture)
python bench.py > $O/${T}_bench_single.json 2> $O/${T}_bench_single.err
python tools/profile_step.py --steps 3 > $O/${T}_plain_single.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches_single_fp32.csv \
      python tools/profile_step.py --steps 3 > $O/${T}_ncu_single.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gemm_sk_kernel|gemm_fwd2_kernel|critic_head_kernel|policy_head_kernel|step_tail_kernel|policy_grad_kernel|replay_gather_kernel" --launch-skip 32 --launch-count 16 -o $O/${T}_single_fp32 -f \
    python tools/profile_step.py --steps 3 > $O/${T}_full_single.log 2>&1
exit 0
fi
if [ "$PHASE" != "ncu" ]; then
python bench.py > $O/${T}_bench_single.json 2> $O/${T}_bench_single.err
python bench.py --impl reference --steps 200 --warmup 5 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err
python bench.py --seeds-per-gpu 64 --steps 100 --warmup 5 > $O/${T}_bench_64seeds.json 2> $O/${T}_bench_64seeds.err
python bench.py --seeds-per-gpu 8 --steps 200 --warmup 10 > $O/${T}_bench_8seeds.json 2> $O/${T}_bench_8seeds.err
python bench.py --algo poac --steps 1000 --warmup 20 > $O/${T}_bench_poac.json 2> $O/${T}_bench_poac.err
python bench.py --algo goac --steps 1000 --warmup 20 > $O/${T}_bench_goac.json 2> $O/${T}_bench_goac.err
python bench.py --gemm-path tf32x3 --steps 1000 --warmup 20 > $O/${T}_bench_single_tf32x3.json 2> $O/${T}_bench_single_tf32x3.err
# launch lists (per-launch gpu__time_duration; cold-cache, serialised)
python tools/profile_step.py --steps 3 > $O/${T}_plain_single.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches_single_fp32.csv \
      python tools/profile_step.py --steps 3 > $O/${T}_ncu_single.log 2>&1
python tools/profile_step.py --steps 2 --seeds 64 --gemm-path tf32 > $O/${T}_plain_64.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches_64seeds_tf32.csv \
      python tools/profile_step.py --steps 2 --seeds 64 --gemm-path tf32 > $O/${T}_ncu_64.log 2>&1
python tools/bench_explore.py > $O/${T}_explore.json 2> $O/${T}_explore.err
fi
if [ "$PHASE" != "bench" ]; then
# full captures: the last step's launches of each configuration
ncu --set full --clock-control none --import-source on -k regex:"gemm_sk_kernel|gemm_fwd2_kernel|critic_head_kernel|policy_head_kernel|step_tail_kernel|policy_grad_kernel|replay_gather_kernel" --launch-skip 32 --launch-count 16 -o $O/${T}_single_fp32 -f \
    python tools/profile_step.py --steps 3 > $O/${T}_full_single.log 2>&1
ncu --set full --clock-control none -k regex:"gemm_ws_kernel|adam_stream_kernel|critic_head_kernel|policy_head_kernel|policy_grad_kernel|replay_gather_kernel" --launch-skip 19 --launch-count 19 -o $O/${T}_64seeds_tf32 -f \
    python tools/profile_step.py --steps 2 --seeds 64 --gemm-path tf32 > $O/${T}_full_64.log 2>&1
tail -2 $O/${T}_full_64.log
fi
