#!/bin/bash
# Evidence run of a round (one GPU-box session): full GPU suite, smoke, bench (all legs + cpu baseline), in-stream profile,
# ncu launch list of one step, ncu --set full of the last layers + tail of a step (GEMM variants AND the HBM-bound kernels),
# GPU damping sweep.   usage: tools/final_round.sh <tag> [skip_tests] [skip_damping]
# (the .ncu-rep files are summarised on the box and left there: gpurun_out/ is capped at 64 MiB)
tag=${1:-rX}
mkdir -p gpurun_out
if [ -z "$2" ]; then
  timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_gpu_tests.log 2>&1; tail -5 gpurun_out/${tag}_gpu_tests.log
fi
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -3 gpurun_out/${tag}_bench.err; cut -c1-300 gpurun_out/${tag}_bench.json
OPB_PROFILE_DUMP=1 timeout 300 python bench.py --no-cpu-baseline --no-extra --steps 3 --warmup 3 > /dev/null 2> gpurun_out/${tag}_instream_profile.txt; tail -24 gpurun_out/${tag}_instream_profile.txt
# one step = 75 launches once the object prologue is cached (first step: 85); list the second and third step
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 85 -c 150 --csv --log-file gpurun_out/${tag}_launches_ncu.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${tag}_ncu_list.log 2>&1
# matching kernels per step: 69 (79 in the first); take the last ~27 of the third step (and the first of the next): GATs layer 9, self layer 10, cross layer 11, tail
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:gemm_tc_kernel|kv_state_h|gats_aggregate|kv_state_reduce|in_stats_final" -s 190 -c 30 -o /tmp/${tag}_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${tag}_ncu_full.log 2>&1; tail -2 gpurun_out/${tag}_ncu_full.log
# the reports stay on the box (gpurun_out/ is capped at 64 MiB): summarise them here
python tools/ncu_summary.py /tmp/${tag}_full.ncu-rep gpurun_out/${tag}_ncu_summary.json "ncu --set full --clock-control none, matcher step at B=32 1024x7000: 30 launches (GATs layer 9, layers 10-11, tail)" | tail -32
if [ -z "$3" ]; then
timeout 900 python tools/damping_sweep.py --out gpurun_out/${tag}_damping_sweep.json > gpurun_out/${tag}_damping_sweep.log 2>&1; tail -12 gpurun_out/${tag}_damping_sweep.log
fi
# SuperPoint: launch list of one batch and ncu --set full of its 20 launches (second batch of sp_bench.py --once)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_superpoint_launches_ncu.csv python tools/sp_bench.py --once > gpurun_out/${tag}_sp_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -s 20 -c 20 -o /tmp/${tag}_superpoint_full python tools/sp_bench.py --once > gpurun_out/${tag}_sp_ncu_full.log 2>&1; tail -2 gpurun_out/${tag}_sp_ncu_full.log
python tools/ncu_summary.py /tmp/${tag}_superpoint_full.ncu-rep gpurun_out/${tag}_superpoint_ncu_summary.json "ncu --set full --clock-control none, one SuperPoint batch (B=8, 512x512): its 20 launches" | tail -22
ls -la gpurun_out | tail -12
