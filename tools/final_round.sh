#!/bin/bash
# Final evidence run of a round (one GPU-box session): bench with cpu baseline, in-stream profile, ncu launch list,
# ncu --set full of two layers' worth of the tensor-core kernels, smoke.   usage: tools/final_round.sh <tag>
tag=${1:-rX}
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log
timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -2 gpurun_out/${tag}_bench.err; cut -c1-400 gpurun_out/${tag}_bench.json
OPB_PROFILE_DUMP=1 timeout 200 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > /dev/null 2> gpurun_out/${tag}_instream_profile.txt; tail -24 gpurun_out/${tag}_instream_profile.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 260 --csv --log-file gpurun_out/${tag}_launches_ncu.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_list.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:gemm_tc_kernel|kv_state_h" -s 40 -c 10 -o gpurun_out/${tag}_tc python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1; tail -2 gpurun_out/${tag}_ncu_full.log
ls -la gpurun_out | tail -8
