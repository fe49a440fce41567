#!/bin/bash
# SuperPoint GPU session: parity tests, the bench leg, ncu launch list + --set full of one batch.   usage: tools/sp_round.sh <tag>
tag=${1:-spX}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_superpoint_gpu.py -m gpu -q -s > gpurun_out/${tag}_tests.log 2>&1; tail -15 gpurun_out/${tag}_tests.log
timeout 300 python tools/sp_bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -3 gpurun_out/${tag}_bench.err; cut -c1-3000 gpurun_out/${tag}_bench.json
if [ -z "$2" ]; then
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${tag}_launches_ncu.csv python tools/sp_bench.py --once > gpurun_out/${tag}_ncu_list.log 2>&1
# second forward of --once: launches 20..39 (20 per batch); full set on its GEMM + helper kernels
timeout 600 ncu --set full --clock-control none --import-source on -s 20 -c 20 -o /tmp/${tag}_full python tools/sp_bench.py --once > gpurun_out/${tag}_ncu_full.log 2>&1; tail -2 gpurun_out/${tag}_ncu_full.log
python tools/ncu_summary.py /tmp/${tag}_full.ncu-rep gpurun_out/${tag}_ncu_summary.json "ncu --set full --clock-control none, one SuperPoint batch (B=8, 512x512): its 20 launches" | tail -22
fi
ls -la gpurun_out | tail -8
