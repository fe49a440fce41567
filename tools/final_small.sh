#!/bin/bash
# Short evidence refresh after a change to the extractor: its parity tests, the full bench line, its ncu launch list + summary.
tag=${1:-rX}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_superpoint_gpu.py -m gpu -q > gpurun_out/${tag}_superpoint_tests.log 2>&1; tail -3 gpurun_out/${tag}_superpoint_tests.log
timeout 700 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -3 gpurun_out/${tag}_bench.err; cut -c1-200 gpurun_out/${tag}_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_superpoint_launches_ncu.csv python tools/sp_bench.py --once > gpurun_out/${tag}_sp_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none -s 20 -c 20 -o /tmp/${tag}_superpoint_full python tools/sp_bench.py --once > gpurun_out/${tag}_sp_ncu_full.log 2>&1; tail -2 gpurun_out/${tag}_sp_ncu_full.log
python tools/ncu_summary.py /tmp/${tag}_superpoint_full.ncu-rep gpurun_out/${tag}_superpoint_ncu_summary.json "ncu --set full --clock-control none, one SuperPoint batch (B=8, 512x512): its 20 launches" | tail -22
