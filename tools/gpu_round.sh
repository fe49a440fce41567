#!/bin/bash
# One GPU-box session: GEMM core unit test first (bounded), then the parity suite, bench, ncu.
# usage: tools/gpu_round.sh <tag>
tag=${1:-rX}
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gemm_cores" -s > gpurun_out/${tag}_gemm.log 2>&1
rc=$?; tail -12 gpurun_out/${tag}_gemm.log
if [ $rc -ne 0 ]; then echo "GEMM core test failed (rc=$rc): stopping"; exit 0; fi
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/${tag}_tests.log 2>&1; tail -25 gpurun_out/${tag}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log
timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -3 gpurun_out/${tag}_bench.err; cat gpurun_out/${tag}_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --frames 4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 30 -c 3 -o gpurun_out/${tag}_gemm_tc python bench.py --frames 4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1; tail -3 gpurun_out/${tag}_ncu_full.log
ls -la gpurun_out | tail -12
