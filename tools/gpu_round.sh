#!/bin/bash
# One GPU-box session: GEMM core unit test first (bounded), then the parity suite, smoke, bench with the in-stream profile.
# usage: tools/gpu_round.sh <tag> [pytest -k expression]
tag=${1:-rX}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gemm_cores or kv_state_row" -s > gpurun_out/${tag}_gemm.log 2>&1
rc=$?; tail -12 gpurun_out/${tag}_gemm.log
if [ $rc -ne 0 ]; then echo "GEMM core test failed (rc=$rc): stopping"; exit 0; fi
timeout 1500 python -m pytest tests -m gpu -q -s ${2:+-k "$2"} > gpurun_out/${tag}_tests.log 2>&1; tail -40 gpurun_out/${tag}_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log
OPB_PROFILE_DUMP=1 timeout 400 python bench.py --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -30 gpurun_out/${tag}_bench.err; cut -c1-600 gpurun_out/${tag}_bench.json
