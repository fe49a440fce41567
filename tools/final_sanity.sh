#!/bin/bash
# Last check of a round on one box: the whole GPU suite, smoke(), the default bench line.   usage: tools/final_sanity.sh <tag>
tag=${1:-rX}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_gpu_tests.log 2>&1; tail -3 gpurun_out/${tag}_gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log
timeout 700 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -3 gpurun_out/${tag}_bench.err; cut -c1-240 gpurun_out/${tag}_bench.json
OPB_PROFILE_DUMP=1 timeout 300 python bench.py --no-cpu-baseline --no-extra --steps 3 --warmup 3 > /dev/null 2> gpurun_out/${tag}_instream_profile.txt; tail -22 gpurun_out/${tag}_instream_profile.txt
