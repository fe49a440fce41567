mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests.log 2>&1; tail -3 gpurun_out/r2_gpu_tests.log
timeout 700 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -3 gpurun_out/r2_bench.err; cut -c1-400 gpurun_out/r2_bench.json
OPB_PROFILE_DUMP=1 timeout 300 python bench.py --no-cpu-baseline --no-extra --steps 3 --warmup 3 > /dev/null 2> gpurun_out/r2_instream_profile.txt; tail -24 gpurun_out/r2_instream_profile.txt
