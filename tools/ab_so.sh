#!/bin/bash
# A/B of two builds of the library on one GPU box: the in-tree .so against build/libopb_prev.so (same ABI), in-stream profile each.
# usage: tools/ab_so.sh <tag>
tag=${1:-ab}
mkdir -p gpurun_out
cp onepose_b200/libonepose_b200.so /tmp/cur.so
for round in 1 2; do
  for v in cur prev; do
    if [ $v = prev ]; then cp build/libopb_prev.so onepose_b200/libonepose_b200.so; else cp /tmp/cur.so onepose_b200/libonepose_b200.so; fi
    OPB_PROFILE_DUMP=1 timeout 300 python bench.py --no-cpu-baseline --no-extra --steps 10 --warmup 5 > gpurun_out/${tag}_${v}${round}.json 2> gpurun_out/${tag}_${v}${round}.err
    echo "== $v $round"; grep -E "epi10|total|ms_per_step" gpurun_out/${tag}_${v}${round}.err | head -8; cut -c1-160 gpurun_out/${tag}_${v}${round}.json
  done
done
cp /tmp/cur.so onepose_b200/libonepose_b200.so
