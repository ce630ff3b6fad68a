#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
TAG=${1:-r02f}
bash tools/gpu_r02c.sh $TAG
python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 2000 $OUT/${TAG}_bench.err; head -c 9000 $OUT/${TAG}_bench.json
