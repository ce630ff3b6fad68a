#!/bin/bash
# Runs on the GPU box (under gpurun): bench line, ncu launch list of the same command, one
# `ncu --set full` capture of the dominant kernel (small batch: ncu replays ~40 passes and
# saves/restores the workspace each time), a limited-section application-replay capture of a
# FULL batch launch (592 frames, the shape the bench times), and the clocks.  Output in gpurun_out/.
set -u
R=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 3 --warmup 3 > $OUT/bench_$R.json 2> $OUT/bench_$R.err
cat $OUT/bench_$R.json
# every launch with its device time (cold cache, serialised: compare SHARES)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file $OUT/launches_$R.csv python bench.py --steps 2 --warmup 3 > $OUT/launches_$R.log 2>&1
# full capture of one tile-kernel launch (16 frames)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fiasco_tile_kernel -s 3 -c 1 \
    -f -o $OUT/prof_$R python bench.py --steps 1 --warmup 3 --batch 16 > $OUT/prof_$R.log 2>&1
# the full batch, application replay
SECTIONS="--section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section SchedulerStats --section ComputeWorkloadAnalysis --section InstructionStats --section SourceCounters" \
    bash tools/prof_batch.sh $R
ls -la $OUT
