#!/bin/bash
# ncu of ONE full-batch launch (592 frames, 128 threads x 4 per SM), application replay so
# that no 50 GB save/restore happens; sections limited to keep the number of passes small
R=${1:-r01c}
SECTIONS=${SECTIONS:-"--section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section SchedulerStats --section ComputeWorkloadAnalysis --section InstructionStats --section SourceCounters"}
mkdir -p gpurun_out
FB200_NT=128 timeout 1500 ncu --replay-mode application --clock-control none --cache-control none \
  $SECTIONS --import-source on \
  -k regex:fiasco_tile_kernel -s 1 -c 1 -f -o gpurun_out/prof_batch_$R python tools/sweep_nt.py 4 37 > gpurun_out/prof_batch_$R.log 2>&1
grep Mpx gpurun_out/prof_batch_$R.log | tail -1
