#!/bin/bash
# BASELINE config 5 through the UNCHANGED reference command line front end linked against our libfiasco
# (fiasco_b200/lib/cfiasco): 30 frames 720x576, q=20, pattern ippp; wall time of the whole process (PGM
# files on disk -> .fco on disk), three runs, md5 of the stream (reference: e9d88f99690abf5b88c449478ff1dbf3).
set -u
D=$(mktemp -d)
python - "$D" <<'PY'
import os, sys
sys.path.insert(0, "oracle")
import gen_frames
for i, f in enumerate(gen_frames.video(30, 720, 576)):
    gen_frames.write_pnm(os.path.join(sys.argv[1], "f%02d.pgm" % i), f)
open(os.path.join(sys.argv[1], "small.fco"), "w").write("Fiasco\n")
PY
export FIASCO_DATA=$D FIASCO_IMAGES=$D
for i in 1 2 3; do
  s=$(date +%s.%N)
  fiasco_b200/lib/cfiasco --progress-meter=0 -V 0 -q 20 --pattern=ippp -o $D/out.fco $D/f[0-2][0-9].pgm 2>/dev/null
  rc=$?
  e=$(date +%s.%N)
  echo "{\"impl\": \"reference CLI on libfiasco (B200)\", \"run\": $i, \"rc\": $rc, \"wall_s\": $(echo "$e - $s" | bc -l 2>/dev/null || python -c "print($e - $s)"), \"fco_md5\": \"$(md5sum $D/out.fco | cut -d' ' -f1)\", \"bytes\": $(stat -c %s $D/out.fco)}"
done
rm -rf $D
