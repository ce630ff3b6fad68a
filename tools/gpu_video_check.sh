#!/bin/bash
# First GPU session of the motion path (it was built in a session without GPU minutes and has only
# run under the thread emulator so far): parity tests of the predicted frames, BASELINE config 5
# (30 frames 720x576, IPPP) through tools/encode_video.py with its md5 next to the CPU reference's,
# the launch list of that run and one full ncu capture of a P-frame launch.  Output in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_video_check.sh r02'
set -u
R=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_predicted.py -x -q > $OUT/video_tests_$R.log 2>&1
tail -5 $OUT/video_tests_$R.log
# config 5 on one GPU, twice (the second run has warm contexts / page cache)
for i in 1 2; do
  timeout 600 python tools/encode_video.py --frames 30 --width 720 --height 576 --pattern ippp \
      --out $OUT/c5_$R.fco > $OUT/video_c5_${R}_run$i.json 2> $OUT/video_c5_${R}_run$i.err
  cat $OUT/video_c5_${R}_run$i.json
done
# the reference binary on the same frames (one core), for the md5 and the CPU time
if [ -x oracle/_ref/cfiasco ]; then
  python - <<'PY' > $OUT/video_c5_ref_$R.json 2>&1
import hashlib, json, os, subprocess, sys, tempfile, time
sys.path.insert(0, "oracle")
import gen_frames
with tempfile.TemporaryDirectory() as tmp:
    names = []
    for i, f in enumerate(gen_frames.video(30, 720, 576)):
        names.append(os.path.join(tmp, "f%02d.pgm" % i)); gen_frames.write_pnm(names[-1], f)
    env = dict(os.environ, FIASCO_DATA=os.path.abspath("oracle/_ref/data"), FIASCO_IMAGES=tmp)
    t0 = time.perf_counter()
    subprocess.run(["oracle/_ref/cfiasco", "--progress-meter=0", "-V", "0", "-q", "20", "--pattern=ippp", "-o",
                    os.path.join(tmp, "ref.fco")] + names, env=env, check=True)
    dt = time.perf_counter() - t0
    print(json.dumps({"impl": "reference", "cores": 1, "wall_s": dt, "mpixels_per_s": 30 * 720 * 576 / 1e6 / dt,
                      "fco_md5": hashlib.md5(open(os.path.join(tmp, "ref.fco"), "rb").read()).hexdigest()}))
PY
  cat $OUT/video_c5_ref_$R.json
fi
# every launch of the video run with its device time
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv \
    --log-file $OUT/video_launches_$R.csv python tools/encode_video.py --frames 8 --pattern ippp \
    > $OUT/video_launches_$R.log 2>&1
# one full capture of a P-frame launch (the second tile-kernel launch of a 2-frame run)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fiasco_tile_kernel -s 1 -c 1 \
    -f -o $OUT/video_prof_$R python tools/encode_video.py --frames 2 --pattern ip > $OUT/video_prof_$R.log 2>&1
ls -la $OUT
