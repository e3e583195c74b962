set -e
RSDET_TUNING=1 python -m rs_detection_b200.build --force > /dev/null
ncu --set full --clock-control none --import-source on -k regex:"roi_align_fwd" -s 6 -c 2 -o gpurun_out/r2b_fwd77 python tools/roi_sweep.py --paths 1,9 --reps 1 > gpurun_out/r2b_fwd77.log 2>&1
tail -2 gpurun_out/r2b_fwd77.log
