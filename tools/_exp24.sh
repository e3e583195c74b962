RSDET_TUNING=1 python -m rs_detection_b200.build --force > /dev/null
python tools/roi_sweep.py --paths 1 --check 2>&1 | tail -1
for v in 1062 882 1242 1082; do echo "variant $v (warps, CTAs/SM, taps)"; RSDET_ROI_PVAR=$v timeout 120 python tools/roi_sweep.py --paths 1 --check 2>&1 | tail -1; done
python -m rs_detection_b200.build --force > /dev/null
