RSDET_TUNING=1 python -m rs_detection_b200.build --force > /dev/null
python tools/roi_sweep.py --paths 1 2>&1 | tail -1
echo "pipelined lists, static columns"; RSDET_ROI_Q=2 timeout 120 python tools/roi_sweep.py --paths 1 --check 2>&1 | tail -1
RSDET_ROI_Q=2 timeout 120 python tools/roi_sweep.py --paths 1 2>&1 | tail -1
python -m rs_detection_b200.build --force > /dev/null
