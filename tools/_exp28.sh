RSDET_TUNING=1 python -m rs_detection_b200.build --force > /dev/null
python tools/roi_sweep.py --paths 1 2>&1 | tail -1
echo "no locality order"; RSDET_ROI_NOORDER=1 python tools/roi_sweep.py --paths 1 --check 2>&1 | tail -1
python -m rs_detection_b200.build --force > /dev/null
