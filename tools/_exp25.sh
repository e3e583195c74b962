RSDET_TUNING=1 python -m rs_detection_b200.build --force > /dev/null
python tools/roi_sweep.py --paths 1 2>&1 | tail -1
for d in 4 3 2; do echo "drop 1 of $d"; RSDET_ROI_DBG_DROP=$d python tools/roi_sweep.py --paths 1 2>&1 | tail -1; done
echo "gather skipped"; RSDET_ROI_DBG_DROP=1 RSDET_ROI_DBG_SKIP_MAIN=1 python tools/roi_sweep.py --paths 1 2>&1 | tail -1
python -m rs_detection_b200.build --force > /dev/null
