set -e
RSDET_TUNING=1 python -m rs_detection_b200.build --force > /dev/null
python tools/roi_sweep.py --paths 9,1 --check 2>&1 | tail -2
for f in 1 2 3; do echo "load flavour $f"; RSDET_ROI_LD=$f python tools/roi_sweep.py --paths 1 --check 2>&1 | tail -1; done
echo "7 warps"; RSDET_ROI_WARPS=7 python tools/roi_sweep.py --paths 1 2>&1 | tail -1
