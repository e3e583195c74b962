set -e
RSDET_TUNING=1 python -m rs_detection_b200.build --force > /dev/null
python tools/roi_sweep.py --paths 9,1 --check 2>&1 | tail -2
RSDET_ROI_WARPS=7 python tools/roi_sweep.py --paths 1 --check 2>&1 | tail -1
python -m pytest tests/test_gpu_roi_align.py -x -q 2>&1 | tail -2
RSDET_ROI_WARPS=7 python -m pytest tests/test_gpu_roi_align.py -x -q 2>&1 | tail -2
