set -e
RSDET_TUNING=1 python -m rs_detection_b200.build --force > /dev/null
echo "baseline"; python tools/roi_sweep.py --paths 1 2>&1 | tail -1
for pad in 20 56 130; do echo "pad smem +$pad KB"; RSDET_ROI_PAD_SMEM=$pad python tools/roi_sweep.py --paths 1 2>&1 | tail -1; done
for m in 0x3ff 0xffff 0x3ffff; do echo "mask $m"; RSDET_ROI_DBG_MASK=$m python tools/roi_sweep.py --paths 1 2>&1 | tail -1; done
echo "skip gather"; RSDET_ROI_DBG_SKIP_MAIN=1 python tools/roi_sweep.py --paths 1 2>&1 | tail -1
echo "skip lists+gather"; RSDET_ROI_DBG_SKIP_MAIN=2 python tools/roi_sweep.py --paths 1 2>&1 | tail -1
echo "mask 0x3ff + 1 tile"; RSDET_ROI_DBG_MASK=0x3ff python tools/roi_sweep.py --paths 1 --tiles 1 2>&1 | tail -1
