RSDET_TUNING=1 python -m rs_detection_b200.build --force > /dev/null
for pad in 0 20; do echo "pad $pad"; RSDET_ROI_PAD_SMEM=$pad python bench.py --steps 30 --warmup 5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['components']['roi_fwd_kernel_ms_per_tile'])"; done
python -m rs_detection_b200.build --force > /dev/null
