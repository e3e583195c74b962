python -m pytest tests/test_gpu_roi_align.py -q 2>&1 | tail -1
python bench.py > gpurun_out/bench_r3h.json 2> gpurun_out/bench_r3h.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r3h.json').read().strip().splitlines()[-1]); r=d['roofline']; print(d['value'], d['ms_per_step'], r['frac'], r['launch_ms'], r['single_tile_launch']['frac'], d['e2e']['value'], d['verified']['ok'], d['verified']['checksum_out_tile0'])"
ncu --set full --clock-control none -k regex:"roi_align_fwd77p" -s 0 -c 1 -o gpurun_out/r3d_final python tools/profile_step.py --tiles 8 --roi-batch 8 --reps 1 --what roi > gpurun_out/r3d_final_ncu.log 2>&1
ncu -i gpurun_out/r3d_final.ncu-rep --page raw --csv > gpurun_out/r3d_final_raw.csv 2>/dev/null; python tools/ncu_summary.py gpurun_out/r3d_final_raw.csv | grep -E "time_duration|dram__bytes|sector_hit"
