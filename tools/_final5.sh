python bench.py > gpurun_out/bench_r3g.json 2> gpurun_out/bench_r3g.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r3g.json').read().strip().splitlines()[-1]); r=d['roofline']; print(d['value'], d['ms_per_step'], r['frac'], r['launch_ms'], r['single_tile_launch']['frac'], d['e2e']['value'], d['verified'], d['gpu_launches'])"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3c_launches.csv python tools/profile_step.py --tiles 8 --roi-batch 8 --reps 1 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r3c_launches.csv 8 > gpurun_out/r3c_launches_summary.txt; head -9 gpurun_out/r3c_launches_summary.txt
ncu --set full --clock-control none -k regex:"roi_align_fwd77p" -s 0 -c 1 -o gpurun_out/r3c_final python tools/profile_step.py --tiles 8 --roi-batch 8 --reps 1 --what roi > gpurun_out/r3c_final_ncu.log 2>&1
