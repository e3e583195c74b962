ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3_launches.csv python tools/profile_step.py --tiles 2 --reps 2 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r3_launches.csv 4 > gpurun_out/r3_launches_summary.txt; head -16 gpurun_out/r3_launches_summary.txt
ncu --set full --clock-control none --import-source on -k regex:"roi_align_fwd77|mc_class_sort|mc_fast_output|reduce_ov_pipe|ov_filter|ov_clip|transpose_prep|roi_geometry" -s 10 -c 10 -o gpurun_out/r3_final python tools/profile_step.py --tiles 2 --reps 2 > gpurun_out/r3_final_ncu.log 2>&1
python bench.py > gpurun_out/bench_r3_n1.json 2> gpurun_out/bench_r3_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r3_ref.json 2> gpurun_out/bench_r3_ref.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r3_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['components']['multiclass_nms_ms_per_tile'], d['verified']['ok'])"
tail -c 300 gpurun_out/bench_r3_ref.json
