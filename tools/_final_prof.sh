set -x
python tools/sweep.py > gpurun_out/r2_sweep.json 2> gpurun_out/r2_sweep.err; tail -c 300 gpurun_out/r2_sweep.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_final_launches.csv python tools/profile_step.py --tiles 2 --reps 2 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2_final_launches.csv 4 > gpurun_out/r2_final_launches_summary.txt; cat gpurun_out/r2_final_launches_summary.txt
ncu --set full --clock-control none --import-source on -k regex:"roi_align_fwd77|mc_class_sort|mc_fast_output|reduce_ov_staged|ov_filter|ov_clip|transpose_prep|roi_order|roi_geometry" -s 11 -c 11 -o gpurun_out/r2_final python tools/profile_step.py --tiles 2 --reps 2 > gpurun_out/r2_final_ncu.log 2>&1
tail -2 gpurun_out/r2_final_ncu.log
python bench.py > gpurun_out/bench_r2_final_n1.json 2> gpurun_out/bench_r2_final_n1.err
python bench.py --impl reference > gpurun_out/bench_r2_final_ref.json 2> gpurun_out/bench_r2_final_ref.err
tail -c 400 gpurun_out/bench_r2_final_ref.json
