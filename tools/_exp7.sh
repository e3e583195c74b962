ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches.csv python tools/profile_step.py --tiles 2 --reps 2 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2b_launches.csv 4
