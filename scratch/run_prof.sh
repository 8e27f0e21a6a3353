# launch list of the bench step + ncu --set full of the named kernels (summarised by scripts/summarize_profiles.py <tag>)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 1 --no-extras > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_kernels python scripts/prof_conv.py > gpurun_out/prof_kernels.log 2>&1
ls -la gpurun_out/launches.csv gpurun_out/prof_kernels.ncu-rep
