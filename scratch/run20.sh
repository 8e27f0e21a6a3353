timeout 300 python scripts/prof_tc.py stem agg c24 2>&1 | tail -10 > gpurun_out/prof_tc_now2.txt
timeout 200 python bench.py --steps 200 --warmup 5 --cpu-seconds 1 --no-extras > gpurun_out/s20.json 2> gpurun_out/s20.err
timeout 200 python scratch/small_layers.py 2>&1 | head -6 > gpurun_out/sl_20.txt
python - <<'P'
import json
d=json.load(open('gpurun_out/s20.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'])
P
