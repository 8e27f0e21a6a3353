cp esmstereo_b200/plans/b200.txt /tmp/merged.txt
timeout 200 python bench.py --steps 200 --warmup 10 --cpu-seconds 1 --no-extras > gpurun_out/s14_merged.json 2> gpurun_out/s14_merged.err
cp scratch/b200_new.txt esmstereo_b200/plans/b200.txt
timeout 200 python bench.py --steps 200 --warmup 10 --cpu-seconds 1 --no-extras > gpurun_out/s14_new.json 2> gpurun_out/s14_new.err
cp /tmp/merged.txt esmstereo_b200/plans/b200.txt
timeout 200 python bench.py --steps 200 --warmup 10 --cpu-seconds 1 --no-extras > gpurun_out/s14_merged2.json 2> gpurun_out/s14_merged2.err
python - <<'P'
import json
for i in ('merged','new','merged2'):
    d=json.load(open('gpurun_out/s14_%s.json'%i)); print(i, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('autotune_calls'))
P
