for i in 1 2 3; do
timeout 200 python bench.py --steps 200 --warmup 10 --cpu-seconds 1 --no-extras > gpurun_out/s13_$i.json 2> gpurun_out/s13_$i.err
done
python - <<'P'
import json
for i in (1,2,3):
    d=json.load(open('gpurun_out/s13_%d.json'%i)); print(i, d['value'], d['ms_per_step'], d['e2e']['value'])
P
nvidia-smi --query-gpu=name,clocks.sm,clocks.mem,power.draw,temperature.gpu --format=csv
