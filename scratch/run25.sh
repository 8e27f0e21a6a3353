timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > gpurun_out/s25_t.log; cat gpurun_out/s25_t.log
grep -q passed gpurun_out/s25_t.log || exit 1
grep -q failed gpurun_out/s25_t.log && exit 1


timeout 200 python bench.py --steps 200 --warmup 5 --cpu-seconds 1 > gpurun_out/s25.json 2> gpurun_out/s25.err
python - <<'P'
import json
d=json.load(open('gpurun_out/s25.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity'), d.get('autotune_calls'))
P
