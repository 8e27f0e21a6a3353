for i in 1 2; do
ESM_LIB=$PWD/scratch/libesm_nopdl.so timeout 200 python bench.py --steps 200 --warmup 5 --cpu-seconds 1 --no-extras > gpurun_out/s21_nopdl$i.json 2> gpurun_out/s21_nopdl$i.err
timeout 200 python bench.py --steps 200 --warmup 5 --cpu-seconds 1 --no-extras > gpurun_out/s21_cur$i.json 2> gpurun_out/s21_cur$i.err
done
python - <<'P'
import json
for k in ('nopdl1','cur1','nopdl2','cur2'):
    d=json.load(open('gpurun_out/s21_%s.json'%k)); print(k, d['value'], d['ms_per_step'], d['e2e']['value'])
P
