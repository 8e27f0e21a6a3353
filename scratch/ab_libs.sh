# A/B of library builds on one box: bench config B with each library named on the command line (diagnostics)
# usage: bash scratch/ab_libs.sh base h1 h20     ("base" = the product library)
for t in "$@" "$1"; do
  if [ "$t" = base ]; then lib=""; else lib="/root/repo/esmstereo_b200/csrc/libesm_b200_$t.so"; fi
  ESM_LIB=$lib python bench.py --steps 200 --warmup 5 --cpu-seconds 1 --no-extras > gpurun_out/ab_$t.json 2> gpurun_out/ab_$t.err
  python - <<P
import json
d=json.loads(open('gpurun_out/ab_$t.json').read().strip().splitlines()[-1]); print('$t', round(d['value'],2), round(d['ms_per_step'],4), d['clocks']['sm_mhz'])
P
done
