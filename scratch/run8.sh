python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 --cpu-seconds 1 --no-extras > gpurun_out/s8_n2.json 2> gpurun_out/s8_n2.err
NCCL_MAX_CTAS=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 5 --cpu-seconds 1 --no-extras > gpurun_out/s8_n2_cta1.json 2> gpurun_out/s8_n2_cta1.err
python - <<'P'
import json
for f in ('gpurun_out/s8_n2.json','gpurun_out/s8_n2_cta1.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d.get('rank_spread'))
    except Exception as e: print(f,'ERR',e)
P
