N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 100 --warmup 5 --cpu-seconds 1 --no-extras > gpurun_out/s19_n${N}_p2p.json 2> gpurun_out/s19_n${N}_p2p.err
#ESM_GATHER=nccl python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 100 --warmup 5 --cpu-seconds 1 --no-extras > gpurun_out/s19_n${N}_nccl.json 2> gpurun_out/s19_n${N}_nccl.err
python - <<P
import json
for k in ('p2p','nccl'):
    try:
        d=json.loads(open('gpurun_out/s19_n${N}_%s.json'%k).read().strip().splitlines()[-1]); print($N, k, d['value'], d['ms_per_step'], d.get('rank_spread'), d['config'].get('gather'))
    except Exception as e: print(k,'ERR',e)
P
grep -h "bench\]" gpurun_out/s19_n${N}_p2p.err | head -3
