N=$1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/bench_n$N.err | tail -1 > gpurun_out/bench_n$N.json; python -c "
import json; d=json.load(open('gpurun_out/bench_n$N.json')); print('N=$N', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['clocks'])"
