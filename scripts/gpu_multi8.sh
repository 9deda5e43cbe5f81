nvidia-smi topo -m > gpurun_out/topo8_r2j.txt 2>&1
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --workload cfg3 --sharded --table-cells 1048576 > gpurun_out/bench_sharded_8gpu_r2j.json 2> gpurun_out/bench_sharded_8gpu_r2j.err; echo sharded rc=$?; tail -2 gpurun_out/bench_sharded_8gpu_r2j.err | cut -c1-300
python -c "
import json; d=json.loads(open('gpurun_out/bench_sharded_8gpu_r2j.json').read().strip().splitlines()[-1]); print(d['value']/1e6, d['per_rank'], d['every_cell_exactly_once']['ok'])"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_8gpu_r2j.json 2> gpurun_out/bench_8gpu_r2j.err; echo bench8 rc=$?
python -c "
import json; d=json.loads(open('gpurun_out/bench_8gpu_r2j.json').read().strip().splitlines()[-1]); e=d['e2e']; print('value', d['value']/1e6, d['ms_per_step_per_rank'], 'e2e', e['value']/1e6, 'coo', e['coo_route']['value']/1e6, 'packed', e['packed_route']['value']/1e6, 'loader', {k: round(v/1e6,2) for k,v in e.get('loader',{}).items() if k.startswith('cells')})"
