python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "rc=$?"; tail -c 1500 gpurun_out/r2_bench_n8.err | grep -v "^\*\|OMP_NUM" ; python -c "
import json
lines=[l for l in open('gpurun_out/r2_bench_n8.json').read().split('\n') if l.startswith('{')]
d=json.loads(lines[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(json.dumps(d['c3_sharded'])); print(json.dumps(d['c5_many_view']))"
