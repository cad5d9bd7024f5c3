python -m pytest tests/test_gpu_multi.py tests/test_gpu_abi.py -m gpu -x -q 2>&1 | tail -6
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "rc=$?"; tail -c 2500 gpurun_out/r2_bench_n2.err; wc -c gpurun_out/r2_bench_n2.json; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n2.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(json.dumps(d['c3_sharded'])); print(json.dumps(d['c5_many_view']))"
