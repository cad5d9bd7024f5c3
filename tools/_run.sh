python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 tools/bench_configs.py c3 --timeline 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -16
