python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python tools/kbench.py sweep > gpurun_out/r2_kb1.txt 2>&1; cat gpurun_out/r2_kb1.txt | tail -4
ncu --set full --clock-control none --import-source on -k regex:meshlet_test_direct -s 6 -c 4 -o gpurun_out/r2_prof1 python bench.py --steps 2 --warmup 1 --step-only --no-cpu-baseline > gpurun_out/r2_ncu1.log 2>&1; tail -3 gpurun_out/r2_ncu1.log
