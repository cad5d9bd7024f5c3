echo "== production lib"; timeout 300 python tools/kbench.py default 2>&1 | tail -1
for f in _variants/lib_ec128.so _variants/lib_ec64.so; do
  echo "== $f"; ORBIT_B200_LIB=$PWD/$f timeout 300 python tools/kbench.py default 2>&1 | tail -1
done
for f in lib_trace lib_trace_ec128; do
ORBIT_B200_LIB=$PWD/_variants/$f.so timeout 300 python tools/trace_frame.py > gpurun_out/r2_trace9_$f.txt 2>&1; grep -A200 "run 1" gpurun_out/r2_trace9_$f.txt | grep -E "^\s+\[|frame:|tiles done|cta done" | cut -c1-170
done
