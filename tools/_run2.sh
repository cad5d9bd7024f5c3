for f in _variants/lib_r2_c4.so _variants/lib_r2_c3.so _variants/lib_r8_c2.so _variants/lib_r4_c2.so; do
  echo "== $f"; ORBIT_B200_LIB=$PWD/$f timeout 300 python tools/kbench.py default 2>&1 | tail -1
done
