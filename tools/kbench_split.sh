#!/bin/bash
# test-only / emit-only timings of the meshlet stage (debug knob; results are NOT valid outputs)
for skip in 0 1 2; do echo "ORBIT_DEBUG_SKIP=$skip"; ORBIT_DEBUG_SKIP=$skip timeout 300 python tools/kbench.py default 2>&1 | tail -1; done
