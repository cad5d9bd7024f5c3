#!/bin/bash
# Box probe (SURVEY §7 step 0 / VERDICT r1 item 7): what the GPU box has for running the reference's own shaders.
out=gpurun_out/box_probe.txt
{
echo "== date"; date -u
echo "== nproc"; nproc
echo "== nvidia-smi -L"; nvidia-smi -L
echo "== driver"; nvidia-smi --query-gpu=driver_version,name,memory.total --format=csv
echo "== topo"; nvidia-smi topo -m 2>&1 | head -20
echo "== vulkan / GL driver files"
for pat in 'libGLX_nvidia.so*' 'nvidia_icd.json' 'libnvidia-glvkspirv*' 'libvulkan*' 'libvulkan_lvp*' '*icd*.json' 'libnvidia-glcore*' 'libEGL_nvidia*' 'libnvidia-eglcore*' 'libnvidia-gpucomp*' 'libnvidia-glsi*'; do
  echo "-- $pat"; find / -xdev -name "$pat" 2>/dev/null | head -10
done
echo "== ldconfig nvidia"; ldconfig -p | grep -i -E 'nvidia|vulkan|cuda' | head -40
echo "== /usr/lib/x86_64-linux-gnu libnvidia*"; ls /usr/lib/x86_64-linux-gnu | grep -i nvidia | head -60
echo "== toolchains"; for t in cargo rustc glslangValidator glslc spirv-dis vulkaninfo ncu; do printf "%s: " $t; command -v $t || echo missing; done
echo "== mounts with nvidia"; grep -i nvidia /proc/mounts | head -40
} > $out 2>&1
