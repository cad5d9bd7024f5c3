#!/bin/bash
# Makes the NVIDIA Vulkan ICD loadable on the GPU box without a Vulkan loader or X11: libGLX_nvidia.so.0 links against
# libX11 / libXext (absent in the image); a headless compute client never calls into them, so empty stand-ins with the
# right SONAMEs and the undefined symbols (as functions returning 0) are enough. Output: $1 (directory to prepend to
# LD_LIBRARY_PATH). Test infrastructure only (oracle/vk).
set -e
out=${1:-/tmp/orbit_vkstubs}
mkdir -p $out
icd=$(ls /usr/local/nvidia/lib/libGLX_nvidia.so.0 /usr/lib/libGLX_nvidia.so.0 /usr/lib/x86_64-linux-gnu/libGLX_nvidia.so.0 2>/dev/null | head -1)
[ -z "$icd" ] && { echo "no libGLX_nvidia.so.0"; exit 1; }
needed=$(readelf -d $icd | grep NEEDED | sed 's/.*\[\(.*\)\]/\1/')
syms=$(nm -D --undefined-only $icd | awk '{print $2}' | grep -E '^(X|x|_X)' | sed 's/@.*//' | sort -u)
{
  for s in $syms; do echo "void* $s(void) { return 0; }"; done
} > $out/stubs.c
for lib in $needed; do
  if ! ldconfig -p | grep -q "$lib" && [ ! -e /usr/local/nvidia/lib/$lib ]; then
    gcc -shared -fPIC -o $out/$lib -Wl,-soname,$lib $out/stubs.c
    echo "stub $lib ($(echo $syms | wc -w) symbols)"
  fi
done
echo "icd=$icd"
