#!/bin/bash
# Builds tuning variants of libslb200.so into 3dscanner-graycode_b200/lib/variants/ (decode kernel knobs).
set -e
cd "$(dirname "$0")/../3dscanner-graycode_b200"
mkdir -p lib/variants
for v in "$@"; do
  name=$(echo "$v" | tr -c 'A-Za-z0-9=' '_' )
  flags=$(echo "$v" | sed 's/,/ -D/g; s/^/-D/')
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC,-pthread $flags \
       -o lib/variants/libslb200_$name.so csrc/sl_decode.cu csrc/sl_cloud.cu csrc/sl_frames.cu csrc/sl_outlier.cu csrc/sl_peer.cu csrc/sl_host.cu csrc/sl_capi.cu &
done
wait
ls -la lib/variants
