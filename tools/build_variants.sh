#!/bin/bash
# Timing-only / experiment builds of the stream kernel v2: tools/build_variants.sh name:-DFLAG[,-DFLAG] ...
# -> vap_realtime_b200/libvapb200_<name>.so (the other objects are the ones of the regular build)
set -e
cd "$(dirname "$0")/../vap_realtime_b200/csrc"
FL="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}; flags=${flags//,/ }
  nvcc $FL $flags -c fused_tf2.cu -o /tmp/fused_tf2_$name.o
  nvcc -shared -o ../libvapb200_$name.so kernels_simt.o gemm_tc.o fused_tf.o /tmp/fused_tf2_$name.o vapb_api.o -gencode arch=compute_100a,code=sm_100a -lcudart_static -ldl -lrt -lpthread
  echo "built libvapb200_$name.so ($flags)"
done
