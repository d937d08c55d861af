#!/bin/bash
# A/B builds of libdlwp_b200.so with parts of the sliding-window kernel compiled out (same ABI; load with DLWP_B200_LIB=...):
#   libdlwp_b200_nopeers.so   no peer-memory halo stores in the epilogue
#   libdlwp_b200_noprobe.so   no barrier probes one row ahead in the MMA issuer
#   libdlwp_b200_plain.so     neither
cd "$(dirname "$0")/../dlwp_b200/csrc" || exit 1
NV=/usr/local/cuda/bin/nvcc
FL="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a"
for v in nopeers:-DDLWP_SW_NO_PEERS noprobe:-DDLWP_SW_NO_PROBE "plain:-DDLWP_SW_NO_PEERS -DDLWP_SW_NO_PROBE"; do
  name=${v%%:*}; defs=${v#*:}
  mkdir -p /tmp/dlwp_$name
  for f in conv conv_tc conv_sw_net_a conv_sw_net_b conv_sw_net_basic conv_sw_bf16 conv_fused elementwise plan train; do
    $NV $FL $defs -c $f.cu -o /tmp/dlwp_$name/$f.o &
  done
  wait
  $NV -shared -o libdlwp_b200_$name.so /tmp/dlwp_$name/*.o -cudart static -gencode arch=compute_100a,code=sm_100a
  echo built libdlwp_b200_$name.so
done
