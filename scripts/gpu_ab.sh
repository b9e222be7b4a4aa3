#!/bin/bash
# A/B of two builds of the library (zstd_jni_b200/lib/libzstdb200.so vs libzstdb200_a.so): usage gpu_ab.sh <level> [level...]
for v in "" _a; do
  export ZSTDB200_LIBRARY=$PWD/zstd_jni_b200/lib/libzstdb200$v.so
  echo "== variant '$v'"
  for L in "$@"; do timeout 300 python scripts/gpu_enc.py 8192 2 $L 2>&1 | grep -E "parity|^rep 1" | cut -c1-200; done
done
