for v in "" _a; do
  export ZSTDB200_LIBRARY=$PWD/zstd_jni_b200/lib/libzstdb200$v.so
  echo "== variant '$v'"
  timeout 300 python scripts/gpu_dec_classes.py 148 2>&1 | grep -o "class.*n=148\|k_dec_exec:[0-9.]*" | paste - - | tr '\n' ' '; echo
  timeout 300 python scripts/gpu_dec.py 8192 3 2>&1 | grep -E "^rep|ok"
done
