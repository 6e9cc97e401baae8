#!/bin/bash
ls /dev/nvidia* 2>&1 | tr '\n' ' '; echo
nvidia-smi -q | grep -i "persistence mode" | sort | uniq -c
for v in unset 0 0,1; do
  if [ $v = unset ]; then unset CUDA_VISIBLE_DEVICES; else export CUDA_VISIBLE_DEVICES=$v; fi
  python - <<PY
import ctypes, time
t0 = time.time(); L = ctypes.CDLL("libcuda.so.1"); t1 = time.time(); rc = L.cuInit(0); t2 = time.time()
n = ctypes.c_int(); L.cuDeviceGetCount(ctypes.byref(n))
print("CUDA_VISIBLE_DEVICES=$v: dlopen %.2fs cuInit %.2fs rc=%d devices=%d" % (t1 - t0, t2 - t1, rc, n.value))
PY
done
