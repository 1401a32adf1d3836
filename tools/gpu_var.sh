#!/bin/bash
mkdir -p gpurun_out
for lib in default tools/libv_tri12.so tools/libv_tri16.so; do
  if [ "$lib" == "default" ]; then unset SSFM_LIB_PATH; else export SSFM_LIB_PATH=$PWD/$lib; fi
  echo "== $lib"; timeout 200 python tools/tri_probe.py 200000 2>&1 | grep "rep 2"
done
bash tools/gpu_libs.sh tools/libv_small6.so tools/libv_solve6.so
