#!/bin/bash
# Build a variant of libpvqt.so with extra nvcc defines, for A/B kernel experiments on the GPU box:
#   scripts/build_variant.sh mb5 -DPVQT_FFT_MIN_BLOCKS=5     ->  pitchvis_b200/lib/libpvqt_mb5.so
#   PVQT_LIB=pitchvis_b200/lib/libpvqt_mb5.so python bench.py
# Only the .cu files are recompiled with the defines; the host objects of the last regular build are reused.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
B=pitchvis_b200/_build/var_$name
mkdir -p $B
objs=""
for f in pitchvis_b200/csrc/*.cu; do
  o=$B/$(basename $f).o
  extra=""
  case $(basename $f) in analysis_kernels.cu|agc_kernels.cu|chroma_kernels.cu|spectrogram_kernels.cu) extra="-fmad=false";; esac
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
       $extra "$@" -I include -I pitchvis_b200/csrc -c $f -o $o &
  objs="$objs $o"
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o pitchvis_b200/lib/libpvqt_$name.so $objs pitchvis_b200/_build/*.cpp.o
echo pitchvis_b200/lib/libpvqt_$name.so
