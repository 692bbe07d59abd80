#!/bin/sh
# tools/ab_build.sh NAME [nvcc -D flags...] -- experimental libpwntrace build into pwnfps_b200/ab/libpwntrace_NAME.so
# (git-ignored, travels to the GPU box); time it with PWN_LIBPWNTRACE=... python tools/ab_time.py
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p pwnfps_b200/ab
nvcc -ccbin "$(command -v g++)" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false \
    -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC -Xptxas -v "$@" -Iinclude -Ipwnfps_b200/csrc \
    -c pwnfps_b200/csrc/pwntrace.cu -o pwnfps_b200/ab/pwntrace_$name.o 2> pwnfps_b200/ab/ptxas_$name.log
nvcc -ccbin "$(command -v g++)" -shared -gencode arch=compute_100a,code=sm_100a -o pwnfps_b200/ab/libpwntrace_$name.so \
    pwnfps_b200/ab/pwntrace_$name.o pwnfps_b200/build/pwn_hostmath.o -lcudart
rm -f pwnfps_b200/ab/pwntrace_$name.o
grep -c "spill stores" pwnfps_b200/ab/ptxas_$name.log >/dev/null
echo "built pwnfps_b200/ab/libpwntrace_$name.so"
