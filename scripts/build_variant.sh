#!/bin/bash
# usage: [SRC=stft_v3.cu] build_variant.sh <name> <extra nvcc flags for $SRC (default stft_features.cu)...>  -> build/variants/libsonar_<name>.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants build/v_$name
python -c "import __graft_entry__ as g; g.build()" >/dev/null
cp build/*.o build/v_$name/
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-fno-fast-math,-ffp-contract=off "$@" -c sonido-sonar_b200/csrc/${SRC:-stft_features.cu} -o build/v_$name/$(basename ${SRC:-stft_features.cu} .cu).o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/libsonar_$name.so build/v_$name/*.o -lcudart_static -lpthread -ldl -lrt
echo built build/variants/libsonar_$name.so
