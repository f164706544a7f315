# usage: tools/variants.sh name "-DLASV_BIN_U=3 -DLASV_BIN_CTAS=3"   -> lasv_b200/variants/liblasv_b200_<name>.so
# (a variant differs from the library in graph_kernels.cu's compile-time knobs only; LASV_B200_LIBRARY selects it)
set -e
cd "$(dirname "$0")/../lasv_b200/csrc"
mkdir -p ../variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr $2 -Xptxas -v -c graph_kernels.cu -o ../variants/graph_kernels_$1.o 2> ../variants/graph_kernels_$1.ptxas.log
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/liblasv_b200_$1.so ../variants/graph_kernels_$1.o supernode_kernels.o cleaning_kernels.o group_kernels.o streamed_insert.o ranking_kernels.o capi.o seq_reader.o -lz
grep -A2 "build_kernelILi2ELi3" ../variants/graph_kernels_$1.ptxas.log | grep -E "Used|spill"
