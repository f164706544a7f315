# usage: tools/variants.sh name "-DLASV_BIN_U=3 -DLASV_BIN_CTAS=3" [file]   -> lasv_b200/variants/liblasv_b200_<name>.so
# (a variant differs from the library in the compile-time knobs of ONE source file, graph_kernels.cu by default;
#  LASV_B200_LIBRARY selects it)
set -e
cd "$(dirname "$0")/../lasv_b200/csrc"
f=${3:-graph_kernels}
mkdir -p ../variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr $2 -Xptxas -v -c $f.cu -o ../variants/${f}_$1.o 2> ../variants/${f}_$1.ptxas.log
objs=""; for o in graph_kernels supernode_kernels cleaning_kernels group_kernels streamed_insert ranking_kernels fastq_parse capi seq_reader gz_parallel; do
  if [ $o = $f ]; then objs="$objs ../variants/${f}_$1.o"; else objs="$objs $o.o"; fi; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/liblasv_b200_$1.so $objs -lz
grep -A2 "build_kernelILi2ELi3\|sg_insert_kernelILi2" ../variants/${f}_$1.ptxas.log | grep -E "Used|spill"
