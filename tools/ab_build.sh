# usage: bash tools/ab_build.sh <out.so> <source.cu> [extra nvcc flags...]   -- rebuilds ONE translation unit with extra flags and links it
# with the other objects of the last full build into <out.so> (A/B experiments: TSFF_LIB_PATH=<out.so> python tools/bench_*.py)
set -e
OUT=$1; SRC=$2; shift 2
D=tsadar_b200/_lib
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v -Iinclude "$@" -c tsadar_b200/csrc/$SRC -o /tmp/ab_$(basename $OUT)_$SRC.o 2>&1 | grep -A2 "${AB_GREP:-k_table_bwdE}" | grep -E "registers|spill" || true
OBJS=$(ls $D/*.cu.o | grep -v "/$SRC.o")
nvcc -shared -o $OUT $OBJS /tmp/ab_$(basename $OUT)_$SRC.o -gencode arch=compute_100a,code=sm_100a -lcudart
echo built $OUT
