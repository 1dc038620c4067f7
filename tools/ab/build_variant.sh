#!/bin/bash
# build_variant.sh NAME -DFOO=1 ...: libfrx_NAME.so = the in-tree objects with frx_solve.cu recompiled under the given macros
set -e
cd "$(dirname "$0")/../.."
name=$1; shift
B=fractulus_b200/_build
nvcc -O3 -std=c++17 -lineinfo -ccbin /usr/bin/g++ -Xcompiler -fPIC -Xcompiler -ffp-contract=off -gencode arch=compute_100a,code=sm_100a "$@" -c fractulus_b200/csrc/frx_solve.cu -o tools/ab/frx_solve_$name.o
objs=$(ls $B/*.o | grep -v frx_solve.o)
nvcc -shared -ccbin /usr/bin/g++ -o tools/ab/libfrx_$name.so $objs tools/ab/frx_solve_$name.o -gencode arch=compute_100a,code=sm_100a
cuobjdump -res-usage tools/ab/libfrx_$name.so 2>/dev/null | grep -A1 "bldl" | grep -o "bldlILi[0-9]\|REG:[0-9]*\|STACK:[0-9]*" | paste - - - | tr '\n' ' '; echo
