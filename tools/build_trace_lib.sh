#!/bin/bash
# builds fractulus_b200/_build/libfrx_trace.so = the production sources + -DFRX_K2_TRACE (tools/trace_k2.py)
set -e
cd "$(dirname "$0")/../fractulus_b200/_build"
nvcc -O3 -std=c++17 -lineinfo -ccbin /usr/bin/g++ -Xcompiler -fPIC -Xcompiler -ffp-contract=off -gencode arch=compute_100a,code=sm_100a \
  -I../../include -DFRX_K2_TRACE -c ../csrc/frx_history.cu -o /tmp/frx_history_trace.o
nvcc -shared -ccbin /usr/bin/g++ -o libfrx_trace.so frx_api.o frx_weights.o /tmp/frx_history_trace.o frx_solve.o frx_ode.o frx_compose.o frx_bench.o -gencode arch=compute_100a,code=sm_100a
