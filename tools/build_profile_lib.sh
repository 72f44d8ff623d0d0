#!/bin/bash
# Instrumented build of K7 only (per-phase cycle counters, -DDTO_TDB_PROFILE) linked with the regular objects into
# build/libdto_b200_prof.so; use with DTO_B200_LIB=build/libdto_b200_prof.so python tools/tdb_phase_profile.py
set -e
cd "$(dirname "$0")/.."
mkdir -p build/prof
nvcc -DDTO_TDB_PROFILE -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I include \
     -c directtrajopt.jl_b200/csrc/tdb_dmma.cu -o build/prof/tdb_dmma.o
objs=$(ls directtrajopt.jl_b200/lib/obj/*.o | grep -v tdb_dmma)
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/libdto_b200_prof.so $objs build/prof/tdb_dmma.o
echo build/libdto_b200_prof.so
