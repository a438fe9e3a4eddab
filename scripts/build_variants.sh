#!/bin/bash
# A/B builds of the NUTS kernel occupancy knob: advancedhmc.jl_b200/_variants/libahmc_b200_minb{N}.so
set -e
cd "$(dirname "$0")/.."
python advancedhmc.jl_b200/build.py
P=advancedhmc.jl_b200
mkdir -p $P/_variants
for m in "$@"; do
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DAHMC_NUTS_MINB=$m \
      -c $P/csrc/ahmc_nuts.cu -o $P/_obj/ahmc_nuts_minb$m.o && \
    nvcc -shared -o $P/_variants/libahmc_b200_minb$m.so $P/_obj/ahmc_api.o $P/_obj/ahmc_leapfrog.o $P/_obj/ahmc_adapt.o $P/_obj/ahmc_nuts_minb$m.o \
      -gencode arch=compute_100a,code=sm_100a -cudart shared ) &
done
wait
ls -la $P/_variants
