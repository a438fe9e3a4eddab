#!/bin/bash
# A/B builds of the NUTS kernel: advancedhmc.jl_b200/_variants/libahmc_b200_<tag>.so, selected at run time with
# AHMC_B200_LIB=<path>.  A tag is a '+'-joined list of knobs:
#   minbN      -DAHMC_NUTS_MINB=N          resident 4-warp CTAs per SM the register cap aims at (default 3)
#   fastdraw   -DAHMC_NUTS_FASTDRAW=1      lane-parallel variates, log-free (m, w) weights, deferred sum_alpha
#   altlayout1 -DAHMC_NUTS_ALT_LAYOUT=1    two chains per warp for 32 < D <= 128
#   altlayout2 -DAHMC_NUTS_ALT_LAYOUT=2    four chains per warp for 32 < D <= 128
#   fulltile   -DAHMC_NUTS_FULLTILE=1      extra instantiation with a compile-time D for D == G * E
#   reloadcoef -DAHMC_NUTS_RELOAD_COEF=1   model / metric coefficients re-read where used instead of held in registers
# e.g.  scripts/build_variants.sh fastdraw fastdraw+altlayout1 fastdraw+minb4
# The three NUTS translation units are recompiled, everything else is reused from the default build.
set -e
cd "$(dirname "$0")/.."
python advancedhmc.jl_b200/build.py
P=advancedhmc.jl_b200
O=${AHMC_OBJ_DIR:-/tmp/ahmc_b200_obj}
mkdir -p $P/_variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -O3"
for tag in "$@"; do
  DEF=""
  for knob in ${tag//+/ }; do
    case $knob in
      minb*) DEF="$DEF -DAHMC_NUTS_MINB=${knob#minb}" ;;
      fastdraw) DEF="$DEF -DAHMC_NUTS_FASTDRAW=1" ;;
      altlayout1|altlayout) DEF="$DEF -DAHMC_NUTS_ALT_LAYOUT=1" ;;
      altlayout2) DEF="$DEF -DAHMC_NUTS_ALT_LAYOUT=2" ;;
      fulltile) DEF="$DEF -DAHMC_NUTS_FULLTILE=1" ;;
      reloadcoef) DEF="$DEF -DAHMC_NUTS_RELOAD_COEF=1" ;;
      *) echo "unknown knob $knob"; exit 1 ;;
    esac
  done
  ( for tu in ahmc_nuts ahmc_nuts_var ahmc_nuts_adapt; do
      nvcc $FLAGS $DEF -c $P/csrc/$tu.cu -o $O/${tu}_$tag.o &
    done; wait
    nvcc -shared -o $P/_variants/libahmc_b200_$tag.so $O/ahmc_api.o $O/ahmc_leapfrog.o $O/ahmc_adapt.o $O/ahmc_multinomial.o \
      $O/ahmc_dense.o $O/ahmc_nuts_$tag.o $O/ahmc_nuts_var_$tag.o $O/ahmc_nuts_adapt_$tag.o \
      -gencode arch=compute_100a,code=sm_100a -cudart shared ) &
done
wait
ls -la $P/_variants
