#!/bin/bash
# A/B builds of the NUTS kernel: advancedhmc.jl_b200/_variants/libahmc_b200_<tag>.so, selected at run time with
# AHMC_B200_LIB=<path>.  Usage: scripts/build_variants.sh minb4 fastdraw ...   (tags: minbN -> -DAHMC_NUTS_MINB=N,
# fastdraw -> -DAHMC_NUTS_FASTDRAW=1, altlayout -> -DAHMC_NUTS_ALT_LAYOUT=1, fastdraw_altlayout -> both).  The three NUTS translation units are recompiled, the rest is reused.
set -e
cd "$(dirname "$0")/.."
python advancedhmc.jl_b200/build.py
P=advancedhmc.jl_b200
O=${AHMC_OBJ_DIR:-/tmp/ahmc_b200_obj}
mkdir -p $P/_variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -O3"
for tag in "$@"; do
  case $tag in
    minb*) DEF="-DAHMC_NUTS_MINB=${tag#minb}" ;;
    fastdraw) DEF="-DAHMC_NUTS_FASTDRAW=1" ;;
    altlayout) DEF="-DAHMC_NUTS_ALT_LAYOUT=1" ;;
    fastdraw_altlayout) DEF="-DAHMC_NUTS_FASTDRAW=1 -DAHMC_NUTS_ALT_LAYOUT=1" ;;
    fastdraw_altlayout2) DEF="-DAHMC_NUTS_FASTDRAW=1 -DAHMC_NUTS_ALT_LAYOUT=2" ;;
    *) echo "unknown tag $tag"; exit 1 ;;
  esac
  ( for tu in ahmc_nuts ahmc_nuts_var ahmc_nuts_adapt; do
      nvcc $FLAGS $DEF -c $P/csrc/$tu.cu -o $O/${tu}_$tag.o &
    done; wait
    nvcc -shared -o $P/_variants/libahmc_b200_$tag.so $O/ahmc_api.o $O/ahmc_leapfrog.o $O/ahmc_adapt.o $O/ahmc_multinomial.o \
      $O/ahmc_dense.o $O/ahmc_nuts_$tag.o $O/ahmc_nuts_var_$tag.o $O/ahmc_nuts_adapt_$tag.o \
      -gencode arch=compute_100a,code=sm_100a -cudart shared ) &
done
wait
ls -la $P/_variants
