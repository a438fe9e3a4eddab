#!/bin/bash
# A/B builds of the NUTS kernel (K3) and the dense tile kernel (K4): advancedhmc.jl_b200/_variants/libahmc_b200_<tag>.so, selected at run time with
# AHMC_B200_LIB=<path>.  A tag is a '+'-joined list of knobs:
#   minbN      -DAHMC_NUTS_MINB=N          resident 4-warp CTAs per SM the register cap aims at (default 3)
#   fastdraw   -DAHMC_NUTS_FASTDRAW=1      lane-parallel variates, log-free (m, w) weights, deferred sum_alpha
#   altlayout1 -DAHMC_NUTS_ALT_LAYOUT=1    two chains per warp for 32 < D <= 128
#   altlayout2 -DAHMC_NUTS_ALT_LAYOUT=2    four chains per warp for 32 < D <= 128
#   fulltile   -DAHMC_NUTS_FULLTILE=1      extra instantiation with a compile-time D for D == G * E
#   reloadcoef -DAHMC_NUTS_RELOAD_COEF=1   model / metric coefficients re-read where used instead of held in registers
#   densepad   -DAHMC_DENSE_PADDED_A=1     K4: padded matrices stored with the stage's leading dimension (one bulk copy per chunk)
#   denserel   -DAHMC_DENSE_MBAR_RELEASE=1 K4: stages released through mbarriers instead of a CTA barrier per chunk
#   dense3     -DAHMC_DENSE_STAGES=3       K4: three pipeline stages (needs denserel)
# e.g.  scripts/build_variants.sh fastdraw fastdraw+altlayout1 fastdraw+minb4 densepad+denserel+dense3
# Only the translation units a knob touches are recompiled (NUTS knobs: the three NUTS units; dense knobs: ahmc_dense.cu
# and ahmc_api.cu, which sizes the padded matrices); everything else is reused from the default build.
set -e
cd "$(dirname "$0")/.."
python advancedhmc.jl_b200/build.py
P=advancedhmc.jl_b200
O=${AHMC_OBJ_DIR:-/tmp/ahmc_b200_obj}
mkdir -p $P/_variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -O3"
for tag in "$@"; do
  DEF=""; NUTS=0; DENSE=0
  for knob in ${tag//+/ }; do
    case $knob in
      minb*) DEF="$DEF -DAHMC_NUTS_MINB=${knob#minb}"; NUTS=1 ;;
      fastdraw) DEF="$DEF -DAHMC_NUTS_FASTDRAW=1"; NUTS=1 ;;
      altlayout1|altlayout) DEF="$DEF -DAHMC_NUTS_ALT_LAYOUT=1"; NUTS=1 ;;
      altlayout2) DEF="$DEF -DAHMC_NUTS_ALT_LAYOUT=2"; NUTS=1 ;;
      fulltile) DEF="$DEF -DAHMC_NUTS_FULLTILE=1"; NUTS=1 ;;
      reloadcoef) DEF="$DEF -DAHMC_NUTS_RELOAD_COEF=1"; NUTS=1 ;;
      densepad) DEF="$DEF -DAHMC_DENSE_PADDED_A=1"; DENSE=1 ;;
      denserel) DEF="$DEF -DAHMC_DENSE_MBAR_RELEASE=1"; DENSE=1 ;;
      dense3) DEF="$DEF -DAHMC_DENSE_STAGES=3"; DENSE=1 ;;
      *) echo "unknown knob $knob"; exit 1 ;;
    esac
  done
  ( OBJS=""
    for tu in ahmc_api ahmc_leapfrog ahmc_adapt ahmc_multinomial ahmc_dense ahmc_nuts ahmc_nuts_var ahmc_nuts_adapt; do
      re=0
      case $tu in
        ahmc_nuts*) re=$NUTS ;;
        ahmc_dense|ahmc_api) re=$DENSE ;;
      esac
      if [ $re = 1 ]; then
        nvcc $FLAGS $DEF -c $P/csrc/$tu.cu -o $O/${tu}_$tag.o &
        OBJS="$OBJS $O/${tu}_$tag.o"
      else
        OBJS="$OBJS $O/$tu.o"
      fi
    done; wait
    nvcc -shared -o $P/_variants/libahmc_b200_$tag.so $OBJS -gencode arch=compute_100a,code=sm_100a -cudart shared ) &
done
wait
ls -la $P/_variants
