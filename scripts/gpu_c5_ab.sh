#!/bin/bash
# C5 rate (65536 chains on one GPU, 5 transitions) with the in-tree library and with each library given as argument
run() { python scripts/run_configs.py c5 --scale 8 2>&1 | tail -2 | cut -c1-400; }
echo "== lib: in-tree"; run
for lib in "$@"; do echo "== lib: $lib"; AHMC_B200_LIB=$PWD/$lib run; done
