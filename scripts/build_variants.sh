#!/bin/bash
# Builds the default library and the tuning variant(s) next to it (see ROUND_NOTES.md):
#   lib/libiife.so         ptxas' own register choice
#   lib/tuned/libiife.so   64 registers for the SELL SpMV and the slot-plan PtAP kernels (minBlocks = 4)
# Select a variant at run time with IIFE_LIB=/abs/path/to/libiife.so.
set -e
cd "$(dirname "$0")/../interpolation-based-immersed-fea_b200/csrc"
make -j8
make -j8 BUILD=build_tuned LIB=../lib/tuned/libiife.so EXTRA_NVCCFLAGS="-DIIFE_SLOT_MINBLOCKS=4 -DIIFE_SELL_MINBLOCKS=4"
