#!/bin/bash
# Builds the reference's OWN fm_demod_benchmark driver (src/fm_demod_benchmark.cpp + src/app.cpp +
# src/rds_decoder, all unmodified) against the Broadcast_FM_Demod shim and libfmgpu.so.
#
#   build_overlay.sh <reference_root> <repo_root> <out_binary>
#
# An overlay of symlinks to <reference_root>/src is made under a temp dir with the three shim files
# laid over src/fm_demod/, so `#include "fm_demod/broadcast_fm_demod.h"` inside the reference's
# app.cpp resolves to the shim.  No reference source is copied into the repository.
set -euo pipefail
REF=${1:?reference root}; REPO=${2:?repo root}; OUT=${3:?output binary}
SHIM=$REPO/fm_radio_b200/csrc/shim
OVER=$(mktemp -d /tmp/fmgpu_overlay.XXXXXX)
trap 'rm -rf "$OVER"' EXIT
(cd "$REF/src" && find . -type d) | while read -r d; do mkdir -p "$OVER/$d"; done
(cd "$REF/src" && find . -type f) | while read -r f; do ln -s "$REF/src/$f" "$OVER/$f"; done
rm -f "$OVER"/fm_demod/broadcast_fm_demod.h "$OVER"/fm_demod/broadcast_fm_demod.cpp \
      "$OVER"/fm_demod/bpsk_synchroniser.h "$OVER"/fm_demod/bpsk_synchroniser.cpp
ln -s "$SHIM/fm_demod/broadcast_fm_demod.h" "$SHIM/fm_demod/broadcast_fm_demod.cpp" "$SHIM/fm_demod/bpsk_synchroniser.h" "$OVER/fm_demod/"
mkdir -p "$(dirname "$OUT")"
CXX=${CXX:-g++}
${CC:-gcc} -O2 -c -o "$OVER/getopt.o" "$OVER/getopt/getopt.c"
$CXX -std=c++17 -O2 -w -DNDEBUG -I"$OVER" -I"$REPO/include" -o "$OUT" \
    "$OVER/fm_demod_benchmark.cpp" "$OVER/app.cpp" "$OVER/fm_demod/broadcast_fm_demod.cpp" \
    "$OVER/dsp/calculate_fft_mag.cpp" \
    "$OVER/rds_decoder/crc10.cpp" "$OVER/rds_decoder/rds_database_decoder_handler.cpp" \
    "$OVER/rds_decoder/rds_decoder.cpp" "$OVER/rds_decoder/rds_group_sync.cpp" "$OVER/getopt.o" \
    -L"$REPO/fm_radio_b200" -lfmgpu -Wl,-rpath,"\$ORIGIN/.." -lpthread
# test driver of the shim's GUI-facing spectrum getters (tests/shim/spectra_check.cpp), next to the benchmark driver
if [ -f "$REPO/tests/shim/spectra_check.cpp" ]; then
    $CXX -std=c++17 -O2 -w -DNDEBUG -I"$OVER" -I"$REPO/include" -o "$(dirname "$OUT")/shim_spectra_check" \
        "$REPO/tests/shim/spectra_check.cpp" "$OVER/fm_demod/broadcast_fm_demod.cpp" "$OVER/dsp/calculate_fft_mag.cpp" \
        -L"$REPO/fm_radio_b200" -lfmgpu -Wl,-rpath,"\$ORIGIN/.." -lpthread
fi
echo "built $OUT"
