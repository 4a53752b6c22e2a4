#!/bin/bash
# Build instrumented / experimental variants of libdsabf.so side by side (tools/variants/libdsabf_<name>.so).
# usage: tools/build_variants.sh name1:"-DFOO=1 -DBAR=2" name2:"..." ...
# The product build (dsabeamformer_b200/csrc/libdsabf.so) is untouched; pick a variant with DSABF_LIB=...
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/variants
for spec in "$@"; do
    name="${spec%%:*}"; flags="${spec#*:}"
    ( make -s -C dsabeamformer_b200/csrc -j4 BUILD=build_var_$name LIB=../../tools/variants/libdsabf_$name.so EXTRA="$flags" 2>&1 | grep -v deprecated || true ) &
done
wait
ls -la tools/variants
