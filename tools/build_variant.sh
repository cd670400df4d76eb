#!/bin/bash
# Build a second copy of the library with ONE translation unit replaced (A/B timing on one box through $ANYREF_SAM_LIB):
#   tools/build_variant.sh <name> <csrc file> [extra nvcc flags]   ->  anyref_b200/libanyref_sam_<name>.so
# The other objects come from anyref_b200/build/ (run `python -m anyref_b200.build` first).
set -e
cd "$(dirname "$0")/../anyref_b200"
name=$1; src=$2; shift 2
base=$(basename "$src")
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
  -I csrc -c "$src" -o "/tmp/${name}_${base}.o" 2>/dev/null
nvcc -shared -o "libanyref_sam_${name}.so" $(ls build/*.o | grep -v "/${base}.o") "/tmp/${name}_${base}.o" -lcudart
echo "anyref_b200/libanyref_sam_${name}.so"
