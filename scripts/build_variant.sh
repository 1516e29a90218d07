#!/bin/bash
# Build libgtc.so from another git revision (or the working tree with extra -D flags) into build/ab/libgtc_<name>.so, for
# same-box A/B runs through GTC_LIB_PATH (GPUs of the pool differ by a few per cent, so A/B across gpurun calls is noise).
#   scripts/build_variant.sh <name> [git-ref|WORK] [extra nvcc flags...]
set -e
NAME=$1; REF=${2:-WORK}; shift; shift || true
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/guitar-tablature-classification_b200/csrc; INC=$ROOT/include
if [ "$REF" != "WORK" ]; then
  TMP=$(mktemp -d); git -C $ROOT archive $REF guitar-tablature-classification_b200/csrc include | tar -x -C $TMP
  SRC=$TMP/guitar-tablature-classification_b200/csrc; INC=$TMP/include
fi
OUT=$ROOT/build/ab; mkdir -p $OUT/$NAME
pids=()
for f in $SRC/*.cu; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I$INC "$@" -c $f -o $OUT/$NAME/$(basename ${f%.cu}).o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libgtc_$NAME.so $OUT/$NAME/*.o
echo built $OUT/libgtc_$NAME.so
