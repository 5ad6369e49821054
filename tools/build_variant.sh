#!/bin/bash
# tools/build_variant.sh NAME [sed-expr ...]: builds build/variants/NAME.so from a copy of csrc/ with the sed
# expressions applied to fw_step.cu and fw_device.cuh (kernel A/B experiments; load with FWB200_LIB=...).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; shift
TMP=$(mktemp -d)
mkdir -p "$TMP/pkg" "$ROOT/build/variants"
cp -r "$ROOT/include" "$TMP/"
cp -r "$ROOT/tum_adlr_deep_reinforcement_learning_b200/csrc" "$TMP/pkg/"
for e in "$@"; do sed -i "$e" "$TMP/pkg/csrc/fw_step.cu" "$TMP/pkg/csrc/fw_device.cuh"; done
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -Xptxas -v "$TMP/pkg/csrc/fw_ppo.cu" \
  -o "$ROOT/build/variants/$NAME.so" "$TMP/pkg/csrc/fw_step.cu" "$TMP/pkg/csrc/fw_gae.cu" "$TMP/pkg/csrc/fw_replay.cu" "$TMP/pkg/csrc/fw_comm.cu" > "$ROOT/build/variants/$NAME.log" 2>&1
grep -A3 "Compiling entry function" "$ROOT/build/variants/$NAME.log" | grep -v "^--" | paste - - - - | sed 's/ptxas info    : //g' \
  | sed "s/Compiling entry function '_Z[0-9]*//" | grep "head_kernelIdLb1ELb0\|init_kernelIdLb1\|attempt_kernelIdLb1" \
  | sed 's/Function properties for [^ \t]*//; s/EEEvNS.*sm_100a.//' | cut -c1-200
rm -rf "$TMP"
