#!/bin/bash
# tools/build_flags.sh NAME [-DFLAG ...]: builds build/variants/NAME.so from the tree as it is with extra nvcc flags
# (kernel A/B experiments behind preprocessor toggles; load with FWB200_LIB=...).  Prints registers / spills of the step kernels.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; shift
C="$ROOT/tum_adlr_deep_reinforcement_learning_b200/csrc"
mkdir -p "$ROOT/build/variants"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -Xptxas -v "$@" \
  -o "$ROOT/build/variants/$NAME.so" "$C/fw_step.cu" "$C/fw_gae.cu" "$C/fw_ppo.cu" "$C/fw_replay.cu" "$C/fw_comm.cu" > "$ROOT/build/variants/$NAME.log" 2>&1
grep -A3 "Compiling entry function" "$ROOT/build/variants/$NAME.log" | grep -v "^--" | paste - - - - | sed 's/ptxas info    : //g' \
  | sed "s/Compiling entry function '_Z[0-9N]*//" | grep "head_kernelIdLb1ELb0\|init_kernelIdLb1ELb0\|attempt_kernelIdLb1ELi32ELb0" \
  | sed 's/Function properties for [^ \t]*//; s/EEEvNS.*sm_100a.//' | cut -c1-220
