#!/bin/bash
# developer A/B on one GPU: tools/gpu_ab.sh <tag> <lib_a> <lib_b> ...   (per-phase stage times of the C2 stage per library)
tag=$1; shift
mkdir -p gpurun_out
for rep in 1 2; do
  for lib in "$@"; do
    SMCB200_LIB=$PWD/smc_jl_b200/$lib timeout 300 python tools/ab_stage.py >> gpurun_out/${tag}_ab.jsonl 2>> gpurun_out/${tag}_ab.err
  done
done
cat gpurun_out/${tag}_ab.jsonl
