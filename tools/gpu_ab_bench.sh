#!/bin/bash
# developer A/B through bench.py: tools/gpu_ab_bench.sh <tag> "<bench args>" <lib_a> <lib_b> ...
tag=$1; shift; bargs=$1; shift
mkdir -p gpurun_out
for lib in "$@"; do
  SMCB200_LIB=$PWD/smc_jl_b200/$lib timeout 600 python bench.py $bargs --no-match --no-cpu 2>> gpurun_out/${tag}_abb.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(json.dumps({'lib':'$lib','cfg':d['config']['name'],'ms_per_step':d['ms_per_step'],'phase':d['phase_ms_per_step'],'e2e':d['e2e']['value'],'value':d['value']}))" | tee -a gpurun_out/${tag}_abb.jsonl
done
