#!/bin/bash
# multi-GPU evidence on one box: tools/gpu_multi.sh <tag> <ngpus> [configs...]   (the sharding parity check, then bench.py per config)
tag=$1; n=$2; shift; shift
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR tests/multigpu_check.py > gpurun_out/${tag}_multigpu_check_n$n.txt 2>&1; tail -4 gpurun_out/${tag}_multigpu_check_n$n.txt
for c in "$@"; do
  steps=100; [ "$c" != "c2" ] && steps=30
  timeout 600 $TR bench.py --gpus $n --config $c --steps $steps > gpurun_out/${tag}_bench_${c}_n$n.json 2> gpurun_out/${tag}_bench_${c}_n$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_bench_${c}_n$n.json"))
    print("$c n=$n", "ms/step %.4f"%d["ms_per_step"], "value %.4g"%d["value"], "e2e %.4g"%d["e2e"]["value"], d["phase_ms_per_step"], "match", d.get("ess_match") and {k:d["ess_match"][k] for k in ("max_rel","cloud_bitexact","resamples")})
except Exception as e:
    print("$c n=$n FAILED", e); print(open("gpurun_out/${tag}_bench_${c}_n$n.err").read()[-1500:])
PY
done
