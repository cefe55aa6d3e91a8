#!/bin/bash
# Short evidence refresh on ONE B200 (after a change that leaves the kernels' bodies alone): GPU tests, smoke, bench lines of every
# BASELINE config, ncu launch list of bench.py.  tools/profile_r02.sh adds the `ncu --set full` captures.
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $out/${tag}_gpu_tests.log
cat $out/${tag}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -1 $out/${tag}_smoke.log
python bench.py > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err
for c in c3 c4 c5; do python bench.py --config $c --steps 30 > $out/${tag}_bench_$c.json 2> $out/${tag}_bench_$c.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-match --no-cpu > $out/${tag}_launches_bench.log 2>&1
head -c 400 $out/${tag}_bench_c2.json; echo
