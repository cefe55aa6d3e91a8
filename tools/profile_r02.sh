#!/bin/bash
# Round-2 evidence run on ONE B200 (tools/profile_r02.sh <tag>): GPU tests, the bench lines of every BASELINE config, the
# ncu launch list of bench.py and one `ncu --set full` capture of every kernel of a C2 / C3 / C4 stage.
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $out/${tag}_gpu_tests.log
cat $out/${tag}_gpu_tests.log
python bench.py > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err
for c in c3 c4 c5; do python bench.py --config $c --steps 30 > $out/${tag}_bench_$c.json 2> $out/${tag}_bench_$c.err; done
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-match > $out/${tag}_launches_bench.log 2>&1
for c in c2 c3 c4; do
  ncu --set full --clock-control none --profile-from-start off -f -o /tmp/${tag}_full_$c \
      python tools/profile_kernels.py --config $c > $out/${tag}_full_$c.log 2>&1
  ncu -i /tmp/${tag}_full_$c.ncu-rep --page raw --csv > $out/${tag}_full_$c.raw.csv      # the report itself exceeds the 64 MiB return limit
  grep captured $out/${tag}_full_$c.log
done
head -c 600 $out/${tag}_bench_c2.json; echo; du -sh $out
