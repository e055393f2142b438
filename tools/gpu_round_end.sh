#!/bin/bash
# Everything the round's evidence needs from ONE GPU (run under gpurun): tests, bench line, ncu launch list of the same
# command, one ncu --set full capture of each of the largest kernels.  Outputs under gpurun_out/ with the given tag.
tag=${1:-r2s4}
out=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $out/${tag}_gputest.txt
python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
tail -c 400 $out/${tag}_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/${tag}_launches_bench_cfg3.csv \
    python bench.py --steps 2 --warmup 3 --no-thal --no-large --no-cpu-baseline > $out/${tag}_ncu_bench.log 2>&1
for kn in fast_scatter:2 encode_keys_packed:0 part_extend:2 fast_csr:0; do
  ncu --set full --import-source on --clock-control none -k regex:${kn%%:*} -s ${kn##*:} -c 1 -o $out/${tag}_${kn%%:*}_cfg3 python tools/prof_step.py cfg3 1 > $out/${tag}_ncu_${kn%%:*}.log 2>&1
done
python tools/prof_step.py cfg3 6 > $out/${tag}_prof_step_cfg3.txt 2>&1
ncu --set full --import-source on --clock-control none -k regex:thal_dimer_thread -c 1 -o $out/${tag}_thal_thread python tools/prof_thal4.py 200 > $out/${tag}_ncu_thal.log 2>&1
cat $out/${tag}_gputest.txt
head -c 600 $out/${tag}_bench_n1.json
