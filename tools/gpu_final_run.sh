tag=r2s6; out=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $out/${tag}_gputest.txt
python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
tail -c 300 $out/${tag}_bench_n1.err
ncu --set full --import-source on --clock-control none -k regex:thal_dimer_thread -c 1 -o $out/${tag}_thal_thread python tools/prof_thal4.py 200 > $out/${tag}_ncu_thal.log 2>&1
python tools/prof_thal4.py 2000 > $out/${tag}_thal_2000rows.txt 2>&1
cat $out/${tag}_gputest.txt; head -c 400 $out/${tag}_bench_n1.json; cat $out/${tag}_thal_2000rows.txt
