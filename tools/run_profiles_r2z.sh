# How the r2z profiles (final build of round 2) were produced on one B200, run from the repo root through gpurun:
# GPU tests, smoke, bench, ncu launch list, ncu --set full of the headline chain and of the exact scan.
R=r2z
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/${R}_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 >> gpurun_out/${R}_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_ref.json 2>/dev/null
python bench.py --steps 3 --warmup 3 --quick --no-cpu-baseline > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 3 --warmup 3 --quick --no-cpu-baseline > gpurun_out/ncu_${R}_list.log 2>&1
python profiles/prof_headline.py > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:gemm_kernel|grouptop_tail_kernel|prep_queries' -s 15 -c 6 -f -o gpurun_out/prof_${R}_headline python profiles/prof_headline.py > gpurun_out/ncu_${R}_headline.log 2>&1
python profiles/prof_scan.py > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_tma -s 4 -c 4 -f -o gpurun_out/prof_${R}_scan python profiles/prof_scan.py > gpurun_out/ncu_${R}_scan.log 2>&1
python tools/sweep_collect.py > gpurun_out/${R}_sweep_collect.txt 2>&1
cat gpurun_out/${R}_tests.log; cut -c1-400 gpurun_out/${R}_bench.json
