# How the r1k profiles were produced on one B200 (run from the repo root through gpurun): bench, ncu launch list,
# ncu --set full of the candidate pass and the tail, reference arm, churn tool, sweeps, GPU tests.
R=r1k
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${R}_list.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:gemm_kernel|grouptop_tail_kernel' -s 12 -c 4 -f -o gpurun_out/prof_$R python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${R}_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 9 -c 1 -f -o gpurun_out/prof_${R}_q256 python profiles/prof_gemm.py 256 > gpurun_out/ncu_${R}_q256.log 2>&1
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${R}_ref.json 2>/dev/null
make -C tools >/dev/null 2>&1
./tools/gf_demo --rows 10000000 --threads 4 --adds 10 --add-size 10000 --deletes 10 --delete-size 10000 --coalesce-us 100 --rounds 300 2>/dev/null | tail -1 > gpurun_out/${R}_churn_10M.json
./tools/gf_demo --rows 10000000 --threads 4 --coalesce-us 100 --rounds 300 2>/dev/null | tail -1 > gpurun_out/${R}_nochurn_10M.json
./tools/gf_demo --rows 1000000 --threads 5 --batch 2 --limit 1 --adds 50 --deletes 200 --rounds 2000 2>/dev/null | tail -1 > gpurun_out/${R}_demo_shape.json
python tools/sweep_batches.py > gpurun_out/${R}_sweep.txt 2>&1
python tools/sweep_onepass.py > gpurun_out/${R}_onepass_sweep.txt 2>&1
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
cut -c1-300 gpurun_out/bench_$R.json
