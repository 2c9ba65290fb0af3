# Multi-GPU part of the r1k profiles (gpurun --gpus 8): 2-GPU sharded tests, bench at N = 8 / 4 / 2, 100 M rows on 8 GPUs.
R=r1k
timeout 300 python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -2
for n in 8 4 2; do timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 200 --warmup 10 > gpurun_out/${R}_bench_n$n.json 2> gpurun_out/${R}_bench_n$n.err; done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 50 --warmup 5 --rows-per-gpu 12500000 --queries 64 > gpurun_out/${R}_bench_100M_n8.json 2> gpurun_out/${R}_bench_100M_n8.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 50 --warmup 5 --rows-per-gpu 12500000 --queries 10 > gpurun_out/${R}_bench_100M_q10_n8.json 2> gpurun_out/${R}_bench_100M_q10_n8.err
for f in gpurun_out/${R}_bench_n8.json gpurun_out/${R}_bench_n4.json gpurun_out/${R}_bench_n2.json gpurun_out/${R}_bench_100M_n8.json gpurun_out/${R}_bench_100M_q10_n8.json; do cut -c1-230 $f; done
