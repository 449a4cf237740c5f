mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 --profile-out gpurun_out/per_op_c2_n8.csv > gpurun_out/bench_c2_n8.json 2> gpurun_out/bench_c2_n8.err
grep -v "^NCCL" gpurun_out/bench_c2_n8.json | cut -c1-600; tail -3 gpurun_out/bench_c2_n8.err
