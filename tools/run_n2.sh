mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_distributed_gpu.py -m gpu -q -s 2>&1 | grep -E "^\[2-GPU|passed|failed" | tail -6) > gpurun_out/tests_n2.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_c2_n2.json 2> gpurun_out/bench_c2_n2.err
tail -4 gpurun_out/tests_n2.log; cut -c1-400 gpurun_out/bench_c2_n2.json; tail -3 gpurun_out/bench_c2_n2.err
