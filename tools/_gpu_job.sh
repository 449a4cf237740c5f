mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01_tests_gpu_final.log 2>&1; tail -3 gpurun_out/r01_tests_gpu_final.log
timeout 600 python bench.py --profile-out gpurun_out/r01_per_op_c2_final.csv > gpurun_out/r01_bench_c2_final.json 2> gpurun_out/bench_final.err; tail -c 2500 gpurun_out/r01_bench_c2_final.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_bench_reference_arm.json 2> gpurun_out/bench_ref.err; tail -c 900 gpurun_out/r01_bench_reference_arm.json
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r01_launches_c2_mb512.csv python bench.py --steps 1 --warmup 3 --global-batch 512 --microbatch 512 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo rc=$?; wc -l gpurun_out/r01_launches_c2_mb512.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 --launch-skip 4 --launch-count 4 -o /tmp/gemm_full python tools/ncu_probe.py > gpurun_out/ncu_gemm.log 2>&1; echo rc=$?
ncu -i /tmp/gemm_full.ncu-rep --page raw --csv > gpurun_out/r01_ncu_gemm_full_raw.csv
S=$(stat -c %s /tmp/gemm_full.ncu-rep); echo size=$S; if [ "$S" -lt 25000000 ]; then cp /tmp/gemm_full.ncu-rep gpurun_out/r01_gemm_full.ncu-rep; fi
timeout 200 python tools/stock_torch_gpu.py --workload c2 --batch 128 --iters 4 > gpurun_out/r01_stock_torch_gpu_c2.json 2> gpurun_out/stock.err; tail -c 600 gpurun_out/r01_stock_torch_gpu_c2.json; tail -3 gpurun_out/stock.err
