mkdir -p gpurun_out
{
echo "=== default (4 epilogue warps)"; python tools/gemm_micro.py
echo "=== default debug=1"; MD_GEMM_DEBUG=1 python tools/gemm_micro.py
echo "=== 8 epilogue warps"; MD_LIB_PATH=$PWD/variants/libmicrodit_b200_epi8.so python tools/gemm_micro.py
} > gpurun_out/gemm_micro_v3.log 2>&1
cat gpurun_out/gemm_micro_v3.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k gemm > gpurun_out/tests6.log 2>&1; tail -3 gpurun_out/tests6.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/prof_c2_v14.csv > gpurun_out/bench_c2_v14.json 2>gpurun_out/bench_c2_v14.err; tail -c 1500 gpurun_out/bench_c2_v14.json
