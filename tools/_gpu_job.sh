mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -x -q > gpurun_out/tests7.log 2>&1; tail -3 gpurun_out/tests7.log
for f in 0 1; do
  MD_FUSE_EXPERT_ACT=$f timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/prof_c2_v15_f$f.csv > gpurun_out/bench_c2_v15_f$f.json 2>gpurun_out/bench_c2_v15.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_c2_v15_f$f.json").read().strip().splitlines()[-1])
print("FUSE_EXPERT_ACT=$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"])
PY
done
