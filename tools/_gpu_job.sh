mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/tests5.log 2>&1; tail -5 gpurun_out/tests5.log
for ts in 0 1; do
  MD_GEMM_TMA_STORE=$ts timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/prof_c2_v13_ts$ts.csv > gpurun_out/bench_c2_v13_ts$ts.json 2>gpurun_out/bench_c2_v13.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_c2_v13_ts$ts.json").read().strip().splitlines()[-1])
print("TMA_STORE=$ts", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"])
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ --launch-skip 8 --launch-count 8 -o /tmp/attn_full python tools/ncu_probe.py > gpurun_out/ncu_attn.log 2>&1; echo rc=$?
ncu -i /tmp/attn_full.ncu-rep --page raw --csv > gpurun_out/r01_ncu_attn_full_raw.csv
ncu -i /tmp/attn_full.ncu-rep --page source --csv -k regex:attn_bwd_small > gpurun_out/r01_ncu_attn_small_source.csv 2>/dev/null
S=$(stat -c %s /tmp/attn_full.ncu-rep); echo size=$S; if [ "$S" -lt 30000000 ]; then cp /tmp/attn_full.ncu-rep gpurun_out/r01_attn_full.ncu-rep; fi
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r01_launches_c2_mb512.csv python bench.py --steps 1 --warmup 3 --global-batch 512 --microbatch 512 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo rc=$?; wc -l gpurun_out/r01_launches_c2_mb512.csv
