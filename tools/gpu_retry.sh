#!/bin/bash
# tools/gpu_retry.sh <timeout_s> '<command>' : like tools/gpu.sh but retries while the pod answers busy (exit 3)
cd "$(dirname "$0")/.."
python -m micro_diffusion_b200.build >/dev/null || exit 1
T=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
