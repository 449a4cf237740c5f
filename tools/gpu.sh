#!/bin/bash
# build in-tree, then run a command on the GPU box:  tools/gpu.sh <timeout_s> '<command>'
set -e
cd "$(dirname "$0")/.."
python -m micro_diffusion_b200.build >/dev/null
T=$1; shift
exec /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
