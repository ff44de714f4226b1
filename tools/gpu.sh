#!/bin/bash
# usage: tools/gpu.sh <timeout-seconds> '<command>'   (rebuilds libwbk.so first so the snapshot is never stale)
set -e
cd "$(dirname "$0")/.."
python -m wavebreaking_b200._build >/dev/null
T=$1; shift
exec /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
