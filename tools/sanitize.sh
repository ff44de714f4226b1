#!/bin/bash
# compute-sanitizer pass over the CUDA paths (run on a GPU box): writes gpurun_out/r2_compute_sanitizer.txt
out=gpurun_out/r2_compute_sanitizer.txt
mkdir -p gpurun_out
run() {  # tool, label, command...
  tool=$1; shift; label=$1; shift
  echo "--- $tool: $label" >> $out
  timeout 400 compute-sanitizer --tool $tool --print-limit 5 "$@" 2>&1 | grep -v "^=========     \|^$" | grep -E "COMPUTE-SANITIZER|SUMMARY|ERROR|Error|hazard|Hazard|Invalid|ntime|smoke|batch 1 |passed|failed|labels" | tail -8 >> $out
}
echo "compute-sanitizer on the round-2 code (B200), $(git rev-parse --short HEAD 2>/dev/null)" > $out
run memcheck "721x1440, 3 time steps, two batches (tools/ncu_kernels.py 3)" python tools/ncu_kernels.py 3
run memcheck "noisy 721x1440 workload, 4 time steps x 2 batches, arenas regrown (tools/stress_probe.py 4 1.5 0 2)" python tools/stress_probe.py 4 1.5 0 2
run memcheck "__graft_entry__.smoke()" python -c "import __graft_entry__ as g; g.smoke()"
run memcheck "tracking kernels + smoothing variants (pytest -m gpu: test_tracking, test_smooth_stream)" python -m pytest tests/test_tracking.py tests/test_smooth_stream.py -q -m gpu -x
run racecheck "721x1440, 2 time steps (tools/ncu_kernels.py 2)" python tools/ncu_kernels.py 2
run racecheck "__graft_entry__.smoke()" python -c "import __graft_entry__ as g; g.smoke()"
run synccheck "__graft_entry__.smoke()" python -c "import __graft_entry__ as g; g.smoke()"
cat $out
