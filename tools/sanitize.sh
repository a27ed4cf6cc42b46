#!/bin/bash
# compute-sanitizer over the smoke shapes of every hand-synchronised kernel (run on a GPU box, e.g. gpurun -- 'bash tools/sanitize.sh').
# memcheck: out-of-bounds / misaligned accesses incl. shared memory and TMA destinations; racecheck: shared-memory hazards between
# the gather warps, the epilogue and the A/B panel writers.  Logs: profiles/r02_sanitizer_{memcheck,racecheck}.txt
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 330 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitize_smoke.py > gpurun_out/r02_sanitizer_$tool.txt 2>&1
  echo "[$tool] rc=$?" >> gpurun_out/r02_sanitizer_$tool.txt
  tail -4 gpurun_out/r02_sanitizer_$tool.txt
done
