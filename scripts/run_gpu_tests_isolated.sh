#!/bin/bash
# Run every GPU test in its own process (a sticky CUDA error then cannot cascade) and summarise.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
names=$(python -m pytest tests -m gpu --collect-only -q 2>/dev/null | grep "::")
for t in $names; do
  out=$(timeout 300 python -m pytest "$t" -x -q 2>&1 | tail -25)
  if echo "$out" | grep -q " passed"; then echo "PASS $t"; else echo "FAIL $t"; echo "$out" | grep -E "Error|error|assert|mismatch|differ|row " | head -12; fi
done
