#!/bin/bash
# compute-sanitizer over the kernels added this round (lean ring / register forms, TMA-staged dense)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for tool in memcheck racecheck; do
echo "=== $tool: smoke (small-grid lean ring kernels, routing)"
timeout 600 compute-sanitizer --tool $tool --print-limit 5 python __graft_entry__.py smoke 2>&1 | grep -E "smoke ok|ERROR SUMMARY|Error|error|hazard" | head -8
echo "=== $tool: dense K1d/K2d (B=2500, T=9)"
timeout 900 compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_dense_gpu.py -q -x -k "is_taken" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Error|hazard" | head -8
echo "=== $tool: lean K1s/K2s large-grid forms (B=2501)"
timeout 900 compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_lean_gpu.py -q -x -k "forward_only or other_cotangents" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Error|hazard" | head -8
done
