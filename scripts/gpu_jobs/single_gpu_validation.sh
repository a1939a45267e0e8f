#!/bin/bash
# One-GPU validation of a tree before it is handed over: the GPU test-suite, the opt-in W8A8 decode comparison and smoke().
#     gpurun --timeout 900 -- 'bash scripts/gpu_jobs/single_gpu_validation.sh'
# Every step runs under its own timeout; results that should be kept are written under gpurun_out/ (copy them to profiles/).
set -u
mkdir -p gpurun_out
echo "== GPU tests"; timeout 700 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -v "W921\|NCCL version" | tail -15 | cut -c1-400
echo "== decode bf16 vs fp8 (7B shapes)"; timeout 300 python scripts/bench_decode_fp8.py 32 384 16 64 128 2>&1 | grep "^{" | tee gpurun_out/decode_fp8_vs_bf16.jsonl
echo "== W8A8 decode step, call-by-call against PyTorch"; timeout 100 python scripts/diag_fp8_decode.py 2>&1 | grep "^{" | tee gpurun_out/fp8_decode_in_situ_check.jsonl | tail -1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | cut -c1-300
