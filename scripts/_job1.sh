set -u
mkdir -p gpurun_out
timeout 100 python scripts/diag_fp8_decode.py 2>&1 | grep "^{" | tee gpurun_out/diag_fp8_decode.jsonl | cut -c1-300
