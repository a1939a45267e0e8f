set -u
mkdir -p gpurun_out
echo "== changed GPU tests"; timeout 500 python -m pytest tests/test_fp8_gpu.py tests/test_moe_gpu.py tests/test_sampling_gpu.py tests/test_gemm_gpu.py tests/test_ops_gpu.py -q --tb=short 2>&1 | grep -v "W921\|NCCL version" | tail -15 | cut -c1-400
echo "== decode bf16 vs fp8 (7B shapes, 32 layers)"; timeout 300 python scripts/bench_decode_fp8.py 32 384 16 64 2>&1 | grep "^{" | tee gpurun_out/decode_fp8_vs_bf16.jsonl
echo "== decode bf16 vs fp8, 2 and 8 layers (drift vs depth)"; for L in 2 8; do timeout 200 python scripts/bench_decode_fp8.py $L 384 16 2>&1 | grep "^{" | tee -a gpurun_out/decode_fp8_vs_bf16.jsonl; done
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | cut -c1-300
