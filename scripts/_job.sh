set -u
mkdir -p gpurun_out
echo "== tp decode"; timeout 600 python -m pytest tests/test_nvls_gpu.py -x -q -k "tp_decode" --tb=short 2>&1 | grep -v "W921\|NCCL version" | tail -30
echo "== fused tp collectives"; timeout 600 python -m pytest tests/test_comm_gpu.py -x -q -k "fused_tp_gemm_collectives" --tb=short 2>&1 | grep -v "W921\|NCCL version" | tail -40
echo "== pp graph"; timeout 600 python -m pytest tests/test_parallel_gpu.py -x -q -k "pipelined_generation" --tb=short 2>&1 | grep -v "W921\|NCCL version" | tail -40
