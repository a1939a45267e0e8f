set -u
mkdir -p gpurun_out
echo "== fp8 tests"; timeout 110 python -m pytest tests/test_fp8_gpu.py -q --tb=short -k "generation or emulation" 2>&1 | grep -v "W921\|NCCL version" | tail -25 | cut -c1-500
