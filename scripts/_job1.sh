set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
echo "== ncu attn fwd"; timeout 300 $NCU -k regex:attn_fwd_kernel -c 1 -f -o gpurun_out/attn_fwd python scripts/ncu_attn.py fwd > /dev/null 2>&1; ls -la gpurun_out/attn_fwd.ncu-rep 2>&1 | cut -c1-120
echo "== ncu attn bwd"; timeout 300 $NCU -k regex:attn_bwd_kernel -c 2 -f -o gpurun_out/attn_bwd python scripts/ncu_attn.py bwd > /dev/null 2>&1; ls -la gpurun_out/attn_bwd.ncu-rep 2>&1 | cut -c1-120
echo "== ncu glu gemm"; timeout 300 $NCU -k regex:gemm_2cta_kernel -c 1 -f -o gpurun_out/gemm_glu python scripts/ncu_glu.py > /dev/null 2>&1; ls -la gpurun_out/gemm_glu.ncu-rep 2>&1 | cut -c1-120
echo "== bench n1 (gen phases)"; REAL_GEN_TIMING=1 timeout 900 python bench.py --gpus 1 --steps 2 --warmup 2 --verbose 2> gpurun_out/bench_n1c.err | tee gpurun_out/bench_n1_r2c.json | cut -c1-160; grep "warmup" gpurun_out/bench_n1c.err | cut -c1-400
echo "== bench n1 master runtime"; timeout 1200 python bench.py --runtime master --gpus 1 --steps 2 --warmup 2 2> gpurun_out/bench_master.err | grep "^{" | tee gpurun_out/bench_n1_master_runtime.json | cut -c1-900; tail -5 gpurun_out/bench_master.err | cut -c1-300
