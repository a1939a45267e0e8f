set -u
mkdir -p gpurun_out
echo "== pytest gpu (1 GPU)"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
echo "== train profile"; timeout 300 python scripts/profile_train.py 4 > gpurun_out/train_profile_r2b.txt 2>&1; head -32 gpurun_out/train_profile_r2b.txt | cut -c1-150
echo "== decode timeline"; timeout 200 python scripts/profile_decode_timeline.py 4 16 384 > gpurun_out/decode_timeline_b16_r2b.jsonl 2>/dev/null; head -1 gpurun_out/decode_timeline_b16_r2b.jsonl
echo "== bench n1"; timeout 900 python bench.py --gpus 1 --steps 2 --warmup 3 --verbose 2> gpurun_out/bench_n1.err | tee gpurun_out/bench_n1_r2b.json; tail -4 gpurun_out/bench_n1.err
