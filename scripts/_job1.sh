set -u
mkdir -p gpurun_out
echo "== pytest gpu (1 GPU)"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for t in 1 6 10; do echo "== decode timeline B=16 ctas_per_sm=$t"; REAL_DECODE_CTAS_PER_SM=$t timeout 200 python scripts/profile_decode_timeline.py 4 16 384 > gpurun_out/decode_timeline_b16_split$t.jsonl 2>/dev/null; head -1 gpurun_out/decode_timeline_b16_split$t.jsonl; done
echo "== bench n1 with N=8-like per-GPU shapes (16 prompts)"; timeout 900 python bench.py --gpus 1 --prompts 16 --steps 2 --warmup 2 --offload-frozen off --verbose --profile-mfc actor_train,critic_train 2> gpurun_out/bench_p16.err | cut -c1-300; grep -v "^W09" gpurun_out/bench_p16.err | cut -c1-200 | tail -80
