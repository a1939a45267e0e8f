"""Tiny random ReaLModel vs the HuggingFace implementation of the same family (CPU, fp32): logits must agree."""
import pytest
import torch

from realhf_b200.models import hf_io
from realhf_b200.models.real_model import ReaLModel

FAMILIES = ["llama", "gpt2", "qwen2", "gemma", "mistral", "mixtral"]


@pytest.mark.parametrize("fam", FAMILIES)
def test_logits_match_hf(fam):
    torch.manual_seed(0)
    spec = hf_io.family(fam)
    cfg = spec.make_test_config()
    model = ReaLModel(cfg, dtype=torch.float32).instantiate(seed=3)
    model.eval()
    hf = hf_io.to_hf_model(model, fam).eval()
    lens = [5, 17, 9]
    ids = torch.randint(0, cfg.vocab_size, (sum(lens),))
    cu = torch.tensor([0, 5, 22, 31], dtype=torch.int32)
    with torch.no_grad():
        out = model(input_ids=ids, cu_seqlens=cu, max_seqlen=max(lens)).logits
        off = 0
        for l in lens:
            ref = hf(input_ids=ids[off:off + l].unsqueeze(0)).logits[0]
            torch.testing.assert_close(out[off:off + l], ref, atol=2e-4, rtol=1e-3)
            off += l


@pytest.mark.parametrize("fam", ["llama", "gpt2", "gemma"])
def test_hf_roundtrip_save_load(fam, tmp_path):
    torch.manual_seed(0)
    cfg = hf_io.family(fam).make_test_config()
    model = ReaLModel(cfg, dtype=torch.float32).instantiate(seed=5)
    hf_io.save_to_hf(model, fam, str(tmp_path))
    import transformers
    hf = transformers.AutoModelForCausalLM.from_pretrained(str(tmp_path)).eval()
    m2 = hf_io.from_hf(fam, str(tmp_path), dtype=torch.float32)
    for k, v in model.state_dict().items():
        torch.testing.assert_close(m2.state_dict()[k], v)
    ids = torch.randint(0, cfg.vocab_size, (12,))
    cu = torch.tensor([0, 12], dtype=torch.int32)
    m2.eval()
    with torch.no_grad():
        torch.testing.assert_close(m2(input_ids=ids, cu_seqlens=cu).logits, hf(input_ids=ids[None]).logits[0], atol=2e-4, rtol=1e-3)


def test_critic_init_from_actor(tmp_path):
    cfg = hf_io.family("llama").make_test_config()
    actor = ReaLModel(cfg, dtype=torch.float32).instantiate(seed=5)
    hf_io.save_to_hf(actor, "llama", str(tmp_path))
    critic = hf_io.from_hf("llama", str(tmp_path), is_critic=True, init_critic_from_actor=True, dtype=torch.float32)
    assert critic.p[f"{cfg.n_layers + 1}.head.weight"].shape == (1, cfg.hidden_dim)
    ids = torch.randint(0, cfg.vocab_size, (7,))
    out = critic(input_ids=ids, cu_seqlens=torch.tensor([0, 7], dtype=torch.int32))
    assert out.values.shape == (7,)


@pytest.mark.parametrize("fam", ["llama", "gpt2", "qwen2", "mixtral"])
def test_greedy_generation_matches_hf_generate(fam):
    """Greedy decoding through the packed prefill + KV-cache decode loop produces the tokens that the HuggingFace model
    picks step by step (cache-free forward + argmax; parity: tests/model/test_generate.py, which accepts >= 0.8 agreement
    on real checkpoints -- tiny fp32 models agree exactly)."""
    from realhf_b200.api.model import GenerationHyperparameters
    from realhf_b200.models import generation as gen
    torch.manual_seed(0)
    cfg = hf_io.family(fam).make_test_config()
    model = ReaLModel(cfg, dtype=torch.float32).instantiate(seed=11)
    model.eval()
    hf = hf_io.to_hf_model(model, fam).eval()
    plens = [6, 9, 4]
    prompts = [torch.randint(3, cfg.vocab_size, (n,)) for n in plens]
    cu = torch.tensor([0] + list(torch.tensor(plens).cumsum(0)), dtype=torch.int32)
    g = GenerationHyperparameters(max_new_tokens=7, min_new_tokens=7, greedy=True)
    out, _ = gen.generate(model, torch.cat(prompts), cu, g, eos_id=None, pad_id=0)
    agree, total = 0, 0
    for i, p in enumerate(prompts):
        seq = p.clone()
        with torch.no_grad():
            for _ in range(7):
                seq = torch.cat([seq, hf(input_ids=seq.unsqueeze(0)).logits[0, -1].argmax().view(1)])
        ref = seq[p.numel():]
        agree += int((out.tokens[i, : ref.numel()] == ref).sum())
        total += ref.numel()
    assert agree == total, (agree, total)
