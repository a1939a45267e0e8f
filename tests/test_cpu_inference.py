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


@pytest.mark.parametrize("fam", ["llama", "gpt2", "qwen2", "gemma", "mistral", "mixtral"])
@pytest.mark.parametrize("critic", [False, True])
def test_saved_model_config_fast_path_equals_the_hf_config_path(fam, critic, tmp_path, monkeypatch):
    """Checkpoints written by this framework carry `real_model_config.json`; workers read it instead of instantiating the HF config
    class.  It must give exactly the config the HF route gives, fall back for a critic initialised from an actor checkpoint and for
    foreign / older files, and the generic fast-tokenizer route must return the class and ids AutoTokenizer returns."""
    import dataclasses
    import json
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import fixtures
    from realhf_b200.api import data as data_api
    from realhf_b200.models import hf_io
    d = str(tmp_path / "ckpt")
    cfg, tok, words = fixtures.make_checkpoint(d, fam, is_critic=critic)
    calls = []
    real = hf_io.load_hf_config
    monkeypatch.setattr(hf_io, "load_hf_config", lambda p: (calls.append(p), real(p))[1])
    fast = hf_io.config_from_hf_path(fam, d, is_critic=critic)
    assert not calls, "the saved config was not used"
    monkeypatch.setenv("REAL_TRUST_SAVED_MODEL_CONFIG", "0")
    slow = hf_io.config_from_hf_path(fam, d, is_critic=critic)
    assert len(calls) == 1
    assert dataclasses.asdict(fast) == dataclasses.asdict(slow) == dataclasses.asdict(cfg)
    monkeypatch.setenv("REAL_TRUST_SAVED_MODEL_CONFIG", "1")
    if not critic:   # critic initialised from an actor checkpoint: head type differs from the saved one -> HF route
        c2 = hf_io.config_from_hf_path(fam, d, is_critic=True)
        assert len(calls) == 2 and c2.is_critic
    fn = os.path.join(d, "real_model_config.json")
    saved = json.load(open(fn))
    json.dump({k: v for k, v in saved.items() if k != "_family"}, open(fn, "w"))       # a file written before the family tag existed
    hf_io.config_from_hf_path(fam, d, is_critic=critic)
    assert len(calls) == (3 if not critic else 2)
    t1 = data_api.load_hf_tokenizer(d)
    import transformers
    t2 = transformers.AutoTokenizer.from_pretrained(d)
    text = " ".join(words[:7])
    assert type(t1) is type(t2) and t1(text).input_ids == t2(text).input_ids and t1.pad_token_id == t2.pad_token_id
