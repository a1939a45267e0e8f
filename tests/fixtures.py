"""Shared test fixtures: a tiny tokenizer trained on the fly and tiny HF checkpoints written with our own saver."""
import json
import os
import random

import torch


def make_tokenizer(save_dir: str, vocab_size: int = 128):
    from tokenizers import Tokenizer, models, pre_tokenizers, trainers
    from transformers import PreTrainedTokenizerFast
    rng = random.Random(0)
    words = ["".join(rng.choice("abcdefghij") for _ in range(rng.randint(2, 5))) for _ in range(60)]
    corpus = [" ".join(rng.choice(words) for _ in range(12)) for _ in range(200)]
    tok = Tokenizer(models.WordLevel(unk_token="[UNK]"))
    tok.pre_tokenizer = pre_tokenizers.Whitespace()
    tok.train_from_iterator(corpus, trainers.WordLevelTrainer(vocab_size=vocab_size, special_tokens=["[PAD]", "[EOS]", "[UNK]"]))
    fast = PreTrainedTokenizerFast(tokenizer_object=tok, pad_token="[PAD]", eos_token="[EOS]", unk_token="[UNK]")
    fast.save_pretrained(save_dir)
    return fast, words


def make_checkpoint(save_dir: str, family: str = "gpt2", is_critic: bool = False, seed: int = 1):
    from realhf_b200.models import hf_io
    from realhf_b200.models.real_model import ReaLModel
    cfg = hf_io.family(family).make_test_config()
    if is_critic:
        cfg.is_critic, cfg.tied_embedding = True, False
    m = ReaLModel(cfg, dtype=torch.float32).instantiate(seed=seed)
    tok, words = make_tokenizer(save_dir, vocab_size=cfg.vocab_size)
    hf_io.save_to_hf(m, family, save_dir, tokenizer=tok)
    return cfg, tok, words


def write_sft_dataset(path: str, words, n: int = 64, seed: int = 0):
    rng = random.Random(seed)
    with open(path, "w") as f:
        for i in range(n):
            p = " ".join(rng.choice(words) for _ in range(rng.randint(3, 8)))
            a = " ".join(rng.choice(words[:5]) for _ in range(rng.randint(4, 10)))
            f.write(json.dumps(dict(id=i, prompt=p + " ", answer=a)) + "\n")


def write_prompt_dataset(path: str, words, n: int = 64, seed: int = 0):
    rng = random.Random(seed)
    with open(path, "w") as f:
        for i in range(n):
            f.write(json.dumps(dict(id=i, prompt=" ".join(rng.choice(words) for _ in range(rng.randint(3, 8))))) + "\n")


def write_pair_dataset(path: str, words, n: int = 32, seed: int = 0):
    rng = random.Random(seed)
    with open(path, "w") as f:
        for i in range(n):
            mk = lambda: " ".join(rng.choice(words) for _ in range(rng.randint(3, 8)))
            f.write(json.dumps(dict(id=i, prompt=mk() + " ", pos_answers=[mk(), mk()], neg_answers=[mk(), mk()])) + "\n")
