"""Datasets: `prompt`, `prompt_answer`, `rw_pair`.

Parity: `realhf/impl/dataset/{prompt_dataset,prompt_answer_dataset,rw_paired_dataset}.py`.  Each reads
JSON / JSONL (or takes a `dataset_builder` callable, which the tests use), tokenises once at construction and
yields one-item `SequenceSample`s that `PackedDataLoader` gathers into packed batches.
"""

from __future__ import annotations

import itertools
from typing import Callable, Dict, List, Optional

import torch
import torch.utils.data

from realhf_b200.api.data import DatasetUtility, SequenceSample, load_shuffle_split_dataset, register_dataset


class PromptDataset(torch.utils.data.Dataset):
    """{"prompt": str, "id": ...} -> key `packed_prompts` (left-truncated to max_length)."""

    def __init__(self, util: DatasetUtility, max_length: Optional[int] = None, dataset_path: Optional[str] = None,
                 dataset_builder: Optional[Callable[[], List[Dict]]] = None, pad_to_max_length: bool = False):
        self.util = util
        data = load_shuffle_split_dataset(util, dataset_path, dataset_builder)
        tok = util.tokenizer
        enc = tok([x["prompt"] for x in data], truncation=True, max_length=max_length, padding=False,
                  return_length=True, return_attention_mask=False)
        self.ids = [x["id"] for x in data]
        self.prompts = enc["input_ids"]
        if pad_to_max_length:  # fixed shapes for benchmarking (reference: api/quickstart/dataset.py:88-92)
            assert max_length is not None
            self.prompts = [([tok.pad_token_id] * (max_length - len(p)) + p)[-max_length:] for p in self.prompts]
        self.lengths = [len(p) for p in self.prompts]

    def __len__(self):
        return len(self.prompts)

    def __getitem__(self, i):
        return SequenceSample.from_default(ids=[self.ids[i]], seqlens=[self.lengths[i]],
                                           data=dict(packed_prompts=torch.tensor(self.prompts[i], dtype=torch.long)))


class PromptAnswerDataset(torch.utils.data.Dataset):
    """{"prompt", "answer"} -> `packed_input_ids` (prompt + answer + EOS) and `prompt_mask` (True on prompt tokens)."""

    def __init__(self, util: DatasetUtility, max_length: int, dataset_path: Optional[str] = None,
                 dataset_builder: Optional[Callable[[], List[Dict]]] = None, pad_to_max_length: bool = False):
        self.util = util
        data = load_shuffle_split_dataset(util, dataset_path, dataset_builder)
        tok = util.tokenizer
        seqs = [x["prompt"] + x["answer"] + tok.eos_token for x in data]
        self.ids = [x["id"] for x in data]
        enc = tok(seqs, truncation=True, max_length=max_length, padding=False, return_length=True, return_attention_mask=False)
        penc = tok([x["prompt"] for x in data], truncation=True, max_length=max_length, padding=False,
                   return_length=True, return_attention_mask=False)
        self.tokens = enc["input_ids"]
        self.prompt_lens = [min(len(p), len(t)) for p, t in zip(penc["input_ids"], self.tokens)]
        if pad_to_max_length:
            for j, t in enumerate(self.tokens):
                self.tokens[j] = t + [tok.eos_token_id] * (max_length - len(t))

    def __len__(self):
        return len(self.tokens)

    def __getitem__(self, i):
        t = torch.tensor(self.tokens[i], dtype=torch.long)
        pm = torch.zeros(len(t), dtype=torch.bool)
        pm[: self.prompt_lens[i]] = True
        return SequenceSample.from_default(ids=[self.ids[i]], seqlens=[len(t)], data=dict(packed_input_ids=t, prompt_mask=pm))


class RewardModelingPairedDataset(torch.utils.data.Dataset):
    """{"prompt", "pos_answers": [...], "neg_answers": [...]} -> one item = up to `max_pairs_per_prompt` (pos, neg)
    pairs laid out [pos0, neg0, pos1, neg1, ...] under key `packed_input_ids` (+ `prompt_mask` for DPO)."""

    def __init__(self, util: DatasetUtility, max_length: int, max_pairs_per_prompt: int = 2, dataset_path: Optional[str] = None,
                 dataset_builder: Optional[Callable[[], List[Dict]]] = None):
        self.util = util
        data = load_shuffle_split_dataset(util, dataset_path, dataset_builder)
        tok = util.tokenizer
        self.ids, self.items, self.prompt_lens = [], [], []
        rng = __import__("random").Random(util.seed)
        for x in data:
            n = min(len(x["pos_answers"]), len(x["neg_answers"]), max_pairs_per_prompt)
            if n == 0:
                continue
            pairs = list(zip(x["pos_answers"], x["neg_answers"]))
            rng.shuffle(pairs)
            texts = list(itertools.chain.from_iterable((x["prompt"] + p + tok.eos_token, x["prompt"] + q + tok.eos_token)
                                                       for p, q in pairs[:n]))
            enc = tok(texts, truncation=True, max_length=max_length, padding=False, return_attention_mask=False)["input_ids"]
            plen = len(tok(x["prompt"], truncation=True, max_length=max_length, padding=False)["input_ids"])
            self.ids.append(x["id"])
            self.items.append(enc)
            self.prompt_lens.append(plen)

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        seqs = self.items[i]
        lens = [len(s) for s in seqs]
        ids = torch.tensor(list(itertools.chain.from_iterable(seqs)), dtype=torch.long)
        pm = torch.zeros(len(ids), dtype=torch.bool)
        off = 0
        for l in lens:
            pm[off: off + min(self.prompt_lens[i], l)] = True
            off += l
        return SequenceSample(keys=["packed_input_ids", "prompt_mask"], ids=[self.ids[i]],
                              seqlens=dict(packed_input_ids=[lens], prompt_mask=[lens]),
                              trailing_shapes=dict(packed_input_ids=(), prompt_mask=()),
                              dtypes=dict(packed_input_ids=torch.long, prompt_mask=torch.bool),
                              data=dict(packed_input_ids=ids, prompt_mask=pm))


register_dataset("prompt", PromptDataset)
register_dataset("prompt_answer", PromptAnswerDataset)
register_dataset("rw_pair", RewardModelingPairedDataset)
