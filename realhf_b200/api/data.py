"""Packed variable-length batches and dataset plumbing.

`SequenceSample` has the reference's surface (`realhf/api/core/data_api.py:95-596`): per-key nested
`seqlens`, 1-D packed `data`, `gather / split / unpack / meta / update_ / from_default /
remap_keys_`.  It is a plain slotted class with an explicit `validate()` instead of a pydantic
model, so constructing one on the hot path costs nothing (the reference has to monkey-patch
`__init__` to switch validation off).
"""

from __future__ import annotations

import dataclasses
import json
import os
import random
from contextlib import contextmanager
from typing import Any, Callable, Dict, Hashable, Iterable, List, Optional, Sequence, Set, Tuple, Union

import numpy as np
import torch
import torch.utils.data

from realhf_b200.api import config as config_api
from realhf_b200.base import datapack

# ---------------------------------------------------------------------------------------------- SequenceSample


@dataclasses.dataclass
class SequenceSplitSpec:
    partitions: Optional[List[Tuple[int, int]]] = None
    sizes: Optional[List[int]] = None

    def __post_init__(self):
        if (self.partitions is None) == (self.sizes is None):
            raise ValueError("give exactly one of `partitions` / `sizes`")
        if self.partitions is None:
            off, parts = 0, []
            for s in self.sizes:
                parts.append((off, off + s))
                off += s
            self.partitions = parts
        else:
            self.sizes = [e - s for s, e in self.partitions]


_VALIDATE = True

# key -> per-item sequence length rule used by `from_default` (L = length of the main sequence)
_LEN_ONE = {"seq_no_eos_mask", "greedy_seq_no_eos_mask", "loss_mask", "rewards", "greedy_rewards", "scores",
            "group_factor", "pos_input_lens"}
_LEN_FULL = {"input_ids", "packed_seq", "seq", "prompt_mask",
             "greedy_prompt_mask", "packed_input_ids", "greedy_packed_input_ids", "values", "packed_prompts"}
_LEN_MINUS1 = {"packed_logprobs", "logprobs", "packed_ref_logprobs", "ref_logprobs", "old_logp", "ref_logp",
               "advantages", "ppo_loss_mask", "kl_rewards", "returns",
               # bit-packed [L-1, V/8] here (aligned with the log-probs); the reference ships a bool [L, V]
               "packed_logits_mask", "logits_mask"}


class SequenceSample:
    """A batch of data items; every key holds >=1 variable-length sequences per item, packed 1-D."""

    __slots__ = ("keys", "trailing_shapes", "dtypes", "ids", "seqlens", "data", "metadata")

    def __init__(self, keys: Iterable[str], trailing_shapes: Dict[str, Any], dtypes: Dict[str, Any],
                 ids: List[Hashable], seqlens: Dict[str, List[List[int]]],
                 data: Optional[Dict[str, Optional[torch.Tensor]]] = None,
                 metadata: Optional[Dict[str, List[Any]]] = None):
        self.keys: Set[str] = set(keys)
        self.trailing_shapes = trailing_shapes
        self.dtypes = dtypes
        self.ids = ids
        self.seqlens = seqlens
        self.data = data
        self.metadata = metadata if metadata is not None else {}
        if _VALIDATE:
            self.validate()

    # ---- validation
    def validate(self):
        if len(self.ids) != len(set(self.ids)):
            raise ValueError(f"ids contain duplicates: {self.ids}")
        n = len(self.ids)
        for k in (self.seqlens, self.trailing_shapes, self.dtypes):
            if set(k.keys()) != self.keys:
                raise KeyError(f"keys mismatch: {self.keys} vs {set(k.keys())}")
        if self.data is not None and set(self.data.keys()) != self.keys:
            raise KeyError(f"data keys {set(self.data.keys())} != {self.keys}")
        for k, lens in self.seqlens.items():
            if len(lens) != n:
                raise ValueError(f"seqlens[{k}] has {len(lens)} items, expected {n}")
            for l in lens:
                if not isinstance(l, list) or not all(isinstance(x, int) for x in l):
                    raise TypeError(f"seqlens[{k}] must be List[List[int]]")
        for k, v in self.metadata.items():
            if not isinstance(v, list) or len(v) != n:
                raise ValueError(f"metadata[{k}] must be a list of length {n}")
        if self.data is not None:
            for k, v in self.data.items():
                if v is None:
                    continue
                total = sum(sum(l) for l in self.seqlens[k])
                exp = (total, *tuple(self.trailing_shapes[k] or ()))
                if tuple(v.shape) != exp:
                    raise ValueError(f"key {k}: data shape {tuple(v.shape)} != {exp}")
                if v.dtype != self.dtypes[k]:
                    raise ValueError(f"key {k}: dtype {v.dtype} != {self.dtypes[k]}")

    @classmethod
    @contextmanager
    def disable_validation(cls):
        global _VALIDATE
        old, _VALIDATE = _VALIDATE, False
        try:
            yield
        finally:
            _VALIDATE = old

    # ---- basic
    @property
    def bs(self) -> int:
        return len(self.ids)

    def __repr__(self):
        return f"SequenceSample(bs={self.bs}, keys={sorted(self.keys)}, has_data={self.data is not None})"

    def __getstate__(self):
        return {s: getattr(self, s) for s in self.__slots__}

    def __setstate__(self, st):
        for k, v in st.items():
            setattr(self, k, v)

    def total_len(self, key: str) -> int:
        return sum(sum(l) for l in self.seqlens[key])

    # ---- gather / split
    @classmethod
    def gather(cls, samples: List["SequenceSample"], keys: Optional[Iterable[str]] = None) -> "SequenceSample":
        keys = set(samples[0].keys if keys is None else keys)
        seqlens = {k: [l for s in samples for l in s.seqlens[k]] for k in keys}
        data = None
        if samples[0].data is not None:
            data = {k: (torch.cat([s.data[k] for s in samples], dim=0) if samples[0].data[k] is not None else None)
                    for k in keys}
        ids = [i for s in samples for i in s.ids]
        metadata = {k: [x for s in samples for x in s.metadata[k]] for k in samples[0].metadata}
        with cls.disable_validation():
            return cls(keys=keys, dtypes={k: samples[0].dtypes[k] for k in keys},
                       trailing_shapes={k: samples[0].trailing_shapes[k] for k in keys}, ids=ids, seqlens=seqlens,
                       data=data, metadata=metadata)

    def _get_split_key(self) -> str:
        return max(sorted(self.keys), key=lambda k: self.total_len(k))

    def get_split_spec(self, k: int, key: Optional[str] = None, min_size: int = 1) -> SequenceSplitSpec:
        """Token-balanced contiguous split into k parts of at least `min_size` sequences.  `min_size` is a preference (room
        for micro-batches / pipeline stages downstream): when the batch cannot honour it, it is relaxed to what k equal parts
        allow -- the engines cope with fewer sequences than micro-batches, an exception here would only kill the run."""
        key = key or self._get_split_key()
        lens = [sum(l) for l in self.seqlens[key]]
        min_size = max(1, min(min_size, len(lens) // max(k, 1)))
        return SequenceSplitSpec(partitions=datapack.min_abs_diff_partition(lens, k, min_size))

    def split_with_spec(self, spec: SequenceSplitSpec) -> List["SequenceSample"]:
        out = []
        off = {k: 0 for k in self.keys}
        for s, e in spec.partitions:
            new_lens = {k: v[s:e] for k, v in self.seqlens.items()}
            n_tok = {k: sum(sum(l) for l in v) for k, v in new_lens.items()}
            new_data = None
            if self.data is not None:
                new_data = {k: (v[off[k]: off[k] + n_tok[k]] if v is not None else None) for k, v in self.data.items()}
            for k in self.keys:
                off[k] += n_tok[k]
            with self.disable_validation():
                out.append(SequenceSample(keys=self.keys, dtypes=self.dtypes, trailing_shapes=self.trailing_shapes,
                                          ids=self.ids[s:e], seqlens=new_lens, data=new_data,
                                          metadata={k: v[s:e] for k, v in self.metadata.items()}))
        return out

    def split(self, k: int, key: Optional[str] = None, min_size: int = 1) -> List["SequenceSample"]:
        return self.split_with_spec(self.get_split_spec(k, key, min_size))

    def unpack(self) -> List["SequenceSample"]:
        return self.split_with_spec(SequenceSplitSpec(partitions=[(i, i + 1) for i in range(self.bs)]))

    def select(self, indices: Sequence[int]) -> "SequenceSample":
        """Items at arbitrary positions (used for shuffled PPO minibatches)."""
        items = self.unpack()
        return SequenceSample.gather([items[i] for i in indices])

    # ---- device
    def to_device(self, device) -> "SequenceSample":
        if self.data is not None:
            self.data = {k: (v.to(device, non_blocking=True) if v is not None else None) for k, v in self.data.items()}
        return self

    def cuda(self):
        return self.to_device("cuda")

    # ---- metadata-only view / update
    def meta(self) -> "SequenceSample":
        with self.disable_validation():
            return SequenceSample(keys=self.keys, trailing_shapes=self.trailing_shapes, dtypes=self.dtypes,
                                  ids=self.ids, seqlens=self.seqlens, data=None, metadata=self.metadata)

    def update_(self, other: "SequenceSample"):
        assert self.ids == other.ids, (self.ids, other.ids)
        self.keys = self.keys | other.keys
        self.trailing_shapes.update(other.trailing_shapes)
        self.dtypes.update(other.dtypes)
        self.seqlens.update(other.seqlens)
        if self.data is not None and other.data is not None:
            self.data.update(other.data)
        self.metadata.update(other.metadata)

    def remap_keys_(self, remap: Dict[str, str]):
        for k in list(self.keys):
            if k in remap:
                nk = remap[k]
                self.seqlens[nk] = self.seqlens.pop(k)
                self.trailing_shapes[nk] = self.trailing_shapes.pop(k)
                self.dtypes[nk] = self.dtypes.pop(k)
                if self.data is not None:
                    self.data[nk] = self.data.pop(k)
        self.keys = {remap.get(k, k) for k in self.keys}

    # ---- constructors
    @staticmethod
    def _resolve_seqlen_from_key(key: str, seqlens: List[int]) -> List[List[int]]:
        if key in _LEN_ONE:
            return [[1] for _ in seqlens]
        if key in _LEN_FULL:
            return [[l] for l in seqlens]
        if key in _LEN_MINUS1:
            return [[l - 1] for l in seqlens]
        raise NotImplementedError(f"no default length rule for key `{key}`; build the SequenceSample explicitly")

    @classmethod
    def from_default(cls, seqlens: List[int], ids: List[Hashable], data: Dict[str, torch.Tensor],
                     metadata: Optional[Dict[str, Any]] = None) -> "SequenceSample":
        """One sequence per item; per-key lengths follow the rule table above."""
        if seqlens and isinstance(seqlens[0], list):
            assert all(len(l) == 1 for l in seqlens)
            seqlens = [l[0] for l in seqlens]
        seqlens = [int(l) for l in seqlens]
        keys = set(data.keys())
        return cls(keys=keys, ids=list(ids), seqlens={k: cls._resolve_seqlen_from_key(k, seqlens) for k in keys},
                   trailing_shapes={k: (tuple(v.shape[1:]) if v is not None else None) for k, v in data.items()},
                   dtypes={k: (v.dtype if v is not None else None) for k, v in data.items()}, data=data,
                   metadata=metadata or {})

    # ---- helpers used by engines
    def flat_seqlens(self, key: str) -> List[int]:
        return [x for l in self.seqlens[key] for x in l]

    def cu_seqlens(self, key: str, device=None) -> torch.Tensor:
        lens = self.flat_seqlens(key)
        cu = torch.zeros(len(lens) + 1, dtype=torch.int32)
        cu[1:] = torch.tensor(lens, dtype=torch.int32).cumsum(0)
        return cu.to(device) if device is not None else cu


@dataclasses.dataclass
class DataBatchMeta:
    dp_rank: int
    meta_sample: Optional[SequenceSample]
    epoch: int
    is_final_batch: bool


# ---------------------------------------------------------------------------------------------- datasets


@dataclasses.dataclass
class DatasetUtility:
    seed: int
    dp_rank: int
    world_size: int
    tokenizer: Any

    def __post_init__(self):
        if self.tokenizer is not None and getattr(self.tokenizer, "pad_token_id", None) is None:
            if getattr(self.tokenizer, "eos_token_id", None) is None:
                raise ValueError("the tokenizer needs an eos token")
            self.tokenizer.pad_token_id = self.tokenizer.eos_token_id


def load_shuffle_split_dataset(util: DatasetUtility, dataset_path: Optional[str],
                               dataset_builder: Optional[Callable[[], List[Dict]]] = None) -> List[Dict]:
    """Read json/jsonl (or call the builder), add ids, shuffle with the seed, take this dp rank's slice."""
    if dataset_path is not None:
        if dataset_path.endswith(".jsonl"):
            with open(dataset_path) as f:
                data = [json.loads(line) for line in f if line.strip()]
        elif dataset_path.endswith(".json"):
            with open(dataset_path) as f:
                data = json.load(f)
        else:
            raise NotImplementedError(f"unknown dataset extension: {dataset_path}")
    else:
        assert dataset_builder is not None
        data = dataset_builder()
    for i, d in enumerate(data):
        d.setdefault("id", i)
    n = len(data)
    rng = np.random.RandomState(util.seed)
    perm = rng.permutation(n)
    bounds = np.linspace(0, n, util.world_size + 1).astype(int)
    sub = perm[bounds[util.dp_rank]: bounds[util.dp_rank + 1]]
    return [data[i] for i in sub]


ALL_DATASET_CLASSES: Dict[str, Callable] = {}
ALL_DATALOADER_CLASSES: Dict[str, Callable] = {}


def register_dataset(name: str, cls):
    if name in ALL_DATASET_CLASSES:
        raise KeyError(f"dataset `{name}` already registered")
    ALL_DATASET_CLASSES[name] = cls


def register_dataloader(name: str, fn):
    ALL_DATALOADER_CLASSES[name] = fn


def make_dataset(cfg: Union[str, config_api.DatasetAbstraction], seed: int, dp_rank: int, world_size: int,
                 tokenizer_or_path, experiment_name: str = "", trial_name: str = "", cache_root: Optional[str] = None):
    if isinstance(cfg, str):
        cfg = config_api.DatasetAbstraction(type_=cfg)
    tokenizer = load_hf_tokenizer(tokenizer_or_path) if isinstance(tokenizer_or_path, str) else tokenizer_or_path
    util = DatasetUtility(seed, dp_rank, world_size, tokenizer)
    if cache_root is None:
        return ALL_DATASET_CLASSES[cfg.type_](util=util, **cfg.args)
    # optional on-disk cache of the tokenised shard (parity: data_api.py:677-741): keyed by everything that determines it
    import hashlib
    import pickle
    tok_id = tokenizer_or_path if isinstance(tokenizer_or_path, str) else getattr(tokenizer, "name_or_path", type(tokenizer).__name__)
    src_stamp = []
    for v in cfg.args.values():
        if isinstance(v, str) and os.path.exists(v):
            st = os.stat(v)
            src_stamp.append((v, st.st_size, int(st.st_mtime)))
    key = hashlib.sha1(repr((cfg.type_, sorted(cfg.args.items(), key=str), seed, dp_rank, world_size, tok_id, src_stamp)).encode()).hexdigest()[:20]
    path = os.path.join(cache_root, "datasets", f"{cfg.type_}-{key}.pkl")
    if os.path.exists(path):
        try:
            with open(path, "rb") as f:
                ds = pickle.load(f)
            ds.util = util
            return ds
        except Exception:
            pass  # stale / incompatible cache entry: rebuild
    ds = ALL_DATASET_CLASSES[cfg.type_](util=util, **cfg.args)
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        tmp = path + f".tmp{os.getpid()}"
        saved_util, ds.util = getattr(ds, "util", None), None  # the tokenizer is not part of the cached payload
        with open(tmp, "wb") as f:
            pickle.dump(ds, f)
        ds.util = saved_util
        os.replace(tmp, path)
    except Exception:
        ds.util = util
    return ds


def make_dataloader(cfg: Union[str, config_api.DataLoaderAbstraction], dataset, **overrides) -> torch.utils.data.DataLoader:
    if isinstance(cfg, str):
        cfg = config_api.DataLoaderAbstraction(type_=cfg)
    return ALL_DATALOADER_CLASSES[cfg.type_](dataset, **{**cfg.args, **overrides})


def load_hf_tokenizer(path: str, fast: bool = True, padding_side: Optional[str] = None):
    import transformers
    kw = {"padding_side": padding_side} if padding_side else {}
    tok = None
    if fast and os.path.isfile(os.path.join(path, "tokenizer.json")):
        # a generic `tokenizers`-backed tokenizer (no model-specific python class): construct the class AutoTokenizer would pick,
        # without AutoTokenizer's detour through the model config classes (their first import costs a worker ~6 s)
        try:
            with open(os.path.join(path, "tokenizer_config.json")) as f:
                cls_name = json.load(f).get("tokenizer_class")
        except (OSError, ValueError):
            cls_name = None
        if cls_name in (None, "PreTrainedTokenizerFast", "TokenizersBackend"):
            tok = transformers.PreTrainedTokenizerFast.from_pretrained(path, **kw)
    if tok is None:
        tok = transformers.AutoTokenizer.from_pretrained(path, use_fast=fast, trust_remote_code=True, **kw)
    if tok.pad_token_id is None:
        tok.pad_token_id = tok.eos_token_id
    return tok


def PackedDataLoader(dataset, batch_size: int = 512, shuffle: bool = True, **kw):
    return torch.utils.data.DataLoader(dataset, batch_size=batch_size, shuffle=shuffle,
                                       collate_fn=SequenceSample.gather, **kw)


def PackedEvalDataLoader(dataset, batch_size: int = 128, **kw):
    return torch.utils.data.DataLoader(dataset, batch_size=batch_size, shuffle=False,
                                       collate_fn=SequenceSample.gather, **kw)


register_dataloader("packed", PackedDataLoader)
register_dataloader("packed_eval", PackedEvalDataLoader)
register_dataloader("iterable_dataset_loader", lambda ds, **kw: ds)
