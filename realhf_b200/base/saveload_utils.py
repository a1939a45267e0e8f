"""Checkpoint file helpers shared by the HF reader / writer and the tools.  Parity: `realhf/base/saveload_utils.py`
(split_state_dict_into_shards, copy_hf_configs, load_safetensor)."""

from __future__ import annotations

import os
import shutil
from typing import Dict, List

import torch

# files of a HuggingFace model directory that are not weights: configs, tokenizer, generation defaults, custom code
HF_AUX_PATTERNS = ("config.json", "generation_config.json", "tokenizer", "vocab", "merges.txt", "special_tokens_map.json",
                   "added_tokens.json", "spiece.model", "sentencepiece", "real_model_config.json")


def split_state_dict_into_shards(sd: Dict[str, torch.Tensor], max_bytes: int) -> List[Dict[str, torch.Tensor]]:
    """Cut a state dict into consecutive (key-sorted) groups of at most `max_bytes` each (a single larger tensor gets its own
    group): the files of a sharded checkpoint."""
    files, cur, size = [], {}, 0
    for k in sorted(sd):
        n = sd[k].numel() * sd[k].element_size()
        if cur and size + n > max_bytes:
            files.append(cur)
            cur, size = {}, 0
        cur[k] = sd[k]
        size += n
    if cur:
        files.append(cur)
    return files


def load_weight_file(fn: str) -> Dict[str, torch.Tensor]:
    """One checkpoint file, `.safetensors` or a pickled `.bin` (read with `weights_only=True`), on the CPU."""
    if fn.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(fn)
    return torch.load(fn, map_location="cpu", weights_only=True)


load_safetensor = load_weight_file


def copy_hf_configs(src_model_dir: str, dst_model_dir: str) -> List[str]:
    """Copy everything of a HF model directory that is NOT a weight file (config, tokenizer files, generation config, custom
    modelling code) so that a directory of freshly written weights becomes loadable by `transformers`.  Returns the copied names."""
    os.makedirs(dst_model_dir, exist_ok=True)
    copied = []
    for fn in sorted(os.listdir(src_model_dir)):
        src = os.path.join(src_model_dir, fn)
        if not os.path.isfile(src):
            continue
        is_weight = fn.endswith((".safetensors", ".bin", ".pt", ".pth")) or fn.endswith(".index.json")
        if is_weight:
            continue
        if fn.endswith((".json", ".txt", ".model", ".py", ".tiktoken")) or any(fn.startswith(p) for p in HF_AUX_PATTERNS):
            shutil.copy2(src, os.path.join(dst_model_dir, fn))
            copied.append(fn)
    return copied
