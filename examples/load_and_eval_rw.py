"""Load a reward-model (critic) checkpoint saved by a `rw` / `ppo` experiment in ONE process and score sequences with it.

    python examples/load_and_eval_rw.py --path <fileroot>/checkpoints/<user>/<exp>/<trial>/default/epoch1epochstep10globalstep10 \
        --family llama [--device cuda] [--text "some text to score" ...]

Critic checkpoints keep the HuggingFace layout but their output head is `[1, hidden]`, so `transformers` cannot load them as a
causal LM (same caveat as the reference, docs quickstart "reward modelling"); this script is the supported way to use one
outside a training run.  It works on CPU too (`--device cpu`, fp32).
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a source checkout

from realhf_b200.models import hf_io  # noqa: E402


def load_reward_model(path: str, family: str, device: str = "cuda", dtype=None):
    dtype = dtype or (torch.bfloat16 if device.startswith("cuda") else torch.float32)
    # is_critic=True: scalar head, untied from the embedding; the checkpoint already IS a critic, so nothing is re-initialised
    return hf_io.from_hf(family, path, is_critic=True, init_critic_from_actor=False, dtype=dtype, device=device).eval()


@torch.no_grad()
def score(model, sequences):
    """sequences: list of 1-D LongTensors.  Returns per-token values (list of tensors) and the score of every sequence
    (value at its last token, the convention of the paired reward-modelling loss)."""
    dev = next(model.parameters()).device
    lens = [int(s.numel()) for s in sequences]
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
    out = model(input_ids=torch.cat(sequences).to(dev), cu_seqlens=cu, max_seqlen=max(lens))
    values = out.values.float().cpu()
    per_seq = list(values.split(lens))
    return per_seq, torch.stack([v[-1] for v in per_seq])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--path", required=True)
    ap.add_argument("--family", default="llama")
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    ap.add_argument("--text", nargs="*", default=None)
    a = ap.parse_args()
    model = load_reward_model(a.path, a.family, a.device)
    if a.text:
        import transformers
        tok = transformers.AutoTokenizer.from_pretrained(a.path)
        seqs = [torch.tensor(tok(t)["input_ids"], dtype=torch.long) for t in a.text]
    else:  # no text: random token ids, just to show the shapes
        g = torch.Generator().manual_seed(0)
        seqs = [torch.randint(0, model.config.vocab_size, (n,), generator=g) for n in (17, 256, 64)]
    per_token, scores = score(model, seqs)
    for i, (v, s) in enumerate(zip(per_token, scores)):
        print(f"sequence {i}: {v.numel()} tokens, score {s.item():+.4f}")


if __name__ == "__main__":
    main()
