"""SFT, paired reward modelling, DPO and generation interfaces.

Parity: `interface/sft_interface.py` (:19-165), `rw_interface.py` (:25-155), `dpo_interface.py` (:75-219),
`gen_interface.py` (:17-153).
"""

from __future__ import annotations

import dataclasses
import fcntl
import functools
import json
import os
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from realhf_b200.api.data import SequenceSample
from realhf_b200.api.model import GenerationHyperparameters, Model, ModelInterface, register_interface
from realhf_b200.interfaces import functional as IF
from realhf_b200.interfaces.ppo import _dp_group, _engine_ctx, _mb_prompt, _save_hf
from realhf_b200.models import generation as gen
from realhf_b200.models.real_model import ModelOutput

# ------------------------------------------------------------------------------------------- SFT


def sft_loss_from_output(out: ModelOutput, mb: SequenceSample):
    """Masked next-token NLL: summed over answer tokens, divided by their count in this micro-batch."""
    seqlens = mb.flat_seqlens("packed_input_ids")
    ids = mb.data["packed_input_ids"]
    rows, labels = IF.shifted_rows_and_labels(seqlens, ids)
    logp = out.logprobs(labels, None, 1.0, rows)
    loss_mask = (~mb.data["prompt_mask"].bool()).index_select(0, rows + 1).float()
    n = loss_mask.sum().clamp(min=1)
    loss = -(logp * loss_mask).sum() / n
    with torch.no_grad():
        stat = dict(loss=loss.detach(), n_tokens=loss_mask.sum(), nll_sum=-(logp.detach() * loss_mask).sum())
    return loss, stat


@dataclasses.dataclass
class SFTInterface(ModelInterface):
    enable_save: bool = True

    def train_step(self, model: Model, data: SequenceSample, n_mbs=None) -> Dict:
        eng = model.module
        eng.train()
        st = eng.train_batch(data, sft_loss_from_output, version_steps=model.version.global_step, num_micro_batches=n_mbs)
        model.inc_version()
        dev = data.data["packed_input_ids"].device
        g = IF.dp_reduce_stats({"nll_sum": st["nll_sum"] * (n_mbs or 1), "n_tokens": st["n_tokens"] * (n_mbs or 1)},
                               _dp_group(model), dev)
        loss = g["nll_sum"] / max(g["n_tokens"], 1.0)
        return dict(loss=loss, ppl=float(torch.tensor(loss).exp()), n_tokens=int(g["n_tokens"]),
                    grad_norm=float(st["grad_norm"]), lr=float(st["lr"]))

    @torch.no_grad()
    def evaluate(self, model: Model, eval_dataloader) -> Dict:
        eng = model.module
        eng.eval()
        dev = model.device
        tot = {"nll_sum": torch.zeros((), device=dev), "n_tokens": torch.zeros((), device=dev)}
        for batch in eval_dataloader:
            batch = batch.to_device(dev)
            st = eng.eval_batch(batch, sft_loss_from_output)  # statistics exist on the last pipeline stage only
            tot["nll_sum"] += st.get("nll_sum", 0.0)
            tot["n_tokens"] += st.get("n_tokens", 0.0)
        g = IF.dp_reduce_stats(tot, _dp_group(model), dev)
        loss = g["nll_sum"] / max(g["n_tokens"], 1.0)
        return dict(loss=loss, ppl=float(torch.tensor(loss).exp()))

    def save(self, model: Model, save_dir: str):
        if self.enable_save:
            _save_hf(model, save_dir)

    def _mock_train_step(self, model: Model, data: SequenceSample):
        n = data.total_len("packed_input_ids")
        dev = data.data["packed_input_ids"].device
        data.update_(SequenceSample.from_default(ids=data.ids, seqlens=data.flat_seqlens("packed_input_ids"),
                                                 data=dict(prompt_mask=torch.zeros(n, dtype=torch.bool, device=dev))))
        return data


# ------------------------------------------------------------------------------------------- paired reward model


def _paired_rw_loss(out: ModelOutput, mb: SequenceSample):
    """Each item holds 2k sequences [pos, neg, pos, neg, ...]; score = value at the last token of each sequence."""
    seqlens = mb.flat_seqlens("packed_input_ids")
    dev = mb.data["packed_input_ids"].device
    ends = IF.seq_end_indices(seqlens, dev)
    scores = out.values.index_select(0, ends)
    pos, neg = scores[0::2], scores[1::2]
    group_factor = torch.tensor([1.0 / (len(l) // 2) for l in mb.seqlens["packed_input_ids"] for _ in range(len(l) // 2)],
                                device=dev)
    loss = -(F.logsigmoid(pos - neg) * group_factor).sum()
    with torch.no_grad():
        stat = dict(loss=loss.detach(), correct=(pos > neg).float().sum(), total=torch.tensor(float(pos.numel()), device=dev),
                    pos_score=pos.sum(), neg_score=neg.sum(), n_groups=torch.tensor(float(mb.bs), device=dev))
    return loss / max(mb.bs, 1), stat


@dataclasses.dataclass
class PairedRewardInterface(ModelInterface):
    enable_save: bool = True
    output_scaling: float = 1.0
    output_bias: float = 0.0

    @torch.no_grad()
    def inference(self, model: Model, data: SequenceSample, n_mbs=None) -> Optional[SequenceSample]:
        eng = model.module
        eng.eval()

        def last_token_score(out: ModelOutput, mb: SequenceSample):
            ends = IF.seq_end_indices(mb.flat_seqlens("packed_input_ids"), out.hidden.device)
            return out.values.index_select(0, ends)

        scores = eng.forward(data, num_micro_batches=n_mbs, post_hook=last_token_score)
        if scores is None:
            return None
        scores = (scores.float() - self.output_bias) * self.output_scaling
        # one score per sequence; items that hold a group of sequences (GRPO, paired data) get a group of scores
        with SequenceSample.disable_validation():
            return SequenceSample(keys=["rewards"], ids=data.ids, trailing_shapes=dict(rewards=()), dtypes=dict(rewards=torch.float32),
                                  seqlens=dict(rewards=[[1] * len(l) for l in data.seqlens["packed_input_ids"]]),
                                  data=dict(rewards=scores))

    def train_step(self, model: Model, data: SequenceSample, n_mbs=None) -> Dict:
        eng = model.module
        eng.train()
        st = eng.train_batch(data, _paired_rw_loss, version_steps=model.version.global_step, num_micro_batches=n_mbs)
        model.inc_version()
        k = n_mbs or 1
        dev = data.data["packed_input_ids"].device
        g = IF.dp_reduce_stats({x: st[x] * k for x in ("loss", "correct", "total", "pos_score", "neg_score", "n_groups")},
                               _dp_group(model), dev)
        tot = max(g["total"], 1.0)
        return dict(loss=g["loss"] / max(g["n_groups"], 1.0), acc=g["correct"] / tot, pos_score=g["pos_score"] / tot,
                    neg_score=g["neg_score"] / tot, grad_norm=float(st["grad_norm"]))

    @torch.no_grad()
    def evaluate(self, model: Model, eval_dataloader) -> Dict:
        eng = model.module
        eng.eval()
        dev = model.device
        tot = {k: torch.zeros((), device=dev) for k in ("loss", "correct", "total", "pos_score", "neg_score", "n_groups")}
        for batch in eval_dataloader:
            st = eng.eval_batch(batch.to_device(dev), _paired_rw_loss)
            for k in tot:
                tot[k] += st.get(k, 0.0)  # empty on all but the last pipeline stage
        g = IF.dp_reduce_stats(tot, _dp_group(model), dev)
        t = max(g["total"], 1.0)
        return dict(loss=g["loss"] / max(g["n_groups"], 1.0), acc=g["correct"] / t, pos_score=g["pos_score"] / t,
                    neg_score=g["neg_score"] / t)

    def save(self, model: Model, save_dir: str):
        if self.enable_save:
            _save_hf(model, save_dir)

    def _mock_train_step(self, model: Model, data: SequenceSample):
        return _mock_pairs(data)


def _mock_pairs(data: SequenceSample, with_seqlogp: bool = False) -> SequenceSample:
    """Profiling inputs for the paired interfaces: consecutive sequences become (chosen, rejected) pairs of one item."""
    lens = data.flat_seqlens("packed_input_ids")
    if len(lens) % 2:
        raise ValueError("paired interfaces are profiled with an even number of sequences")
    ids = data.data["packed_input_ids"]
    dev = ids.device
    n = len(lens) // 2
    grouped = [[lens[2 * i], lens[2 * i + 1]] for i in range(n)]
    pm = torch.zeros(sum(lens), dtype=torch.bool, device=dev)
    off = 0
    for l in lens:
        pm[off: off + max(1, l // 4)] = True
        off += l
    keys = dict(packed_input_ids=(ids, grouped), prompt_mask=(pm, grouped))
    if with_seqlogp:
        keys["seqlogp"] = (-torch.rand(len(lens), device=dev) * 10, [[1, 1]] * n)
    with SequenceSample.disable_validation():
        return SequenceSample(keys=list(keys), ids=[f"pair{i}" for i in range(n)], seqlens={k: v[1] for k, v in keys.items()},
                              trailing_shapes={k: () for k in keys}, dtypes={k: v[0].dtype for k, v in keys.items()},
                              data={k: v[0] for k, v in keys.items()})


# ------------------------------------------------------------------------------------------- DPO


def _answer_logp_sums(out: ModelOutput, mb: SequenceSample) -> torch.Tensor:
    """Per-sequence sum of answer-token log-probs; `prompt_lens` metadata gives each pair's shared prompt length."""
    seqlens = mb.flat_seqlens("packed_input_ids")
    dev = mb.data["packed_input_ids"].device
    rows, labels = IF.shifted_rows_and_labels(seqlens, mb.data["packed_input_ids"])
    logp = out.logprobs(labels, None, 1.0, rows)
    mask = (~mb.data["prompt_mask"].bool()).index_select(0, rows + 1).float()
    seq_id = torch.repeat_interleave(torch.arange(len(seqlens), device=dev), torch.tensor([l - 1 for l in seqlens], device=dev))
    return torch.zeros(len(seqlens), device=dev).index_add_(0, seq_id, logp * mask)


def _dpo_loss_from_output(out: ModelOutput, mb: SequenceSample, *, beta: float):
    pi = _answer_logp_sums(out, mb)
    ref = mb.data["seqlogp"].float()
    loss, pos, neg, kl = IF.dpo_loss(pi, ref, beta)
    n = torch.tensor(float(pi.numel() // 2), device=pi.device)
    return loss, dict(loss=loss.detach() * n, pos_score=pos, neg_score=neg, kl=kl, n_pairs=n)


@dataclasses.dataclass
class DPOInterface(ModelInterface):
    beta: float = 0.1
    enable_save: bool = True

    @torch.no_grad()
    def inference(self, model: Model, data: SequenceSample, n_mbs=None) -> Optional[SequenceSample]:
        eng = model.module
        eng.eval()
        seqlogp = eng.forward(data, num_micro_batches=n_mbs, post_hook=_answer_logp_sums)
        if seqlogp is None:
            return None
        with SequenceSample.disable_validation():
            return SequenceSample(keys=["seqlogp"], ids=data.ids, dtypes=dict(seqlogp=torch.float32),
                                  trailing_shapes=dict(seqlogp=()), data=dict(seqlogp=seqlogp.float()),
                                  seqlens=dict(seqlogp=[[1] * len(l) for l in data.seqlens["packed_input_ids"]]))

    def train_step(self, model: Model, data: SequenceSample, n_mbs=None) -> Dict:
        eng = model.module
        eng.train()
        st = eng.train_batch(data, functools.partial(_dpo_loss_from_output, beta=self.beta),
                             version_steps=model.version.global_step, num_micro_batches=n_mbs)
        model.inc_version()
        k = n_mbs or 1
        dev = data.data["packed_input_ids"].device
        g = IF.dp_reduce_stats({x: st[x] * k for x in ("loss", "pos_score", "neg_score", "kl", "n_pairs")}, _dp_group(model), dev)
        n = max(g["n_pairs"], 1.0)
        return dict(loss=g["loss"] / n, pos_score=g["pos_score"] / n, neg_score=g["neg_score"] / n, kl=g["kl"] / (2 * n),
                    grad_norm=float(st["grad_norm"]))

    def _mock_inference(self, model: Model, data: SequenceSample):
        return _mock_pairs(data)

    def _mock_train_step(self, model: Model, data: SequenceSample):
        return _mock_pairs(data, with_seqlogp=True)

    def save(self, model: Model, save_dir: str):
        if self.enable_save:
            _save_hf(model, save_dir)


# ------------------------------------------------------------------------------------------- generation only


@dataclasses.dataclass
class GenerationInterface(ModelInterface):
    output_file: Optional[str] = None
    generation_config: Dict = dataclasses.field(default_factory=dict)

    def __post_init__(self):
        g = self.generation_config
        self.gconfig = g if isinstance(g, GenerationHyperparameters) else GenerationHyperparameters(**g)

    @torch.no_grad()
    def generate(self, model: Model, data: SequenceSample, n_mbs=None) -> Optional[SequenceSample]:
        eng = model.module
        eng.eval()
        plens = data.flat_seqlens("packed_prompts")
        x = SequenceSample.from_default(ids=data.ids, seqlens=plens, data=dict(packed_input_ids=data.data["packed_prompts"]))
        outs = eng.generate(x, tokenizer=model.tokenizer, gconfig=self.gconfig, num_micro_batches=n_mbs)
        if outs is None:
            return None
        dev = data.data["packed_prompts"].device
        records, all_tokens, all_lens = [], [], []
        for mb, o in IF.pair_generation_outputs(x, outs):
            ids, cu, _ = _mb_prompt(mb, dev)
            cu_l = cu.tolist()
            toks, glens = o.tokens.tolist(), o.gen_lens.tolist()
            for i, sid in enumerate(mb.ids):
                p = ids[cu_l[i]:cu_l[i + 1]].tolist()
                a = toks[i][: glens[i]]
                rec = dict(id=sid, prompt_ids=p, answer_ids=a)
                if hasattr(model.tokenizer, "decode"):
                    rec["prompt"] = model.tokenizer.decode(p, skip_special_tokens=True)
                    rec["answer"] = model.tokenizer.decode(a, skip_special_tokens=True)
                records.append(rec)
            all_tokens.append(torch.cat([o.tokens[i, : glens[i]] for i in range(len(glens))]))
            all_lens += glens
        ctx = _engine_ctx(model)
        if self.output_file is not None and (ctx is None or ctx.is_dp_head):
            os.makedirs(os.path.dirname(os.path.abspath(self.output_file)), exist_ok=True)
            with open(self.output_file, "a") as f:
                fcntl.flock(f, fcntl.LOCK_EX)
                try:
                    for r in records:
                        f.write(json.dumps(r, ensure_ascii=False) + "\n")
                finally:
                    fcntl.flock(f, fcntl.LOCK_UN)
        with SequenceSample.disable_validation():
            return SequenceSample(keys=["gen_tokens"], ids=data.ids, dtypes=dict(gen_tokens=torch.long),
                                  trailing_shapes=dict(gen_tokens=()), data=dict(gen_tokens=torch.cat(all_tokens)),
                                  seqlens=dict(gen_tokens=[[int(l)] for l in all_lens]))


register_interface("sft", SFTInterface)
register_interface("paired_rw", PairedRewardInterface)
register_interface("dpo", DPOInterface)
register_interface("generation", GenerationInterface)
