"""Autoregressive generation: packed prefill -> dense KV cache -> CUDA-graph decode loop -> sampling.

Parity: `realhf/impl/model/nn/real_llm_generate.py` (genstep :26-141, generate :252-368,
concat_prompt_to_generation_output :451-530) and `utils/logits_warper.py`.  Differences by design:
  * termination is checked on the device and read back only every `sync_every` steps (and never before
    `min_new_tokens`), instead of one host sync per token (reference :123-129);
  * the "filtered by top-k/top-p" logits mask is produced bit-packed ([T, V/8] uint8), 8x smaller than the
    reference's bool [T, V];
  * the KV cache is laid out [B, n_kv, S, hd] so every (sequence, head) stream is contiguous for the
    split-KV decode kernel, which also applies RoPE and appends K/V in place.
"""

from __future__ import annotations

import os

import dataclasses
from typing import List, Optional, Tuple

import torch

from realhf_b200.api.model import GenerationHyperparameters
from realhf_b200.models.real_model import ReaLModel
from realhf_b200.ops import functional as OF
from realhf_b200.ops import launches
from realhf_b200.parallel import tp as TP


_SAMPLE_SEED = [0x5DEECE66D]
LAST_TIMING: dict = {}   # REAL_GEN_TIMING=1: device ms of the phases of the last `generate` call (prefill / capture / decode loop)


def seed_sampling(seed: int):
    _SAMPLE_SEED[0] = int(seed) % (1 << 62)


@dataclasses.dataclass
class GenerationOutput:
    tokens: torch.Tensor        # [B, n_gen] generated ids (pad after EOS)
    logprobs: torch.Tensor      # [B, n_gen] fp32 log-prob of each generated token under the sampling distribution
    mask_bits: Optional[torch.Tensor]  # [B, n_gen, ceil(V/8)] uint8, bit set = token was filtered out; None if disabled
    gen_lens: torch.Tensor      # [B] number of generated tokens including EOS
    no_eos: torch.Tensor        # [B] bool: hit max_new_tokens without EOS


def _filter_logits(logits: torch.Tensor, g: GenerationHyperparameters) -> torch.Tensor:
    """top-k then top-p on fp32 logits [B, V]; filtered entries -> -inf."""
    V = logits.shape[-1]
    neg = torch.finfo(logits.dtype).min
    if g.top_k < V:
        kth = torch.topk(logits, g.top_k, dim=-1).values[..., -1:]
        logits = logits.masked_fill(logits < kth, neg)
    if g.top_p < 1.0:
        sl, si = torch.sort(logits, descending=True, dim=-1)
        cp = torch.softmax(sl, dim=-1).cumsum(-1)
        remove = cp - torch.softmax(sl, dim=-1) >= g.top_p  # keep the first token that crosses top_p
        remove[..., 0] = False
        sl = sl.masked_fill(remove, neg)
        logits = torch.empty_like(logits).scatter_(-1, si, sl)
    return logits


def fused_sampler_ok(device, vocab_size: int) -> bool:
    """The fused sampling kernel stages one fp32 logits row in shared memory (<= 200 KB: vocabularies up to 51200)."""
    return torch.device(device).type == "cuda" and vocab_size * 4 <= 200 * 1024


def genstep(logits: torch.Tensor, g: GenerationHyperparameters, step: int, eos_id: Optional[int], pad_id: int,
            unfinished: torch.Tensor, generator: Optional[torch.Generator] = None, want_mask: bool = True):
    """One sampling step on full-vocab logits [B, V].  Returns (next_tokens, logprob, mask_bits | None, unfinished).

    CUDA tensors take the fused sampling kernel (`csrc/sampling.cu`); `generator` selects the PyTorch reference path
    (reproducible with torch RNG), which is also what CPU tensors use."""
    if logits.is_cuda and generator is None and logits.shape[1] * 4 <= 200 * 1024:
        from realhf_b200.ops import lib
        need_mask = want_mask and not g.force_no_logits_mask and not g.greedy
        seed = int(_SAMPLE_SEED[0])
        _SAMPLE_SEED[0] = (seed * 6364136223846793005 + 1442695040888963407) % (1 << 62)
        nxt, lp, mb = lib().sample(logits, unfinished, g.top_k, g.top_p, 1.0 / g.temperature, eos_id if eos_id is not None else -1,
                                   eos_id is not None and step < g.min_new_tokens, g.greedy, pad_id, seed, step, need_mask)
        if eos_id is not None:
            unfinished = unfinished & (nxt != eos_id)
        return nxt, lp, (mb if need_mask else None), unfinished
    x = logits.float()
    if eos_id is not None and step < g.min_new_tokens:
        x[:, eos_id] = torch.finfo(x.dtype).min
    if not g.greedy:
        x = x / g.temperature
        x = _filter_logits(x, g)
    lp_all = torch.log_softmax(x, dim=-1)
    if g.greedy:
        nxt = x.argmax(dim=-1)
    else:
        nxt = torch.multinomial(lp_all.exp(), 1, generator=generator).squeeze(-1)
    lp = lp_all.gather(-1, nxt.unsqueeze(-1)).squeeze(-1)
    mask_bits = None
    if want_mask and not g.force_no_logits_mask and not g.greedy:
        mask_bits = OF.pack_mask_bits(x == torch.finfo(x.dtype).min)
    nxt = torch.where(unfinished, nxt, torch.full_like(nxt, pad_id))
    lp = torch.where(unfinished, lp, torch.zeros_like(lp))
    if eos_id is not None:
        unfinished = unfinished & (nxt != eos_id)
    return nxt, lp, mask_bits, unfinished


class DecodeState:
    """Static buffers of one decode session (fixed addresses => CUDA-graph replayable)."""

    def __init__(self, model: ReaLModel, B: int, S: int):
        c = model.config
        nq, nkv = model._local_heads()
        dev, dt = model.device, model.dtype
        nb = model.n_local_blocks()
        # [B, nkv, S, hd] storage viewed as logical [B, S, nkv, hd]
        self.k = [torch.zeros(B, nkv, S, c.head_dim, dtype=dt, device=dev).permute(0, 2, 1, 3) for _ in range(nb)]
        self.v = [torch.zeros(B, nkv, S, c.head_dim, dtype=dt, device=dev).permute(0, 2, 1, 3) for _ in range(nb)]
        self.cache_lens = torch.zeros(B, dtype=torch.int32, device=dev)
        self.input_ids = torch.zeros(B, dtype=torch.long, device=dev)
        self.hidden_in = torch.zeros(B, c.hidden_dim, dtype=dt, device=dev) if not model.is_first_stage else None
        self.out: Optional[torch.Tensor] = None
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.graph_launches = 0
        self.graph_sig = None
        self.B, self.S = B, S
        # in-graph sampling: per-row step counters, liveness flags, device-resident seed and the [B, n_gen] history buffers the
        # sampling kernel writes at column `step`
        self.step_rows = torch.zeros(B, dtype=torch.int32, device=dev)
        self.unfinished = torch.ones(B, dtype=torch.bool, device=dev)
        self.seed = torch.zeros(1, dtype=torch.long, device=dev)
        self.tok_hist = self.lp_hist = self.mask_hist = None

    def prepare_hist(self, n_gen: int, V: int, need_mask: bool):
        dev = self.cache_lens.device
        if self.tok_hist is None or self.tok_hist.shape[1] != n_gen or (need_mask and self.mask_hist is None):
            self.tok_hist = torch.zeros(self.B, n_gen, dtype=torch.long, device=dev)
            self.lp_hist = torch.zeros(self.B, n_gen, dtype=torch.float32, device=dev)
            self.mask_hist = torch.zeros(self.B, n_gen, (V + 7) // 8, dtype=torch.uint8, device=dev) if need_mask else None
            self.graph = None  # a captured graph points at the old buffers

    def drop_hist(self):
        self.tok_hist = self.lp_hist = self.mask_hist = None

    def fill_from_prefill(self, kv: List[Tuple[torch.Tensor, torch.Tensor]], cu_seqlens: torch.Tensor, lens: torch.Tensor):
        """Scatter the packed prefill K/V of every block into the dense caches."""
        dev = cu_seqlens.device
        T = int(kv[0][0].shape[0])
        tok = torch.arange(T, device=dev)
        seq = torch.searchsorted(cu_seqlens[1:].long().contiguous(), tok, right=True)
        pos = tok - cu_seqlens.long()[seq]
        for li, (k, v) in enumerate(kv):
            self.k[li][seq, pos] = k
            self.v[li][seq, pos] = v
        self.cache_lens.copy_(lens.int())


def _final_logits(model: ReaLModel, hidden: torch.Tensor) -> torch.Tensor:
    """[B, H] -> full-vocab fp32-able logits on every TP rank."""
    f8 = model._fp8.get("head") if model._fp8_active and model._fp8 else None
    if f8 is not None and hidden.shape[0] <= 128:
        return f8(hidden)
    lg = OF.linear(hidden, model.head_weight())
    if model.ctx.tp_size > 1:
        lg = TP._gather_last_dim(lg, model.ctx)
    return lg


@torch.no_grad()
def generate(model: ReaLModel, input_ids: torch.Tensor, cu_seqlens: torch.Tensor, g: GenerationHyperparameters,
             eos_id: Optional[int], pad_id: int, generator: Optional[torch.Generator] = None,
             state: Optional[DecodeState] = None, sync_every: int = 16) -> Tuple[GenerationOutput, DecodeState]:
    """Generate for a packed batch of prompts on a single pipeline stage (pp == 1)."""
    assert model.is_first_stage and model.is_last_stage, "pipelined generation goes through engine.pipe_runner"
    dev = model.device
    if generator is None and model.ctx.tp_size > 1 and not g.greedy and not fused_sampler_ok(dev, model.config.vocab_size):
        # every TP rank samples from the same (gathered) distribution and must draw the same token: the PyTorch sampling
        # path (CPU tensors, and CUDA vocabularies too large for the fused sampler's shared-memory row: Llama-3, Qwen2, Gemma)
        # needs a stream shared by the group; the fused CUDA sampler is seeded identically on all ranks already
        generator = model.shared_generator(dev)
    cu = cu_seqlens.int()
    B = cu.numel() - 1
    lens = (cu[1:] - cu[:-1])
    max_prompt = int(lens.max())
    S = max_prompt + g.max_new_tokens
    was_training = model.training
    model.eval()
    sp_saved, model.sequence_parallel = model.sequence_parallel, False  # token-sharded activations make no sense for decode
    timing = os.environ.get("REAL_GEN_TIMING", "0") == "1" and dev.type == "cuda"
    marks = []

    def mark(name):
        if timing:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append((name, e))
    mark("start")
    # ---- prefill
    kv: List[Tuple[torch.Tensor, torch.Tensor]] = []
    out = model(input_ids=input_ids, cu_seqlens=cu, max_seqlen=max_prompt, kv_sink=kv)
    last = (cu[1:] - 1).long()
    logits = _final_logits(model, out.hidden.index_select(0, last))
    if state is None or state.B != B or state.S < S:
        state = DecodeState(model, B, S)
    state.fill_from_prefill(kv, cu, lens)
    del kv, out
    unfinished = torch.ones(B, dtype=torch.bool, device=dev)
    toks, lps, masks = [], [], []
    nxt, lp, mb, unfinished = genstep(logits, g, 0, eos_id, pad_id, unfinished, generator)
    toks.append(nxt); lps.append(lp); masks.append(mb)
    # ---- decode loop
    use_graph = g.use_cuda_graph and dev.type == "cuda"
    # opt-in W8A8 decode (ops/fp8.py): the prefill above ran in bf16; from here on block linears and the head read e4m3 copies
    use_fp8 = (g.fp8_weights or os.environ.get("REAL_GEN_FP8", "0") == "1") and B <= 128 and model.fp8_decode_supported()
    if use_fp8:
        model.enable_fp8_decode()
    V = logits.shape[1]
    # the fused sampler can run as the tail of the captured step: then a decode step is ONE graph replay and nothing else
    # (no eager copy of the next token, no cache_lens += 1, no sampling launch, no per-step python bookkeeping)
    in_graph = use_graph and generator is None and V * 4 <= 200 * 1024 and os.environ.get("REAL_GEN_SAMPLE_IN_GRAPH", "1") == "1"
    need_mask = mb is not None

    def one_step():
        h = model.decode_step(state.input_ids, state.k, state.v, state.cache_lens)
        return _final_logits(model, h)

    sig = (in_graph, g.max_new_tokens, g.min_new_tokens, g.top_k, g.top_p, g.temperature, g.greedy, need_mask, eos_id, pad_id,
           int(model.flat_param.data_ptr()), use_fp8)
    if state.graph is not None and state.graph_sig != sig:
        state.graph = None  # sampling options changed, or the flat parameter buffer moved (realloc / offload reload / ZeRO-3)
    if in_graph:
        state.prepare_hist(g.max_new_tokens, V, need_mask)
        state.tok_hist[:, 0] = nxt
        state.lp_hist[:, 0] = lp
        if need_mask:
            state.mask_hist[:, 0] = mb
        state.step_rows.fill_(1)
        state.unfinished.copy_(unfinished)
        seed = int(_SAMPLE_SEED[0])
        _SAMPLE_SEED[0] = (seed * 6364136223846793005 + 1442695040888963407) % (1 << 62)
        state.seed.fill_(seed)
    mark("prefill")
    if use_graph and state.graph is None:
        state.input_ids.copy_(nxt)
        lens_backup = state.cache_lens.clone()
        s = torch.cuda.Stream(dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            one_step()  # warm-up outside capture (allocations, lazy inits)
        torch.cuda.current_stream(dev).wait_stream(s)
        state.cache_lens.copy_(lens_backup)
        graph = torch.cuda.CUDAGraph()
        launches.begin_capture()
        with torch.cuda.graph(graph):
            state.out = one_step()
            if in_graph:
                from realhf_b200.ops import lib
                lib().sample_graph(state.out, state.tok_hist, state.lp_hist, state.mask_hist if need_mask else None, state.unfinished,
                                   state.step_rows, state.input_ids, state.cache_lens, state.seed, g.min_new_tokens if eos_id is not None else 0,
                                   g.top_k, g.top_p, 1.0 / g.temperature, eos_id if eos_id is not None else -1, g.greedy, pad_id)
        state.graph_launches = launches.end_capture()
        state.graph = graph
        state.graph_sig = sig
    mark("capture")
    step = 1
    if in_graph:
        state.input_ids.copy_(nxt)
        while step < g.max_new_tokens:
            state.graph.replay()
            launches.count_replay(state.graph_launches)
            step += 1
            if eos_id is not None and step >= g.min_new_tokens and step % sync_every == 0 and not bool(state.unfinished.any()):
                break
        take = (lambda t: t) if g.force_cudagraph_recapture else (lambda t: t.clone())  # a kept graph keeps writing these buffers
        tokens = take(state.tok_hist[:, :step])
        logprobs = take(state.lp_hist[:, :step])
        mask_bits = take(state.mask_hist[:, :step]) if need_mask else None
        if g.force_cudagraph_recapture:
            state.drop_hist()
    else:
        while step < g.max_new_tokens:
            state.input_ids.copy_(nxt)
            if use_graph:
                state.graph.replay()
                launches.count_replay(state.graph_launches)
                logits = state.out
            else:
                logits = one_step()
            state.cache_lens += 1
            nxt, lp, mb, unfinished = genstep(logits, g, step, eos_id, pad_id, unfinished, generator)
            toks.append(nxt); lps.append(lp); masks.append(mb)
            step += 1
            if eos_id is not None and step >= g.min_new_tokens and step % sync_every == 0 and not bool(unfinished.any()):
                break
        tokens = torch.stack(toks, 1)
        logprobs = torch.stack(lps, 1)
        mask_bits = torch.stack(masks, 1) if masks[0] is not None else None
    mark("decode")
    if use_fp8:  # a graph kept for the next call holds pointers into the e4m3 buffers: then they stay (re-quantised in place)
        model.disable_fp8_decode(free=not (use_graph and state.graph is not None and not g.force_cudagraph_recapture))
    if timing:
        torch.cuda.synchronize(dev)
        LAST_TIMING.clear()
        LAST_TIMING.update({f"{b[0]}_ms": round(a[1].elapsed_time(b[1]), 2) for a, b in zip(marks[:-1], marks[1:])})
        LAST_TIMING.update(steps=step, B=B, ms_per_decode_step=round(marks[-2][1].elapsed_time(marks[-1][1]) / max(step - 1, 1), 4))
    n_gen = tokens.shape[1]
    if eos_id is not None:
        is_eos = tokens == eos_id
        first = torch.where(is_eos.any(1), is_eos.float().argmax(1) + 1, torch.full((B,), n_gen, device=dev))
        no_eos = ~is_eos.any(1)
    else:
        first = torch.full((B,), n_gen, device=dev)
        no_eos = torch.ones(B, dtype=torch.bool, device=dev)
    if g.force_cudagraph_recapture and state.graph is not None:
        state.graph = None
        state.out = None
    model.train(was_training)
    model.sequence_parallel = sp_saved
    return GenerationOutput(tokens, logprobs, mask_bits, first.long(), no_eos), state


def concat_prompt_to_generation_output(prompt_ids: torch.Tensor, prompt_cu: torch.Tensor, out: GenerationOutput):
    """Build packed `seq = prompt + generation` per sample without Python loops over samples.

    Returns (packed_seq [sum L], seq_lens [B], packed_logprobs [sum (L-1)], packed_mask_bits [sum L, V/8] | None,
    prompt_mask [sum L] bool).  Log-probs / masks are aligned the reference way: entry t of a sequence belongs to the
    prediction of token t+1, zero over the prompt part (real_llm_generate.py:451-530)."""
    dev = prompt_ids.device
    B = out.tokens.shape[0]
    plens = (prompt_cu[1:] - prompt_cu[:-1]).long()
    glens = out.gen_lens.long()
    slens = plens + glens
    cu = torch.zeros(B + 1, dtype=torch.long, device=dev)
    cu[1:] = slens.cumsum(0)
    total = int(cu[-1])
    tok = torch.arange(total, device=dev)
    seq = torch.searchsorted(cu[1:].contiguous(), tok, right=True)
    pos = tok - cu[seq]
    in_prompt = pos < plens[seq]
    packed = torch.empty(total, dtype=prompt_ids.dtype, device=dev)
    packed[in_prompt] = prompt_ids[(prompt_cu.long()[seq] + pos)[in_prompt]]
    gpos = (pos - plens[seq]).clamp(min=0)
    packed[~in_prompt] = out.tokens[seq[~in_prompt], gpos[~in_prompt]]
    # logprobs: length L-1 per sequence; position t (0-based) predicts token t+1
    cu1 = cu - torch.arange(B + 1, device=dev)
    total1 = int(cu1[-1])
    tok1 = torch.arange(total1, device=dev)
    seq1 = torch.searchsorted(cu1[1:].contiguous(), tok1, right=True)
    pos1 = tok1 - cu1[seq1]
    is_gen1 = pos1 >= plens[seq1] - 1
    gp1 = (pos1 - (plens[seq1] - 1)).clamp(min=0)
    lp = torch.zeros(total1, dtype=torch.float32, device=dev)
    lp[is_gen1] = out.logprobs[seq1[is_gen1], gp1[is_gen1]]
    mask_bits = None
    if out.mask_bits is not None:
        mask_bits = torch.zeros(total1, out.mask_bits.shape[-1], dtype=torch.uint8, device=dev)
        mask_bits[is_gen1] = out.mask_bits[seq1[is_gen1], gp1[is_gen1]]
    return packed, slens, lp, mask_bits, in_prompt


class InflightBatchingGenerator:
    """Continuous (in-flight) batching over a fixed number of decode slots.

    Parity: `InflightBatchingGenerator` (nn/real_llm_generate.py:664-882; not used by the reference's interfaces either).
    Prompts wait in a queue; whenever a slot's sequence ends (EOS or `max_new_tokens`) its result is emitted and the slot
    is refilled by a prefill of the next prompt, so short answers do not hold the batch hostage.  The KV cache, the
    per-slot lengths and the next-token buffer are the same static `DecodeState` buffers the plain loop uses, i.e. a
    decode step is one `decode_step` call (CUDA-graph capturable: refills only rewrite buffer contents)."""

    def __init__(self, model: ReaLModel, g: GenerationHyperparameters, eos_id: Optional[int], pad_id: int, n_slots: int,
                 max_prompt_len: int):
        assert model.is_first_stage and model.is_last_stage
        self.model, self.g, self.eos_id, self.pad_id, self.B = model, g, eos_id, pad_id, n_slots
        self.state = DecodeState(model, n_slots, max_prompt_len + g.max_new_tokens)
        self.slot_seq = [-1] * n_slots                 # which request a slot serves
        self.slot_tokens: List[List[int]] = [[] for _ in range(n_slots)]
        self.slot_logprobs: List[List[float]] = [[] for _ in range(n_slots)]
        self.unfinished = torch.zeros(n_slots, dtype=torch.bool, device=model.device)

    @torch.no_grad()
    def _refill(self, slot: int, req_id: int, prompt: torch.Tensor):
        m, st = self.model, self.state
        cu = torch.tensor([0, prompt.numel()], dtype=torch.int32, device=m.device)
        kv: List[Tuple[torch.Tensor, torch.Tensor]] = []
        out = m(input_ids=prompt, cu_seqlens=cu, max_seqlen=int(prompt.numel()), kv_sink=kv)
        logits = _final_logits(m, out.hidden[-1:])
        n = prompt.numel()
        for li, (k, v) in enumerate(kv):
            st.k[li][slot, :n] = k
            st.v[li][slot, :n] = v
        st.cache_lens[slot] = n
        one = torch.ones(1, dtype=torch.bool, device=m.device)
        nxt, lp, _, _ = genstep(logits, self.g, 0, self.eos_id, self.pad_id, one, want_mask=False)
        st.input_ids[slot] = nxt[0]
        self.slot_seq[slot] = req_id
        self.slot_tokens[slot] = [int(nxt[0])]
        self.slot_logprobs[slot] = [float(lp[0])]
        self.unfinished[slot] = True

    @torch.no_grad()
    def generate(self, prompts: List[torch.Tensor]):
        """prompts: list of 1-D token tensors.  Returns per request (tokens, logprobs, ended_with_eos) in request order."""
        m, st, g = self.model, self.state, self.g
        was_training = m.training
        m.eval()
        sp_saved, m.sequence_parallel = m.sequence_parallel, False
        results: List[Optional[Tuple[List[int], List[float], bool]]] = [None] * len(prompts)
        queue = list(range(len(prompts)))
        n_done = 0

        def retire(slot: int):
            nonlocal n_done
            rid = self.slot_seq[slot]
            toks = self.slot_tokens[slot]
            ended = self.eos_id is not None and len(toks) > 0 and toks[-1] == self.eos_id
            results[rid] = (toks, self.slot_logprobs[slot], ended)
            self.slot_seq[slot] = -1
            self.unfinished[slot] = False
            n_done += 1

        while n_done < len(prompts):
            for slot in range(self.B):
                if self.slot_seq[slot] < 0 and queue:
                    rid = queue.pop(0)
                    self._refill(slot, rid, prompts[rid].to(m.device))
                    toks = self.slot_tokens[slot]
                    if (self.eos_id is not None and toks[-1] == self.eos_id and g.min_new_tokens <= 1) or g.max_new_tokens <= 1:
                        retire(slot)
            active = [s for s in range(self.B) if self.slot_seq[s] >= 0]
            if not active:
                continue
            h = m.decode_step(st.input_ids, st.k, st.v, st.cache_lens)
            logits = _final_logits(m, h)
            st.cache_lens += self.unfinished.int()
            # per-slot step index decides whether EOS is still suppressed
            nxt, lp, _, _ = genstep(logits, g, min(len(self.slot_tokens[s]) for s in active), self.eos_id, self.pad_id,
                                    self.unfinished, want_mask=False)
            st.input_ids.copy_(nxt)
            nxt_h, lp_h = nxt.tolist(), lp.tolist()
            for s in active:
                self.slot_tokens[s].append(nxt_h[s])
                self.slot_logprobs[s].append(lp_h[s])
                n = len(self.slot_tokens[s])
                if n >= g.max_new_tokens or (self.eos_id is not None and nxt_h[s] == self.eos_id and n >= g.min_new_tokens):
                    retire(s)
        m.train(was_training)
        m.sequence_parallel = sp_saved
        return results
