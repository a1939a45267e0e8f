"""Pipeline-parallel execution: inference, 1F1B training and ring generation schedules.

Parity: `realhf/impl/model/backend/pipe_runner.py` + `parallelism/pipeline_parallel/{static_schedule,instruction,
p2p}.py`.  The reference interprets static instruction lists with *blocking* `dist.send/recv` plus a 1-element
"terminate" message after every activation (p2p.py:49-66, pipe_runner.py:545-636).  Here the three schedules are
written directly as loops over micro-batches, every transfer is a non-blocking `isend/irecv` on NCCL (gloo on CPU)
posted as early as its data dependency allows — receives for micro-batch i+1 are in flight while micro-batch i
computes — and generation needs no termination message because all stages run the same number of steps (the device-side
EOS check is read back every `sync_every` steps on the last stage and broadcast with the next tokens).
"""

from __future__ import annotations

import collections
from typing import Any, Callable, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from realhf_b200.api.data import SequenceSample
from realhf_b200.api.model import GenerationHyperparameters
from realhf_b200.models import generation as gen
from realhf_b200.models.real_model import ReaLModel


def _mb_inputs(mb: SequenceSample, device, key: str = "packed_input_ids"):
    lens = mb.flat_seqlens(key)
    cu = torch.zeros(len(lens) + 1, dtype=torch.int32)
    cu[1:] = torch.tensor(lens, dtype=torch.int32).cumsum(0)
    return mb.data[key], cu.to(device), max(lens), sum(lens)


class PipelineRunner:
    def __init__(self, engine):
        self.engine = engine
        self.m: ReaLModel = engine.module
        self.ctx = engine.ctx
        self.pp, self.rank = self.ctx.pp_size, self.ctx.pp_rank
        self.first, self.last = self.m.is_first_stage, self.m.is_last_stage
        self.prev, self.next = self.ctx.pp_prev(), self.ctx.pp_next()

    # ------------------------------------------------------------------ helpers
    def _hidden_shape(self, n_tokens: int):
        t = self.ctx.tp_size if self.m.sequence_parallel else 1
        return (n_tokens // t, self.m.config.hidden_dim)

    def _split(self, input_: SequenceSample, n: int) -> List[SequenceSample]:
        n = max(1, min(n, input_.bs))
        return input_.split(n)

    def _irecv(self, shape, dtype, src):
        buf = torch.empty(shape, dtype=dtype, device=self.m.device)
        return buf, dist.irecv(buf, src)

    def _prep(self, mb: SequenceSample):
        """(ids, cu, max_seqlen, padded token count, n_pad): with sequence parallelism every stage pads identically."""
        ids, cu, mx, T = _mb_inputs(mb, self.m.device)
        n_pad = 0
        if self.m.sequence_parallel:
            from realhf_b200.engine.engine import pad_for_sp
            ids, cu, mx, n_pad = pad_for_sp(ids, cu, mx, self.ctx.tp_size)
        return ids, cu, mx, T + n_pad, n_pad

    def _stage_forward(self, ids, cu, mx, hidden, n_pad: int = 0):
        if self.first:
            out = self.m(input_ids=ids, cu_seqlens=cu, max_seqlen=mx)
        else:
            out = self.m(hidden=hidden, cu_seqlens=cu, max_seqlen=mx)
        if self.last and n_pad:
            out.hidden = out.hidden[: out.hidden.shape[0] - n_pad]
        return out

    # ------------------------------------------------------------------ inference
    @torch.no_grad()
    def forward(self, input_: SequenceSample, n_mbs: int, post_hook: Optional[Callable], aggregate_fn: Callable = torch.cat):
        mbs = self._split(input_, self.pp * n_mbs)
        metas = [self._prep(mb) for mb in mbs]
        outs, pending_send = [], []
        nxt = None
        if not self.first:
            nxt = self._irecv(self._hidden_shape(metas[0][3]), self.m.dtype, self.prev)
        for i, (mb, (ids, cu, mx, T, n_pad)) in enumerate(zip(mbs, metas)):
            hidden = None
            if not self.first:
                buf, h = nxt
                h.wait()
                hidden = buf
                if i + 1 < len(mbs):
                    nxt = self._irecv(self._hidden_shape(metas[i + 1][3]), self.m.dtype, self.prev)
            out = self._stage_forward(ids, cu, mx, hidden, n_pad)
            if self.last:
                outs.append(post_hook(out, mb) if post_hook is not None else out.logits)
            else:
                x = out.contiguous()
                pending_send.append((x, dist.isend(x, self.next)))
        for _, h in pending_send:
            h.wait()
        if not self.last:
            return None
        return aggregate_fn(outs) if len(outs) > 1 else outs[0]

    @torch.no_grad()
    def eval_batch(self, input_: SequenceSample, loss_fn: Callable, n_mbs: int):
        stats: Dict[str, Any] = collections.defaultdict(float)
        n = [0]

        def hook(out, mb):
            _, st = loss_fn(out, mb)
            for k, v in st.items():
                stats[k] = stats[k] + v
            n[0] += 1
            return torch.zeros(1, device=self.m.device)

        self.forward(input_, n_mbs, hook)
        return {k: v / max(n[0], 1) for k, v in stats.items()} if self.last else {}

    # ------------------------------------------------------------------ training: 1F1B
    def train_batch(self, input_: SequenceSample, loss_fn: Callable, n_mbs: int) -> Dict[str, Any]:
        mbs = self._split(input_, 2 * self.pp * n_mbs)
        M = len(mbs)
        metas = [self._prep(mb) for mb in mbs]
        optim = self.engine.optim
        stats: Dict[str, Any] = collections.defaultdict(float)
        saved: Dict[int, Tuple[Optional[torch.Tensor], torch.Tensor]] = {}
        sends = []
        n_warm = min(self.pp - self.rank - 1, M)
        dev, dt = self.m.device, self.m.dtype

        # Point-to-point plumbing.  In steady state a stage has a send and a receive pending towards the SAME neighbour in
        # opposite directions (activation out / gradient in, gradient out / activation in); posted as separate isend / irecv
        # they can deadlock on NCCL once a message exceeds the p2p FIFO (both sides block in their send while the matching
        # receive is queued behind it).  Each such pair is therefore ONE `batch_isend_irecv` group, which NCCL progresses
        # concurrently (Megatron's send_forward_recv_backward / send_backward_recv_forward; the reference orders blocking
        # send / recv per stage parity instead, pipe_runner.py:650-753).
        def exchange(send_t, send_to, recv_shape, recv_from):
            ops, buf = [], None
            if send_t is not None:
                ops.append(dist.P2POp(dist.isend, send_t, send_to))
            if recv_shape is not None:
                buf = torch.empty(recv_shape, dtype=dt, device=dev)
                ops.append(dist.P2POp(dist.irecv, buf, recv_from))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
            return buf

        def compute_fwd(i, hidden):
            ids, cu, mx, T, n_pad = metas[i]
            if hidden is not None:
                hidden = hidden.requires_grad_(True)
            out = self._stage_forward(ids, cu, mx, hidden, n_pad)
            if self.last:
                loss, st = loss_fn(out, mbs[i])
                for k, v in st.items():
                    stats[k] = stats[k] + (v.detach() if torch.is_tensor(v) else v) / M
                saved[i] = (hidden, optim.scale_loss(loss / M))
                return None
            x = out.contiguous()
            saved[i] = (hidden, x)
            return x.detach()

        def compute_bwd(i, g):
            hidden, out = saved.pop(i)
            if self.last:
                out.backward()
            else:
                torch.autograd.backward(out, g)
            return None if self.first else hidden.grad.contiguous()

        hshape = lambda i: self._hidden_shape(metas[i][3])
        f = b = 0
        # warm-up: forwards only (activation in from prev, activation out to next: one direction per neighbour)
        for _ in range(n_warm):
            h = exchange(None, None, None if self.first else hshape(f), self.prev)
            x = compute_fwd(f, h)
            sends.append((x, dist.isend(x, self.next)))
            f += 1
        # steady state: one forward, one backward
        h = None
        if f < M and not self.first:
            h = exchange(None, None, hshape(f), self.prev)
        while f < M:
            x = compute_fwd(f, h)
            f += 1
            # activation out + gradient in, same neighbour: one group
            g = None if self.last else exchange(x, self.next, saved[b][1].shape, self.next)
            gi = compute_bwd(b, g)
            b += 1
            # gradient out + next activation in, same neighbour: one group
            h = exchange(gi, self.prev, hshape(f) if (f < M and not self.first) else None, self.prev)
        # cool-down: backwards only
        while b < M:
            g = None if self.last else exchange(None, None, saved[b][1].shape, self.next)
            gi = compute_bwd(b, g)
            b += 1
            if gi is not None:
                sends.append((gi, dist.isend(gi, self.prev)))
        for _, hnd in sends:
            hnd.wait()
        # statistics live on the last stage: share them along the pipe so every stage returns the same dict
        keys = sorted(stats) if self.last else None
        obj = [keys]
        dist.broadcast_object_list(obj, src=self.ctx.global_rank(pipe=self.pp - 1, data=self.ctx.dp_rank, model=self.ctx.tp_rank),
                                   group=self.ctx.pp_group)
        keys = obj[0]
        vec = torch.stack([torch.as_tensor(stats[k], dtype=torch.float32, device=self.m.device).reshape(()) for k in keys]) \
            if self.last else torch.zeros(len(keys), device=self.m.device)
        dist.broadcast(vec, src=self.ctx.global_rank(pipe=self.pp - 1, data=self.ctx.dp_rank, model=self.ctx.tp_rank), group=self.ctx.pp_group)
        return {k: vec[j] for j, k in enumerate(keys)}

    # ------------------------------------------------------------------ generation: ring last -> first
    @torch.no_grad()
    def generate(self, input_: SequenceSample, g: GenerationHyperparameters, eos_id, pad_id, n_mbs: int,
                 sync_every: int = 16) -> List[gen.GenerationOutput]:
        m, dev = self.m, self.m.device
        mbs = self._split(input_, self.pp * n_mbs)
        M = len(mbs)
        metas = [_mb_inputs(mb, dev) for mb in mbs]
        was_training = m.training
        m.eval()
        sp_saved, m.sequence_parallel = m.sequence_parallel, False
        H = m.config.hidden_dim
        states, unfinished, toks, lps, masks = [], [], [[] for _ in range(M)], [[] for _ in range(M)], [[] for _ in range(M)]
        last_global = lambda: self.ctx.global_rank(pipe=self.pp - 1, data=self.ctx.dp_rank, model=self.ctx.tp_rank)
        first_global = lambda: self.ctx.global_rank(pipe=0, data=self.ctx.dp_rank, model=self.ctx.tp_rank)
        sends = []
        # PyTorch sampling path (CPU tensors) under TP: the ranks of the last stage must draw the same tokens (generation.generate)
        sample_gen = m.shared_generator(dev) if (self.ctx.tp_size > 1 and not g.greedy and not gen.fused_sampler_ok(dev, m.config.vocab_size)) else None
        # ---- prefill, micro-batch by micro-batch
        for i, (ids, cu, mx, T) in enumerate(metas):
            B = cu.numel() - 1
            lens = cu[1:] - cu[:-1]
            kv = []
            hidden = None
            if not self.first:
                buf, h = self._irecv(self._hidden_shape(T), m.dtype, self.prev)
                h.wait()
                hidden = buf
            out = m(input_ids=ids, cu_seqlens=cu, max_seqlen=mx, kv_sink=kv) if self.first else \
                m(hidden=hidden, cu_seqlens=cu, max_seqlen=mx, kv_sink=kv)
            st = gen.DecodeState(m, B, mx + g.max_new_tokens)
            st.fill_from_prefill(kv, cu, lens)
            states.append(st)
            unfinished.append(torch.ones(B, dtype=torch.bool, device=dev))
            if self.last:
                logits = gen._final_logits(m, out.hidden.index_select(0, (cu[1:] - 1).long()))
                nxt, lp, mb_, unfinished[i] = gen.genstep(logits, g, 0, eos_id, pad_id, unfinished[i], sample_gen)
                toks[i].append(nxt); lps[i].append(lp); masks[i].append(mb_)
                if self.pp > 1:
                    sends.append((nxt, dist.isend(nxt, first_global())))
            else:
                x = out.contiguous()
                sends.append((x, dist.isend(x, self.next)))
        # ---- decode steps.  Each stage captures ONE CUDA graph per micro-batch (the KV-cache addresses differ per micro-batch):
        # its blocks' decode step, the cache-length bump and, on the last stage, the LM head; a token round of a stage is then
        # [receive into a static buffer] -> graph replay -> [sample / send] (reference: pipe_runner.py:422-445, one graph per
        # micro-batch for the same reason)
        use_graph = g.use_cuda_graph and dev.type == "cuda"

        def stage_step(st):
            h = m.decode_step(st.input_ids if self.first else None, st.k, st.v, st.cache_lens, hidden=None if self.first else st.hidden_in)
            out = gen._final_logits(m, h) if self.last else h
            st.cache_lens += 1
            return out

        if use_graph and g.max_new_tokens > 1:
            from realhf_b200.ops import launches
            for st in states:
                lens_backup = st.cache_lens.clone()
                side = torch.cuda.Stream(dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    stage_step(st)  # warm-up outside capture (allocations, lazy inits); its KV write is overwritten by the real step
                torch.cuda.current_stream(dev).wait_stream(side)
                st.cache_lens.copy_(lens_backup)
                st.graph = torch.cuda.CUDAGraph()
                launches.begin_capture()
                with torch.cuda.graph(st.graph):
                    st.out = stage_step(st)
                st.graph_launches = launches.end_capture()
        last_send = [None] * M  # the static output buffer of a micro-batch may only be rewritten once its previous send completed
        for step in range(1, g.max_new_tokens):
            for i in range(M):
                st = states[i]
                B = st.B
                if self.first:
                    dist.recv(st.input_ids, last_global())
                else:
                    dist.recv(st.hidden_in, self.prev)
                if use_graph:
                    if last_send[i] is not None:
                        last_send[i].wait()
                    st.graph.replay()
                    launches.count_replay(st.graph_launches)
                    out = st.out
                else:
                    out = stage_step(st)
                if self.last:
                    nxt, lp, mb_, unfinished[i] = gen.genstep(out, g, step, eos_id, pad_id, unfinished[i], sample_gen)
                    toks[i].append(nxt); lps[i].append(lp); masks[i].append(mb_)
                    if step + 1 < g.max_new_tokens:
                        sends.append((nxt, dist.isend(nxt, first_global())))
                else:
                    x = out if use_graph else out.contiguous()
                    hnd = dist.isend(x, self.next)
                    last_send[i] = hnd
                    sends.append((x, hnd))
            if len(sends) > 4 * M:
                for _, hnd in sends[: -2 * M]:
                    hnd.wait()
                sends = sends[-2 * M:]
        for _, hnd in sends:
            hnd.wait()
        m.train(was_training)
        m.sequence_parallel = sp_saved
        if not self.last:
            return None
        outs = []
        for i in range(M):
            tokens = torch.stack(toks[i], 1)
            n_gen = tokens.shape[1]
            B = tokens.shape[0]
            if eos_id is not None:
                is_eos = tokens == eos_id
                first = torch.where(is_eos.any(1), is_eos.float().argmax(1) + 1, torch.full((B,), n_gen, device=dev))
                no_eos = ~is_eos.any(1)
            else:
                first = torch.full((B,), n_gen, device=dev)
                no_eos = torch.ones(B, dtype=torch.bool, device=dev)
            outs.append(gen.GenerationOutput(tokens, torch.stack(lps[i], 1), torch.stack(masks[i], 1) if masks[i][0] is not None else None,
                                             first.long(), no_eos))
        return outs
