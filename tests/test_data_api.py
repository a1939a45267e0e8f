"""SequenceSample / DFG / datasets / datapack unit tests (parity: reference tests/data/*)."""
import pickle
import random
import sys, os

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fixtures  # noqa: E402

from realhf_b200.api.config import ModelInterfaceAbstraction, ModelInterfaceType, ModelName, ModelShardID
from realhf_b200.api.data import (DatasetUtility, PackedDataLoader, SequenceSample, make_dataset)
from realhf_b200.api.dfg import MFCDef, OffloadHook, ParamReallocHook, build_graph, topological_levels
from realhf_b200.base import datapack
from realhf_b200.base.topology import ProcessTopology


def _sample_single(bs, seed=0):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(1, 50, (bs,), generator=g).tolist()
    return SequenceSample.from_default(seqlens=lens, ids=[f"id{i}" for i in range(bs)],
                                       data=dict(packed_input_ids=torch.randint(0, 100, (sum(lens),), generator=g),
                                                 packed_logprobs=torch.randn(sum(lens) - bs, generator=g),
                                                 rewards=torch.randn(bs, generator=g)),
                                       metadata=dict(tag=[f"t{i}" for i in range(bs)]))


def _sample_multi(bs, seed=0):
    """Several sequences per item for one key (e.g. grouped responses), one for another."""
    rng = random.Random(seed)
    sl = {"resp": [[rng.randint(1, 9) for _ in range(rng.randint(1, 4))] for _ in range(bs)],
          "prompt": [[rng.randint(1, 9)] for _ in range(bs)]}
    data = {k: torch.arange(sum(sum(l) for l in v)).float() for k, v in sl.items()}
    return SequenceSample(keys=["resp", "prompt"], ids=list(range(bs)), seqlens=sl, trailing_shapes=dict(resp=(), prompt=()),
                          dtypes=dict(resp=torch.float32, prompt=torch.float32), data=data)


@pytest.mark.parametrize("dp", [1, 2, 3, 4, 8, 15, 16])
@pytest.mark.parametrize("maker", [_sample_single, _sample_multi])
def test_gather_split_identity(dp, maker):
    s = maker(32)
    parts = s.split(dp)
    assert len(parts) == dp and sum(p.bs for p in parts) == s.bs
    back = SequenceSample.gather(parts)
    assert back.ids == s.ids and back.seqlens == s.seqlens
    for k in s.keys:
        assert torch.equal(back.data[k], s.data[k])
    un = SequenceSample.gather(s.unpack())
    for k in s.keys:
        assert torch.equal(un.data[k], s.data[k])
    m = s.meta()
    assert m.data is None and m.seqlens == s.seqlens
    pickle.loads(pickle.dumps(m))


def test_split_is_token_balanced():
    lens = [100] + [1] * 99
    s = SequenceSample.from_default(seqlens=lens, ids=list(range(100)), data=dict(packed_input_ids=torch.zeros(sum(lens), dtype=torch.long)))
    parts = s.split(2)
    tok = [sum(p.flat_seqlens("packed_input_ids")) for p in parts]
    assert max(tok) == 100 and min(tok) == 99


def test_validation_and_remap_update():
    with pytest.raises(ValueError):
        SequenceSample.from_default(seqlens=[3], ids=[0], data=dict(packed_input_ids=torch.zeros(4, dtype=torch.long)))
    with pytest.raises(ValueError):
        SequenceSample.from_default(seqlens=[3, 3], ids=[0, 0], data=dict(packed_input_ids=torch.zeros(6, dtype=torch.long)))
    s = _sample_single(4)
    s.remap_keys_({"packed_input_ids": "seq"})
    assert "seq" in s.keys and "packed_input_ids" not in s.keys and "seq" in s.data
    other = SequenceSample.from_default(seqlens=s.flat_seqlens("seq"), ids=s.ids, data=dict(values=torch.zeros(s.total_len("seq"))))
    s.update_(other)
    assert "values" in s.keys


def test_partition_native_matches_python():
    rng = np.random.RandomState(0)
    for _ in range(50):
        n = rng.randint(1, 200)
        k = rng.randint(1, min(n, 16) + 1)
        ms = rng.randint(1, max(2, n // k + 1))
        if n < k * ms:
            continue
        nums = rng.randint(1, 1000, size=n).tolist()
        a = datapack.partition_balanced(nums, k, ms)
        b = datapack._partition_balanced_py(nums, k, ms)
        assert a == b
        sizes = [a[i + 1] - a[i] for i in range(k)]
        assert all(s >= ms for s in sizes) and a[0] == 0 and a[-1] == n
        # optimal max part: no contiguous k-split does better (brute force on small n)
    order, diff = datapack.reorder_to_balanced_batches([5, 1, 3, 2, 8, 7, 4, 4], 2)
    assert sorted(order.tolist()) == list(range(8))
    assert datapack.merge_intervals([(0, 2), (2, 5), (7, 9), (9, 10)]) == [(0, 5), (7, 10)]


def _ppo_rpcs():
    A = lambda t: ModelInterfaceAbstraction(t)
    T = ModelInterfaceType
    return [
        MFCDef("actor_gen", 8, T.GENERATE, A("ppo_actor"), "actor", input_keys=("packed_prompts",),
               output_keys=("seq_no_eos_mask", "packed_input_ids", "packed_logprobs", "prompt_mask")),
        MFCDef("rew_inf", 8, T.INFERENCE, A("paired_rw"), "reward", input_keys=("packed_input_ids",), output_keys=("rewards",)),
        MFCDef("ref_inf", 8, T.INFERENCE, A("ppo_actor"), "ref", input_keys=("packed_input_ids",), output_keys=("packed_ref_logprobs",)),
        MFCDef("critic_inf", 8, T.INFERENCE, A("ppo_critic"), "critic", input_keys=("packed_input_ids", "seq_no_eos_mask"), output_keys=("values",)),
        MFCDef("actor_train", 8, T.TRAIN_STEP, A("ppo_actor"), "actor",
               input_keys=("packed_input_ids", "packed_logprobs", "packed_ref_logprobs", "rewards", "values", "prompt_mask", "seq_no_eos_mask")),
        MFCDef("critic_train", 8, T.TRAIN_STEP, A("ppo_critic"), "critic",
               input_keys=("packed_input_ids", "packed_logprobs", "packed_ref_logprobs", "rewards", "values", "prompt_mask", "seq_no_eos_mask")),
    ]


def test_dfg_ppo_structure_and_pickle():
    rpcs = _ppo_rpcs()
    G = build_graph(rpcs)
    by = {r.name: r for r in rpcs}
    assert by["actor_gen"].is_src and not by["actor_gen"].is_dst
    assert by["actor_train"].is_dst and by["critic_train"].is_dst
    assert {p.name for p in by["actor_train"].parents} == {"actor_gen", "rew_inf", "ref_inf", "critic_inf"}
    assert topological_levels(G)[0] == ["actor_gen"] and set(topological_levels(G)[1]) == {"rew_inf", "ref_inf", "critic_inf"}
    assert G.graph["dataset_keys"] == ["packed_prompts"]
    assert by["actor_gen"].data_producers["rewards"] == ModelName("reward", 0)
    assert not by["actor_gen"].is_dst_of_model_role and by["actor_train"].is_dst_of_model_role
    by["actor_gen"].add_pre_hook(ParamReallocHook(source=ModelName("actor", 0)))
    by["ref_inf"].add_post_hook(OffloadHook())
    with pytest.raises(ValueError):
        by["ref_inf"].add_pre_hook(OffloadHook())
    clone = pickle.loads(pickle.dumps(rpcs))
    assert [r.name for r in clone] == [r.name for r in rpcs] and clone[0].children[0].name in {"rew_inf", "ref_inf", "critic_inf"}
    with pytest.raises(ValueError):
        build_graph(rpcs + [MFCDef("dup", 8, ModelInterfaceType.INFERENCE, ModelInterfaceAbstraction("x"), "m", output_keys=("values",))])


def test_remax_like_key_remaps_build_a_dag():
    A = lambda t: ModelInterfaceAbstraction(t)
    T = ModelInterfaceType
    rpcs = [
        MFCDef("sample_gen", 4, T.GENERATE, A("g"), "actor", input_keys=("packed_prompts",), output_keys=("packed_input_ids", "prompt_mask")),
        MFCDef("greedy_gen", 4, T.GENERATE, A("g"), "actor", input_keys=("packed_prompts",), output_keys=("greedy_packed_input_ids",),
               output_key_remap={"packed_input_ids": "greedy_packed_input_ids"}),
        MFCDef("rew", 4, T.INFERENCE, A("r"), "reward", input_keys=("packed_input_ids",), output_keys=("rewards",)),
        MFCDef("greedy_rew", 4, T.INFERENCE, A("r"), "reward", input_keys=("greedy_packed_input_ids",), output_keys=("greedy_rewards",),
               input_key_remap={"greedy_packed_input_ids": "packed_input_ids"}, output_key_remap={"rewards": "greedy_rewards"}),
        MFCDef("train", 4, T.TRAIN_STEP, A("t"), "actor", input_keys=("packed_input_ids", "rewards", "greedy_rewards", "prompt_mask")),
    ]
    G = build_graph(rpcs)
    assert len(topological_levels(G)) == 3


def test_shard_id_roundtrip_and_topology():
    topo = ProcessTopology(2, 2, 2)
    seen = set()
    for r in range(8):
        sid = ModelShardID.from_parallelism_rank(ModelName("actor", 1), topo, r)
        assert ModelShardID.parse(repr(sid)) == sid and sid.parallelism_rank == r
        seen.add((sid.pp_rank, sid.dp_rank, sid.tp_rank))
    assert len(seen) == 8
    assert topo.get_axis_comm_lists("model") == [[0, 1], [2, 3], [4, 5], [6, 7]]  # tp = consecutive ranks
    assert topo.get_axis_comm_lists("pipe") == [[0, 4], [1, 5], [2, 6], [3, 7]]


def test_datasets_through_packed_loader(tmp_path):
    import realhf_b200.datasets  # noqa: F401
    tok, words = fixtures.make_tokenizer(str(tmp_path / "tok"))
    rng = random.Random(0)
    mk = lambda n: " ".join(rng.choice(words) for _ in range(n))
    prompts = lambda: [dict(id=i, prompt=mk(rng.randint(2, 9))) for i in range(40)]
    pa = lambda: [dict(id=i, prompt=mk(4) + " ", answer=mk(rng.randint(2, 9))) for i in range(40)]
    pairs = lambda: [dict(id=i, prompt=mk(4) + " ", pos_answers=[mk(3), mk(5)], neg_answers=[mk(4), mk(2)]) for i in range(40)]
    from realhf_b200.api.config import DatasetAbstraction
    seen_ids = set()
    for dp_rank in range(2):
        ds = make_dataset(DatasetAbstraction("prompt", args=dict(max_length=8, dataset_builder=prompts)), 1, dp_rank, 2, tok)
        assert len(ds) == 20
        for batch in PackedDataLoader(ds, batch_size=7, shuffle=False):
            batch.validate()
            assert max(batch.flat_seqlens("packed_prompts")) <= 8
            seen_ids |= set(batch.ids)
    assert len(seen_ids) == 40
    ds = make_dataset(DatasetAbstraction("prompt_answer", args=dict(max_length=32, dataset_builder=pa)), 1, 0, 1, tok)
    b = next(iter(PackedDataLoader(ds, batch_size=8)))
    b.validate()
    assert b.data["prompt_mask"].dtype == torch.bool and b.data["prompt_mask"].any() and not b.data["prompt_mask"].all()
    ds = make_dataset(DatasetAbstraction("rw_pair", args=dict(max_length=32, dataset_builder=pairs, max_pairs_per_prompt=2)), 1, 0, 1, tok)
    b = next(iter(PackedDataLoader(ds, batch_size=4)))
    b.validate()
    assert all(len(l) == 4 for l in b.seqlens["packed_input_ids"])
    parts = b.meta().split(2)
    assert sum(p.bs for p in parts) == 4


def test_dataset_disk_cache(tmp_path):
    """make_dataset(cache_root=...) reuses the tokenised shard and rebuilds it when the source file changes."""
    import json
    import types

    import realhf_b200.datasets  # noqa: F401
    from realhf_b200.api import data as data_api
    from realhf_b200.api.config import DatasetAbstraction

    class Tok:
        name_or_path = "toy"
        eos_token_id, pad_token_id, eos_token = 1, 0, "</s>"
        calls = 0

        def __call__(self, texts, **kw):
            Tok.calls += 1
            ids = [[3 + (ord(c) % 50) for c in t][: kw.get("max_length") or 10 ** 9] for t in texts]
            return types.SimpleNamespace(input_ids=ids) if False else {"input_ids": ids, "length": [len(i) for i in ids]}

    path = tmp_path / "p.jsonl"
    path.write_text("\n".join(json.dumps({"id": i, "prompt": "hello world %d" % i}) for i in range(8)))
    cfg = DatasetAbstraction("prompt", args=dict(dataset_path=str(path), max_length=16))
    d1 = data_api.make_dataset(cfg, 1, 0, 1, Tok(), cache_root=str(tmp_path / "cache"))
    n_calls = Tok.calls
    d2 = data_api.make_dataset(cfg, 1, 0, 1, Tok(), cache_root=str(tmp_path / "cache"))
    assert Tok.calls == n_calls, "second construction must come from the cache"
    assert len(d1) == len(d2) == 8 and d2[0].ids == d1[0].ids
    path.write_text(path.read_text() + "\n" + json.dumps({"id": 99, "prompt": "one more"}))
    d3 = data_api.make_dataset(cfg, 1, 0, 1, Tok(), cache_root=str(tmp_path / "cache"))
    assert len(d3) == 9 and Tok.calls > n_calls


def test_split_relaxes_an_infeasible_min_size():
    """`min_size` expresses room for micro-batches downstream; a batch that is too small for it is split as evenly as it
    can be instead of raising (the engines clamp their micro-batch count to the number of sequences)."""
    import torch
    from realhf_b200.api.data import SequenceSample
    lens = [5, 7, 3, 9, 4, 6, 8, 2]
    s = SequenceSample.from_default(ids=list(range(8)), seqlens=lens, data=dict(packed_input_ids=torch.arange(sum(lens))))
    parts = s.split(2, min_size=8)          # 2 parts of >= 8 from 8 items is impossible: falls back to >= 4
    assert [p.bs for p in parts] == [4, 4] and sum(p.bs for p in parts) == 8
    parts = s.split(3, min_size=100)
    assert sorted(p.bs for p in parts) == [2, 3, 3] or all(p.bs >= 2 for p in parts)
    assert torch.equal(torch.cat([p.data["packed_input_ids"] for p in parts]), s.data["packed_input_ids"])
    import pytest
    with pytest.raises(ValueError):
        s.split(9)                           # more parts than sequences stays an error
