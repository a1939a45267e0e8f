"""RLHF functionals against literal re-implementations of the textbook formulas (CPU): PPO clipped losses, DPO loss, KL-shaped
rewards + GAE on packed sequences, masked normalisation, running mean/std, KL controllers."""
import math
import os

import pytest
import torch

from realhf_b200.interfaces import functional as IF


def test_actor_loss_matches_formula():
    torch.manual_seed(0)
    n = 50
    lp, old, adv = torch.randn(n) * 0.1 - 2, torch.randn(n) * 0.1 - 2, torch.randn(n)
    mask = torch.rand(n) > 0.3
    loss, st = IF.actor_loss_fn(lp, old, adv, 0.2, mask)
    tot, cnt, clipped = 0.0, 0, 0
    for i in range(n):
        if not mask[i]:
            continue
        r = math.exp(float(lp[i] - old[i]))
        l1, l2 = -float(adv[i]) * r, -float(adv[i]) * min(max(r, 0.8), 1.2)
        tot += max(l1, l2)
        clipped += l1 < l2
        cnt += 1
    assert abs(float(loss) - tot / cnt) < 1e-5 and abs(float(st["clip_ratio"]) - clipped / cnt) < 1e-6


@pytest.mark.parametrize("kind", ["mse", "huber"])
def test_critic_loss_matches_formula(kind):
    torch.manual_seed(1)
    n = 40
    v, old, tgt = torch.randn(n), torch.randn(n), torch.randn(n) * 2
    loss, _ = IF.critic_loss_fn(v, old, tgt, 0.2, None, kind)
    f = (lambda a, b: 0.5 * (a - b) ** 2) if kind == "mse" else (lambda a, b: 0.5 * (a - b) ** 2 if abs(a - b) <= 10 else 10 * (abs(a - b) - 5))
    ref = sum(max(f(float(v[i]), float(tgt[i])), f(float(old[i]) + min(max(float(v[i] - old[i]), -0.2), 0.2), float(tgt[i]))) for i in range(n)) / n
    assert abs(float(loss) - ref) < 1e-5


def test_dpo_loss_matches_formula():
    torch.manual_seed(2)
    pi, ref = torch.randn(8) - 5, torch.randn(8) - 5
    loss, pos, neg, kl = IF.dpo_loss(pi, ref, 0.1)
    tot = 0.0
    for i in range(4):
        x = 0.1 * ((float(pi[2 * i]) - float(pi[2 * i + 1])) - (float(ref[2 * i]) - float(ref[2 * i + 1])))
        tot += -math.log(1 / (1 + math.exp(-x)))
    assert abs(float(loss) - tot / 4) < 1e-5
    assert abs(float(pos) - 0.1 * float((pi[0::2] - ref[0::2]).sum())) < 1e-5


def test_packed_rewards_and_gae_matches_python_recursion():
    torch.manual_seed(3)
    seqlens = [5, 3, 8]
    n1 = sum(l - 1 for l in seqlens)
    old_lp, ref_lp = -torch.rand(n1), -torch.rand(n1)
    values = torch.randn(sum(seqlens))
    scores = torch.tensor([0.7, -1.2, 3.0])
    no_eos = torch.tensor([False, True, False])
    kl_ctl, clip, gamma, lam = 0.1, 2.0, 0.99, 0.95
    adv, ret, kl_r, rew = IF.packed_rewards_and_gae(old_lp, ref_lp, scores, values, seqlens, no_eos, kl_ctl, clip, gamma, lam)
    off1 = off = 0
    for s, L in enumerate(seqlens):
        T = L - 1
        v = values[off: off + L].tolist()
        r = [-kl_ctl * float(old_lp[off1 + t] - ref_lp[off1 + t]) for t in range(T)]
        if not bool(no_eos[s]):                                # truncated sequences get no terminal score (reference:
            r[-1] += max(-clip, min(clip, float(scores[s])))  # ppo_functional.py:306-308), they bootstrap from the last value
        nxt_last = v[L - 1] if bool(no_eos[s]) else 0.0        # truncated sequences bootstrap from the last value
        a = [0.0] * T
        last = 0.0
        for t in reversed(range(T)):
            nv = v[t + 1] if t < T - 1 else nxt_last
            delta = r[t] + gamma * nv - v[t]
            last = delta + gamma * lam * last
            a[t] = last
        torch.testing.assert_close(adv[off1: off1 + T], torch.tensor(a), atol=1e-5, rtol=1e-5)
        torch.testing.assert_close(ret[off1: off1 + T], torch.tensor(a) + torch.tensor(v[:T]), atol=1e-5, rtol=1e-5)
        off1 += T
        off += L


def test_masked_normalization_and_running_stats():
    torch.manual_seed(4)
    x, m = torch.randn(100) * 3 + 1, torch.rand(100) > 0.5
    y = IF.masked_normalization(x, m)
    assert abs(float(y[m].mean())) < 1e-4 and abs(float(y[m].std(unbiased=False)) - 1) < 1e-2
    ma = IF.MovingAverageRunningMeanStd()
    chunks = [torch.randn(50) * 2 + 3 for _ in range(4)]
    for c in chunks:
        ma.update(c)
    allx = torch.cat(chunks)
    mean, std = ma.mean_std()
    assert abs(float(mean) - float(allx.mean())) < 1e-5 and abs(float(std) - float(allx.std(unbiased=False))) < 1e-4
    torch.testing.assert_close(ma.denormalize(ma.normalize(allx)), allx, atol=1e-4, rtol=1e-4)
    ex = IF.ExponentialRunningMeanStd(beta=0.9)
    for c in chunks:
        ex.update(c)
    m2, s2 = ex.mean_std()
    assert abs(float(m2) - 3) < 0.5 and abs(float(s2) - 2) < 0.5  # debiased: close to the true moments after 4 updates


def test_kl_controllers():
    fixed = IF.FixedKLController(0.1)
    fixed.update(5.0, 10)
    assert fixed.value == 0.1
    ad = IF.AdaptiveKLController(0.1, target=6.0, horizon=100)
    ad.update(12.0, 10)      # KL twice the target -> coefficient grows (clipped proportional error 0.2 * n / horizon)
    assert abs(ad.value - 0.1 * (1 + 0.2 * 10 / 100)) < 1e-9
    ad.update(0.0, 10)
    assert ad.value < 0.1 * (1 + 0.2 * 10 / 100)


def test_interface_state_roundtrip_kl_controller_and_value_normaliser():
    """What a recover run restores: the adaptive KL coefficient and the running value statistics."""
    from realhf_b200.interfaces import functional as IF
    from realhf_b200.interfaces.ppo import PPOActorInterface, PPOCriticInterface
    a = PPOActorInterface(adaptive_kl_ctl=True, kl_ctl=0.1, adaptive_kl_target=6.0, adaptive_kl_horizon=100)
    a.kl_adapter.update(12.0, n_steps=10)
    assert a.kl_adapter.value != 0.1
    b = PPOActorInterface(adaptive_kl_ctl=True, kl_ctl=0.1, adaptive_kl_target=6.0, adaptive_kl_horizon=100)
    b.load_state_dict(a.state_dict())
    assert b.kl_adapter.value == a.kl_adapter.value
    for kind in ("exp", "ma"):
        c = PPOCriticInterface(value_norm=True, value_norm_type=kind)
        x = torch.randn(257) * 3 + 5
        c.rms.update(x)
        c.rms.update(x * 0.5)
        d = PPOCriticInterface(value_norm=True, value_norm_type=kind)
        d.load_state_dict(c.state_dict())
        y = torch.randn(33)
        torch.testing.assert_close(d.rms.normalize(y), c.rms.normalize(y))
        c.rms.update(y)
        d.rms.update(y)    # updates keep working on restored (host) tensors
        torch.testing.assert_close(d.rms.denormalize(y), c.rms.denormalize(y))
    assert isinstance(c.rms, IF.MovingAverageRunningMeanStd)


def test_interface_states_are_written_with_the_recover_states_and_restored(tmp_path):
    from realhf_b200.interfaces.ppo import PPOActorInterface, PPOCriticInterface
    from realhf_b200.interfaces.basic import SFTInterface
    from realhf_b200.system.model_worker import load_interface_state, save_interface_states
    a = PPOActorInterface(adaptive_kl_ctl=True, kl_ctl=0.1)
    a.kl_adapter.update(20.0, n_steps=50)
    c = PPOCriticInterface(value_norm=True)
    c.rms.update(torch.arange(10.0))
    root = str(tmp_path / "ckpt")
    save_interface_states({"actor_train": a, "critic_train": c, "sft": SFTInterface()}, root, tag="3")
    assert sorted(os.listdir(root)) == ["interface_actor_train.pt", "interface_critic_train.pt"]  # stateless interfaces write nothing
    a2, c2 = PPOActorInterface(adaptive_kl_ctl=True, kl_ctl=0.1), PPOCriticInterface(value_norm=True)
    assert load_interface_state("actor_train", a2, root) and load_interface_state("critic_train", c2, root)
    assert not load_interface_state("ref_inf", PPOActorInterface(), root)
    assert a2.kl_adapter.value == a.kl_adapter.value
    torch.testing.assert_close(c2.rms.normalize(torch.ones(3)), c.rms.normalize(torch.ones(3)))
