"""Model check of the mbarrier protocol of the tcgen05 attention kernels (csrc/attn_fwd_tcgen05.cu, attn_bwd_tcgen05.cu).

The three roles of a CTA (TMA producer, MMA issuer, softmax threads) are written as coroutines that perform exactly the waits,
arrives, expect_tx and commits of the kernels (same barrier names, same parities); TMA completions and tensor-pipe completions
are asynchronous events.  A randomised scheduler interleaves everything over many seeds and checks
  * no deadlock (every role finishes) and no parity aliasing (a wait never needs a phase that is two completions away),
  * no buffer hazard: a shared-memory stage / TMEM buffer is never overwritten before its consumer released it and never
    read before it was produced.
This is the part of the kernels that cannot be unit-tested without hardware and whose failure mode on hardware is a hang."""
import random

import pytest


class Bar:
    def __init__(self, count):
        self.init, self.count, self.tx, self.phase = count, count, 0, 0

    def _check(self):
        if self.count == 0 and self.tx == 0:
            self.phase += 1
            self.count = self.init

    def arrive(self):
        assert self.count > 0, "more arrivals than the barrier expects in one phase"
        self.count -= 1
        self._check()

    def expect_tx(self, n):      # mbarrier.arrive.expect_tx
        self.tx += n
        self.arrive()

    def complete_tx(self, n):
        self.tx -= n
        assert self.tx >= 0
        self._check()

    def passed(self, parity):   # mbarrier.try_wait.parity
        return (self.phase & 1) != parity


class Sim:
    """Roles are generators yielding ("wait", bar_name, parity); everything else they do is immediate."""

    def __init__(self, seed, bars):
        self.rng = random.Random(seed)
        self.bars = {k: Bar(c) for k, c in bars.items()}
        self.pipe = []      # tensor pipe: in-order list of callables (MMA effects, commits)
        self.tma = []       # outstanding TMA loads: callables, complete in any order
        self.state = {}     # buffer name -> "free" | "loading" | "full"

    # ---- buffer hazard tracking
    def expect(self, buf, want, what):
        got = self.state.get(buf, "free")
        assert got == want, f"{what}: buffer {buf} is {got}, expected {want}"

    def set(self, buf, st):
        self.state[buf] = st

    def run(self, roles):
        blocked = {name: None for name in roles}
        gens = dict(roles)
        steps = 0
        while gens or self.pipe or self.tma:
            steps += 1
            assert steps < 200000, "livelock"
            choices = []
            for name, g in gens.items():
                w = blocked[name]
                if w is None or self.bars[w[0]].passed(w[1]):
                    choices.append(("role", name))
            if self.pipe:
                choices.append(("pipe", None))
            for i in range(len(self.tma)):
                choices.append(("tma", i))
            if not choices:
                waiting = {n: blocked[n] for n in gens}
                raise AssertionError(f"deadlock: {waiting}, phases { {k: b.phase for k, b in self.bars.items()} }")
            kind, arg = self.rng.choice(choices)
            if kind == "pipe":
                self.pipe.pop(0)()
            elif kind == "tma":
                self.tma.pop(arg)()
            else:
                blocked[arg] = None
                try:
                    req = next(gens[arg])
                    assert req[0] == "wait"
                    blocked[arg] = (req[1], req[2])
                except StopIteration:
                    del gens[arg]


# ------------------------------------------------------------------------------------------------ forward kernel
def fwd_protocol(sim: Sim, n_kv: int):
    B = sim.bars
    KV = 100  # bytes of one K or V stage (any positive number)

    def producer():
        B["Q_FULL"].expect_tx(50)
        sim.expect("Q", "free", "Q load"); sim.set("Q", "loading")
        sim.tma.append(lambda: (sim.set("Q", "full"), B["Q_FULL"].complete_tx(50)))
        for j in range(n_kv):
            st, par = j & 1, (j >> 1) & 1
            for nm in ("K", "V"):
                yield ("wait", f"{nm}_EMPTY{st}", par ^ 1)
                B[f"{nm}_FULL{st}"].expect_tx(KV)
                sim.expect(f"{nm}{st}", "free", f"{nm} load of tile {j}"); sim.set(f"{nm}{st}", "loading")
                sim.tma.append(lambda nm=nm, st=st: (sim.set(f"{nm}{st}", "full"), B[f"{nm}_FULL{st}"].complete_tx(KV)))

    def mma():
        def issue_s(j):
            st, par = j & 1, (j >> 1) & 1
            yield ("wait", f"K_FULL{st}", par)
            yield ("wait", f"S_EMPTY{st}", par ^ 1)

            def exec_s():
                sim.expect("Q", "full", "S mma"); sim.expect(f"K{st}", "full", f"S mma {j}")
                sim.expect(f"S{st}", "free", f"S mma {j} overwrites S buffer"); sim.set(f"S{st}", "full")
            sim.pipe.append(exec_s)
            sim.pipe.append(lambda: (sim.set(f"K{st}", "free"), B[f"K_EMPTY{st}"].arrive()))
            sim.pipe.append(lambda: B[f"S_FULL{st}"].arrive())
        yield ("wait", "Q_FULL", 0)
        yield from issue_s(0)
        for j in range(n_kv):
            if j + 1 < n_kv:
                yield from issue_s(j + 1)
            st, par = j & 1, (j >> 1) & 1
            yield ("wait", "P_FULL", j & 1)
            yield ("wait", f"V_FULL{st}", par)
            yield ("wait", f"O_EMPTY{st}", par ^ 1)

            def exec_pv(st=st, j=j):
                sim.expect("P", "full", f"PV mma {j}"); sim.expect(f"V{st}", "full", f"PV mma {j}")
                sim.expect(f"T{st}", "free", f"PV mma {j} overwrites T buffer"); sim.set(f"T{st}", "full")
            sim.pipe.append(exec_pv)
            sim.pipe.append(lambda st=st: (sim.set(f"V{st}", "free"), B[f"V_EMPTY{st}"].arrive()))
            sim.pipe.append(lambda: (sim.set("P", "free"), B["P_EMPTY"].arrive()))
            sim.pipe.append(lambda st=st: B[f"O_FULL{st}"].arrive())

    def softmax(w):
        def fold(t):
            b = t & 1
            yield ("wait", f"O_FULL{b}", (t >> 1) & 1)
            sim.expect(f"T{b}", "full", f"fold {t}")
            sim.state[f"T{b}_readers"] = sim.state.get(f"T{b}_readers", 0) + 1
            if sim.state[f"T{b}_readers"] == 4:   # the buffer is free once all four warps arrived
                sim.state[f"T{b}_readers"] = 0
                sim.set(f"T{b}", "free")
            B[f"O_EMPTY{b}"].arrive()
        for j in range(n_kv):
            b = j & 1
            yield ("wait", f"S_FULL{b}", (j >> 1) & 1)
            sim.expect(f"S{b}", "full", f"softmax {j} pass 1")
            if j > 0:
                yield ("wait", "P_EMPTY", (j - 1) & 1)
            sim.expect(f"S{b}", "full", f"softmax {j} pass 2")
            sim.expect("P", "free" if w == 0 or sim.state.get("P_writers", 0) == 0 else "free", f"P write {j}")
            sim.state["P_writers"] = sim.state.get("P_writers", 0) + 1
            sim.state[f"S{b}_readers"] = sim.state.get(f"S{b}_readers", 0) + 1
            if sim.state[f"S{b}_readers"] == 4:
                sim.state[f"S{b}_readers"] = 0
                sim.set(f"S{b}", "free")
            if sim.state["P_writers"] == 4:
                sim.state["P_writers"] = 0
                sim.set("P", "full")
            B[f"S_EMPTY{b}"].arrive()
            B["P_FULL"].arrive()
            if j > 0:
                yield from fold(j - 1)
        yield from fold(n_kv - 1)

    roles = {"producer": producer(), "mma": mma()}
    for w in range(4):
        roles[f"softmax{w}"] = softmax(w)
    return roles


FWD_BARS = {"Q_FULL": 1, "P_FULL": 4, "P_EMPTY": 1}
for _st in (0, 1):
    FWD_BARS.update({f"K_FULL{_st}": 1, f"K_EMPTY{_st}": 1, f"V_FULL{_st}": 1, f"V_EMPTY{_st}": 1, f"S_FULL{_st}": 1,
                     f"S_EMPTY{_st}": 4, f"O_FULL{_st}": 1, f"O_EMPTY{_st}": 4})


@pytest.mark.parametrize("n_kv", [1, 2, 3, 4, 5, 8])
def test_forward_protocol_has_no_deadlock_or_buffer_hazard(n_kv):
    for seed in range(150):
        sim = Sim(seed, FWD_BARS)
        sim.run(fwd_protocol(sim, n_kv))
        assert all(v in ("free", "full", 0) for v in sim.state.values())


# ------------------------------------------------------------------------------------------------ backward kernel (both passes)
def bwd_protocol(sim: Sim, n_steps: int):
    B = sim.bars

    def producer():
        B["R_FULL"].expect_tx(60)
        sim.set("R", "loading")
        sim.tma.append(lambda: (sim.set("R", "full"), B["R_FULL"].complete_tx(60)))
        for t in range(n_steps):
            b, par = t & 1, (t >> 1) & 1
            yield ("wait", f"X_EMPTY{b}", par ^ 1)
            B[f"X_FULL{b}"].expect_tx(40)
            sim.expect(f"X{b}", "free", f"X load of step {t}"); sim.set(f"X{b}", "loading")
            done = {"n": 0}

            def part(b=b, done=done):   # four TMA boxes complete independently
                done["n"] += 1
                if done["n"] == 4:
                    sim.set(f"X{b}", "full")
                B[f"X_FULL{b}"].complete_tx(10)
            for _ in range(4):
                sim.tma.append(part)

    def mma():
        def issue_sdp(t):
            b, par = t & 1, (t >> 1) & 1
            yield ("wait", f"X_FULL{b}", par)
            yield ("wait", f"S_EMPTY{b}", par ^ 1)

            def exec_sdp():
                sim.expect("R", "full", "S/dP mma"); sim.expect(f"X{b}", "full", f"S/dP mma {t}")
                sim.expect(f"S{b}", "free", f"S/dP mma {t} overwrites the S/dP buffers"); sim.set(f"S{b}", "full")
            sim.pipe.append(exec_sdp)
            sim.pipe.append(lambda: B[f"S_FULL{b}"].arrive())
        yield ("wait", "R_FULL", 0)
        yield from issue_sdp(0)
        for t in range(n_steps):
            if t + 1 < n_steps:
                yield from issue_sdp(t + 1)
            b, par = t & 1, (t >> 1) & 1
            yield ("wait", f"P_FULL{b}", par)

            def exec_acc(b=b, t=t):
                sim.expect(f"P{b}", "full", f"accumulating mma {t}"); sim.expect(f"X{b}", "full", f"accumulating mma {t}")
            sim.pipe.append(exec_acc)
            sim.pipe.append(lambda b=b: (sim.set(f"X{b}", "free"), B[f"X_EMPTY{b}"].arrive()))
            sim.pipe.append(lambda b=b: (sim.set(f"P{b}", "free"), B[f"P_EMPTY{b}"].arrive()))
        sim.pipe.append(lambda: B["ACC_FULL"].arrive())

    def threads(w):
        for t in range(n_steps):
            b, par = t & 1, (t >> 1) & 1
            yield ("wait", f"S_FULL{b}", par)
            yield ("wait", f"P_EMPTY{b}", par ^ 1)
            sim.expect(f"S{b}", "full", f"threads read S/dP {t}")
            sim.expect(f"P{b}", "free", f"threads write P buffers {t}")
            k = f"cnt{b}"
            sim.state[k] = sim.state.get(k, 0) + 1
            if sim.state[k] == 4:
                sim.state[k] = 0
                sim.set(f"S{b}", "free")
                sim.set(f"P{b}", "full")
            B[f"S_EMPTY{b}"].arrive()
            B[f"P_FULL{b}"].arrive()
        yield ("wait", "ACC_FULL", 0)

    roles = {"producer": producer(), "mma": mma()}
    for w in range(4):
        roles[f"threads{w}"] = threads(w)
    return roles


BWD_BARS = {"R_FULL": 1, "ACC_FULL": 1}
for _b in (0, 1):
    BWD_BARS.update({f"X_FULL{_b}": 1, f"X_EMPTY{_b}": 1, f"S_FULL{_b}": 1, f"S_EMPTY{_b}": 4, f"P_FULL{_b}": 4, f"P_EMPTY{_b}": 1})


@pytest.mark.parametrize("n_steps", [1, 2, 3, 4, 7, 12])
def test_backward_protocol_has_no_deadlock_or_buffer_hazard(n_steps):
    for seed in range(150):
        sim = Sim(seed, BWD_BARS)
        sim.run(bwd_protocol(sim, n_steps))


def test_the_model_checker_catches_a_wrong_parity():
    """Sanity of the checker itself: a producer that waits on the wrong phase is reported as a hazard or a deadlock."""
    def broken(sim, n_steps):
        roles = bwd_protocol(sim, n_steps)

        def bad_producer():
            B = sim.bars
            B["R_FULL"].expect_tx(60)
            sim.set("R", "loading")
            sim.tma.append(lambda: (sim.set("R", "full"), B["R_FULL"].complete_tx(60)))
            for t in range(n_steps):
                b = t & 1
                yield ("wait", f"X_EMPTY{b}", 1)   # BUG: constant parity instead of ((t >> 1) & 1) ^ 1
                B[f"X_FULL{b}"].expect_tx(40)
                sim.expect(f"X{b}", "free", f"X load of step {t}"); sim.set(f"X{b}", "loading")
                sim.tma.append(lambda b=b: (sim.set(f"X{b}", "full"), B[f"X_FULL{b}"].complete_tx(40)))
        roles["producer"] = bad_producer()
        return roles
    failures = 0
    for seed in range(40):
        sim = Sim(seed, BWD_BARS)
        try:
            sim.run(broken(sim, 6))
        except AssertionError:
            failures += 1
    assert failures > 0
