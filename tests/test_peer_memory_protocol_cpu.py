"""Model checks of the flag protocols of the peer-memory kernels (csrc/comm_common.cuh, nvls.cu, gemm_tcgen05.cu stream-K).

These are the parts of the multi-GPU / multi-CTA kernels whose failure on hardware is a hang or a silently stale read, and which no
single-process numerics test can reach.  Every actor (a thread block of one rank, a CTA of the stream-K GEMM) is a coroutine that
performs the same flag writes, spins and data accesses as the kernel, in the same order; a randomised scheduler interleaves them
over many seeds and the data carry version numbers, so that
  * a deadlock shows up as "nobody can make progress",
  * a read of data that is not the version the algorithm expects (stale, or overwritten early by a fast peer) is an assertion,
  * counters that must be back at zero for the next CUDA-graph replay are checked at the end.
Each protocol is also run in a deliberately broken variant to show that the checker sees the hazard the real design avoids."""
import random

import pytest


def run_actors(actors, rng, max_steps=2_000_000):
    """actors: list of generators.  A generator yields a zero-argument predicate to wait on (re-evaluated until true) or None to
    just give way.  Returns when all finished; raises on deadlock."""
    waiting = [None] * len(actors)
    alive = list(range(len(actors)))
    for _ in range(max_steps):
        if not alive:
            return
        runnable = [i for i in alive if waiting[i] is None or waiting[i]()]
        assert runnable, f"deadlock: {len(alive)} actors blocked"
        i = rng.choice(runnable)
        try:
            waiting[i] = next(actors[i])
        except StopIteration:
            alive.remove(i)
    raise AssertionError("did not terminate")


# ----------------------------------------------------------------------------------- epoch barrier + two-region all-reduce


class Rank:
    def __init__(self, world, n_blocks):
        self.epoch = [0] * n_blocks                                  # per-block epoch counter in this rank's own pad
        self.flags = [[0] * world for _ in range(n_blocks)]         # flags[block][src_rank], written by the peers
        self.region = [None, None]                                   # version of the partial sums held in each data region
        self.ar_done = -1                                            # last all-reduce kernel of which every block has finished
        self.blocks_done = {}                                        # call -> number of finished blocks


def _i32(x):
    x &= 0xFFFFFFFF
    return x - (1 << 32) if x & 0x80000000 else x


def block_barrier(ranks, r, b, start=0):
    """csrc/comm_common.cuh::block_barrier: advance the block's epoch, store it into slot [b][r] of every peer's pad (release),
    spin until every slot of the own pad reached it (acquire).  Epochs are compared as wrapped 32-bit differences."""
    me = ranks[r]
    me.epoch[b] = (me.epoch[b] + 1) & 0xFFFFFFFF
    e = me.epoch[b]
    for p in ranks:
        p.flags[b][r] = e
        yield None                                                   # the stores to different peers are separate events
    yield lambda: all(_i32(f - e) >= 0 for f in me.flags[b])


def tp_decode_rank(ranks, r, n_calls, n_blocks, regions, pdl_guard, log):
    """One rank's stream of a tensor-parallel decode: GEMM_k writes its partial sums into region k % regions of symmetric memory,
    then the all-reduce + norm kernel (nvls_ar_add_rmsnorm_kernel) waits for the GEMM (griddepcontrol.wait), crosses the epoch
    barrier block by block, and reads region k % regions of EVERY rank through the multicast address.  There is no trailing barrier.
    With programmatic dependent launch GEMM_{k+1} is launched early but writes only after every block of kernel k has finished."""
    me = ranks[r]

    def gemm(k):
        if pdl_guard:
            yield lambda: me.ar_done >= k - 1
        me.region[k % regions] = k
        yield None

    def ar_block(k, b):
        yield lambda: me.region[k % regions] == k                    # pdl_wait: THIS rank's producing GEMM is complete
        yield from block_barrier(ranks, r, b)
        for p_i, p in enumerate(ranks):                              # multimem.ld_reduce touches every rank's copy
            got = p.region[k % regions]
            log.append((k, r, b, p_i, got))
            assert got == k, f"call {k}: rank {r} block {b} read version {got} of rank {p_i}'s region {k % regions}"
            yield None
        me.blocks_done[k] = me.blocks_done.get(k, 0) + 1
        if me.blocks_done[k] == n_blocks:
            me.ar_done = k

    actors = []
    # the stream of rank r: kernels are launched in order; blocks of one kernel run concurrently, and a kernel's blocks only start
    # once the previous kernel of the same kind has finished (stream order)
    def gemm_stream():
        for k in range(n_calls):
            yield from gemm(k)

    def block_stream(b):
        for k in range(n_calls):
            if k > 0:
                yield lambda k=k: me.ar_done >= k - 1                # kernel k starts after kernel k-1 of this stream completed
            yield from ar_block(k, b)

    actors.append(gemm_stream())
    actors += [block_stream(b) for b in range(n_blocks)]
    return actors


@pytest.mark.parametrize("world,n_blocks", [(2, 1), (2, 3), (4, 2), (8, 2)])
def test_two_region_allreduce_never_reads_stale_or_overwritten_partials(world, n_blocks):
    for seed in range(60):
        rng = random.Random(seed)
        ranks = [Rank(world, n_blocks) for _ in range(world)]
        log = []
        actors = [a for r in range(world) for a in tp_decode_rank(ranks, r, 7, n_blocks, regions=2, pdl_guard=True, log=log)]
        run_actors(actors, rng)
        assert len(log) == 7 * world * n_blocks * world
        assert all(rk.epoch == [7] * n_blocks for rk in ranks)


def test_checker_catches_a_single_region_design():
    """With ONE region and no trailing barrier a fast rank's next GEMM overwrites partial sums a slow peer has not read yet."""
    caught = 0
    for seed in range(60):
        rng = random.Random(seed)
        ranks = [Rank(2, 1) for _ in range(2)]
        actors = [a for r in range(2) for a in tp_decode_rank(ranks, r, 5, 1, regions=1, pdl_guard=True, log=[])]
        try:
            run_actors(actors, rng)
        except AssertionError as e:
            assert "read version" in str(e)
            caught += 1
    assert caught > 10


def test_checker_catches_writes_before_the_dependent_launch_wait():
    """Two regions are only enough because GEMM_{k+2} writes after kernel k+1 completed on its own rank (griddepcontrol.wait before
    the first store): a GEMM that stored early could overwrite region k % 2 while a peer still reads call k."""
    caught = 0
    for seed in range(200):
        rng = random.Random(seed)
        ranks = [Rank(2, 1) for _ in range(2)]
        actors = [a for r in range(2) for a in tp_decode_rank(ranks, r, 6, 1, regions=2, pdl_guard=False, log=[])]
        try:
            run_actors(actors, rng)
        except AssertionError as e:
            # either a peer reads the overwritten region, or the version a block waits for has already been replaced (the model's
            # stand-in for "the kernel consumed partial sums of a later call")
            assert "read version" in str(e) or "deadlock" in str(e)
            caught += 1
    assert caught > 0


def test_epoch_comparison_survives_the_32_bit_wrap():
    """The per-block epoch lives in the pad for the life of the process; the spin compares wrapped differences."""
    for start in (0xFFFFFFFD, 0x7FFFFFFE):
        for seed in range(20):
            rng = random.Random(seed)
            ranks = [Rank(4, 1) for _ in range(4)]
            for rk in ranks:
                rk.epoch = [start]
                rk.flags = [[start] * 4]
            passed = [0] * 4

            def actor(r):
                for _ in range(6):
                    yield from block_barrier(ranks, r, 0)
                    passed[r] += 1
                    # nobody may be more than one barrier ahead of anybody else
                    assert max(passed) - min(passed) <= 1
                    yield None

            run_actors([actor(r) for r in range(4)], rng)
            assert passed == [6] * 4 and ranks[0].epoch[0] == (start + 6) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------- stream-K cooperative fix-up


def _cta_of(u, units, G):
    return next(c for c in range(G) if units * c // G <= u < units * (c + 1) // G)


def streamk_cta(g, G, tiles, num_kb, flags, slots, out, order_log, rearm=True):
    """CTA g of gemm_streamk_kernel: the flattened (tile, k-block) space is cut into G equal contiguous ranges.  A range that covers
    part of a tile writes a partial accumulator to one of the CTA's two workspace slots and bumps the tile's `arrived` counter; after
    posting all of its partials the CTA waits, tile by tile, until all S contributors of the tile arrived, sums its 1/S share of the
    tile from everybody's slots, and bumps `done`; the last reader resets both counters (CUDA-graph replay needs them at zero)."""
    units = tiles * num_kb
    u0, u1 = units * g // G, units * (g + 1) // G
    pending = []
    u = u0
    while u < u1:
        tile = u // num_kb
        seg_end = min(u1, (tile + 1) * num_kb)
        whole = (u == tile * num_kb and seg_end == (tile + 1) * num_kb)
        if whole:
            out[tile] = ("whole", g)
        else:
            slot = 0 if (units * g // G) // num_kb == tile else 1
            assert slots[g][slot] is None, f"CTA {g}: workspace slot {slot} still holds an unread partial"
            slots[g][slot] = (tile, seg_end - u)
            yield None
            flags[tile][0] += 1                                      # red.release.gpu.add
            pending.append(tile)
        u = seg_end
        yield None
    for tile in pending:
        yield from _fixup(g, G, tile, num_kb, units, flags, slots, out, order_log, rearm)


def _fixup(g, G, tile, num_kb, units, flags, slots, out, order_log, rearm):
    first, last = _cta_of(tile * num_kb, units, G), _cta_of((tile + 1) * num_kb - 1, units, G)
    S = last - first + 1
    yield lambda: flags[tile][0] >= S                                # ld.acquire.gpu spin
    got = 0
    for m in range(first, last + 1):
        slot = 0 if (units * m // G) // num_kb == tile else 1
        ent = slots[m][slot]
        assert ent is not None and ent[0] == tile, f"CTA {g} reads CTA {m}'s slot {slot} for tile {tile}: holds {ent}"
        got += ent[1]
        yield None
    assert got == num_kb, f"tile {tile}: partials cover {got} of {num_kb} k-blocks"
    assert not isinstance(out.get(tile), tuple), f"tile {tile} was stored whole AND fixed up"
    out.setdefault(tile, []).append(g)
    order_log.append((tile, g))
    flags[tile][1] += 1                                              # atomicAdd(done)
    if flags[tile][1] == S:                                          # last reader: every member has finished reading the slots
        for m in range(first, last + 1):
            slot = 0 if (units * m // G) // num_kb == tile else 1
            slots[m][slot] = None
        if rearm:
            flags[tile][1] = 0
            yield None
            flags[tile][0] = 0


@pytest.mark.parametrize("tiles,num_kb,G", [(16, 32, 148), (43, 32, 148), (3, 7, 5), (5, 4, 3), (125, 16, 148), (2, 9, 4), (8, 3, 8)])
def test_streamk_fixup_terminates_reads_complete_partials_and_rearms(tiles, num_kb, G):
    G = min(G, tiles * num_kb)
    for replay in range(2):                                          # the second pass reuses the counters, like a graph replay
        for seed in range(12):
            rng = random.Random(seed)
            flags = [[0, 0] for _ in range(tiles)] if (replay == 0 or seed == 0) else flags
            slots = [[None, None] for _ in range(G)]
            out, log = {}, []
            run_actors([streamk_cta(g, G, tiles, num_kb, flags, slots, out, log) for g in range(G)], rng)
            assert all(f == [0, 0] for f in flags), "counters must be zero again for the next graph replay"
            assert all(s == [None, None] for s in slots)
            assert set(out) == set(range(tiles))                     # every tile stored, by a whole-tile CTA or by its fix-up group
            for t, v in out.items():
                if not isinstance(v, tuple):
                    assert len(v) == len(set(v)) >= 2


def test_streamk_without_rearm_breaks_the_second_replay():
    tiles, num_kb, G = 3, 7, 5

    def launch(flags, seed):
        slots = [[None, None] for _ in range(G)]
        run_actors([streamk_cta(g, G, tiles, num_kb, flags, slots, {}, [], rearm=False) for g in range(G)], random.Random(seed))

    flags = [[0, 0] for _ in range(tiles)]
    launch(flags, 0)
    assert any(f != [0, 0] for f in flags)
    # stale `arrived` counts let a CTA of the next launch read slots nobody has written yet
    with pytest.raises(AssertionError):
        for seed in range(50):
            launch([list(f) for f in flags], seed)
