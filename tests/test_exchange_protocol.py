"""CPU-only: the step protocol of the fused pixel exchange (localrf_b200/dist.py, DESIGN.md par. 5) under
randomly interleaved simulated ranks.

The render kernel of step s (a) waits until every peer has published step s-1-lag, (b) stores its pixels into
every rank's rotating gathered buffer, (c) publishes step s; the consumer reads the gathered image of step
s-1-lag after launch s (stream order).  `ExchangeSchedule` is the bookkeeping `PixelExchange` really uses
(buffer index, signal / wait step, readable step), so this drives the product's arithmetic -- for N up to 8
ranks and the lags the bench uses -- through adversarial schedules (one very slow rank, bursts) and checks:
  * no deadlock,
  * every image is COMPLETE when read (all N shards carry the expected step),
  * no buffer is overwritten before every rank has read it,
and that the checker has teeth: with one buffer fewer than the protocol's minimum (2 lag + 3) it finds a
corrupted read."""
import random

import pytest

from localrf_b200.dist import ExchangeSchedule


class Rank:
    def __init__(self, r, world, sched, steps):
        self.r, self.world, self.sched, self.steps = r, world, sched, steps
        self.flags = [0] * world                                  # flags[p] = last step peer p published HERE
        self.buffers = [[0] * world for _ in range(sched.n_buf)]   # buffers[b][p] = step whose shard of p sits there
        self.pc, self.cur = 0, (None, None, None)                  # micro-op counter within the current step
        self.reads = 0

    def done(self):
        return self.sched.seq >= self.steps and self.pc == 0


def micro_ops(world):
    """One step of a rank as a list of atomically executed micro-operations (see the module docstring)."""
    return (["launch", "wait"] + [("store", p) for p in range(world)] + [("publish", p) for p in range(world)]
            + ["advance", "close", "read"])


def simulate(world, lag, steps, seed, n_buf=None, slow=None):
    rng = random.Random(seed)
    ranks = []
    for r in range(world):
        s = ExchangeSchedule(lag)
        if n_buf is not None:
            s.n_buf = n_buf                                       # (sabotage: fewer buffers than the protocol needs)
        ranks.append(Rank(r, world, s, steps))
    ops = micro_ops(world)
    weights = [1.0] * world
    if slow is not None:
        weights[slow] = 0.02                                      # one rank 50x slower than the others
    errors = []

    def try_advance(me):
        """Executes the next micro-op of rank `me` if it is not blocked; -> True if progress was made."""
        if me.pc == 0 and me.sched.seq >= me.steps:
            return False                                          # finished
        op = ops[me.pc]
        if op == "launch":
            me.cur = me.sched.next_step()                         # (buffer, signal step, wait step)
        buf, sig, wait = me.cur
        if op == "wait":
            if any(f < wait for f in me.flags):
                return False                                      # prologue spin: blocked
        elif isinstance(op, tuple) and op[0] == "store":
            ranks[op[1]].buffers[buf][me.r] = sig                 # peer store of this rank's shard
        elif isinstance(op, tuple) and op[0] == "publish":
            ranks[op[1]].flags[me.r] = sig                        # release-store of the step flag (after all stores)
        elif op == "advance":
            me.sched.advance()
        elif op == "close":
            if me.sched.lag == 0 and any(f < me.sched.seq for f in me.flags):
                return False                                      # lag 0: the wait kernel closes the SAME step
        elif op == "read":
            g = me.sched.gathered_step()
            if g >= 1:
                got = me.buffers[me.sched.gathered_buffer()]
                if any(v != g for v in got):
                    errors.append((me.r, me.sched.seq, g, list(got)))
                me.reads += 1
        me.pc = (me.pc + 1) % len(ops)
        return True

    guard = 0
    while not all(k.done() for k in ranks):
        order = list(range(world))
        # weighted random pick, falling back over the other ranks so a blocked pick does not stall the loop
        order.sort(key=lambda i: rng.random() / weights[i])
        burst = rng.choice((1, 1, 2, 5, 40))                      # bursts: a rank runs several micro-ops in a row
        progressed = False
        for i in order:
            n = 0
            while n < burst and try_advance(ranks[i]):
                n += 1
            if n:
                progressed = True
                break
        if not progressed:
            return "deadlock", errors, ranks
        guard += 1
        assert guard < 5_000_000
    return "ok", errors, ranks


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("lag", [0, 1, 3])
def test_protocol_complete_reads_no_overwrite_no_deadlock(world, lag):
    for seed in range(6):
        for slow in (None, 0, world - 1):
            status, errors, ranks = simulate(world, lag, steps=40, seed=seed, slow=slow)
            assert status == "ok", (world, lag, seed, slow)
            assert not errors, (world, lag, seed, slow, errors[:3])
            # every rank consumed every image the protocol makes readable within `steps` launches
            assert all(k.reads == 40 - (0 if lag == 0 else 1 + lag) for k in ranks)


def test_a_fast_rank_runs_ahead_by_the_slack_and_no_further():
    """With lag L a rank is never more than L + 1 launches ahead of the slowest peer's published step (that is
    the slack the bench's weak-scaling mode relies on), and it does get that far."""
    world, lag = 4, 3
    sch = [ExchangeSchedule(lag) for _ in range(world)]
    flags = [[0] * world for _ in range(world)]
    # rank 0 runs alone: how many steps can it launch while the others have published nothing?
    launched = 0
    while True:
        _, sig, wait = sch[0].next_step()
        if any(f < wait for f in flags[0]):
            break
        for p in range(world):
            flags[p][0] = sig
        sch[0].advance()
        launched += 1
        assert launched < 100
    assert launched == lag + 1


@pytest.mark.parametrize("lag", [1, 3])
def test_checker_detects_too_few_buffers(lag):
    """2 lag + 3 buffers are the protocol's minimum (the product rotates 2 lag + 4): with 2 lag + 2 a peer may
    overwrite an image that is still to be read -- the simulation must find such a schedule."""
    found = False
    for seed in range(40):
        for slow in (0, 1, None):
            status, errors, _ = simulate(4, lag, steps=40, seed=seed, n_buf=2 * lag + 2, slow=slow)
            if errors:
                found = True
                break
        if found:
            break
    assert found
    # ... and 2 lag + 3 is indeed enough
    for seed in range(10):
        status, errors, _ = simulate(4, lag, steps=40, seed=seed, n_buf=2 * lag + 3, slow=seed % 4)
        assert status == "ok" and not errors


def test_schedule_arithmetic():
    s = ExchangeSchedule(0)
    assert s.n_buf == 2 and s.next_step() == (1, 1, 0) and s.gathered_step() == 0
    s.advance()
    assert s.gathered_step() == 1 and s.gathered_buffer() == 1 and s.next_step() == (0, 2, 0)
    s = ExchangeSchedule(3)
    assert s.n_buf == 10
    seen = []
    for step in range(1, 30):
        buf, sig, wait = s.next_step()
        assert sig == step and wait == max(step - 1 - 3, 0) and buf == step % 10
        s.advance()
        seen.append(s.gathered_step())
    assert seen[:6] == [0, 0, 0, 0, 1, 2] and seen[-1] == 29 - 4
    with pytest.raises(ValueError):
        ExchangeSchedule(-1)


def test_pixel_exchange_fills_the_output_struct_from_the_schedule():
    """PixelExchange.fill_outputs / close_step / gathered() on a hand-made instance (no symmetric memory, no
    device): the pointers and step numbers that reach the kernel follow ExchangeSchedule."""
    import torch
    from localrf_b200 import _lib
    from localrf_b200.dist import PixelExchange
    world, lag, max_rays = 4, 3, 4096
    x = PixelExchange.__new__(PixelExchange)
    x.world, x.rank, x.max_rays = world, 2, max_rays
    x.buf_bytes = (max_rays * 16 + 255) // 256 * 256
    x.sched = ExchangeSchedule(lag)
    x.lag, x.n_buf = x.sched.lag, x.sched.n_buf
    x.ptrs = [0x10000000 * (p + 1) for p in range(world)]
    x.mc_ptr = 0
    x.mem = torch.zeros(x.n_buf * x.buf_bytes + PixelExchange.FLAG_BYTES, dtype=torch.uint8)
    for step in range(1, 25):
        o = _lib.LrfOutputs()
        x.fill_outputs(o, ray_lo=1024)
        buf = step % x.n_buf
        assert o.n_peers == world and o.rank == 2 and o.signal_seq == step and o.wait_seq == max(step - 1 - lag, 0)
        for p in range(world):
            assert o.peer_pix[p] == x.ptrs[p] + buf * x.buf_bytes + 1024 * 16
            assert o.peer_flags[p] == x.ptrs[p] + x.n_buf * x.buf_bytes
        assert not o.mc_pix
        x.close_step(None)                                # lag >= 1: nothing is enqueued
        assert x.seq == step and x.gathered_step() == max(step - 1 - lag, 0)
        g = x.gathered(max_rays)
        assert g.shape == (max_rays, 4)
        assert g.data_ptr() - x.mem.data_ptr() == (x.gathered_step() % x.n_buf) * x.buf_bytes
