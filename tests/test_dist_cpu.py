"""N > 1 host logic on CPU (gloo, world_size 2 and 4): the exchange plan of rocsvxDistPlanExchange, executed
with torch.distributed send/recv on host shards, must equal the oracle's index-bit swap of the full state."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sv_oracle as so
from rocquantum_b200 import capi
from tests import util


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, pairs, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = capi.load("c64")
        m = world.bit_length() - 1
        nl = n - m
        full = util.random_state(n, seed=123).astype(np.complex128)
        shard = full[rank << nl:(rank + 1) << nl].copy()
        lbits = [p[0] for p in pairs]
        gbits = [p[1] for p in pairs]
        cnt = C.c_size_t()
        assert lib.rocsvxDistPlanExchange(nl, world, rank, capi.uarr(lbits), capi.uarr(gbits), len(pairs), None, 0, C.byref(cnt)) == 0
        segs = (capi.ExchangeSeg * max(1, cnt.value))()
        assert lib.rocsvxDistPlanExchange(nl, world, rank, capi.uarr(lbits), capi.uarr(gbits), len(pairs), segs, cnt.value, C.byref(cnt)) == 0
        out = shard.copy()
        reqs, bufs = [], []
        for i in range(cnt.value):
            s = segs[i]
            send = torch.from_numpy(shard[s.sendOffset:s.sendOffset + s.count].copy().view(np.float64))
            recv = torch.empty_like(send)
            reqs.append(dist.isend(send, s.peer, tag=i % 1))
            bufs.append((s, recv, send))
        # receives are posted in the peer's own segment order: match by enumerating segments per peer in order
        for s, recv, _ in bufs:
            dist.recv(recv, s.peer)
            out[s.recvOffset:s.recvOffset + s.count] = recv.numpy().view(np.complex128)
        for r in reqs:
            r.wait()
        o = so.Oracle(n, "c128")
        o.set_state(full)
        for l, g in pairs:
            o.swap_index_bits(l, g)
        want = o.state[rank << nl:(rank + 1) << nl]
        ret[rank] = bool(np.array_equal(out, want))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,pairs", [
    (2, 6, [(4, 5)]),              # top local bit <-> the rank bit: one contiguous half-slice
    (2, 7, [(2, 6)]),              # low local bit: many short runs
    (4, 8, [(5, 6), (4, 7)]),      # both rank bits at once: all-to-all among 4 ranks
    (4, 8, [(5, 7)]),              # one of two rank bits: pairwise
    (4, 9, [(1, 7), (6, 8)]),      # mixed positions
])
def test_exchange_plan_equals_index_bit_swap(world, n, pairs):
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, n, pairs, ret), nprocs=world, join=True)
        assert all(ret.get(r) for r in range(world)), dict(ret)


@pytest.mark.parametrize("world,n,pairs", [
    (2, 6, [(4, 5)]),              # what the engine issues: top local bit <-> the rank bit
    (2, 7, [(2, 6)]),              # low local bit: many short runs
    (2, 3, [(0, 2)]),              # runs of ONE amplitude: the odd split (low rank moves nothing, high rank one element)
    (4, 8, [(4, 6), (5, 7)]),      # both rank bits <-> the two top local bits (the engine's k = 2 exchange)
    (4, 8, [(5, 7)]),              # one of two rank bits: pairwise
    (4, 9, [(1, 7), (6, 8)]),      # mixed positions
    (8, 10, [(4, 7), (5, 8), (6, 9)]),   # k = 3 on 8 ranks: 7 peers per rank
    (8, 10, [(6, 8)]),             # one of three rank bits
])
def test_peer_swap_plan_equals_index_bit_swap(world, n, pairs):
    """ROCQ_EXCHANGE=p2p: the in-place half-swaps of rocsvxDistPlanPeerSwap, played on all ranks' shards in one process
    (any rank order: every amplitude is touched by exactly one rank), equal the oracle's index-bit swap."""
    lib = capi.load("c64")
    m = world.bit_length() - 1
    nl = n - m
    full = util.random_state(n, seed=321).astype(np.complex128)
    shards = [full[r << nl:(r + 1) << nl].copy() for r in range(world)]
    touched = [np.zeros(1 << nl, dtype=np.int32) for _ in range(world)]
    lbits, gbits = [p[0] for p in pairs], [p[1] for p in pairs]
    moved = 0
    for rank in reversed(range(world)):
        cnt = C.c_size_t()
        assert lib.rocsvxDistPlanPeerSwap(nl, world, rank, capi.uarr(lbits), capi.uarr(gbits), len(pairs), None, 0, C.byref(cnt)) == 0
        segs = (capi.ExchangeSeg * max(1, cnt.value))()
        assert lib.rocsvxDistPlanPeerSwap(nl, world, rank, capi.uarr(lbits), capi.uarr(gbits), len(pairs), segs, cnt.value, C.byref(cnt)) == 0
        for i in range(cnt.value):
            s = segs[i]
            assert s.peer != rank and s.count > 0
            mine = slice(s.sendOffset, s.sendOffset + s.count)
            theirs = slice(s.recvOffset, s.recvOffset + s.count)
            tmp = shards[rank][mine].copy()
            shards[rank][mine] = shards[s.peer][theirs]
            shards[s.peer][theirs] = tmp
            touched[rank][mine] += 1
            touched[s.peer][theirs] += 1
            moved += 2 * s.count
    o = so.Oracle(n, "c128")
    o.set_state(full)
    for l, g in pairs:
        o.swap_index_bits(l, g)
    for r in range(world):
        assert touched[r].max() <= 1                                    # nothing is moved twice
        assert np.array_equal(shards[r], o.state[r << nl:(r + 1) << nl])
    assert moved == world * ((1 << nl) - (1 << (nl - len(pairs))))      # every rank keeps 2^-k of its slice
