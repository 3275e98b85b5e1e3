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
