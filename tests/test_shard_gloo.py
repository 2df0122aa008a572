"""CPU, world_size 2, gloo: the multi-GPU path's host logic - static stream partition (atz_host_partition: longest plaintext
first, to the least loaded shard), record gather in stream order by owner, max-over-ranks timing - without a GPU.  The records are the reference candidate sequences of each stream's
header class (pure host code in the C ABI), so the gather moves real, checkable data."""
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _record(i):
    """deterministic stand-in for a stream's search record: head of the candidate order of header type i % 24"""
    import ctypes as C
    import antiz_b200 as az
    L = az.lib()
    c = (C.c_uint8 * 600)(); w = (C.c_uint8 * 600)(); m = (C.c_uint8 * 600)()
    n = L.atz_host_candidate_sequence(i % 24, 0, c, w, m, 600)
    return (i, n, c[0], w[0], m[0])


def _lengths(n):
    return [1000 + (i * 7919) % 250000 for i in range(n)]


def _worker(rank, world, port, nstreams, q):
    import torch.distributed as dist
    from antiz_b200 import shard
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    own = shard.owners(_lengths(nstreams), world)
    mine = {i: _record(i) for i in shard.my_streams(own, rank)}
    merged = shard.gather_records(mine, own, dist)
    t = shard.max_over_ranks(10.0 + rank, dist)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, merged, t))


@pytest.mark.timeout(120)
def test_two_rank_partition_and_gather():
    world, nstreams = 2, 37
    ctx = mp.get_context("spawn")
    q = ctx.Queue(); port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, nstreams, q)) for r in range(world)]
    for p in ps:
        p.start()
    got = [q.get(timeout=90) for _ in range(world)]
    for p in ps:
        p.join(30); assert p.exitcode == 0
    want = [_record(i) for i in range(nstreams)]
    for rank, merged, t in got:
        assert merged == want
        assert t == 11.0            # max over ranks of (10 + rank)


def test_merge_rejects_wrong_owner():
    from antiz_b200 import shard
    own = [0, 1, 0]
    with pytest.raises(ValueError):
        shard.merge([{0: "a", 1: "b"}, {}], own[:2])      # stream 1 belongs to shard 1
    with pytest.raises(ValueError):
        shard.merge([{0: "a"}, {3: "b"}], own)            # out of range / hole
    assert shard.merge([{0: "a", 2: "c"}, {1: "b"}], own) == ["a", "b", "c"]


def test_partition_is_balanced_and_deterministic():
    """atz_host_partition: every stream has exactly one owner, the loads differ by at most one longest stream, and the result
    depends on the lengths only (every context computes it independently)"""
    from antiz_b200 import shard
    ul = _lengths(500)
    for nsh in (1, 2, 3, 8):
        own = shard.owners(ul, nsh)
        assert own == shard.owners(list(ul), nsh) and set(own) <= set(range(nsh))
        load = [sum(u for u, g in zip(ul, own) if g == k) for k in range(nsh)]
        assert max(load) - min(load) <= max(ul) + 4096 * len(ul) // nsh
    assert shard.owners([], 4) == []
    # a sharded scan: streams stay where they were probed unless that shard is more than 2 % above the mean
    probed = [min(3, i * 4 // 500) for i in range(500)]
    own = shard.owners(ul, 4, probed)
    load = [sum(u + 4096 for u, g in zip(ul, own) if g == k) for k in range(4)]
    assert max(load) <= sum(load) / 4 * 1.03
    assert sum(1 for a, b in zip(own, probed) if a != b) < 60
    skew = [0] * 400 + [1] * 50 + [2] * 25 + [3] * 25
    own = shard.owners(ul, 4, skew)
    load = [sum(u + 4096 for u, g in zip(ul, own) if g == k) for k in range(4)]
    assert max(load) <= sum(load) / 4 * 1.03 and all(a == b for a, b in zip(own, skew) if b != 0)
