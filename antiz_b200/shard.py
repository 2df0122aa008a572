"""One container over several ranks (one process per GPU): the host side of SURVEY.md 8(e).

The path shards with no data-path collective.  What crosses ranks is host data only, through whatever byte transport the
launcher provides (torch.distributed object collectives here, plain memcpy between the contexts of one process in uncomp.cpp):
  1. after atz_scan_shard: every rank's probe records (one fixed-size record per candidate the accept logic can act on) to every
     other rank - then atz_scan_finish replays the accept logic identically everywhere and the static partition
     (atz_host_partition: by plaintext length; a stream stays on the shard that probed it where the balance allows) says who owns
     which stream;
  2. after atz_search_shard: the fixed-size per-stream records (+ diff lists) of the streams a rank owns, gathered in stream order.
Pure host logic - usable without a GPU (tests/test_shard_gloo.py runs partition/gather/merge with the gloo backend, world_size 2).
"""


def owners(inflated_lengths, nshards, probed_by=None):
    """owner shard of every accepted stream (the C ABI's atz_host_partition, so Python and uncomp.cpp agree by construction);
    a context that has scanned reports the same list through ctx.owners()"""
    import antiz_b200 as az
    return az.partition(inflated_lengths, nshards, probed_by)


def my_streams(own, shard):
    return [i for i, g in enumerate(own) if g == shard]


def merge(per_shard, own):
    """per_shard[g] = {stream_index: record} of the streams shard g owns -> list of records in stream order.
    Raises if a stream is missing, duplicated, or reported by a shard that does not own it."""
    n = len(own)
    out = [None] * n
    for g, d in enumerate(per_shard):
        for i, rec in d.items():
            if not (0 <= i < n) or own[i] != g or out[i] is not None:
                raise ValueError(f"shard {g} reported stream {i} it does not own (or twice)")
            out[i] = rec
    if any(r is None for r in out):
        raise ValueError("a stream was not reported by any shard")
    return out


def _world(dist):
    return dist.get_world_size() if dist is not None and dist.is_initialized() else 1


def gather_records(local: dict, own, dist=None, group=None):
    """all ranks contribute {stream_index: record}; every rank gets the merged list (rank order = shard order)"""
    if _world(dist) == 1:
        return merge([local], own)
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, local, group=group)
    return merge(parts, own)


def exchange_probes(ctx, rank, dist=None, group=None):
    """step 1 above: hand this rank's probe records to the others and take theirs (no-op for one rank)"""
    if _world(dist) == 1:
        return
    blobs = [None] * dist.get_world_size()
    dist.all_gather_object(blobs, ctx.probe_export(), group=group)
    for g, b in enumerate(blobs):
        if g != rank:
            ctx.probe_import(g, b)


def scan_search(ctx, chunksize, opt, rank=0, world=1, dist=None, group=None):
    """phases 1 and 3 of one container on this rank's shard; the file has been given to ctx (load / load_device / attach).
    Returns the number of streams (the same on every rank)."""
    ctx.scan_shard(chunksize, rank, world)
    exchange_probes(ctx, rank, dist, group)
    n = ctx.scan_finish()
    ctx.search(opt, rank, world)
    return n


FIELDS = ("offset", "streamLength", "inflatedLength", "identBytes", "firstDiffByte", "ndiff", "diff_index", "offsetType", "clevel", "window", "memlevel", "recomp")


def owned_records(ctx, rank, world):
    """{stream index: (record tuple in FIELDS order, diff offsets, diff values)} of the streams this rank owns, and the owner list"""
    import numpy as np
    tab = ctx.stream_table()
    own = ctx.owners()
    idx = np.nonzero(np.asarray(own, dtype=np.uint32) == rank)[0]
    rows = tab[idx][list(FIELDS)].tolist()
    out = {i: (r, (), b"") for i, r in zip(idx.tolist(), rows)}
    if int(tab["ndiff"][idx].sum()):
        offs, vals = ctx.diffs()
        for i in idx[tab["ndiff"][idx] > 0].tolist():
            d, k = int(tab["diff_index"][i]), int(tab["ndiff"][i])
            out[i] = (out[i][0], tuple(offs[d:d + k]), bytes(vals[d:d + k]))
    return out, own


def max_over_ranks(ms: float, dist=None, device=None) -> float:
    """bench timing rule: the step time of a multi-GPU run is the maximum over ranks"""
    if _world(dist) == 1:
        return ms
    import torch
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
