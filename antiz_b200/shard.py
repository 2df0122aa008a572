"""Static partition of the stream x parameter grid over ranks and the host-side gather of the
fixed-size per-stream records.  No data-path collective: every rank scans the same file, searches the streams it
owns, and the records are gathered on the host (torch.distributed object gather here; plain memcpy in uncomp.cpp).
Pure host logic - usable without a GPU (tests/test_shard_gloo.py runs it with the gloo backend, world_size 2)."""


def owner(stream_index: int, nshards: int) -> int:
    """stream i is searched by shard i % nshards (atz_search_shard, csrc/api.cu)"""
    return stream_index % nshards


def my_streams(nstreams: int, shard: int, nshards: int):
    return range(shard, nstreams, nshards)


def merge(per_shard):
    """per_shard[g] = {stream_index: record} of the streams shard g owns -> list of records in stream order.
    Raises if a stream is missing, duplicated, or reported by a shard that does not own it."""
    nshards = len(per_shard)
    n = sum(len(d) for d in per_shard)
    out = [None] * n
    for g, d in enumerate(per_shard):
        for i, rec in d.items():
            if not (0 <= i < n) or owner(i, nshards) != g or out[i] is not None:
                raise ValueError(f"shard {g} reported stream {i} it does not own (or twice)")
            out[i] = rec
    if any(r is None for r in out):
        raise ValueError("a stream was not reported by any shard")
    return out


def gather_records(local: dict, dist=None):
    """all ranks contribute {stream_index: record}; every rank gets the merged list (rank order = shard order)"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return merge([local])
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, local)
    return merge(parts)


def max_over_ranks(ms: float, dist=None, device=None) -> float:
    """bench timing rule: the step time of a multi-GPU run is the maximum over ranks"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return ms
    import torch
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
