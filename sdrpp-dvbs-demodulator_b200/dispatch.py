"""Frame sharding across ranks / GPUs.  FECFRAMEs are independent, so the only multi-GPU logic is
"who decodes which frames" plus reductions of timing and counters for reporting; there is no data-path
collective.  The same contiguous split is used by the C++ in-process dispatcher (csrc/api.cu,
decode_host) and by bench.py's one-process-per-GPU launch."""


def shard_range(nframes, world, rank):
    """Contiguous share [f0, f1) of rank `rank`; the first shares are the larger ones."""
    per = (nframes + world - 1) // world
    f0 = min(nframes, rank * per)
    return f0, min(nframes, f0 + per)


def merge_in_order(shares):
    """shares: list of (f0, array-like rows) from every rank -> rows in submission order."""
    out = []
    for _, rows in sorted(shares, key=lambda s: s[0]):
        out.extend(rows)
    return out


def reduce_max(values, device=None):
    """Element-wise max over ranks of a list of floats (plain list back).  No-op without a process group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    t = torch.tensor(values, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def reduce_sum(values, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    t = torch.tensor(values, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()]
