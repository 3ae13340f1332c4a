"""Multi-GPU plumbing: problems are independent, so the only cross-rank
operations are the partition of the index range and the reduction of timings / counters.
No collective touches the data path."""
import os


def env_rank():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def problem_range(rank, world, batch_per_rank):
    """weak scaling: rank r owns problems [r*B, (r+1)*B) of the global index space."""
    return rank * batch_per_rank, batch_per_rank


def split_range(first, count, parts):
    """static contiguous split of [first, first+count) into `parts` nearly equal ranges"""
    out, base, rem = [], count // parts, count % parts
    at = first
    for p in range(parts):
        n = base + (1 if p < rem else 0)
        out.append((at, n))
        at += n
    return out


def reduce_stats(times, counts, device=None):
    """max over ranks of `times`, sum over ranks of `counts` (lists of floats)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(times), list(counts)
    t = torch.tensor(list(times), dtype=torch.float64, device=device)
    c = torch.tensor(list(counts), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()], [float(x) for x in c.tolist()]
