"""Multi-GPU plumbing for batches of INDEPENDENT local areas: static sharding of the area list over
ranks and the two reductions bench.py needs (max time over ranks, total LM iterations).  No data-path collective:
a local-area optimisation never exchanges data with another one.  Backend-agnostic (NCCL on the GPU box, gloo in tests)."""
import torch
import torch.distributed as dist


def shard_indices(n_items, rank, world_size):
    """Round-robin assignment (item i -> rank i % world): the cost-sorted area list stays balanced across ranks."""
    return list(range(rank, n_items, world_size))


def reduce_timing(elapsed_ms, n_trials, device="cpu"):
    """Returns (max elapsed over ranks, total trials over ranks); identity when torch.distributed is not initialised."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(elapsed_ms), float(n_trials)
    t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    c = torch.tensor([float(n_trials)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return float(t.item()), float(c.item())


def reduce_max_sum(max_values, sum_values, device="cpu"):
    """Element-wise MAX over ranks of `max_values` (timings) and SUM over ranks of `sum_values` (counts); lists of floats in and out.
    Identity when torch.distributed is not initialised.  What bench.py / tools/bench_configs.py use for their multi-rank lines."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in max_values], [float(v) for v in sum_values]
    t = torch.tensor([float(v) for v in max_values], dtype=torch.float64, device=device)
    c = torch.tensor([float(v) for v in sum_values], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return t.tolist(), c.tolist()


def gather_unknowns(local_ids, local_values, device="cpu"):
    """Optional: all ranks receive every rank's optimised unknowns (ids int64 [k], values float64 [k, d]).
    Sizes may differ per rank; padded all_gather."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_ids, local_values
    world = dist.get_world_size()
    n = torch.tensor([local_ids.shape[0]], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    m = int(max(s.item() for s in sizes))
    d = local_values.shape[1]
    ids = torch.full((m,), -1, dtype=torch.int64, device=device)
    vals = torch.zeros((m, d), dtype=torch.float64, device=device)
    ids[:local_ids.shape[0]] = local_ids
    vals[:local_values.shape[0]] = local_values
    all_ids = [torch.empty_like(ids) for _ in range(world)]
    all_vals = [torch.empty_like(vals) for _ in range(world)]
    dist.all_gather(all_ids, ids)
    dist.all_gather(all_vals, vals)
    ids = torch.cat([a[:int(s.item())] for a, s in zip(all_ids, sizes)])
    vals = torch.cat([a[:int(s.item())] for a, s in zip(all_vals, sizes)])
    return ids, vals
