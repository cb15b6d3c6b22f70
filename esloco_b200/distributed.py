"""Multi-GPU plumbing: iterations are independent (esloco.py:149-152), so ranks own disjoint iteration ranges and
the only exchange is one sum-reduce of per-target moment accumulators (plus, when the per-iteration table is
wanted, a gather of the tallies).  Works with the nccl backend on GPUs and with gloo on CPU tensors (tests)."""
import os

import numpy as np


def world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_iterations(n_iterations, rank, world_size):
    """Contiguous block of iteration ids owned by `rank`: sizes differ by at most one, union is range(n)."""
    base, rem = divmod(n_iterations, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def sync_seed(param_dictionary, src=0):
    """Make every rank use rank `src`'s seed.  An INI without `seed` makes parse_config draw a random one
    (config_handler.py:53-57,102) -- per process, so the ranks of a torchrun launch would otherwise build different
    insertion sites (even different target counts with the Poisson option), key different Philox streams and then
    sum each other's moments under rank 0's target names.  Call right after init_process_group, before any
    np.random.seed / set-up.  Returns the common seed."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        box = [param_dictionary.get("seed")]
        dist.broadcast_object_list(box, src=src)
        param_dictionary["seed"] = box[0]
    return param_dictionary.get("seed")


def reduce_moments(acc, dst=0):
    """In-place SUM-reduce of the [T, 8] int64 moment accumulators to rank dst (one collective per run)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(acc, dst=dst, op=dist.ReduceOp.SUM)
    return acc


def gather_iterations(t, n_iterations, dst=0):
    """Gather per-iteration rows (dim 0) from every rank's shard to rank dst in iteration order.  Shards are padded
    to the largest shard so a plain gather works on nccl and gloo alike."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t
    rank, ws = dist.get_rank(), dist.get_world_size()
    sizes = [len(shard_iterations(n_iterations, r, ws)) for r in range(ws)]
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    buf[: t.shape[0]] = t
    out = [torch.empty_like(buf) for _ in range(ws)] if rank == dst else None
    dist.gather(buf, out, dst=dst)
    if rank != dst:
        return None
    return torch.cat([o[:n] for o, n in zip(out, sizes)], dim=0)


def moments_from_tallies(tallies):
    """numpy form of the device kernel k_moments for HOST-side tallies (tests, tiny runs): [n_iter, T, 3] ->
    [T, 8] = n, sum f, sum f^2, sum p, sum p^2, sum b, lo32(sum b^2), hi(sum b^2)."""
    t = np.asarray(tallies).astype(object)
    n_iter, T, _ = t.shape
    out = np.zeros((T, 8), np.int64)
    out[:, 0] = n_iter
    for c, (a, b) in enumerate(((1, 2), (3, 4))):
        out[:, a] = t[:, :, c].sum(0)
        out[:, b] = (t[:, :, c] ** 2).sum(0)
    out[:, 5] = t[:, :, 2].sum(0)
    sq = (t[:, :, 2] ** 2).sum(0) if n_iter else np.zeros(T, object)
    out[:, 6] = [int(x) & 0xFFFFFFFF for x in sq]
    out[:, 7] = [int(x) >> 32 for x in sq]
    return out


def _mean_sd(n, s, s2):
    """mean and sample sd (ddof = 1) from exact integer sums; the variance numerator n*s2 - s^2 is formed in
    Python integers so no cancellation happens before the final division."""
    mean = s / n
    if n < 2:
        return mean, float("nan")
    num = n * s2 - s * s
    return mean, (num / (n * (n - 1))) ** 0.5 if num > 0 else 0.0


def summary_from_moments(mom):
    """[T, 8] moments (possibly summed over ranks / rows of one group) -> dict of float arrays with the columns of
    summary.csv (esloco.py:183-195): mean / sd of full and partial matches, mean / sd / sem / 95 % CI of bases."""
    m = np.asarray(mom).astype(object)  # Python integers: exact products and differences, element by element inside numpy
    n = m[:, 0]
    nan = float("nan")

    def mean_sd(s, s2):  # _mean_sd over the rows
        num, den = n * s2 - s * s, n * (n - 1)
        sd = np.array([nan if k < 2 else ((a / b) ** 0.5 if a > 0 else 0.0) for k, a, b in zip(n, num, den)], dtype=float)
        return (s / n).astype(float), sd

    cols = {}
    cols["mean_full_matches"], cols["sd_full_matches"] = mean_sd(m[:, 1], m[:, 2])
    cols["mean_partial_matches"], cols["sd_partial_matches"] = mean_sd(m[:, 3], m[:, 4])
    cols["mean_bases_on_target"], cols["sd_bases_on_target"] = mean_sd(m[:, 5], m[:, 7] * (1 << 32) + m[:, 6])
    cols["sem_bases_on_target"] = np.array([sd / k ** 0.5 for sd, k in zip(cols["sd_bases_on_target"].tolist(), n)], dtype=float)
    cols["ci_lower_bases_on_target"] = cols["mean_bases_on_target"] - 1.96 * cols["sem_bases_on_target"]
    cols["ci_upper_bases_on_target"] = cols["mean_bases_on_target"] + 1.96 * cols["sem_bases_on_target"]
    return cols
