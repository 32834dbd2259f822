"""Multi-GPU: replicas shard embarrassingly; one all-reduce of a 16-double vector ends the run.

Replica keys derive from the GLOBAL replica id (base_seed + id), so any (rank, world) split reproduces the same
trajectories.  There is no data-path collective: ``allreduce_aggregate`` is the only communication.
"""
import numpy as np

from . import spec as S


def shard(n_total: int, rank: int, world: int):
    """Contiguous split of replica ids [0, n_total) -> (first_id, count) of ``rank``; sizes differ by <= 1."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, extra = divmod(n_total, world)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def aggregate_rows(summary: np.ndarray) -> np.ndarray:
    """Host mirror of csrc dcsim_reduce_kernel (same components, spec.A_*)."""
    a = np.zeros(S.AGG_K)
    fin = summary[:, S.S_JOBS_FINISHED]
    e = summary[:, S.S_TOTAL_ENERGY_J]
    ml = np.divide(summary[:, S.S_LAT_SUM], fin, out=np.zeros_like(fin), where=fin > 0)
    a[S.A_REPLICAS] = len(summary)
    a[S.A_FAILED] = np.count_nonzero((summary[:, S.S_STATUS] != 0) | (summary[:, S.S_DONE] == 0))
    a[S.A_EVENTS] = summary[:, S.S_EVENTS].sum()
    a[S.A_JOBS] = fin.sum()
    a[S.A_ENERGY] = e.sum()
    a[S.A_ENERGY_SQ] = (e * e).sum()
    a[S.A_LAT_SUM] = summary[:, S.S_LAT_SUM].sum()
    a[S.A_MEANLAT_SUM] = ml.sum()
    a[S.A_MEANLAT_SQ] = (ml * ml).sum()
    a[S.A_RNG_WORDS] = summary[:, S.S_RNG_WORDS].sum()
    a[S.A_RUNNING] = np.count_nonzero((summary[:, S.S_STATUS] == 0) & (summary[:, S.S_DONE] == 0))
    return a


def allreduce_aggregate(vec):
    """Sum the per-rank aggregate vectors in place (torch tensor; NCCL over NVLink on GPUs, gloo on CPU)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
    return vec


def finalize(agg) -> dict:
    """Aggregate vector -> batch statistics (means and unbiased variances across replicas)."""
    a = np.asarray(agg, dtype=np.float64)
    n = a[S.A_REPLICAS]
    var = lambda s, sq: float(max(0.0, (sq - s * s / n) / (n - 1))) if n > 1 else 0.0  # noqa: E731
    return {"replicas": int(n), "failed": int(a[S.A_FAILED]), "events": float(a[S.A_EVENTS]),
            "jobs_finished": float(a[S.A_JOBS]), "energy_j_mean": float(a[S.A_ENERGY] / n),
            "energy_j_var": var(a[S.A_ENERGY], a[S.A_ENERGY_SQ]),
            "mean_latency_s_mean": float(a[S.A_MEANLAT_SUM] / n),
            "mean_latency_s_var": var(a[S.A_MEANLAT_SUM], a[S.A_MEANLAT_SQ]),
            "rng_words": float(a[S.A_RNG_WORDS])}
