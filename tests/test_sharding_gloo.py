"""The N>1 path on CPU: world_size-2 gloo.  Shards tile the replica ids, keys follow the global id, and the one
collective (all-reduce of the 16-double aggregate) reproduces the single-process aggregate."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from distributed_cluster_gpus_b200 import scenarios as SC, sharding, spec as S


def test_shard_tiles_the_id_space():
    for n, w in ((65536, 8), (10, 4), (7, 8), (1, 1), (1_048_576, 8)):
        spans = [sharding.shard(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == n
        for (f0, c0), (f1, _) in zip(spans, spans[1:]):
            assert f0 + c0 == f1
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        sharding.shard(8, 8, 8)


def _worker(rank, world, port, n_total, out_path):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle_lib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        blob = SC.to_spec(dict(SC.CFG3, duration=15.0)).to_bytes()
        first, count = sharding.shard(n_total, rank, world)
        rows, _ = oracle_lib.run_batch(blob, count, 500, first)   # stand-in producer of summary rows on CPU
        vec = torch.from_numpy(sharding.aggregate_rows(rows))
        sharding.allreduce_aggregate(vec)
        if rank == 0:
            np.save(out_path, vec.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_allreduce_matches_single_process(tmp_path, oracle):
    n_total, world = 11, 2
    out = str(tmp_path / "agg.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n_total, out), nprocs=world, join=True)
    got = np.load(out)
    blob = SC.to_spec(dict(SC.CFG3, duration=15.0)).to_bytes()
    rows, _ = oracle.run_batch(blob, n_total, 500, 0)
    want = sharding.aggregate_rows(rows)
    assert got[S.A_REPLICAS] == n_total and got[S.A_EVENTS] == want[S.A_EVENTS] and got[S.A_JOBS] == want[S.A_JOBS]
    np.testing.assert_allclose(got, want, rtol=1e-12)
    stats = sharding.finalize(got)
    assert stats["replicas"] == n_total and stats["failed"] == 0 and stats["energy_j_var"] > 0
