"""World-size-2 (gloo, CPU) test of the multi-GPU host logic: each rank takes its shard of the cost-weighted
partition, "searches" it (the oracle stands in for the GPU: same Philox keying by ORIGINAL entry index), the shards
are merged by original index and must equal the unsharded result; timing is reduced with MAX like bench.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    from pathlib import Path
    here = Path(__file__).resolve().parent
    sys.path.insert(0, str(here)); sys.path.insert(0, str(here.parent))
    import cuda_satabsearch_b200 as S
    from _refio import GOLDEN, Oracle, read_packed
    ents = read_packed(GOLDEN / "small586.satsdb")[:120]
    q = {s.name: s for s in read_packed(GOLDEN / "queries.satsdb")}["D1UBIA_"]
    db = S.Database.from_structures([s.name for s in ents], [s.tab for s in ents], [s.dmat for s in ents])
    owner = db.partition(world)
    mine = np.nonzero(owner == rank)[0].astype(np.int32)
    sc, _ = Oracle().search_philox(q, [ents[i] for i in mine], entry_ids=mine, restarts=16, seed=5, query_index=3)
    merged = torch.full((len(ents),), -(2 ** 31), dtype=torch.int32)
    merged[torch.from_numpy(mine.astype(np.int64))] = torch.from_numpy(sc)
    dist.all_reduce(merged, op=dist.ReduceOp.MAX)            # disjoint shards: MAX == gather by original index
    # bench.py's N > 1 end-to-end tail: every rank contributes its scores in DEVICE order (decreasing structure order) padded
    # to the largest shard, one all-gather, rank 0 un-permutes with the plan built once from the ranks' entry indices
    import bench
    orders = np.array([s.n for s in ents])
    dev = mine[np.argsort(-orders[mine], kind="stable")]
    by_orig = dict(zip(mine.tolist(), sc.tolist()))
    cap_t = torch.tensor([len(mine)]); dist.all_reduce(cap_t, op=dist.ReduceOp.MAX)
    cap = int(cap_t.item())
    idx_pad = torch.full((cap,), -1, dtype=torch.int32); idx_pad[:len(dev)] = torch.from_numpy(dev)
    sc_pad = torch.zeros(cap, dtype=torch.int32); sc_pad[:len(dev)] = torch.tensor([by_orig[int(e)] for e in dev], dtype=torch.int32)
    idx_all = [torch.empty_like(idx_pad) for _ in range(world)]
    sc_all = [torch.empty_like(sc_pad) for _ in range(world)]
    dist.all_gather(idx_all, idx_pad)
    dist.all_gather(sc_all, sc_pad)
    take = bench.gather_plan(torch.cat(idx_all).numpy(), len(ents))
    assembled = np.take(torch.cat(sc_all).numpy(), take)
    assert np.array_equal(assembled, merged.numpy()), "all-gather + plan must equal the merge by original index"
    t = torch.tensor([10.0 + rank], dtype=torch.float64)     # per-rank "device time"
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    cnt = torch.tensor([len(mine)]); dist.all_reduce(cnt)
    if rank == 0:
        np.save(out_path, np.concatenate([merged.numpy().astype(np.int64), [int(t.item()), int(cnt.item())]]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_merge_to_unsharded_result(tmp_path, oracle, fixtures):
    out = str(tmp_path / "merged.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    ents = fixtures["small586"][:120]
    q = fixtures["queries_by_name"]["D1UBIA_"]
    want, _ = oracle.search_philox(q, ents, restarts=16, seed=5, query_index=3)
    assert np.array_equal(got[:-2], want)
    assert got[-2] == 11 and got[-1] == len(ents)
