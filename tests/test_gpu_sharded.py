"""GPU parity of the row-sharded catalog (piece 5) over NCCL: world_size 2 on one box.  Skipped
when fewer than 2 GPUs are visible.  Sharded results must equal the single-shard oracle bit for
bit (grid-valued data, exact ties across the shard boundary)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_path as O

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _data(n_items, U, d):
    g = torch.Generator().manual_seed(4242)
    W = torch.randint(-32, 33, (n_items + 2, d), generator=g).float() / 16
    W[0] = 0
    W[n_items // 2 - 10: n_items // 2 + 10] = W[5]          # exact ties straddling the shard boundary
    users = torch.randint(-32, 33, (U, d), generator=g).float() / 16
    target = torch.randint(1, n_items + 1, (U,), generator=g)
    target[:8] = n_items // 2 + torch.arange(-4, 4)
    return W, users, target


def _worker(rank, world, port, n_items, U, d, k, out, counts=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from recguru_b200.catalog import CatalogEvaluator
        W, users, target = _data(n_items, U, d)
        ev = CatalogEvaluator(W.cuda(), k=k, group=dist.group.WORLD)
        per = U // world
        sl = slice(rank * per, (rank + 1) * per) if counts is None else \
            slice(sum(counts[:rank]), sum(counts[:rank + 1]))
        val, idx, rk = ev.rank_topk(users[sl].cuda(), target[sl].cuda())
        met = ev.evaluate_global(users[sl].cuda(), target[sl].cuda(), ks=(5, 10))
        out[rank] = (val.cpu(), idx.cpu(), rk.cpu(), met)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("n_items,U,d,k", [(5000, 512, 128, 10), (1001, 256, 64, 5)])
def test_sharded_catalog_nccl_matches_oracle(built_lib, n_items, U, d, k):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_items, U, d, k, out), nprocs=world, join=True)
    W, users, target = _data(n_items, U, d)
    s = O.catalog_scores(users, W[1:n_items + 1])
    ref_rank = O.catalog_rank(s, target - 1)
    rv, ri = O.catalog_topk(s, k)
    per = U // world
    for r in range(world):
        val, idx, rk, met = out[r]
        sl = slice(r * per, (r + 1) * per)
        assert torch.equal(rk, ref_rank[sl])
        assert torch.equal(idx, ri[sl] + 1) and torch.equal(val, rv[sl])
        assert abs(met[10]["hr"] - O.hit_at_k_batch(ref_rank.tolist(), 10)) < 1e-12
        assert abs(met[5]["ndcg"] - O.NDCG_at_k_batch(ref_rank.tolist(), 5)) < 1e-9


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_catalog_nccl_ragged_users(built_lib):
    """Ranks holding different numbers of users (a last partial batch): padded to the maximum
    inside rank_topk, results sliced back -- bit-equal to the single-shard oracle."""
    world, n_items, d, k = 2, 3000, 128, 10
    counts = [300, 77]
    U = sum(counts)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_items, U, d, k, out, counts), nprocs=world, join=True)
    W, users, target = _data(n_items, U, d)
    s = O.catalog_scores(users, W[1:n_items + 1])
    ref_rank = O.catalog_rank(s, target - 1)
    rv, ri = O.catalog_topk(s, k)
    for r in range(world):
        val, idx, rk, met = out[r]
        sl = slice(sum(counts[:r]), sum(counts[:r + 1]))
        assert rk.shape[0] == counts[r]
        assert torch.equal(rk, ref_rank[sl]) and torch.equal(idx, ri[sl] + 1) and torch.equal(val, rv[sl])
        assert abs(met[10]["hr"] - O.hit_at_k_batch(ref_rank.tolist(), 10)) < 1e-12
