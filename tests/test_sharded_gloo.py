"""The N>1 host logic of the row-sharded catalog (piece 5) on CPU: world_size 2 and 3 over gloo.
The CUDA compute methods of CatalogEvaluator are overridden IN THE TEST by the oracle, so what is
exercised is the product's shard arithmetic and its collective sequence (all-gather users,
all-reduce target score, all-to-all of per-shard top-K / counts, merge order)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_path as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_oracle_evaluator():
    from recguru_b200.catalog import CatalogEvaluator

    class OracleEvaluator(CatalogEvaluator):
        def _prepare_items(self, rows):
            return rows.to(torch.bfloat16)

        def _prepare_users(self, users):
            return users.to(torch.bfloat16)

        def _scores(self, u):
            return O.catalog_scores(u.float(), self.items.float())

        def _local_target_scores(self, u, target):
            s = self._scores(u)
            loc = target - self.id_lo
            ok = (loc >= 0) & (loc < self.n_local)
            out = torch.zeros(u.shape[0])
            idx = torch.arange(u.shape[0])[ok]
            out[ok] = s[idx, loc[ok]]
            return out

        def _local_topk_rank(self, u, target, ts):
            s = self._scores(u)
            ids = torch.arange(self.n_local)[None, :] + self.id_lo
            cnt = (s > ts[:, None]).sum(1) + ((s == ts[:, None]) & (ids < target[:, None])).sum(1)
            k = min(self.k, self.n_local)
            v, i = O.catalog_topk(s, k)
            val = torch.full((u.shape[0], self.k), float("-inf")); idx = torch.full((u.shape[0], self.k), -1)
            val[:, :k] = v; idx[:, :k] = i + self.id_lo
            return val, idx, cnt.to(torch.int64)

        def _local_rank(self, u, target, ts):
            return self._local_topk_rank(u, target, ts)[2]

        def _merge(self, vals, ids, counts):
            G, U, k = vals.shape
            v = vals.permute(1, 0, 2).reshape(U, G * k)
            i = ids.permute(1, 0, 2).reshape(U, G * k)
            key_i = torch.where(i < 0, torch.full_like(i, 1 << 40), i)
            order = torch.argsort(key_i, dim=1, stable=True)              # asc id ...
            v2 = v.gather(1, order)
            order2 = torch.argsort(-v2, dim=1, stable=True)               # ... then desc score, stable
            sel = order.gather(1, order2)[:, :k]
            return v.gather(1, sel), i.gather(1, sel), counts.sum(0)

    return OracleEvaluator


def _worker(rank, world, port, n_items, U, k, out, counts=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(99)
        d = 64
        W = torch.randint(-32, 33, (n_items + 2, d), generator=g).float() / 16
        W[0] = 0
        if n_items > 100:
            W[40:60] = W[7]                                                # exact ties across shards
        users = torch.randint(-32, 33, (U, d), generator=g).float() / 16
        target = torch.randint(1, n_items + 1, (U,), generator=g)
        Ev = _make_oracle_evaluator()
        ev = Ev(W, k=k, group=dist.group.WORLD)
        assert ev.world == world
        if counts is None:
            per = U // world
            sl = slice(rank * per, (rank + 1) * per)
        else:                                   # ragged: rank r holds counts[r] users (possibly none)
            sl = slice(sum(counts[:rank]), sum(counts[:rank + 1]))
        val, idx, rk = ev.rank_topk(users[sl], target[sl])
        rk_only = ev.target_rank(users[sl], target[sl])                            # rank-only exchange
        assert torch.equal(rk_only, rk)
        out[rank] = (val, idx, rk, ev.id_lo, ev.n_local)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_items", [(2, 1000), (3, 301), (2, 5)])
def test_sharded_catalog_matches_single_shard(world, n_items):
    U, k = 12 * world, 10
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, n_items, U, k, out), nprocs=world, join=True)
    # single-shard oracle on the same bytes
    g = torch.Generator().manual_seed(99)
    d = 64
    W = torch.randint(-32, 33, (n_items + 2, d), generator=g).float() / 16
    W[0] = 0
    if n_items > 100:
        W[40:60] = W[7]
    users = torch.randint(-32, 33, (U, d), generator=g).float() / 16
    target = torch.randint(1, n_items + 1, (U,), generator=g)
    s = O.catalog_scores(users, W[1:n_items + 1])
    ref_rank = O.catalog_rank(s, target - 1)
    kk = min(k, n_items)
    rv, ri = O.catalog_topk(s, kk)
    per = U // world
    covered = 0
    for r in range(world):
        val, idx, rk, id_lo, n_local = out[r]
        covered += n_local
        sl = slice(r * per, (r + 1) * per)
        assert torch.equal(rk, ref_rank[sl])
        assert torch.equal(idx[:, :kk], ri[sl] + 1)
        assert torch.equal(val[:, :kk], rv[sl])
    assert covered == n_items


@pytest.mark.parametrize("world,n_items,counts", [(2, 1000, [7, 12]), (4, 5, [3, 0, 5, 2]), (3, 2, [1, 1, 4])])
def test_sharded_ragged_users_and_empty_shards(world, n_items, counts):
    """a last partial batch that differs across ranks, a rank with no users at all, and
    shards that own no items (n_items < world) must neither hang nor mis-slice."""
    U, k = sum(counts), 10
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, n_items, U, k, out, counts), nprocs=world, join=True)
    g = torch.Generator().manual_seed(99)
    d = 64
    W = torch.randint(-32, 33, (n_items + 2, d), generator=g).float() / 16
    W[0] = 0
    if n_items > 100:
        W[40:60] = W[7]
    users = torch.randint(-32, 33, (U, d), generator=g).float() / 16
    target = torch.randint(1, n_items + 1, (U,), generator=g)
    s = O.catalog_scores(users, W[1:n_items + 1])
    ref_rank = O.catalog_rank(s, target - 1)
    kk = min(k, n_items)
    rv, ri = O.catalog_topk(s, kk)
    covered = 0
    for r in range(world):
        val, idx, rk, id_lo, n_local = out[r]
        covered += n_local
        sl = slice(sum(counts[:r]), sum(counts[:r + 1]))
        assert rk.shape[0] == counts[r] and val.shape == (counts[r], k)
        assert torch.equal(rk, ref_rank[sl])
        assert torch.equal(idx[:, :kk], ri[sl] + 1)
        assert torch.equal(val[:, :kk], rv[sl])
        assert (idx[:, kk:] == -1).all()
    assert covered == n_items
