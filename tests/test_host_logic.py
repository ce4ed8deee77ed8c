"""Host-side logic that needs no GPU: synthetic data layout, shard arithmetic, metric folding,
op registration / shape inference, loud failure on CPU tensors."""
import numpy as np
import pytest
import torch

from oracle import ref_path as O


def test_synth_batch_layout_matches_reference_padding():
    from recguru_b200 import synth
    B, L, N, K = 64, 50, 10_000, 5
    b = synth.train_batch(B, L, N, K)
    assert b.enc_in.shape == (B, L) and b.n_items.shape == (B, L * K) and b.enc_in.dtype == torch.int64
    for i in range(B):
        row = b.enc_in[i]
        seq = row[(row != 0) & (row != N + 1)].tolist()
        assert 3 <= len(seq) <= L - 1
        e, di, do = O.seq_padding(seq, L, L, N + 1)      # the reference's own layout
        assert row.tolist() == e.tolist()
        assert b.dec_in[i].tolist() == di.tolist()
        assert b.dec_out[i].tolist() == do.tolist()
    assert torch.equal(b.mask, (b.dec_in != 0).float().view(-1))
    assert int(b.n_items.min()) >= 1 and int(b.n_items.max()) <= N
    b2 = synth.train_batch(B, L, N, K)
    assert torch.equal(b.enc_in, b2.enc_in) and torch.equal(b.n_items, b2.n_items)   # seeded


def test_synth_zipf_and_tables():
    from recguru_b200 import synth
    ids = synth.zipf_ids(200_000, 1000)
    assert int(ids.min()) >= 1 and int(ids.max()) <= 1000
    counts = torch.bincount(ids, minlength=1001).float()
    top = counts.sort(descending=True).values
    assert top[0] > 5 * top[50]                          # heavy head
    W = synth.item_table(100, 64)
    assert W.shape == (102, 64) and float(W[0].abs().max()) == 0.0
    G = synth.item_table(100, 64, grid=True)
    assert bool(((G * 16).round() == G * 16).all()) and float(G.abs().max()) <= 2.0
    # grid scores are exact in any accumulation order: fp32 and fp64 agree bit for bit
    u = synth.user_reps(8, 64, grid=True)
    assert torch.equal((u @ G.t()), (u.double() @ G.double().t()).float())


def test_shard_bounds_cover_catalog():
    from recguru_b200.catalog import shard_bounds
    for n in (1, 7, 1000, 10_000_000):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= -(-n // w)


def test_metrics_from_histogram_matches_reference_loops():
    from recguru_b200.catalog import metrics_from_histogram
    rng = np.random.RandomState(0)
    r = rng.randint(0, 60, size=5000)
    kmax = 30
    hist = np.bincount(np.minimum(r, kmax), minlength=kmax + 1)
    m = metrics_from_histogram(hist, len(r), [1, 5, 10, 20, 30])
    for k in (1, 5, 10, 20, 30):
        assert abs(m[k]["hr"] - O.hit_at_k_batch(r, k)) < 1e-12
        assert abs(m[k]["ndcg"] - O.NDCG_at_k_batch(r, k)) < 1e-12
        assert abs(m[k]["mrr"] - O.mrr_at_k_batch(r, k)) < 1e-12
    m = metrics_from_histogram(np.bincount([0, 3, 12], minlength=31), 3, [10])
    assert abs(m[10]["hr"] - 2 / 3) < 1e-12 and abs(m[10]["mrr"] - 1.25 / 3) < 1e-12
    with pytest.raises(ValueError):
        metrics_from_histogram(hist, len(r), [31])


def test_ops_registered_with_shape_inference(built_lib):
    import recguru_b200.ops  # noqa: F401
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        W = torch.empty(100, 64, device="cuda")
        ids = torch.empty(4, 10, dtype=torch.int64, device="cuda")
        pe = torch.empty(1, 50, 64, device="cuda")
        out, m = torch.ops.recguru.gather_pe_mask(W, ids, pe, 0)
        assert out.shape == (4, 10, 64) and m.shape == (4, 10)
        assert torch.ops.recguru.gather_rows(W, ids, 0).shape == (4, 10, 64)
        h = torch.empty(4, 10, 64, device="cuda")
        neg = torch.empty(4, 50, dtype=torch.int64, device="cuda")
        p, n = torch.ops.recguru.sampled_scores(h, W, ids, neg, 0)
        assert p.shape == (4, 10, 1) and n.shape == (4, 10, 5)
        aux, p, n, dp, dn = torch.ops.recguru.sampled_score_loss(h, W, ids, neg, None, "bpr", 0)
        assert aux.shape == (2,) and dp.shape == (40,) and dn.shape == (40, 5)
        u = torch.empty(33, 64, device="cuda")
        v, i, r = torch.ops.recguru.catalog_topk_rank(u, W.to(torch.bfloat16), torch.empty(33, dtype=torch.int64, device="cuda"), 10, 1)
        assert v.shape == (33, 10) and i.dtype == torch.int64 and r.shape == (33,)
        assert torch.ops.recguru.embedding_bwd(h.view(-1, 64), ids.view(-1), None, 100, 0).shape == (100, 64)


def test_no_cpu_fallback(built_lib):
    """The product path fails loudly on CPU tensors instead of falling back."""
    from recguru_b200 import ops, modules
    W = torch.randn(10, 32)
    ids = torch.zeros(2, 3, dtype=torch.int64)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.raw_gather_pe_mask(W, ids, None, -1)
    with pytest.raises(Exception):
        modules.FusedItemEmbedding(10, 32, padding_idx=0)(ids)      # no CPU kernel registered
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.raw_embedding_bwd([ops.Site(ids, torch.randn(6, 32))], 10, 32)


def test_modules_mirror_reference_surface():
    from recguru_b200 import modules as M
    emb = M.FusedItemEmbedding(12, 32, padding_idx=0)
    assert list(emb.state_dict().keys()) == ["weight"] and emb.padding_idx == 0
    assert float(emb.weight[0].abs().max()) == 0.0
    ref = torch.nn.Embedding(12, 32, padding_idx=0)
    fused = M.FusedItemEmbedding.from_embedding(ref)
    assert fused.weight.data_ptr() == ref.weight.data_ptr()          # shares storage
    pos = M.FusedPositionalEncoding(32, 0.5, max_len=64)
    assert list(pos.state_dict().keys()) == ["pe"] and pos.pe.shape == (1, 64, 32)
    assert torch.equal(pos.pe, O.pe_table(64, 32))
    assert M.FusedBPRLoss.kind == "bpr" and M.FusedBPRLoss_sas.kind == "bpr_sas"


def test_product_never_imports_oracle():
    import re
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent / "recguru_b200"
    for f in root.rglob("*.py"):
        assert not re.search(r"^\s*(from|import)\s+oracle", f.read_text(), flags=re.M), f


def test_alias_table_reproduces_weights():
    """Walker alias table of the negative sampler (recguru_b200/sampler.py): marginal == weights."""
    import numpy as np
    from recguru_b200.sampler import build_alias
    w = np.array([0, 3, 1, 0, 6, 2.5, 0.5], dtype=np.float64)
    prob, alias = build_alias(w)
    n = len(w)
    p = np.zeros(n)
    for i in range(n):
        p[i] += prob[i] / n
        p[alias[i]] += (1 - prob[i]) / n
    assert np.allclose(p, w / w.sum(), atol=1e-6)
    g = np.random.default_rng(0).random(1000)
    prob, alias = build_alias(g)
    p = np.zeros(1000)
    np.add.at(p, np.arange(1000), prob / 1000)
    np.add.at(p, alias, (1 - prob) / 1000)
    assert np.allclose(p, g / g.sum(), atol=1e-7)


def test_sequence_store_host_logic():
    """The packing of the pickles' ragged lists (CSR) and the loud failures of the loader module on a
    box without a GPU."""
    from recguru_b200 import loader
    seqs = [[5, 6, 7], [], [9], list(range(1, 12))]
    items, offsets = loader.pack_sequences(seqs)
    assert items.dtype == np.int64 and offsets.tolist() == [0, 3, 3, 4, 15]
    for s, lo, hi in zip(seqs, offsets[:-1], offsets[1:]):
        assert items[lo:hi].tolist() == s
    e, o = loader.pack_sequences([])
    assert e.size == 0 and o.tolist() == [0]
    with pytest.raises(RuntimeError, match="no CPU path"):
        loader.SequenceStore(seqs, None, None, eos=99, enc_len=6, device="cpu")
    # CPU tensors are refused before anything is launched
    t = torch.zeros(4, dtype=torch.int64)
    with pytest.raises(RuntimeError, match="CUDA"):
        loader.raw_build_batches(t, t, None, None, 1, 6, 6, 99, loader.BATCH_TRAIN, None, None, None)


def test_new_entry_points_refuse_cpu_tensors():
    from recguru_b200 import ops
    w = torch.zeros(8, 4)
    e = torch.zeros(8, dtype=torch.int32)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.raw_table_adam_rows(w, w.clone(), w.clone(), w.clone(), e, 1, 1e-3, 0.9, 0.99, 1e-9, 1)
    with pytest.raises(RuntimeError):
        ops.raw_pad_mask_count(torch.zeros(4, dtype=torch.int64), 0, torch.zeros(4), torch.zeros(3, dtype=torch.int32))
    from recguru_b200.sampler import NegativeSampler
    with pytest.raises(RuntimeError, match="CUDA"):
        NegativeSampler.advance(torch.zeros(1, dtype=torch.int64))


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver times beside ours) runs without a GPU and
    prints ONE JSON line carrying the contract keys: same metric/unit as our arm, impl, a
    cpu_baseline describing the run, and an e2e object with zero transfer bytes."""
    import json, subprocess, sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "train seqs/sec" and j["unit"] == "seqs/s"
    assert j["higher_is_better"] is True and j["n_gpus"] == 1 and j["warmup"] >= 3 and j["value"] > 0
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"] and j["vs_baseline"] is None
