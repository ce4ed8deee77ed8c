"""GPU negative sampler vs the reference's sampling law (data/data_loader.py:289-314): the weights
vector with id 0, the sequence, val and test zeroed, multinomial with replacement.  The random
stream is ours (Philox), so parity is distributional: exclusions are never drawn, every draw is in
range, and the empirical frequencies match the reference weights (chi-square, 5-sigma bound)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def reference_weights(v_size, seq, val, test, item_freq=None, exclude_val=True):
    """The weight vector the reference hands to torch.multinomial: oracle.sampling_weights, pinned to the
    unmodified loader by tests/golden/sampler_weights.npz (test_oracle_golden.py).  exclude_val=True is
    the repaired law NegativeSampler defaults to; False the reference's own (its val exclusion is a no-op)."""
    from oracle import ref_path as O
    w = O.sampling_weights(seq, val, test, v_size, item_freq=None if item_freq is None else item_freq[:v_size],
                           exclude_val=exclude_val)
    w[0] = 0                                      # id 0 is never a negative here (wf[0] == 0 in the tests)
    return w / w.sum()


def make_batch(B, H, v_size, seed):
    g = np.random.default_rng(seed)
    hist = np.zeros((B, H), dtype=np.int64)
    seqs = []
    for b in range(B):
        n = int(g.integers(1, H))
        s = g.integers(1, v_size, n)
        hist[b, H - n:] = s                      # left padded like seq_padding
        seqs.append(s)
    val = g.integers(1, v_size, B)
    test = g.integers(1, v_size, B)
    return hist, seqs, val, test


@pytest.mark.parametrize("weighted", [False, True])
def test_sampler_distribution_and_exclusions(built_lib, weighted):
    from recguru_b200.sampler import NegativeSampler
    B, H, v_size, num_n = 6, 12, 60, 200_000
    hist, seqs, val, test = make_batch(B, H, v_size, 3)
    freq = None
    if weighted:
        freq = np.random.default_rng(5).integers(1, 1000, v_size + 1).astype(np.float64)
    smp = NegativeSampler(v_size, num_n, seed=1234, item_freq=freq)
    out = smp.sample(torch.from_numpy(hist).cuda(), torch.from_numpy(val).cuda(), torch.from_numpy(test).cuda()).cpu().numpy()
    assert out.shape == (B, num_n) and out.min() >= 1 and out.max() < v_size
    for b in range(B):
        p = reference_weights(v_size, seqs[b], val[b], test[b], freq)
        cnt = np.bincount(out[b], minlength=v_size).astype(np.float64)
        assert cnt[p == 0].sum() == 0                                   # exclusions never drawn
        exp = p * num_n
        sigma = np.sqrt(num_n * p * (1 - p))
        assert np.all(np.abs(cnt - exp)[p > 0] < 5.5 * sigma[p > 0] + 1)  # per-id 5.5 sigma
        chi2 = (((cnt - exp) ** 2)[p > 0] / exp[p > 0]).sum()
        dof = int((p > 0).sum()) - 1
        assert chi2 < dof + 6 * np.sqrt(2 * dof)                         # chi-square within 6 sigma


def test_sampler_reference_law_keeps_validation_item(built_lib):
    """exclude_val=False: the reference's exact law (data_loader.py:291-295: `for val in seq` rebinds val, the
    validation item stays a legal negative) -- chi-square against the loader-pinned oracle weights."""
    from recguru_b200.sampler import NegativeSampler
    B, H, v_size, num_n = 6, 40, 300, 200_000
    hist, seqs, val, test = make_batch(B, H, v_size, 21)
    for b in range(B):                                                   # make sure val is not excluded by accident
        while val[b] in set(seqs[b].tolist()) | {int(test[b])}:
            val[b] = val[b] % (v_size - 1) + 1
    smp = NegativeSampler(v_size, num_n, seed=5, exclude_val=False)
    out = smp.sample(torch.from_numpy(hist).cuda(), torch.from_numpy(val).cuda(), torch.from_numpy(test).cuda()).cpu().numpy()
    fixed = NegativeSampler(v_size, num_n, seed=5).sample(torch.from_numpy(hist).cuda(), torch.from_numpy(val).cuda(),
                                                          torch.from_numpy(test).cuda()).cpu().numpy()
    for b in range(B):
        p = reference_weights(v_size, seqs[b], val[b], test[b], exclude_val=False)
        assert p[val[b]] > 0
        cnt = np.bincount(out[b], minlength=v_size).astype(np.float64)
        assert cnt[p == 0].sum() == 0 and cnt[val[b]] > 0               # val IS drawn under the reference's law
        exp = p * num_n
        chi2 = (((cnt - exp) ** 2)[p > 0] / exp[p > 0]).sum()
        dof = int((p > 0).sum()) - 1
        assert chi2 < dof + 6 * np.sqrt(2 * dof)
        assert np.bincount(fixed[b], minlength=v_size)[val[b]] == 0     # ... and never under the repaired one


def test_sampler_is_deterministic_in_seed_and_step(built_lib):
    from recguru_b200.sampler import NegativeSampler
    B, H, v_size, num_n = 33, 50, 5000, 1000
    hist, seqs, val, test = make_batch(B, H, v_size, 9)
    h, v, t = (torch.from_numpy(x).cuda() for x in (hist, val, test))
    a = NegativeSampler(v_size, num_n, seed=7).sample(h, v, t, step=3)
    b = NegativeSampler(v_size, num_n, seed=7).sample(h, v, t, step=3)
    c = NegativeSampler(v_size, num_n, seed=7).sample(h, v, t, step=4)
    d = NegativeSampler(v_size, num_n, seed=8).sample(h, v, t, step=3)
    assert torch.equal(a, b) and not torch.equal(a, c) and not torch.equal(a, d)
    for bi in range(B):                                                   # exclusions at a realistic size
        bad = set(seqs[bi].tolist()) | {int(val[bi]), int(test[bi]), 0}
        assert not (set(a[bi].cpu().tolist()) & bad)


def test_sampler_device_step_counter_matches_immediate_step(built_lib):
    """recguru_sample_negatives_devstep reads the step from device memory (CUDA-graph use): same draws
    as the immediate-step call, and a captured [sample, advance] graph draws fresh negatives on every
    replay."""
    from recguru_b200.sampler import NegativeSampler
    hist, _, val, test = make_batch(8, 20, 500, seed=5)
    h, v, t = (torch.from_numpy(x).cuda() for x in (hist, val, test))
    smp = NegativeSampler(500, 64, seed=11)
    step_dev = torch.tensor([7], dtype=torch.int64, device="cuda")
    a = smp.sample(h, v, t, step_dev=step_dev)
    assert torch.equal(a, smp.sample(h, v, t, step=7))
    smp.advance(step_dev, 2)
    assert int(step_dev.item()) == 9
    out = torch.empty_like(a)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        smp.sample(h, v, t, step_dev=step_dev, out=out)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        smp.sample(h, v, t, step_dev=step_dev, out=out)
        smp.advance(step_dev)
    for s in (9, 10, 11):
        g.replay()
        assert torch.equal(out, smp.sample(h, v, t, step=s)), s
    assert int(step_dev.item()) == 12


def test_sampler_from_the_ragged_store(built_lib):
    """recguru_sample_negatives_csr: exclusions come from the user's WHOLE stored sequence (the
    reference zeroes weights[seq] for the full list, data_loader.py:291-293), val/test are looked up
    per user; for users whose sequence fits the enc_in window the draws are identical to the matrix
    form fed with the device-built enc_in rows."""
    from recguru_b200.loader import SequenceStore
    from recguru_b200.sampler import NegativeSampler
    rng = np.random.default_rng(21)
    v_size, L, num_n = 400, 12, 5000
    lens = [0, 1, 5, 11, 12, 40, 150, 3, 7, 11]
    seqs = [rng.integers(1, v_size, n).tolist() for n in lens]
    val, test = rng.integers(1, v_size, len(seqs)), rng.integers(1, v_size, len(seqs))
    store = SequenceStore(seqs, val, test, eos=v_size, enc_len=L)
    assert store.max_len == 150
    idx = torch.tensor([6, 0, 5, 2, 9, 3, 4, 1, 8, 7], device="cuda")
    smp = NegativeSampler(v_size, num_n, seed=3)
    out = smp.sample_store(store, idx, step=5).cpu().numpy()
    assert out.shape == (len(idx), num_n) and out.min() >= 1 and out.max() < v_size
    for b, u in enumerate(idx.tolist()):
        bad = set(seqs[u]) | {int(val[u]), int(test[u])}
        assert not (set(out[b].tolist()) & bad), (b, u)               # incl. items older than the window
        allowed = v_size - 1 - len(bad - {0})
        assert len(set(out[b].tolist())) > 0.9 * min(allowed, num_n) * (1 - np.exp(-num_n / allowed)) * 0.9
    rows = store.train_batch(idx)
    mat = smp.sample(rows.enc_in, rows.val, rows.test, step=5).cpu().numpy()
    for b, u in enumerate(idx.tolist()):
        if len(seqs[u]) <= L - 1:                                       # whole sequence visible in enc_in
            assert np.array_equal(mat[b], out[b]), (b, u)
    assert not np.array_equal(mat[0], out[0])                           # user 6: 150 items, window of 11
    # device step counter and identity batch
    step_dev = torch.tensor([5], dtype=torch.int64, device="cuda")
    assert np.array_equal(smp.sample_store(store, idx, step_dev=step_dev).cpu().numpy(), out)
    first = smp.sample_store(store, None, B=4, step=5).cpu().numpy()
    assert np.array_equal(first, smp.sample_store(store, torch.arange(4, device="cuda"), step=5).cpu().numpy())
