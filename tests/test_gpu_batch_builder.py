"""On-GPU batch builder (recguru_build_batches) vs the reference loaders' row builders
``seq_padding`` / ``test_seq_gen`` (data/data_loader.py:25-55).  Integer work: bit-exact, against
rows produced by the reference itself (tests/golden/batch_layouts.npz) and against the oracle
restatement on seeded ragged data at cfg2 size."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import ref_path as O

pytestmark = pytest.mark.gpu


def _cases():
    g = load_golden("batch_layouts.npz")
    for c in range(int(g["n_cases"])):
        yield {k[len(f"c{c}_"):]: v for k, v in g.items() if k.startswith(f"c{c}_")}


def _store(c, built_lib):
    from recguru_b200.loader import SequenceStore
    Le, Ld, eos = (int(x) for x in c["dims"])
    off = c["offsets"]
    seqs = [c["items"][off[b]:off[b + 1]].tolist() for b in range(len(off) - 1)]
    return SequenceStore(seqs, c["val"], c["test"], eos, Le, Ld)


def test_train_rows_golden(built_lib):
    for c in _cases():
        st = _store(c, built_lib)
        r = st.train_batch(None)
        assert r.enc_in.dtype == torch.int64 and r.enc_in.shape == c["enc_in"].shape
        assert np.array_equal(r.enc_in.cpu().numpy(), c["enc_in"])
        assert np.array_equal(r.dec_in.cpu().numpy(), c["dec_in"])
        assert np.array_equal(r.dec_out.cpu().numpy(), c["dec_out"])
        assert np.array_equal(r.val.cpu().numpy(), c["val"]) and np.array_equal(r.test.cpu().numpy(), c["test"])


def test_eval_rows_golden(built_lib):
    for c in _cases():
        st = _store(c, built_lib)
        r = st.eval_batch(None)
        assert np.array_equal(r.eval_enc_in.cpu().numpy(), c["eval_enc_in"])
        assert np.array_equal(r.dec_in.cpu().numpy(), c["eval_dec_in"])
        assert np.array_equal(r.test_enc_in.cpu().numpy(), c["test_enc_in"])
        assert np.array_equal(r.t_dec_in.cpu().numpy(), c["t_dec_in"])
        assert np.array_equal(r.val.cpu().numpy(), c["val"]) and np.array_equal(r.test.cpu().numpy(), c["test"])


def test_user_index_selects_and_repeats(built_lib):
    c = list(_cases())[2]
    st = _store(c, built_lib)
    idx = torch.tensor([5, 0, 5, 31, 7, 7, 2], device="cuda")
    r = st.train_batch(idx)
    sel = idx.cpu().numpy()
    assert np.array_equal(r.enc_in.cpu().numpy(), c["enc_in"][sel])
    assert np.array_equal(r.dec_in.cpu().numpy(), c["dec_in"][sel])
    assert np.array_equal(r.dec_out.cpu().numpy(), c["dec_out"][sel])
    assert np.array_equal(r.val.cpu().numpy(), c["val"][sel])
    # in-place refill of static buffers (the CUDA-graph use)
    r2 = st.train_batch(torch.flip(idx, [0]), out=r)
    assert r2.enc_in.data_ptr() == r.enc_in.data_ptr()
    assert np.array_equal(r.dec_out.cpu().numpy(), c["dec_out"][sel[::-1]])


def test_cfg2_sized_batch_vs_oracle(built_lib):
    """B=1024, L=200 ragged Amazon-shaped histories: every row against the oracle."""
    from recguru_b200.loader import SequenceStore
    rng = np.random.default_rng(7)
    L, eos, n_users = 200, 100001, 3000
    lens = np.clip(3 + np.floor(rng.lognormal(np.log(8), 1.0, n_users)), 0, 450).astype(int)
    lens[:4] = [0, 199, 200, 201]
    seqs = [rng.integers(1, eos, n).tolist() for n in lens]
    val, test = rng.integers(1, eos, n_users), rng.integers(1, eos, n_users)
    st = SequenceStore(seqs, val, test, eos, L)
    idx = np.concatenate([np.arange(4), rng.integers(0, n_users, 1020)])
    r = st.train_batch(torch.from_numpy(idx).cuda())
    ev = st.eval_batch(torch.from_numpy(idx).cuda())
    enc, din, dout = (x.cpu().numpy() for x in (r.enc_in, r.dec_in, r.dec_out))
    for b, u in enumerate(idx):
        e, di, do = O.seq_padding(seqs[u], L, L, eos)
        assert np.array_equal(enc[b], e) and np.array_equal(din[b], di) and np.array_equal(dout[b], do)
    ee, edi, te, tdi = (x.cpu().numpy() for x in (ev.eval_enc_in, ev.dec_in, ev.test_enc_in, ev.t_dec_in))
    for b, u in enumerate(idx[:64]):
        a = O.test_seq_gen(seqs[u], L, L, eos, int(val[u]))
        assert np.array_equal(ee[b], a[0]) and np.array_equal(edi[b], a[1])
        assert np.array_equal(te[b], a[2]) and np.array_equal(tdi[b], a[3])
    # size-independent properties: dec_out is dec_in advanced by one item -- except that the
    # reference forces the first TWO decoder positions of both rows to 0 ([0, 0] + enc_in[1:-1]), so
    # a full window loses enc_in[0] in dec_out; the mask of the training step (dec_in != 0) marks
    # clip(len - 1, 0, L - 2) positions
    assert np.array_equal(din[:, 3:], dout[:, 2:-1])
    assert not din[:, :2].any() and not dout[:, :2].any()
    n_live = (din != 0).sum(1)
    assert np.array_equal(n_live, np.clip(lens[idx] - 1, 0, L - 2))


def test_builder_feeds_the_hot_path(built_lib):
    """Rows built on the device drive K1 exactly like host-built rows."""
    from recguru_b200 import ops
    from recguru_b200.loader import SequenceStore
    rng = np.random.default_rng(3)
    L, n_items, d = 20, 300, 64
    seqs = [rng.integers(1, n_items + 1, int(n)).tolist() for n in rng.integers(0, 30, 16)]
    st = SequenceStore(seqs, None, None, n_items + 1, L)
    r = st.train_batch(None)
    host = np.stack([O.seq_padding(s, L, L, n_items + 1)[1] for s in seqs])
    W = torch.randn(n_items + 2, d, device="cuda")
    W[0].zero_()
    pe = O.pe_table(L, d)[0].cuda()
    a, ma = ops.raw_gather_pe_mask(W, r.dec_in, pe, 0)
    b, mb = ops.raw_gather_pe_mask(W, torch.from_numpy(host).cuda(), pe, 0)
    assert torch.equal(a, b) and torch.equal(ma, mb)


def test_builder_errors(built_lib):
    from recguru_b200._lib import RecGuruError
    from recguru_b200.loader import SequenceStore, raw_build_batches, BATCH_TRAIN
    with pytest.raises(RuntimeError):
        SequenceStore([[1, 2]], None, None, 9, 6, device="cpu")
    st = SequenceStore([[1, 2], [3]], [4, 5], [6, 7], 9, 6)
    with pytest.raises(RuntimeError):
        st.train_batch(None, B=3)                       # more rows than stored sequences
    with pytest.raises(RuntimeError):
        st.train_batch(torch.zeros(2, dtype=torch.int64))   # CPU index tensor
    enc = torch.empty(2, 6, dtype=torch.int64, device="cuda")
    dec = torch.empty(2, 8, dtype=torch.int64, device="cuda")
    with pytest.raises(RecGuruError):                   # decoder window longer than the encoder window
        raw_build_batches(st.items, st.offsets, None, None, 2, 6, 8, 9, BATCH_TRAIN, enc, dec, None)
    assert st.train_batch(None, B=0).enc_in.shape == (0, 6)
