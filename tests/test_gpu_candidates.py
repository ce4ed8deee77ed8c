"""GPU parity: the reference's sampled-candidate evaluation (get_scores + rank of column 0,
train_auto.py:164-199,:234-237) as one kernel, against the oracle.  Grid-valued data makes every
fp32 dot product exact and order independent => scores and ranks bit-exact; Gaussian data is
checked within 1e-5 relative (fp32 summation order) and the rank against the kernel's own scores."""
import pytest
import torch

from conftest import load_golden, t
from oracle import ref_path as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cat(built_lib):
    from recguru_b200 import catalog
    return catalog


def grid(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(-32, 33, shape, generator=g).float() / 16.0


@pytest.mark.parametrize("d", [32, 64, 128, 256])
@pytest.mark.parametrize("B,C,V", [(1, 1, 5), (3, 7, 50), (128, 199, 3000), (17, 1000, 700), (4, 19999, 30000)])
def test_candidate_scores_rank_exact_on_grid(cat, d, B, C, V):
    g = torch.Generator().manual_seed(B * 131 + C)
    W = grid((V + 2, d), 3)
    u = grid((B, d), 4)
    target = torch.randint(1, V + 1, (B,), generator=g)
    cand = torch.randint(1, V + 1, (B, C), generator=g)
    cand[0, 0] = target[0]                      # an exact tie: the target must still win it
    ref = O.candidate_scores(u, W, target, cand.reshape(-1), C).reshape(B, C + 1)
    # spec = stable reading of the double argsort: column 0 wins ties (torch.argsort without
    # stable=True leaves tied order unspecified, so the idiom itself is only a bracket check)
    ref_rank = O.catalog_rank(ref, torch.zeros(B, dtype=torch.int64))
    idiom, ties = O.rank_of_column0(ref), (ref[:, 1:] == ref[:, :1]).sum(1)
    assert bool(((idiom >= ref_rank) & (idiom <= ref_rank + ties)).all())
    s, r = cat.get_scores_rank(u.cuda(), W.cuda(), target.cuda(), cand.cuda(), want_scores=True)
    assert torch.equal(s.cpu(), ref)
    assert torch.equal(r.cpu(), ref_rank)
    _, r2 = cat.get_scores_rank(u.cuda(), W.cuda(), target.cuda(), cand.cuda(), want_scores=False)
    assert torch.equal(r2, r)


def test_candidate_scores_gaussian(cat):
    g = torch.Generator().manual_seed(9)
    B, C, V, d = 64, 199, 5000, 128
    W = torch.randn(V + 2, d, generator=g)
    u = torch.randn(B, d, generator=g)
    target = torch.randint(1, V + 1, (B,), generator=g)
    cand = torch.randint(1, V + 1, (B, C), generator=g)
    s, r = cat.get_scores_rank(u.cuda(), W.cuda(), target.cuda(), cand.cuda(), want_scores=True)
    ref = O.candidate_scores(u.double(), W.double(), target, cand.reshape(-1), C).reshape(B, C + 1)
    torch.testing.assert_close(s.cpu().double(), ref, rtol=1e-5, atol=1e-5)
    assert torch.equal(r.cpu(), O.catalog_rank(s.cpu(), torch.zeros(B, dtype=torch.int64)))


def test_candidate_rank_golden(cat):
    """Fixture produced by the unmodified reference (oracle/make_golden.py)."""
    gd = load_golden("eval_rank_metrics.npz")
    C = gd["cand"].shape[1]
    u, W, target, cand = t(gd["u"]), t(gd["weight"]), t(gd["target"]), t(gd["cand"])
    B = u.shape[0]
    s, r = cat.get_scores_rank(u.cuda(), W.cuda(), target.cuda(), cand.reshape(B, C).cuda(), want_scores=True)
    torch.testing.assert_close(s.cpu(), t(gd["scores"]).reshape(B, C + 1), rtol=1e-5, atol=1e-5)
    ref_rank = t(gd["rank"])
    srt = s.cpu()
    no_tie = (srt[:, 1:] == srt[:, :1]).sum(1) == 0
    assert torch.equal(r.cpu()[no_tie], ref_rank[no_tie])
    # with ties the reference's order is unspecified; ours is the lower end of its bracket
    ties = (srt[:, 1:] == srt[:, :1]).sum(1)
    assert bool(((ref_rank - r.cpu() >= 0) & (ref_rank - r.cpu() <= ties)).all())


def test_candidate_rejects_bad_d(cat):
    from recguru_b200._lib import RecGuruError
    with pytest.raises(RecGuruError):
        cat.get_scores_rank(torch.zeros(2, 48).cuda(), torch.zeros(10, 48).cuda(),
                            torch.ones(2, dtype=torch.int64).cuda(), torch.ones(2, 3, dtype=torch.int64).cuda())
