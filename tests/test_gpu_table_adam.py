"""GPU parity of the fused item-table Adam step against torch.optim.Adam itself -- the optimizer the
reference constructs (train_auto.py:354-364) -- run on the CPU in fp32 over several steps, with
rows that never receive a gradient (their moments must still decay) and a changing learning rate
(ScheduledOptim, Transformer/transformer.py:38-51).  Tolerance: the update is the same fp32
formula with a different fused-multiply-add contraction, rtol 2e-6 per step."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("V1,d", [(101, 32), (1000, 128), (257, 64)])
def test_table_adam_matches_torch_adam(built_lib, V1, d):
    from recguru_b200.optim import FusedTableAdam
    g = torch.Generator().manual_seed(V1 + d)
    W0 = torch.randn(V1, d, generator=g) * 0.1
    ref_p = torch.nn.Parameter(W0.clone())
    ref = torch.optim.Adam([ref_p], lr=1e-3, betas=(0.9, 0.99), eps=1e-9)
    w = W0.clone().cuda()
    opt = FusedTableAdam(w, lr=1e-3)
    for step in range(1, 8):
        grad = torch.randn(V1, d, generator=g) * 0.05
        grad[torch.rand(V1, generator=g) < 0.7] = 0.0          # most rows untouched, like a real step
        lr = 1e-3 * min(step ** -0.5, step * 4 ** -1.5)        # scheduled lr
        ref.param_groups[0]["lr"] = lr
        opt.param_groups[0]["lr"] = lr
        ref_p.grad = grad.clone()
        ref.step()
        opt.step(grad.cuda())
        assert torch.allclose(w.cpu(), ref_p.detach(), rtol=2e-5, atol=1e-7), step
        st = ref.state[ref_p]
        assert torch.allclose(opt.state["exp_avg"].cpu(), st["exp_avg"], rtol=1e-5, atol=1e-9)
        assert torch.allclose(opt.state["exp_avg_sq"].cpu(), st["exp_avg_sq"], rtol=1e-5, atol=1e-12)
    assert opt.state["step"] == 7
    sd = opt.state_dict()
    opt2 = FusedTableAdam(w.clone(), lr=1.0)
    opt2.load_state_dict(sd)
    assert opt2.state["step"] == 7 and opt2.param_groups[0]["lr"] == opt.param_groups[0]["lr"]


def test_table_adam_rejects_cpu(built_lib):
    from recguru_b200.optim import FusedTableAdam
    with pytest.raises(RuntimeError):
        FusedTableAdam(torch.zeros(4, 4))


@pytest.mark.parametrize("V1,d", [(101, 32), (1000, 128), (257, 64), (64, 256)])
def test_table_adam_rows_bit_identical_to_dense(built_lib, V1, d):
    """Row-store form: rows stamped with the current epoch carry a gradient, all others
    take g = 0 unread -- even when the row store holds garbage (NaN) there.  Must equal the dense
    kernel on the zero-filled gradient bit for bit, moments included, over several steps."""
    from recguru_b200 import ops
    g = torch.Generator().manual_seed(3 * V1 + d)
    w_a = (torch.randn(V1, d, generator=g) * 0.1).cuda()
    w_b = w_a.clone()
    m_a, v_a, m_b, v_b = (torch.zeros_like(w_a) for _ in range(4))
    epoch_of_row = torch.zeros(V1, dtype=torch.int32, device="cuda")
    for step in range(1, 6):
        live = torch.rand(V1, generator=g) < 0.3
        dense = torch.randn(V1, d, generator=g) * 0.05
        dense[~live] = 0.0
        rows = dense.clone()
        rows[~live] = float("nan")                       # never read
        epoch_of_row[live.cuda()] = step                 # older stamps stay behind: they must not count
        ops.raw_table_adam(w_a, m_a, v_a, dense.cuda(), 1e-3, 0.9, 0.99, 1e-9, step)
        ops.raw_table_adam_rows(w_b, m_b, v_b, rows.cuda(), epoch_of_row, step, 1e-3, 0.9, 0.99, 1e-9, step)
        assert torch.equal(w_a, w_b) and torch.equal(m_a, m_b) and torch.equal(v_a, v_b), step
    assert bool(torch.isfinite(w_b).all())


def test_hotpath_row_store_adam_equals_dense(built_lib):
    """TrainHotPath(dense_grad=False): no zero fill, rows stamped by recguru_embedding_bwd_mark_rows;
    FusedTableAdam.step_rows on it == FusedTableAdam.step on the dense gradient, bit for bit, over
    steps with different batches (stale rows of earlier steps must be ignored)."""
    from recguru_b200.hotpath import TrainHotPath
    from recguru_b200.optim import FusedTableAdam
    from test_gpu_hotpath import make_step
    B, L, d, K, V = 8, 20, 64, 5, 300
    W, pe, *_ = make_step(B, L, d, K, V, seed=0)
    w_dense, w_rows = W.clone().cuda(), W.clone().cuda()
    hp_dense = TrainHotPath(w_dense, pe.cuda(), B, L, K, "bpr_sas")
    hp_rows = TrainHotPath(w_rows, pe.cuda(), B, L, K, "bpr_sas", dense_grad=False)
    hp_rows.dW.fill_(float("nan"))                        # the row store starts as garbage
    opt_dense, opt_rows = FusedTableAdam(w_dense, lr=1e-2), FusedTableAdam(w_rows, lr=1e-2)
    for step in range(4):
        _, _, dec_in, dec_out, n_items, h, g_x0 = (t.cuda() for t in make_step(B, L, d, K, V, seed=step + 1))
        la, _, dWa = hp_dense.step(dec_in, dec_out, n_items, h, g_x0)
        lb, _, dWb = hp_rows.step(dec_in, dec_out, n_items, h, g_x0)
        assert torch.equal(la, lb)
        live = hp_rows.row_epoch == hp_rows.epoch
        # the stamped rows are exactly the rows with a non-structural gradient, and they agree
        assert torch.equal(dWa[live], dWb[live])
        assert not bool(dWa[~live].any())
        assert int(live.sum()) > 0 and not bool(live[0])
        opt_dense.step(dWa)
        opt_rows.step_rows(dWb, hp_rows.row_epoch, hp_rows.epoch)
        assert torch.equal(w_dense, w_rows), step
        assert torch.equal(opt_dense.state["exp_avg_sq"], opt_rows.state["exp_avg_sq"])
    with pytest.raises(RuntimeError):
        hp_rows.capture_step(dec_in, dec_out, n_items, h, g_x0)
    # serial (non-overlapped) schedule takes the other branch of backward()
    hp_serial = TrainHotPath(w_rows.clone(), pe.cuda(), B, L, K, "bpr_sas", dense_grad=False, overlap_plan=False)
    _, _, dWc = hp_serial.step(dec_in, dec_out, n_items, h, g_x0)
    lv = hp_serial.row_epoch == hp_serial.epoch
    hp_check = TrainHotPath(w_rows.clone(), pe.cuda(), B, L, K, "bpr_sas")
    _, _, dWd = hp_check.step(dec_in, dec_out, n_items, h, g_x0)
    assert torch.equal(dWc[lv], dWd[lv]) and not bool(dWd[~lv].any())
