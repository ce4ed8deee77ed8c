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
