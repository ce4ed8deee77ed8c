"""The drop-in against the REAL reference classes, on CPU: MyRec (SASRec and
AutoRec) and MyAuto4Rec_c from /root/reference are instantiated unmodified, their tables / position
modules / losses replaced by the fused modules exactly as INTEGRATION.md section 2 says, and their
own forward() runs under FakeTensorMode on fake CUDA tensors -- every unchanged call site
(AutoEnc4Rec.py:55-133,:175-230; AutoEnc4Rec_cross.py:93-221) must accept what the fused modules
return (shape, dtype, device, differentiability).  register_fake supplies the shapes; no kernel
runs, so no GPU is needed.  Skipped where /root/reference is absent (the GPU box)."""
from pathlib import Path

import pytest
import torch

REF = Path("/root/reference/GURU")
pytestmark = pytest.mark.skipif(not REF.exists(), reason="/root/reference is not present on this box")


@pytest.fixture(scope="module")
def ref():
    from oracle.make_golden import import_reference, Param
    M1, M2, T, LF, MET, DL = import_reference()
    return M1, M2, Param


def to_fake_cuda(model):
    """Parameters and buffers -> fake CUDA tensors (module.to() cannot swap real for fake tensors)."""
    for mod in model.modules():
        for k, v in list(mod._parameters.items()):
            if v is not None:
                mod._parameters[k] = torch.nn.Parameter(torch.empty(v.shape, dtype=v.dtype, device="cuda"),
                                                        requires_grad=v.requires_grad)
        for k, v in list(mod._buffers.items()):
            if v is not None:
                mod._buffers[k] = torch.empty(v.shape, dtype=v.dtype, device="cuda")


class CallCounter:
    def __init__(self, *mods):
        self.n = {}
        for name, m in mods:
            self.n[name] = 0
            m.register_forward_hook(lambda mod, i, o, name=name: self.n.__setitem__(name, self.n[name] + 1))


def fake_batch(B, L, K, n_items):
    ids = lambda *s: torch.randint(1, n_items, s, device="cuda")  # noqa: E731
    return ids(B, L), ids(B, L), ids(B, L), ids(B, L * K)


def check_out(x, shape):
    from torch._subclasses.fake_tensor import FakeTensor
    assert isinstance(x, FakeTensor) and tuple(x.shape) == shape and x.dtype == torch.float32
    assert x.device.type == "cuda" and x.requires_grad


def test_myrec_sas_forward_with_fused_modules(ref):
    from torch._subclasses.fake_tensor import FakeTensorMode
    from recguru_b200.modules import FusedBPRLoss_sas, FusedItemEmbedding, FusedPositionalEncoding
    M1, _, Param = ref
    n_items, L, B, K = 60, 12, 6, 5
    p = Param(vocab_size=n_items + 1, enc_maxlen=L, rec_maxlen=L)
    model = M1.MyRec("cuda", p, sas=True).to(torch.float32)
    keys = set(model.state_dict().keys())
    model.src_emb = FusedItemEmbedding.from_embedding(model.src_emb)
    model.pos_emb = FusedPositionalEncoding.from_module(model.pos_emb)
    assert set(model.state_dict().keys()) == keys                          # checkpoints stay loadable
    assert model.src_emb.padding_idx == p.pad_index
    cnt = CallCounter(("src_emb", model.src_emb), ("pos_emb", model.pos_emb))
    with FakeTensorMode(allow_non_fake_inputs=True):
        to_fake_cuda(model)
        enc_in, dec_in, dec_out, neg = fake_batch(B, L, K, n_items)
        p_logits, n_logits = model(enc_in, dec_in, dec_out, neg)           # MyRec.forward, unchanged
        check_out(p_logits, (B, L, 1))
        check_out(n_logits, (B, L, K))
        mask = (1 - (dec_in == p.pad_index).to(int)).view(-1).to(torch.float32)   # train_auto.py:108-110
        loss = FusedBPRLoss_sas()(p_logits, n_logits, mask=mask)           # train_auto.py:68
        check_out(loss, ())
    assert cnt.n == {"src_emb": 3, "pos_emb": 1}                           # AutoEnc4Rec.py:92,:115,:118 ; :93


def test_myrec_autorec_forward_with_fused_modules(ref):
    from torch._subclasses.fake_tensor import FakeTensorMode
    from recguru_b200.modules import (FusedBPRLoss_sas, FusedItemEmbedding, FusedPositionalEncoding,
                                      FusedSampledCrossEntropyLoss)
    M1, _, Param = ref
    n_items, L, B, Kn, Kb = 80, 10, 5, 30, 5
    p = Param(vocab_size=n_items + 1, enc_maxlen=L, rec_maxlen=L, n_negs=Kn)
    model = M1.MyRec("cuda", p, sas=False).to(torch.float32)
    enc = model.AutoEnc
    enc.src_emb = FusedItemEmbedding.from_embedding(enc.src_emb)
    enc.pos_emb = FusedPositionalEncoding.from_module(enc.pos_emb)
    cnt = CallCounter(("src_emb", enc.src_emb), ("pos_emb", enc.pos_emb))
    with FakeTensorMode(allow_non_fake_inputs=True):
        to_fake_cuda(model)
        enc_in, dec_in, dec_out, neg = fake_batch(B, L, Kn, n_items)
        mask = (1 - (dec_in == p.pad_index).to(int)).view(-1).to(torch.float32)
        logits = model(enc_in, dec_in, dec_out, neg, recon=True)           # MyAuto4Rec.forward
        check_out(logits, (B, L, 1 + Kn))
        label = torch.zeros(B * L, device="cuda").long()
        check_out(FusedSampledCrossEntropyLoss()(logits, label, Kn + 1, mask=mask), ())   # train_auto.py:51
        assert cnt.n == {"src_emb": 4, "pos_emb": 2}                       # AutoEnc4Rec.py:180,:192,:219,:221
        neg5 = torch.randint(1, n_items, (B, L * Kb), device="cuda")
        p_logits, n_logits = model(enc_in, dec_in, dec_out, neg5)          # BPR fine-tune branch :112-133
        check_out(p_logits, (B, L, 1))
        check_out(n_logits, (B, L, Kb))
        check_out(FusedBPRLoss_sas()(p_logits, n_logits, mask=mask), ())


def test_recguru_cross_forward_with_fused_modules(ref):
    from torch._subclasses.fake_tensor import FakeTensorMode
    from recguru_b200.modules import (FusedBPRLoss, FusedItemEmbedding, FusedPositionalEncoding,
                                      FusedSampledCrossEntropyLoss)
    _, M2, Param = ref
    na, nb, L, B, K = 50, 70, 8, 4, 5
    p = Param(vocab_size_a=na + 1, vocab_size_b=nb + 1, enc_maxlen=L, rec_maxlen=L, n_negs=K, n_bpr_neg=K)
    model = M2.MyAuto4Rec_c("cuda", p).to(torch.float32)
    for dom in ("a", "b"):
        emb = FusedItemEmbedding.from_embedding(getattr(model, f"src_emb_{dom}"))
        assert emb.padding_idx is None                                     # AutoEnc4Rec_cross.py:32,:35
        setattr(model, f"src_emb_{dom}", emb)
        setattr(model, f"pos_emb_{dom}", FusedPositionalEncoding.from_module(getattr(model, f"pos_emb_{dom}")))
    with FakeTensorMode(allow_non_fake_inputs=True):
        to_fake_cuda(model)
        for dom, n_items in (("a", na), ("b", nb)):
            enc_in, dec_in, dec_out, neg = fake_batch(B, L, K, n_items)
            mask = (1 - (dec_out == p.pad_index).to(int)).view(-1).to(torch.float32)     # gan_training.py:347-350
            logits = model(enc_in, dec_in, dec_out, neg, dom, mask)        # MyAuto4Rec_c.forward :185-221
            check_out(logits, (B, L, 1 + K))
            label = torch.zeros(B * L, device="cuda").long()
            check_out(FusedSampledCrossEntropyLoss()(logits, label, K + 1, mask=mask), ())   # tools/utils.py:84
            rec = model.recommend_forward(enc_in, dec_in, dom, mask.view(-1, L))         # :149-183
            check_out(rec, (B, L, p.d_model))
            emb = getattr(model, f"src_emb_{dom}")
            rec4 = rec.view(-1, L, 1, p.d_model)                           # tools/utils.py:113-124
            n_emb = emb(neg).view(-1, L, K, p.d_model)
            p_emb = emb(dec_out).view(-1, L, 1, p.d_model)
            p_logits = torch.squeeze(torch.matmul(rec4, p_emb.transpose(2, 3)), 2)
            n_logits = torch.squeeze(torch.matmul(rec4, n_emb.transpose(2, 3)), 2)
            check_out(FusedBPRLoss()(p_logits, n_logits, mask=mask), ())
            u = model.get_seq_embed(enc_in, dom, mask.view(-1, L))[:, -1, :]             # gan_training.py:152-162
            check_out(u, (B, p.d_model))
