"""ORACLE tooling -- generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference):

    python oracle/make_golden.py

It puts /root/reference/GURU on sys.path, stubs ``matplotlib`` (absent here; only imported by
the reference's plotting helpers), builds tiny seeded instances of the reference's own models
(MyRec sas / non-sas, MyAuto4Rec_c), runs them in eval mode (dropout off) on CPU and records
the inputs and outputs at the hot path's two edges with forward hooks.  The fixtures are the
oracle's pin (tests/test_oracle_golden.py) and the GPU kernels' parity targets on the GPU
box, where /root/reference does not exist.
"""
from __future__ import annotations

import os
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference/GURU")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def import_reference():
    if not REF.exists():
        raise SystemExit("reference not mounted; golden fixtures can only be generated in the authoring container")
    sys.path.insert(0, str(REF))
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.use = lambda *a, **k: None
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    import AutoEnc4Rec as M1
    import AutoEnc4Rec_cross as M2
    import Transformer.transformer as T
    import tools.lossfunctions as LF
    import tools.metrics as MET
    import data.data_loader as DL
    return M1, M2, T, LF, MET, DL


class Param:
    """The attributes of config_auto4rec.get_param the models read (config_auto4rec.py:21-237)."""

    def __init__(self, **kw):
        self.pad_index = 0
        self.num_train_neg = 5
        self.n_bpr_neg = 5
        self.enc_maxlen = 12
        self.rec_maxlen = 12
        self.d_ff = 64
        self.d_model = 32
        self.num_heads = 2
        self.d_k = 16
        self.d_v = 16
        self.num_blocks = 1
        self.dropout_rate = 0.5
        self.decoder_neg = True
        self.n_negs = 30
        self.fixed_enc = False
        self.candidate_size = 19
        self.__dict__.update(kw)


def make_batch(g, B, L, n_items, K, eos):
    """Padded sequences in the reference's layout through ITS seq_padding."""
    return g, B, L, n_items, K, eos


def tnp(t):
    return t.detach().cpu().numpy()


def capture(module, store, key):
    def hook(mod, inp, out):
        o = out[0] if isinstance(out, tuple) else out
        if o.requires_grad:
            o.retain_grad()
        store.setdefault(key, []).append((tuple(i for i in inp), o))
    return module.register_forward_hook(hook)


def golden_myrec_sas(M1, LF, DL, path):
    torch.manual_seed(20260101)
    np.random.seed(7)
    n_items, L, B, K = 60, 12, 6, 5
    p = Param(vocab_size=n_items + 1, enc_maxlen=L, rec_maxlen=L, num_train_neg=K)
    model = M1.MyRec("cpu", p, sas=True).to(torch.float32).eval()
    eos = n_items + 1
    rng = np.random.RandomState(3)
    encs, dins, douts = [], [], []
    for b in range(B):
        ln = int(rng.randint(3, L + 3))
        seq = rng.randint(1, n_items + 1, size=ln)
        e, di, do = DL.seq_padding(seq, L, L, eos)
        encs.append(e); dins.append(di); douts.append(do)
    enc_in = torch.from_numpy(np.stack(encs)).long()
    dec_in = torch.from_numpy(np.stack(dins)).long()
    dec_out = torch.from_numpy(np.stack(douts)).long()
    neg = torch.from_numpy(rng.randint(1, n_items + 1, size=(B, L * K))).long()
    mask = (1 - (dec_in == p.pad_index).to(int)).view(-1).to(torch.float32)  # train_auto.py:109-110

    store = {}
    h1 = capture(model.src_emb, store, "emb")
    h2 = capture(model.pos_emb, store, "pos")
    h3 = capture(model.sas_rec, store, "enc")
    p_logits, n_logits = model(enc_in, dec_in, dec_out, neg)
    loss = LF.BPRLoss_sas()(p_logits, n_logits, mask=mask)
    p_logits.retain_grad(); n_logits.retain_grad()
    loss.backward()
    for h in (h1, h2, h3):
        h.remove()
    emb_calls = store["emb"]          # order: dec_in (get_embedding_sas), n_items, dec_out
    assert len(emb_calls) == 3
    x0 = store["pos"][0][1]
    hid = store["enc"][0][1]
    np.savez_compressed(
        path, weight=tnp(model.src_emb.weight), pe=tnp(model.pos_emb.pe[:, :64]), enc_in=tnp(enc_in),
        dec_in=tnp(dec_in), dec_out=tnp(dec_out), n_items=tnp(neg), mask=tnp(mask),
        x0=tnp(x0), h=tnp(hid), p_logits=tnp(p_logits), n_logits=tnp(n_logits), loss=tnp(loss),
        dh=tnp(hid.grad), g_x0=tnp(x0.grad), g_in=tnp(emb_calls[0][1].grad),
        dp=tnp(p_logits.grad), dn=tnp(n_logits.grad), dW=tnp(model.src_emb.weight.grad),
        L=L, K=K, padding_idx=0)


def golden_myrec_recon(M1, LF, DL, path):
    """AutoRec reconstruction step: MyAuto4Rec.forward + SampledCrossEntropyLoss (K = n_negs)."""
    torch.manual_seed(20260102)
    n_items, L, B, K = 80, 10, 5, 30
    p = Param(vocab_size=n_items + 1, enc_maxlen=L, rec_maxlen=L, n_negs=K)
    model = M1.MyRec("cpu", p, sas=False).to(torch.float32).eval()
    eos = n_items + 1
    rng = np.random.RandomState(5)
    encs, dins, douts = [], [], []
    for b in range(B):
        seq = rng.randint(1, n_items + 1, size=int(rng.randint(3, L + 2)))
        e, di, do = DL.seq_padding(seq, L, L, eos)
        encs.append(e); dins.append(di); douts.append(do)
    enc_in = torch.from_numpy(np.stack(encs)).long()
    dec_in = torch.from_numpy(np.stack(dins)).long()
    dec_out = torch.from_numpy(np.stack(douts)).long()
    neg = torch.from_numpy(rng.randint(1, n_items + 1, size=(B, L * K))).long()
    mask = (1 - (dec_in == p.pad_index).to(int)).view(-1).to(torch.float32)

    store = {}
    hooks = [capture(model.AutoEnc.src_emb, store, "emb"), capture(model.AutoEnc.pos_emb, store, "pos"),
             capture(model.AutoEnc.decoder, store, "dec")]
    logits = model(enc_in, dec_in, dec_out, neg, recon=True)       # [B, L, 1+K]
    label = torch.zeros(B * L).long()
    loss = LF.SampledCrossEntropyLoss(reduction="none")(logits, label, K + 1, mask=mask)
    logits.retain_grad()
    loss.backward()
    for h in hooks:
        h.remove()
    emb = store["emb"]            # enc_in, dec_in, n_items, dec_out
    assert len(emb) == 4
    hid = store["dec"][0][1]
    np.savez_compressed(
        path, weight=tnp(model.AutoEnc.src_emb.weight), pe=tnp(model.AutoEnc.pos_emb.pe[:, :64]),
        enc_in=tnp(enc_in), dec_in=tnp(dec_in), dec_out=tnp(dec_out), n_items=tnp(neg), mask=tnp(mask),
        x_enc=tnp(store["pos"][0][1]), x_dec=tnp(store["pos"][1][1]), h=tnp(hid), logits=tnp(logits),
        loss=tnp(loss), dh=tnp(hid.grad), g_enc=tnp(emb[0][1].grad), g_dec=tnp(emb[1][1].grad),
        dlogits=tnp(logits.grad), dW=tnp(model.AutoEnc.src_emb.weight.grad), L=L, K=K, padding_idx=0)


def golden_cross(M2, LF, DL, path):
    """RecGURU generator, domain a: MyAuto4Rec_c.forward (recon, sce) and the BPR path of
    tools/utils.loss_bpr_func restated with the model's own methods; tables have NO padding_idx."""
    torch.manual_seed(20260103)
    na, nb, L, B, K = 50, 70, 8, 4, 5
    p = Param(vocab_size_a=na + 1, vocab_size_b=nb + 1, enc_maxlen=L, rec_maxlen=L, n_negs=K, n_bpr_neg=K)
    model = M2.MyAuto4Rec_c("cpu", p).to(torch.float32).eval()
    rng = np.random.RandomState(11)
    encs, dins, douts = [], [], []
    for b in range(B):
        seq = rng.randint(1, na + 1, size=int(rng.randint(3, L + 2)))
        e, di, do = DL.seq_padding(seq, L, L, na + 1)
        encs.append(e); dins.append(di); douts.append(do)
    enc_in = torch.from_numpy(np.stack(encs)).long()
    dec_in = torch.from_numpy(np.stack(dins)).long()
    dec_out = torch.from_numpy(np.stack(douts)).long()
    neg = torch.from_numpy(rng.randint(1, na + 1, size=(B, L * K))).long()
    mask = (1 - (enc_in == p.pad_index).to(int)).view(-1).to(torch.float32)   # gan_training.py:347-350

    store = {}
    hooks = [capture(model.src_emb_a, store, "emb"), capture(model.pos_emb_a, store, "pos"),
             capture(model.recommend_a, store, "rec")]
    # BPR path (tools/utils.py:90-127, single-device branch)
    rec = model.recommend_forward(enc_in, dec_in, "a", mask.view(-1, L))
    rec4 = rec.view(-1, L, 1, p.d_model)
    n_emb = model.src_emb_a(neg).view(-1, L, K, p.d_model)
    p_emb = model.src_emb_a(dec_out).view(-1, L, 1, p.d_model)
    p_logits = torch.squeeze(torch.matmul(rec4, p_emb.transpose(2, 3)), 2)
    n_logits = torch.squeeze(torch.matmul(rec4, n_emb.transpose(2, 3)), 2)
    d_mask = (1 - (dec_in == p.pad_index).to(int)).view(-1).to(torch.float32)
    loss = LF.BPRLoss()(p_logits, n_logits, mask=d_mask)
    loss.backward()
    for h in hooks:
        h.remove()
    emb = store["emb"]            # enc_in, dec_in, n_items, dec_out
    assert len(emb) == 4
    hid = store["rec"][0][1]
    np.savez_compressed(
        path, weight=tnp(model.src_emb_a.weight), pe=tnp(model.pos_emb_a.pe[:, :64]), enc_in=tnp(enc_in),
        dec_in=tnp(dec_in), dec_out=tnp(dec_out), n_items=tnp(neg), enc_mask=tnp(mask), dec_mask=tnp(d_mask),
        x_enc=tnp(store["pos"][0][1]), x_dec=tnp(store["pos"][1][1]), h=tnp(hid),
        p_logits=tnp(p_logits), n_logits=tnp(n_logits), loss=tnp(loss), dh=tnp(hid.grad),
        g_enc=tnp(emb[0][1].grad), g_dec=tnp(emb[1][1].grad), dW=tnp(model.src_emb_a.weight.grad), L=L, K=K)


def golden_eval(M1, MET, DL, path):
    """get_scores tail + the double-argsort rank idiom + metrics (train_auto.py:164-253)."""
    torch.manual_seed(20260104)
    n_items, d, B, C = 200, 32, 16, 19
    W = torch.randn(n_items + 2, d)
    W[0] = 0
    u = torch.randn(B, d)
    target = torch.randint(1, n_items + 1, (B,))
    cand = torch.randint(1, n_items + 1, (B, C))
    emb = torch.nn.Embedding(n_items + 2, d, padding_idx=0)
    emb.weight.data.copy_(W)
    # train_auto.py:189-199 with model.src_emb == emb and dec_embedding[:, -1, :] == u
    n_embeddings = emb(cand).view(-1, C, d)
    p_embeddings = emb(target).view(-1, d)
    candidate_embeddings = torch.cat([torch.unsqueeze(p_embeddings, 1), n_embeddings], dim=1)
    scores = torch.squeeze(torch.matmul(torch.unsqueeze(u, 1), candidate_embeddings.transpose(1, 2)))
    rank = torch.argsort(torch.argsort(-scores, dim=1), dim=1)[:, 0]
    # tie KAT (SURVEY.md section 8c ii)
    tie_scores = torch.tensor([[1, 2, 1, .5], [3, 3, 3, 3], [0, 1, 1, -1]], dtype=torch.float32)
    tie_rank = torch.argsort(torch.argsort(-tie_scores, dim=1), dim=1)[:, 0]
    ks = [1, 5, 10, 20, 30]
    ranks_kat = [0, 3, 12]
    r_list = list(tnp(rank))
    met = {f"{name}_{k}": fn(r_list, k) for k in ks
           for name, fn in (("hr", MET.hit_at_k_batch), ("ndcg", MET.NDCG_at_k_batch), ("mrr", MET.mrr_at_k_batch))}
    kat = {f"kat_{name}_10": fn(ranks_kat, 10)
           for name, fn in (("hr", MET.hit_at_k_batch), ("ndcg", MET.NDCG_at_k_batch), ("mrr", MET.mrr_at_k_batch))}
    e, di, do = DL.seq_padding([1, 2, 3], 6, 6, 9)
    ee, edi, te, tdi = DL.test_seq_gen([1, 2, 3], 6, 6, 9, 4)
    np.savez_compressed(path, weight=tnp(W), u=tnp(u), target=tnp(target), cand=tnp(cand), scores=tnp(scores),
                        rank=tnp(rank), tie_scores=tnp(tie_scores), tie_rank=tnp(tie_rank),
                        ks=np.array(ks), pad_enc=e, pad_dec_in=di, pad_dec_out=do, tsg_eval_enc=ee,
                        tsg_dec_in=edi, tsg_test_enc=te, tsg_t_dec_in=tdi,
                        **{k: np.float64(v) for k, v in {**met, **kat}.items()})


def golden_pe(T, path):
    pe32 = T.PositionalEncoding(d_model=32, dropout=0.5, max_len=64).pe
    pe128 = T.PositionalEncoding(d_model=128, dropout=0.5, max_len=256).pe
    np.savez_compressed(path, pe32=tnp(pe32), pe128=tnp(pe128))


def main():
    M1, M2, T, LF, MET, DL = import_reference()
    OUT.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(1)  # fixtures independent of the host's thread count
    golden_myrec_sas(M1, LF, DL, OUT / "myrec_sas.npz")
    golden_myrec_recon(M1, LF, DL, OUT / "myrec_recon.npz")
    golden_cross(M2, LF, DL, OUT / "cross_a.npz")
    golden_eval(M1, MET, DL, OUT / "eval_rank_metrics.npz")
    golden_pe(T, OUT / "pe_tables.npz")
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, os.path.getsize(f), "bytes")


if __name__ == "__main__":
    main()
