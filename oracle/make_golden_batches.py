"""ORACLE tooling -- generates tests/golden/batch_layouts.npz by running the UNMODIFIED reference
loaders' row builders (data/data_loader.py:25-36 ``seq_padding``, :39-55 ``test_seq_gen``).

Run in the authoring container only (needs /root/reference):

    python oracle/make_golden_batches.py

Seeded ragged histories covering every branch of the list surgery (empty history, shorter than
the window, exactly window-1 / window / window+1 items, much longer; decoder window equal to and
shorter than the encoder window; the appended validation item pushing a history over the
window).  The fixture pins ``oracle.ref_path.seq_padding`` / ``test_seq_gen``
(tests/test_oracle_golden.py) and is the parity target of ``recguru_build_batches`` on the GPU
box (tests/test_gpu_batch_builder.py), where /root/reference does not exist.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
from make_golden import OUT, import_reference  # noqa: E402

CASES = [(6, 6), (12, 12), (12, 5), (9, 1), (2, 2), (50, 50), (200, 200)]


def main():
    DL = import_reference()[5]
    rng = np.random.default_rng(20261017)
    out = {}
    for c, (Le, Ld) in enumerate(CASES):
        lens = [0, 1, 2, max(Le - 2, 0), Le - 1, Le, Le + 1, 3 * Le + 1] + rng.integers(0, 2 * Le + 2, 24).tolist()
        seqs = [rng.integers(1, 1000, n).tolist() for n in lens]
        val = rng.integers(1, 1000, len(seqs))
        test = rng.integers(1, 1000, len(seqs))
        eos = 1000
        tr = [DL.seq_padding(s, Le, Ld, eos) for s in seqs]
        ev = [DL.test_seq_gen(s, Le, Ld, eos, int(v)) for s, v in zip(seqs, val)]
        out[f"c{c}_dims"] = np.array([Le, Ld, eos], dtype=np.int64)
        out[f"c{c}_items"] = np.array([x for s in seqs for x in s], dtype=np.int64)
        out[f"c{c}_offsets"] = np.cumsum([0] + [len(s) for s in seqs]).astype(np.int64)
        out[f"c{c}_val"], out[f"c{c}_test"] = val.astype(np.int64), test.astype(np.int64)
        for k, name in enumerate(("enc_in", "dec_in", "dec_out")):
            out[f"c{c}_{name}"] = np.stack([r[k] for r in tr]).astype(np.int64)
        for k, name in enumerate(("eval_enc_in", "eval_dec_in", "test_enc_in", "t_dec_in")):
            out[f"c{c}_{name}"] = np.stack([r[k] for r in ev]).astype(np.int64)
    out["n_cases"] = np.array(len(CASES))
    np.savez_compressed(OUT / "batch_layouts.npz", **out)
    print("wrote", OUT / "batch_layouts.npz", {k: v.shape for k, v in out.items() if k.startswith("c2_")})


if __name__ == "__main__":
    main()
