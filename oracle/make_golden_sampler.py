"""ORACLE tooling -- tests/golden/sampler_weights.npz: the weight vectors the UNMODIFIED reference loader
(/root/reference/GURU/data/data_loader.py, pickle_loader.__getitem__, :276-316) hands to
torch.multinomial, captured by intercepting that call, for a handful of users with and without item
frequencies.  Pins oracle.ref_path.sampling_weights (and through it the device sampler's law,
tests/test_gpu_sampler.py).  Run in the authoring container: python oracle/make_golden_sampler.py"""
import pickle
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference/GURU")


def main():
    sys.path.insert(0, str(REF))
    sys.path.insert(0, str(REF / "data"))
    if "matplotlib" not in sys.modules:          # the loaders do not need it; sibling imports might
        sys.modules["matplotlib"] = types.ModuleType("matplotlib")
        sys.modules["matplotlib.pyplot"] = types.ModuleType("matplotlib.pyplot")
    import data_loader as DL
    rng = np.random.RandomState(17)
    v_size, n_users, L, num_n = 300, 6, 20, 5
    seqs = [rng.randint(1, v_size, size=int(rng.randint(3, 15))).tolist() for _ in range(n_users)]
    val = rng.randint(1, v_size, size=n_users).tolist()
    test = rng.randint(1, v_size, size=n_users).tolist()
    wf = rng.randint(0, 50, size=v_size).astype(np.float64)
    wf[0] = 0.0

    class Param:
        enc_maxlen = L
        rec_maxlen = L
        vocab_size_a = v_size
        vocab_size_b = v_size
        candidate_size = 19
    captured = []
    real = torch.multinomial

    def spy(weights, num, replacement=True):
        captured.append(np.asarray(weights.detach().cpu(), dtype=np.float64).copy())
        return real(weights, num, replacement=replacement)

    out = {"seq_flat": np.concatenate(seqs), "seq_len": np.array([len(s) for s in seqs]), "val": np.array(val),
           "test": np.array(test), "wf": wf, "v_size": v_size}
    with tempfile.TemporaryDirectory() as td:
        fn = td + "/d.pkl"
        pickle.dump({"seq": seqs, "val": val, "test": test, "len": [len(s) for s in seqs]}, open(fn, "wb"))
        for name, w in (("uniform", None), ("freq", wf)):
            ds = DL.pickle_loader([fn], Param, num_n, wf=None if w is None else w.copy(), domain="a")
            captured.clear()
            torch.multinomial = spy
            try:
                for i in range(n_users):
                    ds[i]
            finally:
                torch.multinomial = real
            out[f"weights_{name}"] = np.stack(captured)
    np.savez_compressed(ROOT / "tests" / "golden" / "sampler_weights.npz", **out)
    print("wrote sampler_weights.npz", {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
