"""SM clock and board power DURING a sustained loop of one kernel (NVML polled every 5 ms from a
thread): is the catalog sweep clock-limited by the power cap, and at what clock does each variant run?
cycles/tile = time per 128x128 tile per SM x the clock actually held."""
import os
import sys
import threading
import time
from pathlib import Path

import pynvml
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from recguru_b200 import ops  # noqa: E402

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
dev = torch.device("cuda")
d = 128
U = 148 * 4 * 128
g = torch.Generator(device=dev); g.manual_seed(1)
users = ops.raw_cast_bf16(torch.randn(U, d, generator=g, device=dev))


class Poll:
    def __enter__(self):
        self.stop, self.clk, self.pw = False, [], []
        self.t = threading.Thread(target=self.run, daemon=True); self.t.start(); return self

    def run(self):
        while not self.stop:
            self.clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.pw.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
            time.sleep(0.005)

    def __exit__(self, *a):
        self.stop = True; self.t.join()

    def med(self, x):
        x = sorted(x[len(x) // 3:])            # last two thirds: settled
        return x[len(x) // 2] if x else float("nan")


def sustained(label, fn, seconds, tiles_per_sm=None, flops=None):
    fn(); torch.cuda.synchronize()
    n = 0
    with Poll() as p:
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        while time.perf_counter() - t0 < seconds:
            for _ in range(4):
                fn(); n += 1
            torch.cuda.synchronize()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
    clk, pw = p.med(p.clk), p.med(p.pw)
    extra = ""
    if tiles_per_sm:
        extra = f" cycles/tile {ms * 1e-3 / tiles_per_sm * clk * 1e6:7.0f}"
    if flops:
        tf = flops / (ms * 1e-3) / 1e12
        extra += f" {tf:7.1f} TFLOP/s = {flops / (ms * 1e-3) / 148 / (clk * 1e6):6.0f} flop/cycle/SM"
    print(f"{label:44s} {ms:9.3f} ms  SM {clk:5.0f} MHz  {pw:6.0f} W{extra}", flush=True)


def sweep_case(n_items, k, label, seconds=2.0, **env):
    for kk, v in env.items():
        os.environ[kk] = str(v)
    items = ops.raw_cast_bf16(torch.randn(n_items, d, generator=g, device=dev))
    target = torch.randint(1, n_items + 1, (U,), generator=g, device=dev)
    ts = ops.raw_catalog_target_scores(users, items, target, 1)
    bufs = {}
    tiles = (U // 128) * ((n_items + 127) // 128) / 148
    sustained(label, lambda: ops.raw_catalog_topk_rank(users, items, target, ts, k, 1, bufs=bufs), seconds,
              tiles_per_sm=tiles, flops=2.0 * U * n_items * d)
    for kk in env:
        os.environ.pop(kk, None)


a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
b = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
sustained("cuBLAS bf16 8192^3 (torch.matmul)", lambda: torch.matmul(a, b), 3.0, flops=2.0 * 8192 ** 3)
del a, b
sweep_case(1_000_000, 10, "sweep v2 top-10, 1M items")
sweep_case(1_000_000, 0, "sweep v2 rank only, 1M items")
sweep_case(1_000_000, 10, "sweep v1 top-10, 1M items", RECGURU_CATALOG_V=1)
sweep_case(10_000_000, 10, "sweep v2 top-10, 10M items", seconds=3.0)
sweep_case(10_000_000, 0, "sweep v2 rank only, 10M items", seconds=3.0)
time.sleep(1.0)
print("idle clock", pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), "MHz; power limit",
      pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1000.0, "W")
