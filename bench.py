#!/usr/bin/env python
"""bench.py -- RecGURU embedding-and-scoring hot path on B200 (contract: see DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload train|catalog] [--impl reference]

BASELINE.json's metric has three parts; one JSON line carries all of them:
  * primary (``value``): train seqs/sec of the hot-path training step on configs[1]
    (SASRec shape: 100k items, L=200, d=128, B=1024, K=5 BPR negatives) -- gather+PE+mask,
    sampled scoring + BPRLoss_sas forward, its backward, and the one-pass deterministic table
    backward.  N>1: replicas only (the reference replicates the table; nothing is partitioned).
  * ``catalog``: full-catalog eval users/sec -- cfg4 (1M items, N=1 only) and cfg5 (10M items,
    row-sharded over the N GPUs: all-gather users, all-reduce target score, all-to-all + merge).
  * ``gather``: GB/s of the fused gather kernel alone (cold: rotating inputs larger than L2).
``--workload catalog`` makes the cfg5 catalog line the primary one instead.

``--impl reference`` times the reference's own CPU op sequence for the same path (the oracle
port: identical ATen ops; the reference is pure PyTorch and /root/reference does not travel to
the GPU box) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

CFG1 = dict(n_items=10_000, L=50, d=64, B=256)                     # configs[0]: AutoRec (K=30 sce, K=5 bpr_sas)
CFG2 = dict(n_items=100_000, L=200, d=128, B=1024, K=5, kind="bpr_sas")
CFG3 = dict(n_items=500_000, L=100, d=128, B=1024, K=30, kind="sce")  # configs[2]: RecGURU, per domain
CFG4 = dict(n_items=1_000_000, d=128, k=10, users_total=1_000_000)
CFG5 = dict(n_items=10_000_000, d=128, k=10, users_total=1_000_000)
USERS_PER_STEP = 148 * 4 * 128          # 4 full waves of 128-user tiles on 148 SMs


# ----------------------------------------------------------------------------------------------
# plumbing
# ----------------------------------------------------------------------------------------------
TRAFFIC_FILES = ("r02_traffic.json", "r01_traffic.json")


def ncu_traffic(*kernel_prefixes):
    """DRAM bytes per launch (read + write) of the named kernels, summed, from the committed
    `ncu --set full` capture (profiles/r02_traffic.json, else r01; written by scripts/ncu_summary.py)."""
    tab = None
    for fn in TRAFFIC_FILES:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", fn)
        try:
            tab = json.load(open(path))
            break
        except (OSError, ValueError):
            continue
    if tab is None:
        return None
    total, found = 0.0, 0
    for pre in kernel_prefixes:
        for name, rec in tab.items():
            if name.startswith(pre):
                total += rec["dram_bytes_per_launch"]
                found += 1
                break
    return int(total) if found == len(kernel_prefixes) else None


_REAL_STDOUT = None


def emit(line: dict) -> None:
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return dict(hbm=j["hbm_gbs"], tf_burst=j["bf16_tflops"], tf_sust=j["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


class Dist:
    def __init__(self, n_gpus: int):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            # NCCL prints its version banner on stdout at INFO/VERSION level; the bench line must be
            # the only thing there
            # (any level >= VERSION, WARN included, prints it): keep whatever level the launcher
            # asked for but send NCCL's log to stderr
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
            torch.cuda.set_device(self.local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        else:
            torch.cuda.set_device(0)
            self.dist = None
        if self.world != n_gpus and self.rank == 0:
            print(f"# note: --gpus {n_gpus} but WORLD_SIZE={self.world}; using WORLD_SIZE", file=sys.stderr)

    def barrier(self):
        torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def group(self):
        return None if self.dist is None else self.dist.group.WORLD


def timed_loop(D: Dist, fn, steps: int, warmup: int):
    """W untimed steps, then exactly K steps between barrier+synchronize, CUDA events on the
    launching stream, max over ranks.  fn(i) enqueues step i."""
    for i in range(warmup):
        fn(i)
    D.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    D.barrier()
    return D.max_over_ranks(ms)


# ----------------------------------------------------------------------------------------------
# workload 1: training-step hot path (configs[1])
# ----------------------------------------------------------------------------------------------
def make_train_inputs(cfg, n_batches: int, device):
    from recguru_b200 import synth, modules
    W = synth.item_table(cfg["n_items"], cfg["d"]).to(device)
    pe = modules.sinusoid_table(max(cfg["L"], 256), cfg["d"])[0].to(device)
    batches, host = [], []
    for b in range(n_batches):
        tb = synth.train_batch(cfg["B"], cfg["L"], cfg["n_items"], cfg["K"], seed=synth.SEED_DATA + 100 * b)
        h = synth.hidden_states(cfg["B"], cfg["L"], cfg["d"], seed=50 + b)
        g = synth.hidden_states(cfg["B"], cfg["L"], cfg["d"], seed=90 + b)
        host.append(tb)
        batches.append(dict(tb=tb.to(device), h=h.to(device), g=g.to(device)))
    return W, pe, batches, host


def bench_train(D: Dist, args, peaks):
    from recguru_b200 import _lib
    from recguru_b200.hotpath import TrainHotPath
    cfg = CFG2
    dev = torch.device("cuda", torch.cuda.current_device())
    n_batches = 4      # rotating inputs: 4 x (h + g + out) = 4 x 315 MB >> 126 MB L2
    W, pe, batches, host = make_train_inputs(cfg, n_batches, dev)
    hp = TrainHotPath(W, pe, cfg["B"], cfg["L"], cfg["K"], cfg["kind"])
    K, Wm = args.steps, args.warmup

    # ---- per-stage durations: the step's kernels launched one after another on ONE stream,
    # CUDA events between the stages (no overlap here: these are kernel times, not the schedule)
    from recguru_b200 import ops
    n_ev = 6
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(n_ev)] for _ in range(K)]
    n = cfg["B"] * cfg["L"]

    def stage_fns(b):
        tb = b["tb"]
        m = hp.mask.view(n)
        h2 = b["h"].view(n, cfg["d"])

        def f_gather():
            ops.raw_gather_pe_mask(hp.weight, tb.dec_in, hp.pe, hp.pad_index, out=hp.x0, mask_out=hp.mask)

        sm = hp._side_mask.view(n)
        plan_sites = [ops.Site(tb.dec_in, None, live=sm), ops.Site(tb.dec_out, None, live=sm),
                      ops.Site(tb.n_items, None, group=cfg["K"], live=sm)]
        rows = ops.raw_embedding_bwd_plan_live_rows(plan_sites, hp.V1, hp.ws_bwd) if hp.score_rows else None

        def f_fwd():
            if hp.fuse_dh:       # forward + dh in one kernel; sum(mask) and the live-row list come from the plan stage
                ops.raw_sampled_score_loss(b["h"], hp.weight, tb.dec_out, tb.n_items, hp.mask, hp.kind, bufs=hp.sbuf,
                                           skip_masked=True, dh_out=hp.dh, live_count=hp._cnt, live_rows=rows)
            else:
                ops.raw_sampled_score_loss(b["h"], hp.weight, tb.dec_out, tb.n_items, hp.mask, hp.kind, bufs=hp.sbuf,
                                           skip_masked=hp.skip_masked)

        def f_dh():
            ops.raw_sampled_score_bwd_dh(hp.sbuf["dp"], hp.sbuf["dn"], hp.weight, tb.dec_out, tb.n_items,
                                         scale=hp.one, scale2=hp.sbuf["aux"][1:2], out=hp.dh)

        def f_plan():
            if hp.fuse_dh:
                ops.raw_pad_mask_count(tb.dec_in, hp.pad_index, hp._side_mask, hp._cnt)
            else:
                ops.raw_pad_mask(tb.dec_in, hp.pad_index, out=hp._side_mask)
            ops.raw_fill_zero(hp.dW)
            ops.raw_embedding_bwd_plan(plan_sites, hp.V1, 0, hp.ws_bwd, stage=ops.PLAN_COMPACT)
            ops.raw_embedding_bwd_plan(plan_sites, hp.V1, 0, hp.ws_bwd, stage=ops.PLAN_SORT)

        def f_apply():
            sites = [ops.Site(tb.dec_in, b["g"].view(n, cfg["d"]), coef=m, live=m),
                     ops.Site(tb.dec_out, h2, coef=hp.sbuf["dp"], scale=hp.one, scale2=hp.sbuf["aux"][1:2], live=m),
                     ops.Site(tb.n_items, h2, coef=hp.sbuf["dn"].view(-1), scale=hp.one,
                              scale2=hp.sbuf["aux"][1:2], group=cfg["K"], live=m)]
            ops.raw_embedding_bwd_apply(sites, hp.V1, hp.d, hp.dW, hp.ws_bwd, mode=ops.DW_PREZEROED)
        # plan first: it needs the ids only (in the timed step it is forked onto the side stream at
        # the start) and its pad-mask kernel provides sum(mask) for the fused forward
        return [f_plan, f_gather, f_fwd, f_dh, f_apply]

    fns = [stage_fns(b) for b in batches]
    # one eager pass counts the launches of a step; the timing itself replays one small CUDA graph
    # per stage, so that a slow host cannot stretch the gaps between the events
    for f in fns[0]:
        f()
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    hp.step(batches[0]["tb"].dec_in, batches[0]["tb"].dec_out, batches[0]["tb"].n_items, batches[0]["h"],
            batches[0]["g"])
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count() - l0                  # kernels of one TrainHotPath.step
    stage_graphs = None
    if not args.no_graph:
        stage_graphs = []
        for fb in fns:
            gs = []
            for f in fb:
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_):
                    f()
                gs.append(g_)
            stage_graphs.append(gs)

    def step(i):
        bi = i % n_batches
        rec = evs[i - Wm] if i >= Wm else None
        if rec: rec[0].record()
        for si in range(5):
            if stage_graphs is None:
                fns[bi][si]()
            else:
                stage_graphs[bi][si].replay()
            if rec: rec[si + 1].record()

    ms_serial = timed_loop(D, step, K, Wm)
    launches = launches_per_step * K
    # ---- the timed number: TrainHotPath.step -- same kernels, same buffers, the table-backward
    # plan forked onto a side stream under the scoring kernels -- replayed as one CUDA graph per
    # rotating batch (the way a launch-bound chain of short kernels is meant to be driven);
    # eager launches of the same step are timed beside it
    ms_eager = timed_loop(D, lambda i: hp.step(batches[i % n_batches]["tb"].dec_in, batches[i % n_batches]["tb"].dec_out,
                                               batches[i % n_batches]["tb"].n_items, batches[i % n_batches]["h"],
                                               batches[i % n_batches]["g"]), K, Wm)
    # The timed region is exactly K steps (contract).  At 0.15 ms per step K = 20 is 3 ms -- shorter than
    # one nvidia-smi sample and than the time clocks take to settle -- so it is preceded, inside the same
    # clock-sampler window, by >= 60 ms of the same step (untimed), and a sustained figure over >= 400
    # steps is reported beside it.
    def run_step(i):
        b = batches[i % n_batches]
        hp.step(b["tb"].dec_in, b["tb"].dec_out, b["tb"].n_items, b["h"], b["g"])

    if args.no_graph:
        timed_fn = run_step
    else:
        graphs = [hp.capture_step(b["tb"].dec_in, b["tb"].dec_out, b["tb"].n_items, b["h"], b["g"])
                  for b in batches]

        def timed_fn(i):
            graphs[i % n_batches].replay()
    n_sustain = max(400, K)
    with ClockSampler(torch.cuda.current_device()) as cs:
        ms_sustain = timed_loop(D, timed_fn, n_sustain, Wm)      # pre-load: clocks at their under-load state
        ms = timed_loop(D, timed_fn, K, Wm)
        timed_loop(D, timed_fn, n_sustain, 0)                       # keeps the sampler window under load
    dh_name = "sampled_score_bwd_dh_standalone" if hp.fuse_dh else "sampled_score_bwd_dh"
    stage_names = ["embedding_bwd_plan", "gather_pe_mask", "sampled_score_loss_fwd", dh_name, "embedding_bwd_apply"]
    stage_ms = {nme: sum(evs[i][j].elapsed_time(evs[i][j + 1]) for i in range(K)) / K
                for j, nme in enumerate(stage_names)}
    # distinct touched rows (for the algorithmic bytes of the table backward)
    tb0 = batches[0]["tb"]
    m0 = tb0.mask.view(-1) != 0                                   # positions that carry gradient
    live_ids = torch.cat([tb0.dec_in.view(-1)[m0 & (tb0.dec_in.view(-1) != 0)],
                          tb0.dec_out.view(-1)[m0 & (tb0.dec_out.view(-1) != 0)],
                          tb0.n_items.view(-1, cfg["K"])[m0].reshape(-1)])
    live, distinct = int(live_ids.numel()), int(torch.unique(live_ids).numel())
    abytes = hp.algorithmic_bytes(distinct, live, int(m0.sum()))
    abytes_dense = hp.algorithmic_bytes(distinct, None, None)
    stages = {nme: {"ms": round(stage_ms[nme], 4), "algorithmic_mb": round(abytes[nme] / 1e6, 1),
                    "gbs": round(abytes[nme] / 1e6 / stage_ms[nme], 1),
                    "frac_of_hbm_peak": round(abytes[nme] / 1e6 / stage_ms[nme] / peaks["hbm"], 3)}
              for nme in stage_names}
    stages["embedding_bwd_plan"]["schedule"] = ("off the critical path: forked onto a high-priority side stream at "
                                                "the start of the step (own pad-mask kernel), runs under the "
                                                "gather and scoring kernels")
    if hp.fuse_dh:
        stages["sampled_score_loss_fwd"]["fused"] = (
            "forward + d loss/d h in one kernel over the plan's live-row list (recguru_sampled_score_loss_fwd_dh_rows)"
            if hp.score_rows else "forward + d loss/d h in one kernel (recguru_sampled_score_loss_fwd_dh)")
        stages[dh_name]["schedule"] = "NOT launched by the timed step (dh is fused into the forward); timed for reference"
    # roofline: the longest stage ON the critical path of the timed schedule
    critical = [s_ for s_ in stage_names if s_ not in ("embedding_bwd_plan", "sampled_score_bwd_dh_standalone")]
    dom = max(critical, key=lambda s_: stage_ms[s_])
    stage_kernels = {"gather_pe_mask": ("gather_pe_mask_kernel",),
                     "sampled_score_loss_fwd": ("sampled_score_rows_kernel" if hp.score_rows else "sampled_score_fwd",),
                     "sampled_score_bwd_dh": ("sampled_score_bwd_dh",),
                     "sampled_score_bwd_dh_standalone": ("sampled_score_bwd_dh",),
                     "embedding_bwd_apply": ("chunk_reduce_tma_kernel", "span_fixup_kernel")}
    roof = {"bound": "hbm", "kernel": dom, "achieved": stages[dom]["gbs"], "peak": peaks["hbm"], "unit": "GB/s",
            "frac": stages[dom]["frac_of_hbm_peak"], "traffic": ncu_traffic(*stage_kernels[dom]),
            "traffic_source": "profiles/r02_traffic.json (ncu --set full, DRAM read+write per launch)",
            "peak_source": peaks["src"],
            "whole_step": {"algorithmic_mb": round(abytes["total"] / 1e6, 1),
                           "gbs": round(abytes["total"] / 1e6 / (ms / K), 1),
                           "frac": round(abytes["total"] / 1e6 / (ms / K) / peaks["hbm"], 3)},
            "note": "longest stage on the critical path of the timed schedule; stage durations are measured "
                    "with the kernels serialised on one stream; whole_step = all algorithmic bytes of a step "
                    "over the timed step"}

    # ---- the same step on a NON-padded batch: every position live, so nothing is
    # skipped or compacted away -- the seqs/s figure that does not depend on the batch's sparsity, and
    # the table backward (K2) against the HBM roofline where it actually moves its bytes
    gd = torch.Generator().manual_seed(4321)
    dense_tb = batches[0]["tb"]
    dense = dict(dec_in=torch.randint(1, cfg["n_items"] + 1, (cfg["B"], cfg["L"]), generator=gd).to(dev),
                 dec_out=torch.randint(1, cfg["n_items"] + 1, (cfg["B"], cfg["L"]), generator=gd).to(dev),
                 n_items=dense_tb.n_items, h=batches[0]["h"], g=batches[0]["g"])
    if args.no_graph:
        def dense_fn(i):
            hp.step(dense["dec_in"], dense["dec_out"], dense["n_items"], dense["h"], dense["g"])
    else:
        dense_graph = hp.capture_step(dense["dec_in"], dense["dec_out"], dense["n_items"], dense["h"], dense["g"])

        def dense_fn(i):
            dense_graph.replay()
    ms_dense = timed_loop(D, dense_fn, K, Wm) / K
    dm = torch.ones(n, device=dev)
    d_sites_plan = [ops.Site(dense["dec_in"], None, live=dm), ops.Site(dense["dec_out"], None, live=dm),
                    ops.Site(dense["n_items"], None, group=cfg["K"], live=dm)]
    d_sites = [ops.Site(dense["dec_in"], dense["g"].view(n, cfg["d"]), coef=dm, live=dm),
               ops.Site(dense["dec_out"], dense["h"].view(n, cfg["d"]), coef=hp.sbuf["dp"], scale=hp.one,
                        scale2=hp.sbuf["aux"][1:2], live=dm),
               ops.Site(dense["n_items"], dense["h"].view(n, cfg["d"]), coef=hp.sbuf["dn"].view(-1), scale=hp.one,
                        scale2=hp.sbuf["aux"][1:2], group=cfg["K"], live=dm)]
    ms_dplan = timed_loop(D, lambda i: ops.raw_embedding_bwd_plan(d_sites_plan, hp.V1, 0, hp.ws_bwd), max(K, 10), Wm) / max(K, 10)
    ms_dapply = timed_loop(D, lambda i: ops.raw_embedding_bwd_apply(d_sites, hp.V1, hp.d, hp.dW, hp.ws_bwd,
                                                                    mode=ops.DW_PREZEROED), max(K, 10), Wm) / max(K, 10)
    d_live_ids = torch.cat([dense["dec_in"].view(-1), dense["dec_out"].view(-1), dense["n_items"].view(-1)])
    d_distinct = int(torch.unique(d_live_ids).numel())
    ab_dense = hp.algorithmic_bytes(d_distinct, None, None)
    k2_mb = (ab_dense["embedding_bwd_plan"] - 4 * hp.V1 * hp.d + ab_dense["embedding_bwd_apply"]) / 1e6
    dense_batch = {
        "ms_per_step": ms_dense, "value": cfg["B"] * D.world / (ms_dense * 1e-3), "unit": "seqs/s",
        "algorithmic_mb": round(ab_dense["total"] / 1e6, 1),
        "frac_of_hbm_peak": round(ab_dense["total"] / 1e6 / ms_dense / peaks["hbm"], 3),
        "k2_table_backward": {"plan_ms": round(ms_dplan, 4), "apply_ms": round(ms_dapply, 4),
                              "gathered_indices": int(d_live_ids.numel()), "distinct_rows": d_distinct,
                              "algorithmic_mb": round(k2_mb, 1),
                              "frac_of_hbm_peak": round(k2_mb / (ms_dplan + ms_dapply) / peaks["hbm"], 3),
                              "note": "plan + apply launched back to back on one stream (no overlap), "
                                      "zero fill of dW not included in either the bytes or the time"},
        "note": "no pad positions: all 1.43 M gathered indices carry gradient (uniform ids)"}
    del dm, d_sites, d_sites_plan

    # ---- fused dropout: K1 with torch-compatible Philox dropout vs K1 + torch's dropout
    b0 = batches[0]["tb"]
    xd = torch.empty(cfg["B"], cfg["L"], cfg["d"], device=dev)

    def k1_fused(i):
        seed, off, thr = ops.torch_dropout_state(xd.numel(), dev)
        ops.raw_gather_pe_mask_dropout(hp.weight, b0.dec_in, hp.pe, hp.pad_index, 0.5, seed, off, thr, out=xd,
                                       mask_out=hp._main_masks[0])

    def k1_then_torch(i):
        ops.raw_gather_pe_mask(hp.weight, b0.dec_in, hp.pe, hp.pad_index, out=xd, mask_out=hp._main_masks[0])
        torch.nn.functional.dropout(xd, 0.5, True)
    nd = max(K, 20)
    ms_k1f = timed_loop(D, k1_fused, nd, Wm) / nd
    ms_k1t = timed_loop(D, k1_then_torch, nd, Wm) / nd
    fused_dropout = {"k1_with_fused_dropout_ms": round(ms_k1f, 4), "k1_then_torch_dropout_ms": round(ms_k1t, 4),
                     "note": "p = 0.5 (config_auto4rec.py:225); same kept elements and generator offsets as nn.Dropout "
                             "(tests/test_gpu_dropout.py); padded cfg2 batch; not part of the timed step"}
    del xd

    # ---- the drop-in surface: what the reference's UNCHANGED call sites execute --
    # FusedItemEmbedding.lookup_pe_mask + sampled_loss_fused + loss.backward() through
    # torch.ops.recguru.* and torch autograd, eager launches; the table gradient of all lookup sites is
    # reduced in one K2 pass per backward() (ops.enable_one_pass_backward) and lands in weight.grad
    from recguru_b200.modules import FusedItemEmbedding, FusedPositionalEncoding, sampled_loss_fused
    dropin = {}
    for one_pass in (True, False):
        emb = FusedItemEmbedding(hp.V1, hp.d, padding_idx=0, _weight=W.clone(), one_pass_backward=one_pass)
        posm = FusedPositionalEncoding(hp.d, 0.0, max_len=max(cfg["L"], 256)).to(dev).eval()
        hs = [b["h"].clone().requires_grad_(True) for b in batches]

        def dropin_step(i, emb=emb, posm=posm, hs=hs):
            b = batches[i % n_batches]
            h_ = hs[i % n_batches]
            emb.weight.grad = None
            h_.grad = None
            x_, m_ = emb.lookup_pe_mask(b["tb"].dec_in, posm, pad_index=0)
            loss_, _, _ = sampled_loss_fused(h_, emb, b["tb"].dec_out, b["tb"].n_items, m_.view(-1), cfg["kind"])
            torch.autograd.backward([loss_, x_], [None, b["g"]])
        ms_d = timed_loop(D, dropin_step, K, Wm) / K
        rec = {"ms_per_step": round(ms_d, 4), "seqs_per_s": round(cfg["B"] * D.world / (ms_d * 1e-3))}
        if one_pass and not args.no_graph:
            # the same forward + backward() captured as ONE CUDA graph (autograd inside the capture): what
            # the drop-in path costs on the device once the Python / dispatcher / autograd-engine time of
            # the eager loop (several hundred us per step on this host) is out of the picture
            try:
                warm = torch.cuda.Stream()
                warm.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(warm):
                    for w_ in range(3):
                        dropin_step(w_)
                torch.cuda.current_stream().wait_stream(warm)
                dgs = []
                for bi in range(n_batches):
                    gph = torch.cuda.CUDAGraph()
                    emb.weight.grad = None
                    hs[bi].grad = None
                    with torch.cuda.graph(gph):
                        b = batches[bi]
                        x_, m_ = emb.lookup_pe_mask(b["tb"].dec_in, posm, pad_index=0)
                        loss_, _, _ = sampled_loss_fused(hs[bi], emb, b["tb"].dec_out, b["tb"].n_items, m_.view(-1),
                                                         cfg["kind"])
                        torch.autograd.backward([loss_, x_], [None, b["g"]])
                    dgs.append(gph)
                ms_g = timed_loop(D, lambda i: dgs[i % n_batches].replay(), K, Wm) / K
                rec["cuda_graph_ms_per_step"] = round(ms_g, 4)
                rec["cuda_graph_seqs_per_s"] = round(cfg["B"] * D.world / (ms_g * 1e-3))
                del dgs
            except Exception as e:  # noqa: BLE001  (a capture that torch refuses must not kill the bench line)
                rec["cuda_graph_error"] = repr(e)[:200]
                torch.cuda.synchronize()
        dropin["one_pass_backward" if one_pass else "per_site_dense_gradients"] = rec
        del emb, posm, hs
    dropin["ratio_vs_hotpath_eager"] = round(dropin["one_pass_backward"]["ms_per_step"] / (ms_eager / K), 3)
    if "cuda_graph_ms_per_step" in dropin["one_pass_backward"]:
        dropin["ratio_graph_vs_hotpath_graph"] = round(dropin["one_pass_backward"]["cuda_graph_ms_per_step"] / (ms / K), 3)
    dropin["note"] = ("eager launches through torch.ops.recguru.* + autograd (no CUDA graph, no side-stream plan, "
                      "allocations by torch's caching allocator); padding-aware scoring kernel; all lookup sites of "
                      "the step reduced in one row-compacted K2 pass when backward() ends")

    # ---- next row: the optimizer step of the table, timed on its own (NOT part of
    # the step above: north_star stops at the dense gradient)
    from recguru_b200.optim import FusedTableAdam
    w_adam = W.clone()
    opt = FusedTableAdam(w_adam, lr=1e-3)
    ms_adam = timed_loop(D, lambda i: opt.step(hp.dW), max(K, 10), Wm) / max(K, 10)
    adam_mb = 7 * W.numel() * 4 / 1e6
    table_adam = {"ms": round(ms_adam, 4), "algorithmic_mb": round(adam_mb, 1),
                  "gbs": round(adam_mb / ms_adam, 1), "frac_of_hbm_peak": round(adam_mb / ms_adam / peaks["hbm"], 3),
                  "note": "torch.optim.Adam arithmetic for the [V1,d] table in one kernel (read W,m,v,g; write W,m,v); "
                          "not included in the timed step"}
    # sparse-gradient form: dW as a row store (no zero fill), rows stamped by mark_rows, g = 0 elsewhere
    tb0 = batches[0]["tb"]
    m0 = hp.mask.view(n)
    hp.forward_in(tb0.dec_in)
    plan_sites = [ops.Site(tb0.dec_in, None, live=m0), ops.Site(tb0.dec_out, None, live=m0),
                  ops.Site(tb0.n_items, None, group=cfg["K"], live=m0)]
    ops.raw_embedding_bwd_plan(plan_sites, hp.V1, 0, hp.ws_bwd)
    row_epoch = torch.zeros(hp.V1, dtype=torch.int32, device=dev)
    ms_mark = timed_loop(D, lambda i: ops.raw_embedding_bwd_mark_rows(plan_sites, hp.V1, hp.ws_bwd, row_epoch, 1),
                         max(K, 10), Wm) / max(K, 10)
    rows_live = int((row_epoch == 1).sum().item())
    ms_adam_rows = timed_loop(D, lambda i: opt.step_rows(hp.dW, row_epoch, 1), max(K, 10), Wm) / max(K, 10)
    rows_mb = (6 * W.numel() * 4 + rows_live * cfg["d"] * 4 + hp.V1 * 4) / 1e6
    table_adam["row_store"] = {
        "ms": round(ms_adam_rows, 4), "mark_rows_ms_eager_launch": round(ms_mark, 4), "rows_with_gradient": rows_live,
        "algorithmic_mb": round(rows_mb, 1), "gbs": round(rows_mb / ms_adam_rows, 1),
        "frac_of_hbm_peak": round(rows_mb / ms_adam_rows / peaks["hbm"], 3),
        "note": "recguru_table_adam_rows: same arithmetic, gradient rows read only where the table backward "
                "wrote them (TrainHotPath(dense_grad=False)); also removes the 4*V1*d zero fill of dW from the step"}
    del w_adam, opt

    # ---- end to end through the public API with host buffers ----
    # A prefetching loader: the ids of step i+1 are copied (pinned host -> device, side stream)
    # while step i computes; the loss of every step is read back and waited for, as the
    # reference's loop does with loss.item().  All copies happen inside the timed region.
    pin = [dict(dec_in=t.dec_in.pin_memory(), dec_out=t.dec_out.pin_memory(), n_items=t.n_items.pin_memory())
           for t in host]
    # The ids of a batch are < V1 < 2^31: a loader that collates into pinned INT32 buffers halves the
    # per-step host->device bytes (the PCIe-bound part of this loop, and the part that collapses when 8
    # replicas share the host links); recguru_widen_ids restores the int64 layout the ops take, on the
    # copy stream.  The int64-over-PCIe form (what default_collate hands over, data_loader.py:418-421)
    # is timed beside it.
    # ... and into ONE pinned block per step ([dec_in | dec_out | n_items], int32): one H2D copy and one
    # widen kernel instead of three of each
    keys3 = ("dec_in", "dec_out", "n_items")
    sizes3 = [pin[0][k].numel() for k in keys3]
    pin32 = [torch.cat([pb[k].reshape(-1).to(torch.int32) for k in keys3]).pin_memory() for pb in pin]
    d_blk64 = [torch.empty(sum(sizes3), dtype=torch.int64, device=dev) for _ in range(2)]
    d_blk32 = [torch.empty(sum(sizes3), dtype=torch.int32, device=dev) for _ in range(2)]
    d_in = []
    for slot in range(2):                                   # the int64 views the step's kernels read
        o, views = 0, {}
        for k, sz in zip(keys3, sizes3):
            views[k] = d_blk64[slot][o:o + sz].view(pin[0][k].shape)
            o += sz
        d_in.append(views)
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream()
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    state = {"primed": -1, "ids32": True}

    def enqueue_copy(i):
        slot = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])          # buffer free (its previous step finished)
            if state["ids32"]:
                d_blk32[slot].copy_(pin32[i % n_batches], non_blocking=True)
                ops.raw_widen_ids(d_blk32[slot], out=d_blk64[slot])
            else:
                for k in keys3:
                    d_in[slot][k].copy_(pin[i % n_batches][k], non_blocking=True)
            copied[slot].record(copy_stream)

    for ev in consumed:
        ev.record()

    def e2e_step(i):
        if state["primed"] < i:                              # first step of a loop: nothing prefetched yet
            enqueue_copy(i)
        slot = i % 2
        b = batches[i % n_batches]
        cur = torch.cuda.current_stream()
        cur.wait_event(copied[slot])
        enqueue_copy(i + 1)                                  # overlaps this step's kernels
        state["primed"] = i + 1
        if e2e_graphs is None:
            hp.step(d_in[slot]["dec_in"], d_in[slot]["dec_out"], d_in[slot]["n_items"], b["h"], b["g"])
        else:
            e2e_graphs[i % n_batches].replay()
        consumed[slot].record(cur)
        loss_host.copy_(hp.sbuf["aux"][0:1], non_blocking=True)
        cur.synchronize()                                    # the reference reads loss.item() every step

    def restart_e2e(ids32):
        state["primed"], state["ids32"] = -1, ids32
        torch.cuda.synchronize()
        for ev in consumed:
            ev.record()

    # batch b always lands in id slot b % 2 (n_batches is even), so one graph per batch suffices
    for slot in range(2):                       # valid ids in the static buffers while capturing
        for k in d_in[slot]:
            d_in[slot][k].copy_(pin[slot][k])
    # both launch modes are timed (K steps each): eager TrainHotPath.step, and the same step replayed
    # as a CUDA graph; the headline e2e is the faster one and says which it was
    e2e_graphs = None
    ms_e2e_eager = timed_loop(D, e2e_step, K, Wm)
    ms_e2e_graph = ms_e2e_i64 = None
    if not args.no_graph:
        restart_e2e(True)
        e2e_graphs = [hp.capture_step(d_in[b % 2]["dec_in"], d_in[b % 2]["dec_out"], d_in[b % 2]["n_items"],
                                      batches[b]["h"], batches[b]["g"]) for b in range(n_batches)]
        ms_e2e_graph = timed_loop(D, e2e_step, K, Wm)
    # the same loop with the loss of step i read back while step i+1 runs (a training loop that logs one
    # step late): every step's ids still cross PCIe and every step's loss still reaches the host inside
    # the timed region, but the host no longer drains the GPU before it queues the next step.  Reported
    # beside the headline, not instead of it (the reference calls loss.item() in the step it belongs to).
    ms_e2e_lagged = None
    if e2e_graphs is not None:
        loss_host2 = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        loss_ev = [torch.cuda.Event(), torch.cuda.Event()]

        def e2e_step_lagged(i):
            first = state["primed"] < i
            if first:
                enqueue_copy(i)
            slot = i % 2
            cur = torch.cuda.current_stream()
            cur.wait_event(copied[slot])
            enqueue_copy(i + 1)
            state["primed"] = i + 1
            e2e_graphs[i % n_batches].replay()
            consumed[slot].record(cur)
            loss_host2[slot].copy_(hp.sbuf["aux"][0:1], non_blocking=True)
            loss_ev[slot].record(cur)
            if not first:
                loss_ev[slot ^ 1].synchronize()              # the previous step's loss is on the host now

        restart_e2e(True)
        ms_e2e_lagged = timed_loop(D, e2e_step_lagged, K, Wm)
    restart_e2e(False)
    ms_e2e_i64 = timed_loop(D, e2e_step, K, Wm)              # same loop, int64 ids over PCIe
    restart_e2e(True)
    ms_e2e = ms_e2e_eager if ms_e2e_graph is None else min(ms_e2e_eager, ms_e2e_graph)

    # next row: negatives drawn on the device by recguru_sample_negatives instead of
    # arriving over PCIe -- only dec_in/dec_out cross the link.  Reported beside the e2e number,
    # not instead of it (the reference's loader samples on the host).
    from recguru_b200.sampler import NegativeSampler
    smp = NegativeSampler(cfg["n_items"] + 1, cfg["L"] * cfg["K"], seed=1)
    e2e_graphs = None
    state["primed"] = -1
    torch.cuda.synchronize()
    for ev in consumed:
        ev.record()

    def enqueue_copy_ids(i):
        slot = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            for k in ("dec_in", "dec_out"):
                d_in[slot][k].copy_(pin[i % n_batches][k], non_blocking=True)
            copied[slot].record(copy_stream)

    def e2e_sampler_step(i):
        if state["primed"] < i:
            enqueue_copy_ids(i)
        slot = i % 2
        b = batches[i % n_batches]
        cur = torch.cuda.current_stream()
        cur.wait_event(copied[slot])
        enqueue_copy_ids(i + 1)
        state["primed"] = i + 1
        smp.sample(d_in[slot]["dec_in"], out=d_in[slot]["n_items"].view(cfg["B"], -1))
        hp.step(d_in[slot]["dec_in"], d_in[slot]["dec_out"], d_in[slot]["n_items"], b["h"], b["g"])
        consumed[slot].record(cur)
        loss_host.copy_(hp.sbuf["aux"][0:1], non_blocking=True)
        cur.synchronize()

    ms_e2e_smp = timed_loop(D, e2e_sampler_step, K, Wm)
    e2e_sampler = {"value": cfg["B"] * D.world * K / (ms_e2e_smp / 1e3), "unit": "seqs/s",
                   "ms_per_step": ms_e2e_smp / K,
                   "h2d_bytes_per_step": sum(pin[0][k].numel() * 8 for k in ("dec_in", "dec_out")),
                   "note": "negatives sampled on the device (recguru_sample_negatives, same law as "
                           "data_loader.py:289-314); eager launches"}
    # next row (second half): the dataset resident in HBM in ragged form
    # (recguru_b200.loader.SequenceStore), rows built by recguru_build_batches, negatives by the
    # device sampler -- the only per-step host input left is the batch's user indices (8 B each).
    from recguru_b200.loader import SequenceStore
    eos = cfg["n_items"] + 1
    ragged = []
    for t in host:                                      # histories back from the padded rows
        for row in t.enc_in.numpy():
            ragged.append(row[(row != 0) & (row != eos)])
    store = SequenceStore(ragged, None, None, eos, cfg["L"])
    idx_pin = [torch.arange(b * cfg["B"], (b + 1) * cfg["B"], dtype=torch.int64).pin_memory()
               for b in range(n_batches)]
    idx_dev = [torch.empty(cfg["B"], dtype=torch.int64, device=dev) for _ in range(2)]
    rows = [store.train_batch(idx_dev[s].copy_(idx_pin[s])) for s in range(2)]
    for s in range(2):                                   # the builder reproduces the host-built rows
        assert torch.equal(rows[s].dec_in, batches[s]["tb"].dec_in) and \
            torch.equal(rows[s].dec_out, batches[s]["tb"].dec_out), "device-built rows differ from the host rows"

    def e2e_builder_step(i):
        slot = i % 2
        b = batches[i % n_batches]
        idx_dev[slot].copy_(idx_pin[i % n_batches], non_blocking=True)
        r = store.train_batch(idx_dev[slot], out=rows[slot])
        smp.sample(r.enc_in, out=d_in[slot]["n_items"].view(cfg["B"], -1))
        hp.step(r.dec_in, r.dec_out, d_in[slot]["n_items"], b["h"], b["g"])
        loss_host.copy_(hp.sbuf["aux"][0:1], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    bld_graphs = None
    ms_e2e_bld_eager = timed_loop(D, e2e_builder_step, K, Wm)
    ms_e2e_bld_graph = None
    if not args.no_graph:
        # the same chain as ONE CUDA graph per rotating batch: rows, negatives (the sampler reads its
        # step from a device counter that the graph itself advances, so every replay draws fresh
        # negatives), then the step's kernels; the index copy and the loss read-back stay outside
        step_dev = torch.full((1,), 1 << 20, dtype=torch.int64, device=dev)

        def chain(bi):
            slot = bi % 2
            r = store.train_batch(idx_dev[slot], out=rows[slot])
            smp.sample(r.enc_in, out=d_in[slot]["n_items"].view(cfg["B"], -1), step_dev=step_dev)
            smp.advance(step_dev)
            hp.step(r.dec_in, r.dec_out, d_in[slot]["n_items"], batches[bi]["h"], batches[bi]["g"])

        warm = torch.cuda.Stream()
        warm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(warm):
            for bi in range(n_batches):
                chain(bi)
        torch.cuda.current_stream().wait_stream(warm)
        bld_graphs = []
        for bi in range(n_batches):
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph):
                chain(bi)
            bld_graphs.append(gph)

        def e2e_builder_graph_step(i):
            bi = i % n_batches
            idx_dev[bi % 2].copy_(idx_pin[bi], non_blocking=True)
            bld_graphs[bi].replay()
            loss_host.copy_(hp.sbuf["aux"][0:1], non_blocking=True)
            torch.cuda.current_stream().synchronize()

        ms_e2e_bld_graph = timed_loop(D, e2e_builder_graph_step, K, Wm)
    ms_e2e_bld = ms_e2e_bld_eager if ms_e2e_bld_graph is None else min(ms_e2e_bld_eager, ms_e2e_bld_graph)
    e2e_builder = {"value": cfg["B"] * D.world * K / (ms_e2e_bld / 1e3), "unit": "seqs/s",
                   "ms_per_step": ms_e2e_bld / K, "h2d_bytes_per_step": cfg["B"] * 8, "d2h_bytes_per_step": 4,
                   "ms_per_step_eager": ms_e2e_bld_eager / K,
                   "ms_per_step_graph": None if ms_e2e_bld_graph is None else ms_e2e_bld_graph / K,
                   "resident_dataset_bytes": int(store.items.numel() * 8 + store.offsets.numel() * 8),
                   "note": "ragged histories resident in HBM (SequenceStore), rows built on the device "
                           "(recguru_build_batches = seq_padding, data_loader.py:25-36), negatives sampled on the "
                           "device; only the batch's user indices cross PCIe; value = the faster of eager launches and one "
                           "CUDA graph per batch (sampler step read from a device counter the graph advances)"}
    e2e_mode = "eager (TrainHotPath.step)" if ms_e2e == ms_e2e_eager else "cuda_graph (TrainHotPath.capture_step)"
    h2d = pin32[0].numel() * pin32[0].element_size()
    h2d_i64 = sum(v.numel() * v.element_size() for v in pin[0].values())
    seqs = cfg["B"] * D.world
    out = {
        "value": seqs * K / (ms / 1e3), "ms_per_step": ms / K, "stages": stages, "roofline": roof,
        "sustained": {"steps": n_sustain, "ms_per_step": ms_sustain / n_sustain,
                      "value": seqs * n_sustain / (ms_sustain / 1e3), "unit": "seqs/s",
                      "note": "same step, same launch mode, timed over >= 400 steps right before the K-step region "
                              "(the clock sampler covers both)"},
        "e2e": {"value": seqs * K / (ms_e2e / 1e3), "unit": "seqs/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / K,
                "launch_mode": e2e_mode, "ms_per_step_eager": ms_e2e_eager / K,
                "ms_per_step_graph": None if ms_e2e_graph is None else ms_e2e_graph / K,
                "ids_on_host": "int32, one pinned block per step [dec_in | dec_out | n_items] (the loader's collate); one "
                               "H2D copy, widened to int64 on the device by recguru_widen_ids on the copy stream",
                "int64_ids": {"value": seqs * K / (ms_e2e_i64 / 1e3), "ms_per_step": ms_e2e_i64 / K,
                              "h2d_bytes_per_step": h2d_i64,
                              "note": "same loop with int64 ids over PCIe (what default_collate hands over)"},
                "lagged_loss_read": None if ms_e2e_lagged is None else {
                    "value": seqs * K / (ms_e2e_lagged / 1e3), "ms_per_step": ms_e2e_lagged / K,
                    "note": "same loop (graph replay, int32 id block), the loss of step i read back and waited for "
                            "while step i+1 runs -- a loop that logs one step late; every step's copies are still "
                            "inside the timed region"},
                "note": "ids from pinned host memory each step, loss read back each step; h and the "
                        "gradient w.r.t. the input activations come from the (out-of-scope) transformer "
                        "and are device resident"},
        "gpu_launches": int(launches), "clocks": cs.summary(), "loss": float(loss_host.item()),
        "launch_mode": "eager" if args.no_graph else "cuda_graph (one graph per rotating batch)",
        "table_adam": table_adam, "e2e_device_negatives": e2e_sampler, "e2e_device_batches": e2e_builder,
        "dropin_surface": dropin, "dense_batch": dense_batch, "fused_dropout": fused_dropout,
        "ms_per_step_eager": ms_eager / K, "ms_per_step_serialised_stages": ms_serial / K,
        "algorithmic_mb_per_step": round(abytes["total"] / 1e6, 1),
        "table_backward": {"gathered_indices": cfg["B"] * cfg["L"] * (2 + cfg["K"]), "live_indices": live,
                           "distinct_rows": distinct,
                           "algorithmic_mb_if_all_live": round((abytes_dense["embedding_bwd_plan"]
                                                                 + abytes_dense["embedding_bwd_apply"]) / 1e6, 1),
                           "note": "dead indices (pad id or zero coefficient under the pad mask) are compacted "
                                   "away before the sort; roofline numerator counts only their 12 B id+coef"},
    }
    return out, (W, pe, batches)


def bench_gather(D: Dist, args, peaks, W, pe, batches):
    """Gather GB/s in isolation: fused gather+PE+mask, rotating ids/out buffers (> L2)."""
    from recguru_b200 import ops
    cfg = CFG2
    dev = W.device
    outs = [torch.empty(cfg["B"], cfg["L"], cfg["d"], device=dev) for _ in range(4)]
    masks = [torch.empty(cfg["B"], cfg["L"], device=dev) for _ in range(4)]

    def step(i):
        j = i % 4
        ops.raw_gather_pe_mask(W, batches[j]["tb"].dec_in, pe, 0, out=outs[j], mask_out=masks[j])

    ms = timed_loop(D, step, max(args.steps, 20), args.warmup)
    n = cfg["B"] * cfg["L"]
    mb = n * (8 + 8 * cfg["d"] + 4) / 1e6
    per = ms / max(args.steps, 20)
    # dense ids (no padding) variant: every token reads a real row
    ids = torch.randint(1, cfg["n_items"] + 1, (4, cfg["B"], cfg["L"]), device=dev)

    def step2(i):
        j = i % 4
        ops.raw_gather_pe_mask(W, ids[j], pe, 0, out=outs[j], mask_out=masks[j])

    ms2 = timed_loop(D, step2, max(args.steps, 20), args.warmup) / max(args.steps, 20)
    out = {"gbs": round(mb / per, 1), "frac_of_hbm_peak": round(mb / per / peaks["hbm"], 3),
           "ms": round(per, 4), "algorithmic_mb": round(mb, 1),
           "dense_ids": {"gbs": round(mb / ms2, 1), "frac_of_hbm_peak": round(mb / ms2 / peaks["hbm"], 3),
                         "ms": round(ms2, 4)},
           "note": "cfg2's 51 MB table lives in the 126 MB L2: these two are L2 figures, NOT HBM (amazon-shaped "
                   "batch: 93% pad tokens hit one row; dense_ids: a no-padding batch); 4 rotating out buffers "
                   "= 420 MB > L2.  The HBM figures are in `hbm_table`."}
    del outs, masks, ids
    if D.world == 1:
        out["hbm_table"] = bench_gather_hbm(D, args, peaks)
    return out


def bench_gather_hbm(D: Dist, args, peaks):
    """K1 against HBM: a table far larger than L2 -- cfg4's 1M x 128 fp32 =
    512 MB -- with uniform and Zipf(1.05) ids, no padding; 'cold' rotates over 8 id/out sets (each
    call touches 105 MB of rows spread over the 512 MB table and writes a fresh 105 MB output, so
    what a call finds in L2 is at most what the previous seven left), 'hot' repeats one set.  The
    DRAM bytes at the pins come from the ncu capture of this same loop (profiles/r02_traffic.json,
    key gather_hbm_*; scripts/gather_hbm.py)."""
    from recguru_b200 import ops, synth, modules
    dev = torch.device("cuda", torch.cuda.current_device())
    n_items, d, B, L = 1_000_000, 128, CFG2["B"], CFG2["L"]
    n = B * L
    gen = torch.Generator(device=dev).manual_seed(4242)
    W = torch.randn(n_items + 2, d, generator=gen, device=dev)
    pe = modules.sinusoid_table(max(L, 256), d)[0].to(dev)
    n_sets = 8
    outs = [torch.empty(B, L, d, device=dev) for _ in range(n_sets)]
    masks = [torch.empty(B, L, device=dev) for _ in range(n_sets)]
    ids_u = torch.randint(1, n_items + 1, (n_sets, B, L), generator=gen, device=dev)
    ids_z = synth.zipf_ids(n_sets * n, n_items, seed=99).view(n_sets, B, L).to(dev)
    mb = n * (8 + 8 * d + 4) / 1e6
    steps = max(args.steps, 40)
    res = {"table_mb": round(W.numel() * 4 / 1e6, 1), "algorithmic_mb": round(mb, 1), "tokens": n}
    for name, ids in (("uniform", ids_u), ("zipf", ids_z)):
        distinct = int(torch.unique(ids[0]).numel())
        compulsory_mb = (n * (8 + 4 * d + 4) + distinct * 4 * d) / 1e6      # every distinct row once + ids + output

        def cold(i, ids=ids):
            j = i % n_sets
            ops.raw_gather_pe_mask(W, ids[j], pe, 0, out=outs[j], mask_out=masks[j])

        def hot(i, ids=ids):
            ops.raw_gather_pe_mask(W, ids[0], pe, 0, out=outs[0], mask_out=masks[0])

        t_cold = timed_loop(D, cold, steps, args.warmup) / steps
        t_hot = timed_loop(D, hot, steps, args.warmup) / steps
        tr = ncu_traffic(f"gather_hbm_{name}_cold")
        res[name] = {
            "distinct_rows": distinct,
            "cold": {"ms": round(t_cold, 4), "gbs": round(mb / t_cold, 1),
                     "frac_of_hbm_peak": round(mb / t_cold / peaks["hbm"], 3),
                     "compulsory_mb": round(compulsory_mb, 1),
                     "compulsory_frac_of_hbm_peak": round(compulsory_mb / t_cold / peaks["hbm"], 3),
                     "dram_bytes_ncu": tr,
                     "dram_frac_of_hbm_peak": None if tr is None else round(tr / 1e6 / t_cold / peaks["hbm"], 3)},
            "hot": {"ms": round(t_hot, 4), "gbs": round(mb / t_hot, 1),
                    "frac_of_hbm_peak": round(mb / t_hot / peaks["hbm"], 3)},
        }
    res["note"] = ("algorithmic bytes = 8 (id) + 4d (row) + 4d (out) + 4 (mask) per token, cache hits not discounted; "
                   "compulsory = each distinct row read once; dram_bytes_ncu = dram read+write of one launch of the "
                   "cold loop under ncu (cold-cache, serialised) -- the claim against HBM is dram_frac_of_hbm_peak")
    return res



# ----------------------------------------------------------------------------------------------
# workload 1b: the other two step topologies (configs[0] AutoRec, configs[2] RecGURU)
# ----------------------------------------------------------------------------------------------
def _site_stats(pairs, padding_idx):
    """pairs: (ids [n*group], live mask [n], group).  -> (live gathered indices, distinct live ids)."""
    live_ids = []
    for ids, m, group in pairs:
        keep = m.view(-1, 1).expand(-1, group).reshape(-1) != 0
        v = ids.reshape(-1)[keep]
        if padding_idx is not None:
            v = v[v != padding_idx]
        live_ids.append(v)
    allv = torch.cat(live_ids)
    return int(allv.numel()), int(torch.unique(allv).numel())


def bench_topology(D: Dist, args, peaks, which: str):
    """Hot-path step of configs[0] (AutoRec: two input sites, K=30 sampled softmax and K=5 BPR) or
    configs[2] (RecGURU reconstruction: two domains x (enc_in, dec_in, pos, 30 negatives), tables
    without padding_idx, the two domains on two streams), CUDA-graph replay over 4 rotating batches,
    inputs resident; transformer replaced by given h and returned gradients, as in the cfg2 line."""
    from recguru_b200 import synth, modules
    from recguru_b200.hotpath import TrainStep, RecGuruStep
    dev = torch.device("cuda", torch.cuda.current_device())
    K_steps, Wm = args.steps, args.warmup
    n_batches = 4
    out = {}
    if which == "cfg1":
        cfg = CFG1
        variants = [("recon_sce_k30", 30, "sce"), ("bpr_sas_k5", 5, "bpr_sas")]
    else:
        cfg = CFG3
        # reconstruction phase (K = n_negs = 30, sampled softmax) and the BPR phase (K = 5, BPRLoss; tools/utils.py:12,:126)
        variants = [("recon_sce_k30", cfg["K"], cfg["kind"]), ("bpr_k5", 5, "bpr")]
    B, L, d, n_items = cfg["B"], cfg["L"], cfg["d"], cfg["n_items"]
    pe = modules.sinusoid_table(max(L, 256), d)[0].to(dev)
    for vname, K, kind in variants:
        def mk(seed0):
            bs = []
            for b in range(n_batches):
                tb = synth.train_batch(B, L, n_items, K, seed=seed0 + 100 * b).to(dev)
                bs.append(dict(tb=tb, h=synth.hidden_states(B, L, d, seed=seed0 + b).to(dev),
                               g=[(synth.hidden_states(B, L, d, seed=seed0 + 40 + 2 * b + i) * 0.01).to(dev)
                                  for i in range(2)]))
            return bs
        if which == "cfg1":
            W = synth.item_table(n_items, d).to(dev)
            st = TrainStep(W, pe, B, L, K, kind, n_inputs=2, padding_idx=0)
            bs = mk(synth.SEED_DATA)
            graphs = [st.capture_run([b["tb"].enc_in, b["tb"].dec_in], b["tb"].dec_out, b["tb"].n_items, b["h"], b["g"],
                                     loss_mask_ids=b["tb"].dec_in) for b in bs]
            tb0 = bs[0]["tb"]
            m_enc, m_dec = (tb0.enc_in != 0).float(), (tb0.dec_in != 0).float()
            live, distinct = _site_stats([(tb0.enc_in, m_enc, 1), (tb0.dec_in, m_dec, 1), (tb0.dec_out, m_dec, 1),
                                          (tb0.n_items, m_dec, K)], 0)
            ab = st.algorithmic_bytes(distinct, live, int(m_dec.sum()))["total"]
            seqs, tables = B, 1
            steps_obj = [st]
        else:
            Wa = synth.item_table(n_items, d, seed=synth.SEED_WEIGHTS).to(dev)
            Wb = synth.item_table(n_items, d, seed=synth.SEED_WEIGHTS + 1).to(dev)
            recommend = kind != "sce"      # BPR phase: recommend_forward masks dec_in by itself (AutoEnc4Rec_cross.py:150-160)
            rg_ = RecGuruStep(Wa, Wb, pe, B, L, K, kind, recommend=recommend)
            bsa, bsb = mk(synth.SEED_DATA), mk(synth.SEED_DATA + 5000)
            graphs = [rg_.capture_run(a["tb"], b["tb"], a["h"], b["h"], a["g"], b["g"]) for a, b in zip(bsa, bsb)]
            ab = 0
            for st_, b0 in ((rg_.dom[0], bsa[0]["tb"]), (rg_.dom[1], bsb[0]["tb"])):
                m_out = (b0.dec_out != 0).float()
                m_enc = ((b0.dec_in if recommend else b0.enc_in) != 0).float()
                live, distinct = _site_stats([(b0.enc_in, m_out, 1), (b0.dec_in, m_enc, 1), (b0.dec_out, m_out, 1),
                                              (b0.n_items, m_out, K)], None)
                ab += st_.algorithmic_bytes(distinct, live, int(m_out.sum()))["total"]
            seqs, tables = 2 * B, 2
            steps_obj = rg_.dom
        from recguru_b200 import _lib
        l0 = _lib.launch_count()
        graphs[0].replay()                      # graph replays do not pass through the C ABI: count eagerly below
        torch.cuda.synchronize()
        n_sustain = max(200, K_steps)
        with ClockSampler(torch.cuda.current_device()) as cs:
            timed_loop(D, lambda i: graphs[i % n_batches].replay(), n_sustain, Wm)
            ms = timed_loop(D, lambda i: graphs[i % n_batches].replay(), K_steps, Wm)
        per = ms / K_steps
        out[vname] = {"metric": "train seqs/sec", "value": seqs * D.world / (per * 1e-3), "unit": "seqs/s",
                      "ms_per_step": per, "steps": K_steps,
                      "sequences_per_step": seqs, "tables": tables, "K": K, "kind": kind,
                      "loss": float(steps_obj[0].sbuf["aux"][0].item()),
                      "roofline": {"bound": "hbm", "kernel": "whole step", "algorithmic_mb": round(ab / 1e6, 1),
                                   "achieved": round(ab / 1e6 / per, 1), "peak": peaks["hbm"], "unit": "GB/s",
                                   "frac": round(ab / 1e6 / per / peaks["hbm"], 3), "traffic": None,
                                   "note": "all algorithmic bytes of the step (every lookup site, one table backward "
                                           "per table incl. its dense zero fill) over the timed step"},
                      "clocks": cs.summary()}
        del graphs
        torch.cuda.empty_cache()
    conf = {"cfg1": "configs[0] AutoRec single-domain hot-path step: 10k items, L=50, d=64, B=256; input sites enc_in + "
                    "dec_in; reconstruction (K=30 sampled softmax) and BPR fine-tune (K=5 BPRLoss_sas)",
            "cfg3": "configs[2] RecGURU two-domain hot-path step: 2 x 500k items, L=100, d=128, B=1024 per domain, tables "
                    "without padding_idx, 4 lookup sites per domain; reconstruction phase (K=30 sampled softmax) and "
                    "BPR phase (K=5 BPRLoss)"}
    return {"config": {"workload": conf[which], **cfg}, "launch_mode": "cuda_graph (one graph per rotating batch)",
            "variants": out}

# ----------------------------------------------------------------------------------------------
# workload 2: full-catalog eval
# ----------------------------------------------------------------------------------------------
def grid_rows(lo: int, hi: int, d: int, salt: int, dev) -> torch.Tensor:
    """Rows [lo, hi) of a grid-valued matrix (values k/16, |k| <= 32) as a pure function of (row, column,
    salt): any rank can generate any row, so the data -- and every exact score on it -- does not depend on
    how the catalog is sharded."""
    out = torch.empty(hi - lo, d, dtype=torch.float32, device=dev)
    c = torch.arange(d, device=dev, dtype=torch.int64)[None, :]
    for r0 in range(lo, hi, 1 << 20):
        r1 = min(hi, r0 + (1 << 20))
        i = torch.arange(r0, r1, device=dev, dtype=torch.int64)[:, None]
        h = i * 2654435761 + c * 40503 + (i >> 5) * (c + 7) * 977 + salt * 7919
        h = (h ^ (h >> 13)) * 1274126177
        h = h ^ (h >> 16)
        out[r0 - lo:r1 - lo] = ((h % 65) - 32).to(torch.float32) / 16.0
    return out


def catalog_checksums(D: Dist, ev, cfg, u_step: int, G: int, dev):
    """One step of the TIMED path (same evaluator class, same shard layout, same user count, same
    kernels) on grid-valued users/items, where every bf16 x bf16 -> fp32 score is exact in any
    summation order: ranks and top-k are then a function of the data alone, so these checksums must
    be IDENTICAL at N = 1/2/4/8 (SURVEY 8e "shard-count-independent bit-identical") -- and rank 0
    verifies 256 of its users against plain fp32 torch arithmetic over the whole catalog."""
    from recguru_b200 import ops, catalog
    n_items, d, k = cfg["n_items"], cfg["d"], cfg["k"]
    lo, hi = ev.id_lo - 1, ev.id_lo - 1 + ev.n_local
    chk = catalog.CatalogEvaluator.__new__(catalog.CatalogEvaluator)
    chk.__dict__.update(ev.__dict__)
    chk.items = ops.raw_cast_bf16(grid_rows(lo, hi, d, 1, dev))
    u_loc = u_step // G
    u0 = ev.rank * u_loc
    users = grid_rows(u0, u0 + u_loc, d, 2, dev)
    ui = torch.arange(u0, u0 + u_loc, device=dev, dtype=torch.int64)
    target = (ui * 2654435761 + 12345) % n_items + 1
    val, idx, rank = chk.rank_topk(users, target, u_max=u_loc)
    rank_only = chk.target_rank(users, target, u_max=u_loc)
    acc = torch.zeros(2 + 24, dtype=torch.float64, device=dev)
    acc[0] = rank.sum().double()
    acc[1] = val.double().sum()
    for b in range(24):
        acc[2 + b] = ((idx >> b) & 1).sum().double()
    same_rank = torch.tensor([float(torch.equal(rank, rank_only))], dtype=torch.float64, device=dev)
    if D.dist is not None:
        D.dist.all_reduce(acc)
        D.dist.all_reduce(same_rank, op=D.dist.ReduceOp.MIN)
    xor = 0
    for b in range(24):
        xor |= (int(acc[2 + b].item()) & 1) << b
    out = {"users": u_step, "rank_sum": int(acc[0].item()), "topk_id_xor": xor, "topk_val_sum": float(acc[1].item()),
           "rank_only_equals_topk_mode": bool(same_rank.item() == 1.0),
           "data": "grid-valued (k/16) users and items generated from (row, column) hashes: shard-count independent"}
    if D.rank == 0:
        # parity against fp32 torch on the whole catalog for the first 256 users of rank 0
        torch.backends.cuda.matmul.allow_tf32 = False
        nu = min(256, u_loc)
        uu, tt = users[:nu], target[:nu] - 1
        ref_rank = torch.zeros(nu, dtype=torch.int64, device=dev)
        best = torch.full((nu, k), float("-inf"), device=dev)
        tsc = None
        for n0 in range(0, n_items, 1 << 20):
            n1 = min(n_items, n0 + (1 << 20))
            it = grid_rows(n0, n1, d, 1, dev)
            s_ = uu @ it.t()
            if tsc is None:
                tsc = torch.zeros(nu, device=dev)
            inside = (tt >= n0) & (tt < n1)
            if bool(inside.any()):
                tsc[inside] = s_[inside, (tt[inside] - n0)]
            best = torch.cat([best, s_.topk(k, dim=1).values], 1).topk(k, dim=1).values
            del s_, it
        for n0 in range(0, n_items, 1 << 20):
            n1 = min(n_items, n0 + (1 << 20))
            s_ = uu @ grid_rows(n0, n1, d, 1, dev).t()
            cols = torch.arange(n0, n1, device=dev)[None, :]
            ref_rank += (s_ > tsc[:, None]).sum(1) + ((s_ == tsc[:, None]) & (cols < tt[:, None])).sum(1)
            del s_
        vi = val[:nu]
        ii = idx[:nu]
        ties_ok = bool((~(vi[:, 1:] == vi[:, :-1]) | (ii[:, 1:] > ii[:, :-1])).all())
        out["vs_torch_fp32"] = {"users_checked": nu, "ranks_equal": bool(torch.equal(rank[:nu], ref_rank)),
                                "topk_values_equal": bool(torch.equal(vi, best)), "ties_ascending_id": ties_ok}
    return out


def bench_catalog(D: Dist, args, peaks, cfg, name, sharded: bool):
    from recguru_b200 import _lib, ops, catalog
    dev = torch.device("cuda", torch.cuda.current_device())
    G = D.world if sharded else 1
    n_items, d, k = cfg["n_items"], cfg["d"], cfg["k"]
    lo, hi = catalog.shard_bounds(n_items, G, D.rank if sharded else 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1235 + (D.rank if sharded else 0))
    # item slice generated on the device (cfg5 is 5 GB fp32 on the host otherwise); N(0,1) rows
    items = torch.empty(hi - lo, d, dtype=torch.bfloat16, device=dev)
    chunk = 1 << 20
    for r0 in range(0, hi - lo, chunk):
        r1 = min(hi - lo, r0 + chunk)
        items[r0:r1] = ops.raw_cast_bf16(torch.randn(r1 - r0, d, generator=gen, device=dev))
    ev = catalog.CatalogEvaluator.__new__(catalog.CatalogEvaluator)
    ev.k, ev.group, ev.world, ev.rank = k, (D.group() if sharded else None), G, (D.rank if sharded else 0)
    ev.id_lo, ev.n_local, ev.n_items, ev.d, ev.items, ev.id_first = 1 + lo, hi - lo, n_items, d, items, 1
    u_step = USERS_PER_STEP
    u_loc = u_step // G
    n_chunks = 4
    gen.manual_seed(77 + D.rank)
    users = [torch.randn(u_loc, d, generator=gen, device=dev) for _ in range(n_chunks)]
    users_bf = [ops.raw_cast_bf16(u) for u in users]
    targets = [torch.randint(1, n_items + 1, (u_loc,), generator=gen, device=dev) for _ in range(n_chunks)]
    steps = max(3, min(args.steps, 10 if n_items <= 1_000_000 else 4))
    warm = max(3, min(args.warmup, 3))
    res = {}
    parity = catalog_checksums(D, ev, cfg, u_step, G, dev)     # before timing: proves THIS path on THIS box
    torch.cuda.empty_cache()

    def step(i):
        j = i % n_chunks
        res["out"] = ev.rank_topk(users_bf[j], targets[j], u_max=u_loc)

    l0 = _lib.launch_count()
    with ClockSampler(torch.cuda.current_device()) as cs:
        ms = timed_loop(D, step, steps, warm)
    launches = (_lib.launch_count() - l0) * steps // (steps + warm)
    per = ms / steps

    # sweep kernel alone (dominant kernel) with CUDA events, single device view
    ub, tg = (users_bf[0], targets[0]) if G == 1 else (None, None)
    sweep_ms = None
    if G == 1:
        ts = ops.raw_catalog_target_scores(ub, items, tg, ev.id_lo)
        bufs = {}

        def sweep(i):
            ops.raw_catalog_topk_rank(users_bf[i % n_chunks], items, targets[i % n_chunks], ts, k, ev.id_lo, bufs=bufs)
        sweep_ms = timed_loop(D, sweep, steps, warm) / steps
    # rank-only mode (k = 0; what HR/NDCG/MRR need -- CatalogEvaluator.target_rank / evaluate)
    ms_rank = timed_loop(D, lambda i: ev.target_rank(users[i % n_chunks], targets[i % n_chunks], u_max=u_loc),
                         steps, warm) / steps
    flops = 2.0 * u_step * (hi - lo) * d
    rank_only = {"ms_per_step": round(ms_rank, 3), "users_per_s": round(u_step / (ms_rank * 1e-3)),
                 "tflops": round(flops / (ms_rank * 1e-3) / 1e12, 1),
                 "note": "target rank only: the sweep skips the maxima and the top-k lists"}
    t_kernel = sweep_ms if sweep_ms is not None else per
    tflops = flops / (t_kernel * 1e-3) / 1e12
    roof = {"bound": "tensor", "kernel": "catalog_sweep_kernel", "achieved": round(tflops, 1),
            "peak": peaks["tf_sust"], "unit": "TFLOP/s", "frac": round(tflops / peaks["tf_sust"], 3),
            "frac_of_burst_peak": round(tflops / peaks["tf_burst"], 3),
            # DRAM bytes of one sweep launch from the committed ncu capture (cfg4 shape only)
            "traffic": ncu_traffic("catalog_sweep_kernel") if name.startswith("cfg4") and D.world == 1 else None,
            "peak_source": peaks["src"] + " (sustained bf16 cuBLAS)"}

    # end to end: fp32 user reps + targets from pinned host memory, ranks read back
    pin_u = [u.cpu().pin_memory() for u in users[:2]]
    pin_t = [t.cpu().pin_memory() for t in targets[:2]]
    d_u = torch.empty_like(users[0]); d_t = torch.empty_like(targets[0])
    rank_host = torch.empty(u_loc, dtype=torch.int64).pin_memory()

    def e2e(i):
        d_u.copy_(pin_u[i % 2], non_blocking=True); d_t.copy_(pin_t[i % 2], non_blocking=True)
        _, _, rk = ev.rank_topk(d_u, d_t, u_max=u_loc)  # includes the fp32->bf16 cast kernel
        rank_host.copy_(rk, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    ms_e2e = timed_loop(D, e2e, steps, warm) / steps
    met = catalog.rank_metrics(res["out"][2], (5, 10))
    return {
        "metric": "full-catalog eval users/sec", "value": u_step / (per * 1e-3), "unit": "users/s",
        "ms_per_step": per, "steps": steps, "warmup": warm,
        "config": {"workload": name, "n_items": n_items, "d": d, "k": k, "users_per_step": u_step,
                   "users_total_cfg": cfg["users_total"], "sharding": f"rows/{G}" if sharded else "none",
                   "l2": "4 rotating user chunks; item table 256 MB-2.5 GB > L2"},
        "scaling": "strong" if sharded else None,
        "sweep_kernel_ms": sweep_ms, "rank_only": rank_only, "roofline": roof,
        "e2e": {"value": u_step / (ms_e2e * 1e-3), "unit": "users/s",
                "h2d_bytes_per_step": u_loc * (d * 4 + 8), "d2h_bytes_per_step": u_loc * 8, "ms_per_step": ms_e2e},
        "gpu_launches": int(launches), "clocks": cs.summary(),
        "hr10_synthetic": met[10]["hr"], "parity": parity,
    }


# ----------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's ATen op sequence on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_train_step_fn(cfg, B_cpu: int):
    from oracle import ref_path as O
    from recguru_b200 import synth, modules
    W = torch.nn.Parameter(synth.item_table(cfg["n_items"], cfg["d"]))
    pe = modules.sinusoid_table(max(cfg["L"], 256), cfg["d"])
    tb = synth.train_batch(B_cpu, cfg["L"], cfg["n_items"], cfg["K"])
    h = synth.hidden_states(B_cpu, cfg["L"], cfg["d"]).requires_grad_(True)
    g = synth.hidden_states(B_cpu, cfg["L"], cfg["d"], seed=91)
    L, K = cfg["L"], cfg["K"]

    def step():
        W.grad = None; h.grad = None
        x0, mask = O.gather_pe_mask(W, tb.dec_in, pe, 0, padding_idx=0)            # G1 + M1 + P1
        p, n = O.sampled_scores(h, W, tb.dec_out, tb.n_items, L, K, padding_idx=0)  # S1-S3
        loss = O.LOSSES[cfg["kind"]](p, n, mask.view(-1))                           # Lsas
        (loss + (x0 * g).sum()).backward()                                          # B1 (+ dh)
        return float(loss.detach())
    return step


def cpu_catalog_fn(cfg, U_cpu: int):
    from oracle import ref_path as O
    g = torch.Generator().manual_seed(5)
    W = torch.randn(cfg["n_items"], cfg["d"], generator=g)
    u = torch.randn(U_cpu, cfg["d"], generator=g)
    tgt = torch.randint(0, cfg["n_items"], (U_cpu,), generator=g)

    def step():
        s = torch.matmul(u, W.transpose(1, 0))                                     # AutoEnc4Rec.py:230
        rank = torch.argsort(torch.argsort(-s, dim=1), dim=1)[torch.arange(U_cpu), tgt]   # train_auto.py:234-237
        r = rank.cpu().detach().numpy()
        return O.hit_at_k_batch(r, 10), O.NDCG_at_k_batch(r, 10)
    return step


def train_config():
    """`config` of the primary line -- identical in both arms (the reference arm runs the same workload)."""
    return {"workload": "cfg2 SASRec hot-path training step (configs[1]): 100k items, L=200, "
                        "d=128, B=1024/GPU, K=5, BPRLoss_sas; replicas only for N>1",
            "l2": "4 rotating batches (h+g+out = 315 MB each) > 126 MB L2",
            **{k: v for k, v in CFG2.items()}}


def time_cpu(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    return (time.perf_counter() - t0) / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    if args.workload == "catalog":
        cfg, U_cpu = CFG5, 8
        sec = time_cpu(cpu_catalog_fn(cfg, U_cpu), max(1, min(args.steps, 3)), 1)
        val, unit, metric = U_cpu / sec, "users/s", "full-catalog eval users/sec"
        sample = f"{U_cpu} users x {cfg['n_items']} items per step (u@W^T + double argsort + metrics)"
        conf = {"workload": "cfg5 full-catalog eval, 10M items, d=128 (bounded CPU sample)"}
    else:
        cfg, B_cpu = CFG2, CFG2["B"]
        steps = max(1, min(args.steps, 20))
        sec = time_cpu(cpu_train_step_fn(cfg, B_cpu), steps, max(1, min(args.warmup, 3)))
        val, unit, metric = B_cpu / sec, "seqs/s", "train seqs/sec"
        sample = (f"the full batch: {B_cpu} of {cfg['B']} sequences per step (F.embedding+pe+mask, sampled matmuls, "
                  f"BPRLoss_sas, autograd backward incl. dense [V1,d] embedding grads)")
        conf = train_config()
    line = {"impl": "reference", "metric": metric, "value": val, "unit": unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": conf,
            "cpu_baseline": {"value": val, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "catalog"])
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replay")
    ap.add_argument("--only", default="", help="comma list of sections: train,gather,cfg1,cfg3,cfg4,cfg5")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    # The JSON line must be the only thing on stdout.  Libraries write there too (NCCL prints its
    # version banner on fd 1 at any NCCL_DEBUG level >= VERSION), so fd 1 is pointed at stderr for
    # the duration of the run and the line goes to a duplicate of the real stdout.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (recguru_b200 has no CPU path); "
                         "use --impl reference for the CPU reference arm")
    from recguru_b200 import build
    build.build()
    D = Dist(args.gpus)
    peaks = load_peaks()
    only = set(filter(None, args.only.split(",")))

    def want(s):
        return not only or s in only

    line = {}
    train = gather = None
    if want("train") or want("gather"):
        train, (W, pe, batches) = bench_train(D, args, peaks)
        if want("gather"):
            gather = bench_gather(D, args, peaks, W, pe, batches)
        del W, pe, batches
        torch.cuda.empty_cache()
    topo = {}
    for which in ("cfg1", "cfg3"):
        if want(which) and (D.world == 1 or which in only):   # replicas only: measured at N = 1 unless asked for
            topo[which] = bench_topology(D, args, peaks, which)
            torch.cuda.empty_cache()
    cat = {}
    if D.world == 1 and want("cfg4"):
        cat["cfg4"] = bench_catalog(D, args, peaks, CFG4, "cfg4 full-catalog eval 1M users x 1M items, d=128, top-10", False)
        torch.cuda.empty_cache()
    if want("cfg5"):
        cat["cfg5"] = bench_catalog(D, args, peaks, CFG5,
                                    "cfg5 full-catalog eval, 10M items row-sharded, d=128, top-10", D.world > 1)
        torch.cuda.empty_cache()

    cpu = None
    if D.rank == 0 and D.world == 1 and not args.skip_cpu:
        cores = os.cpu_count()
        torch.set_num_threads(cores)
        if args.workload == "train":
            B_cpu = CFG2["B"]
            sec = time_cpu(cpu_train_step_fn(CFG2, B_cpu), 10, 2)
            for cname, ccfg, U_cpu in (("cfg4", CFG4, 64), ("cfg5", CFG5, 16)):
                if cname in cat:
                    csec = time_cpu(cpu_catalog_fn(ccfg, U_cpu), 2, 1)
                    cat[cname]["cpu_baseline"] = {
                        "value": U_cpu / csec, "unit": "users/s", "cores": cores, "kind": "port",
                        "sample": f"{U_cpu} users x {ccfg['n_items']} items per step, 2 steps: the reference's "
                                  "u @ W^T (AutoEnc4Rec.py:230) + double argsort (train_auto.py:234-237) + "
                                  "metrics.*_batch, fp32; users/s is linear in the user count (extrapolated)"}
            cpu = {"value": B_cpu / sec, "unit": "seqs/s", "cores": cores, "kind": "port",
                   "sample": f"the full batch ({B_cpu} sequences) per step, 10 steps, reference ATen op sequence (oracle/ref_path.py)"}
        else:
            U_cpu = 8
            sec = time_cpu(cpu_catalog_fn(CFG5, U_cpu), 2, 1)
            cpu = {"value": U_cpu / sec, "unit": "users/s", "cores": cores, "kind": "port",
                   "sample": f"{U_cpu} users x 10M items per step, 2 steps (u@W^T + double argsort + metrics)"}

    if D.rank == 0:
        if args.workload == "train" and train is not None:
            line = {"metric": "train seqs/sec", "value": train["value"], "unit": "seqs/s", "n_gpus": D.world,
                    "steps": args.steps, "warmup": args.warmup, "ms_per_step": train["ms_per_step"],
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                    "data": "synthetic",
                    "config": train_config(),
                    "roofline": train["roofline"], "stages": train["stages"], "e2e": train["e2e"],
                    "gpu_launches": train["gpu_launches"], "clocks": train["clocks"],
                    "algorithmic_mb_per_step": train["algorithmic_mb_per_step"],
                    "table_backward": train["table_backward"], "launch_mode": train["launch_mode"],
                    "ms_per_step_eager": train["ms_per_step_eager"],
                    "ms_per_step_serialised_stages": train["ms_per_step_serialised_stages"],
                    "sustained": train["sustained"], "dropin_surface": train["dropin_surface"],
                    "dense_batch": train["dense_batch"], "fused_dropout": train["fused_dropout"],
                    "table_adam": train["table_adam"], "e2e_device_negatives": train["e2e_device_negatives"],
                    "e2e_device_batches": train["e2e_device_batches"]}
            if gather is not None:
                line["gather"] = gather
            line["catalog"] = cat
            line["train_topologies"] = topo
            if "cfg5" in cat:       # top-level copies of the strong-scaling metric (the primary line is replicas)
                line["catalog_cfg5_users_per_s"] = cat["cfg5"]["value"]
                line["catalog_cfg5_e2e_users_per_s"] = cat["cfg5"]["e2e"]["value"]
                line["catalog_cfg5_scaling"] = "strong"
        else:
            c = cat.get("cfg5") or cat.get("cfg4")
            if c is None:                   # --only cfg1,cfg3 (a debugging subset): the topology records alone
                emit({"train_topologies": topo})
                if D.dist is not None:
                    D.dist.destroy_process_group()
                return
            line = {"metric": c["metric"], "value": c["value"], "unit": c["unit"], "n_gpus": D.world,
                    "steps": c["steps"], "warmup": c["warmup"], "ms_per_step": c["ms_per_step"],
                    "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
                    "data": "synthetic", "config": c["config"], "roofline": c["roofline"], "e2e": c["e2e"],
                    "gpu_launches": c["gpu_launches"], "clocks": c["clocks"], "catalog": cat}
            if train is not None:
                line["train"] = {k: train[k] for k in ("value", "ms_per_step", "stages", "e2e")}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    if D.dist is not None:
        D.dist.destroy_process_group()


if __name__ == "__main__":
    main()
