#!/usr/bin/env python
"""bench.py — training triples/s of the fused KGE step (BASELINE.json's metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2]

Own arm ("ours"): one step = negative sampling -> fused gather/score/adversarial-loss forward ->
fused atomic-scatter backward -> dense Adam on both tables (gradient zeroing folded in), i.e. the
loop body of mkb/compose/pipeline.py:206-242, on a synthetic batch of the named config.
  * value : whole-job triples/s (positives + negatives scored = B*(1+K) per step per GPU) with the
            step's inputs already in HBM, timed with CUDA events over exactly K steps, max over ranks;
  * e2e   : the same step driven through the public API (ops.fused_adversarial_step + autograd +
            optim.DenseAdam, what compose.Pipeline.learn runs) from PINNED HOST batches, host->device
            copies and the loss read-back inside the timed region;
  * roofline : the dominant kernel (fused backward), algorithmic bytes / CUDA-event time measured
            inside the timed region vs the measured HBM peak in MEASURED_PEAKS.json;
  * cpu_baseline : the UNMODIFIED reference (baseline/_ref, installed by baseline/install_ref.sh) driven
            through its own mkb.compose.Pipeline.learn on this box's host cores on a bounded sample of
            the same workload (rank 0, N=1 only); oracle/torch_port.py if baseline/_ref is absent.
  * reference_on_this_gpu : the same unmodified reference with device='cuda', full batch.

Reference arm (--impl reference): that same reference run, all host threads, bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

# name: (dataset label, model, n_entity, n_relation, n_train, hidden_dim, batch, negatives, gamma)
CONFIGS = {
    "cfg1": ("Wn18rr", "TransE", 40943, 11, 86835, 200, 256, 64, 6.0),
    "cfg2": ("FB15k-237", "RotatE", 14541, 237, 272115, 1000, 1024, 256, 9.0),
    "cfg3": ("FB15k-237", "ComplEx", 14541, 237, 272115, 1000, 1024, 256, 9.0),
    "cfg4": ("Yago3-10", "RotatE", 123182, 37, 1079040, 500, 1024, 256, 24.0),
}
METRIC = "training triples/sec (pos+neg scored)"
UNIT = "triples/s"


def workload_name(cfg):
    ds, model, N, R, T, D, B, K, gamma = CONFIGS[cfg]
    return f"{ds}-shaped synthetic graph, {model} dim={D} batch={B} neg={K} Adversarial+Adam ({cfg})"


def synth_graph(cfg, seed=41):
    """Seeded synthetic triples with the named dataset's entity/relation/train counts."""
    _, _, N, R, T, *_ = CONFIGS[cfg]
    rng = np.random.RandomState(seed)
    tri = np.stack([rng.randint(N, size=T), rng.randint(R, size=T), rng.randint(N, size=T)], 1).astype(np.int64)
    return np.unique(tri, axis=0)


def row_bytes(model, D):
    row_e = 4 * D * (2 if model in ("ComplEx", "RotatE") else 1)
    row_r = 4 * D * (2 if model == "ComplEx" else 1)
    return row_e, row_r


def algorithmic_bytes(cfg):
    """logical gather bytes, every row counted once per use, fp32 tables, int64 ids."""
    _, model, N, R, T, D, B, K, _ = CONFIGS[cfg]
    row_e, row_r = row_bytes(model, D)
    fwd = B * (2 * row_e + row_r) + B * K * row_e + 8 * (3 * B + B * K) + 4 * B + 4 * B * (1 + K) + 4
    bwd = fwd + B * K * row_e + 2 * B * row_e + B * row_r  # re-read (recompute) + atomic row writes
    return fwd, bwd


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Polls NVML for SM clock and throttle reasons while the timed region runs."""

    BAD = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown"}
    NOTE = {0x4: "sw_power_cap"}

    def __init__(self, device_index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = torch.cuda.get_device_properties(device_index).uuid
                self.h = pynvml.nvmlDeviceGetHandleByUUID(f"GPU-{uuid}".encode())
            except Exception:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                idx = int(vis.split(",")[device_index]) if vis else device_index
                self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            # the first NVML queries of a process take milliseconds and serialise with kernel launches in the
            # driver: pay for them here, not inside the timed region (seen as a one-off stall of the first timed
            # steps when 4-8 ranks started polling at t0)
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, repr(e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in {**self.BAD, **self.NOTE}.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def __enter__(self):
        if self.nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's eager-PyTorch CPU path (oracle port)
# ------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "baseline", "_ref")  # the unmodified reference, installed by baseline/install_ref.sh


def _import_reference():
    """The unmodified reference package (raphaelsty/mkb) from the git-ignored baseline/_ref (it travels to the
    GPU box with gpurun).  Returns the module or None when it was never installed."""
    if not os.path.isdir(os.path.join(REF_DIR, "mkb")):
        return None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    try:
        import mkb  # noqa: F401

        return mkb
    except Exception:
        return None


class _Quiet:
    """Pipeline.learn prints its epoch banner to stdout and a tqdm bar to stderr; the bench prints ONE line."""

    def __enter__(self):
        self._o, self._e = sys.stdout, sys.stderr
        sys.stdout = sys.stderr = open(os.devnull, "w")

    def __exit__(self, *a):
        sys.stdout.close()
        sys.stdout, sys.stderr = self._o, self._e


def reference_run(cfg, steps, warmup, budget_s, graph=None, device="cpu", full_batch=False):
    """Times the UNMODIFIED reference through its own public API — mkb.compose.Pipeline.learn over a
    mkb.datasets.Dataset with mkb.models.*, mkb.sampling.NegativeSampling, mkb.losses.Adversarial and
    torch.optim.Adam (compose/pipeline.py:202-244) — on a bounded sample of the config: batches of the first
    B_s positives (all K negatives each), B_s sized so that (steps + warmup) steps fit in ``budget_s`` seconds
    (B_s = B when that fits).  Wall clock around ``learn`` (it ends every step with ``error.item()``)."""
    mkb = _import_reference()
    if mkb is None:
        return None
    from mkb import compose as rcompose, datasets as rdatasets, losses as rlosses, models as rmodels
    from mkb import sampling as rsampling

    ds, mname, N, R, T, D, B, K, gamma = CONFIGS[cfg]
    cores = torch.get_num_threads()
    graph = synth_graph(cfg) if graph is None else graph
    tri = [tuple(r) for r in graph.tolist()]
    ents, rels = {i: i for i in range(N)}, {i: i for i in range(R)}
    torch.manual_seed(42)
    model = getattr(rmodels, mname)(hidden_dim=D, entities=ents, relations=rels, gamma=gamma).to(device)
    sampler = rsampling.NegativeSampling(size=K, train_triples=tri, entities=ents, relations=rels, seed=42)
    opt = torch.optim.Adam(filter(lambda p: p.requires_grad, model.parameters()), lr=5e-5)
    loss = rlosses.Adversarial(alpha=0.5)
    pick = np.random.RandomState(1)

    def learn(bs, n_steps):
        pairs = max(1, (n_steps + 1) // 2)  # one epoch of Dataset = a head-batch and a tail-batch per bs triples
        rows = [tri[i] for i in pick.choice(len(tri), bs * pairs, replace=False)]
        data = rdatasets.Dataset(train=rows, entities=ents, relations=rels, batch_size=bs, shuffle=False,
                                 seed=None, num_workers=0)
        pipe = rcompose.Pipeline(epochs=1, device=device)
        with _Quiet():
            if device != "cpu":
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            pipe.learn(model=model, dataset=data, sampling=sampler, optimizer=opt, loss=loss)
            if device != "cpu":
                torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        return dt, 2 * pairs

    if full_batch:
        bs = B
    else:
        learn(2, 2)  # first touch of the allocator / thread pool
        dt, n = learn(4, 2)
        per_row = dt / (4 * n)
        bs = int(max(2, min(B, budget_s / max(steps + warmup, 1) / max(per_row, 1e-6))))
    if warmup:
        learn(bs, warmup)
    t, n = learn(bs, steps)
    return {
        "value": bs * (1 + K) * n / t, "unit": UNIT, "cores": cores, "kind": "reference",
        "sample": f"{n} steps x {'all' if bs == B else 'first'} {bs} of {B} positives per batch, all {K} negatives "
                  f"each, D={D}: unmodified mkb {mkb.__version__} (baseline/_ref) driven through "
                  f"mkb.compose.Pipeline.learn — Dataset, NegativeSampling.generate, {mname}.forward x2, Adversarial, "
                  f"backward, torch.optim.Adam — torch {torch.__version__} on {device}, {cores} host threads",
        "ms_per_step": 1e3 * t / n, "batch": bs, "steps_run": n,
    }


def port_run(cfg, steps, warmup, budget_s, graph=None):
    """Fallback when baseline/_ref is absent: oracle/torch_port.CpuTrainer (the reference's ATen operator
    sequence, bit-exact with it on the golden vectors) on the same bounded sample."""
    from oracle import torch_port as tp

    ds, model, N, R, T, D, B, K, gamma = CONFIGS[cfg]
    cores = torch.get_num_threads()
    graph = synth_graph(cfg) if graph is None else graph
    tri = [tuple(r) for r in graph.tolist()]
    th, tt = tp.true_sets(tri)
    trainer = tp.CpuTrainer(model, N, R, D, gamma, lr=5e-5, seed=42)
    rng = np.random.RandomState(42)
    pick = np.random.RandomState(1)

    def batch(bs):
        idx = pick.choice(len(tri), bs, replace=False)
        return torch.tensor([tri[i] for i in idx]), torch.full((bs,), 0.3)

    def one(bs, mode):
        s, w = batch(bs)
        t0 = time.perf_counter()
        neg = tp.generate_negatives(rng, s, mode, th, tt, N, K)
        trainer.step(s, w, mode, neg)
        return time.perf_counter() - t0

    one(2, "tail-batch")  # first-touch of the allocator / thread pool
    t_probe = one(4, "head-batch") / 4  # seconds per positive row
    bs = int(max(2, min(B, budget_s / max(steps + warmup, 1) / max(t_probe, 1e-6))))
    for i in range(warmup):
        one(bs, "head-batch" if i % 2 == 0 else "tail-batch")
    t = 0.0
    for i in range(steps):
        t += one(bs, "head-batch" if i % 2 == 0 else "tail-batch")
    value = bs * (1 + K) * steps / t
    return {
        "value": value, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": f"{steps} steps x first {bs} of {B} positives per batch, all {K} negatives each, D={D}: "
                  f"reference sampler + 2 forwards + loss + backward + dense Adam (oracle/torch_port.py, "
                  f"torch {torch.__version__} CPU, {cores} threads; baseline/_ref not installed)",
        "ms_per_step": 1e3 * t / steps, "batch": bs, "steps_run": steps,
    }


def cpu_reference_run(cfg, steps, warmup, budget_s, graph=None):
    """The reference arm / CPU baseline: the unmodified reference when baseline/_ref is installed, else the port."""
    out = reference_run(cfg, steps, warmup, budget_s, graph=graph)
    return out if out is not None else port_run(cfg, steps, warmup, budget_s, graph=graph)


def eager_gpu_run(cfg, graph, dev, steps=5, warmup=2):
    """The reference's eager-PyTorch operator sequence (oracle/torch_port.py) on THIS GPU, full batch — what a
    user gets from the reference with device='cuda' (SURVEY §8(d)).  Its sampler stays on the host as in the
    reference (Python loop over the positives), so two numbers are reported: the whole step and the device
    part alone.  Part of the baseline leg: never the thing shipped."""
    from oracle import torch_port as tp

    ds, model, N, R, T, D, B, K, gamma = CONFIGS[cfg]
    tri = [tuple(r) for r in graph.tolist()]
    th, tt = tp.true_sets(tri)
    trainer = tp.CpuTrainer(model, N, R, D, gamma, lr=5e-5, seed=42, device=dev)
    rng, pick = np.random.RandomState(42), np.random.RandomState(1)
    t_all = t_dev = 0.0
    for i in range(warmup + steps):
        idx = pick.choice(len(tri), B, replace=False)
        s, w = torch.tensor([tri[j] for j in idx]), torch.full((B,), 0.3)
        mode = "head-batch" if i % 2 == 0 else "tail-batch"
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        neg = tp.generate_negatives(rng, s, mode, th, tt, N, K)
        s_d, w_d, neg_d = s.to(dev), w.to(dev), neg.to(dev)
        t1 = time.perf_counter()
        trainer.step(s_d, w_d, mode, neg_d)  # ends with error.item(): synchronises
        t2 = time.perf_counter()
        if i >= warmup:
            t_all += t2 - t0
            t_dev += t2 - t1
    return {"value": B * (1 + K) * steps / t_all, "device_part_only": B * (1 + K) * steps / t_dev, "unit": UNIT,
            "steps": steps, "ms_per_step": 1e3 * t_all / steps, "device_ms_per_step": 1e3 * t_dev / steps,
            "what": "oracle/torch_port.py (the reference's ATen sequence + torch.optim.Adam) on cuda:0, full batch; "
                    "host-side reference sampler included in `value`, excluded in `device_part_only`"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = args.config
    base = cpu_reference_run(cfg, args.steps, args.warmup, budget_s=args.cpu_budget)
    ds, model, N, R, T, D, B, K, gamma = CONFIGS[cfg]
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(cfg), "device": "cpu"},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0



def load_peaks():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        peaks = {}
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    return hbm, src


def ncu_entry(cfg, role):
    """Per-launch DRAM bytes and unit throughputs of this config's kernel from the committed ncu capture
    (profiles/ncu_traffic.json, written by scripts/ncu_summary.py from an `ncu --set full` run of this bench)."""
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[cfg]
        return tr[role], tr.get("source")
    except Exception:
        return None, None


def kernel_roofline(cfg, role, name, algo_bytes, ms, hbm, peak_src):
    """achieved = ALGORITHMIC bytes / CUDA-event time measured live (every gathered row counted once
    per use, so rows served by the L2 count too and `frac` can exceed 1 for an L2-resident table); frac_dram =
    the DRAM bytes ncu counted for one launch of this kernel / the same live time — the real HBM fraction; `bound`
    is the unit ncu saw closest to its peak, not an assertion."""
    ent, src = ncu_entry(cfg, role)
    gbs = algo_bytes / (ms * 1e-3) / 1e9
    out = {"kernel": name, "bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
           "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes, "avg_launch_ms": ms,
           "frac_logical": gbs / hbm, "frac_dram": None}
    if ent:
        traffic = int(ent["dram_read_bytes"] + ent["dram_write_bytes"])
        lim = {"L1/TEX": "l1tex", "L2": "l2", "DRAM": "hbm", "SM": "sm"}.get(ent.get("limiter"), "hbm")
        out.update({
            "bound": lim, "traffic": traffic, "frac_dram": traffic / (ms * 1e-3) / 1e9 / hbm,
            "limiter": {"unit": ent.get("limiter"), "l1tex_pct": ent.get("l1tex_pct"), "l2_pct": ent.get("l2_pct"),
                        "dram_pct_of_nominal": ent.get("dram_pct"), "sm_pct": ent.get("sm_pct"),
                        "l2_hit_pct": ent.get("l2_hit_pct"), "ncu_duration_us": ent.get("duration_us"),
                        "source": f"profiles/ncu_traffic.json <- {src}"},
        })
        if lim != "hbm":
            out["note"] = ("table is L2-resident at this config: the algorithmic (logical gather) bytes are mostly L2 hits, "
                           "so frac > 1 is not an HBM fraction; frac_dram is.  See roofline_hbm_config for the config "
                           "whose table exceeds the L2")
    return out


def hbm_config_run(dev, hbm, peak_src, cfg="cfg4", steps=40, warmup=5):
    """A short device-resident run of config 4 (Yago3-10 shape, RotatE dim 500: entity table 493 MB >> L2) so that
    the bench line carries one set of fractions that are true HBM fractions: per kernel, algorithmic bytes / live
    CUDA-event time and ncu DRAM bytes / the same time, both against the measured copy bandwidth."""
    from mkb_b200 import models, sampling
    from mkb_b200.compose import DeviceTrainer
    from mkb_b200.datasets.dataset import subsampling_weights

    ds, mname, N, R, T, D, B, K, gamma = CONFIGS[cfg]
    graph = synth_graph(cfg)
    torch.manual_seed(42)
    model = getattr(models, mname)(hidden_dim=D, entities={i: i for i in range(N)},
                                   relations={i: i for i in range(R)}, gamma=gamma).to(dev)
    ns = sampling.NegativeSampling(size=K, train_triples=graph, entities=range(N), relations=range(R), seed=42, device=dev)
    trainer = DeviceTrainer(model, ns, lr=5e-5, max_batch=B)
    w_all = subsampling_weights(graph)
    order = np.random.RandomState(43).permutation(len(graph))[: (steps + warmup) * B].reshape(steps + warmup, B)
    samples = torch.from_numpy(graph[order]).to(dev)
    weights = w_all[torch.from_numpy(order)].to(dev)
    modes = ["head-batch" if i % 2 == 0 else "tail-batch" for i in range(steps + warmup)]
    for i in range(warmup):
        trainer.step(samples[i], weights[i], modes[i])
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0.record()
    for i in range(steps):
        trainer.hooks = ev[i]
        trainer.step(samples[warmup + i], weights[warmup + i], modes[warmup + i])
    t1.record()
    torch.cuda.synchronize()
    trainer.hooks = None
    ms = t0.elapsed_time(t1) / steps
    # per-step durations (forward-start to forward-start) as this rank saw them: a transient shows up here
    per_step = [ev[i][0].elapsed_time(ev[i + 1][0]) for i in range(min(steps, 41) - 1)]
    fwd_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    bwd_ms = float(np.mean([e[2].elapsed_time(e[3]) for e in ev]))
    adam_ms = float(np.mean([e[4].elapsed_time(e[5]) for e in ev]))
    fwd_b, bwd_b = algorithmic_bytes(cfg)
    row_e, row_r = row_bytes(mname, D)
    return {
        "workload": workload_name(cfg), "steps": steps, "warmup": warmup, "ms_per_step": ms,
        "value": B * (1 + K) / (ms * 1e-3), "unit": UNIT,
        "fwd": kernel_roofline(cfg, "fwd", "score_neg_kernel<FUSED> (K2)", fwd_b, fwd_ms, hbm, peak_src),
        "bwd": kernel_roofline(cfg, "bwd", "score_bwd_kernel (K3)", bwd_b, bwd_ms, hbm, peak_src),
        "adam": kernel_roofline(cfg, "adam", "adam_kernel x2", 7 * (N * row_e + R * row_r), adam_ms, hbm, peak_src),
    }

# ------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    from mkb_b200 import _native, models, ops, optim, sampling
    from mkb_b200.compose import DeviceTrainer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = world > 1
    if dist:
        import datetime

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # a stuck collective aborts the job after 3 minutes instead of hanging the box
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "1")
        torch.distributed.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    if world != args.gpus and rank == 0:
        print(f"# warning: --gpus {args.gpus} but WORLD_SIZE={world}; reporting n_gpus={world}", file=sys.stderr)

    cfg = args.config
    ds, mname, N, R, T, D, B, K, gamma = CONFIGS[cfg]
    steps, warmup = args.steps, max(args.warmup, 3)
    graph = synth_graph(cfg)

    torch.manual_seed(42)  # identical replicas on every rank
    model = getattr(models, mname)(hidden_dim=D, entities={i: i for i in range(N)},
                                   relations={i: i for i in range(R)}, gamma=gamma).to(dev)
    ns = sampling.NegativeSampling(size=K, train_triples=graph, entities=range(N), relations=range(R),
                                   seed=42 + rank, device=dev, pool=args.pool)
    topts = {"mode": args.mode, "backward": args.backward, "pooled_gemm": args.pooled_gemm,
             "packed_records": args.packed_records, "merged_backward": args.merged_backward,
             "handshake": args.handshake}
    if args.virtual_shards:
        topts = {"mode": "rowshard", "virtual_shards": args.virtual_shards}
    trainer = DeviceTrainer(model, ns, lr=5e-5, max_batch=B, distributed=dist, **topts)

    # this rank's batches: disjoint slices of a seeded permutation of the training triples
    from mkb_b200.datasets.dataset import subsampling_weights

    weights_all = subsampling_weights(graph)
    perm = np.random.RandomState(43).permutation(len(graph))
    n_batches = steps + warmup
    need = n_batches * B * world
    reps = -(-need // len(perm))
    order_all = np.tile(perm, reps)[:need].reshape(n_batches, world, B)  # global batch g = order_all[g]
    order = order_all[:, rank, :]
    host_samples = torch.from_numpy(graph[order]).pin_memory()  # [n_batches, B, 3]
    host_weights = weights_all[torch.from_numpy(order)].pin_memory()
    dev_samples = host_samples.to(dev)
    dev_weights = host_weights.to(dev)
    modes = ["head-batch" if i % 2 == 0 else "tail-batch" for i in range(n_batches)]

    def barrier():
        if dist:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm (value) ----------------
    # W warm-up steps, the last three of them enqueued AFTER the barrier and right before t0: the host-side barrier,
    # the status check and the event set-up leave the GPUs idle for milliseconds (clocks drop, and ranks leave the
    # barrier up to a few ms apart), which a short timed region would otherwise absorb as "slow first steps" —
    # seen as value > e2e per step at 4 / 8 GPUs.  The steps themselves keep the ranks in lockstep (flag handshakes).
    late = min(3, warmup)
    clock_sampler = ClockSampler(local)  # NVML handle + first queries now, polling thread only inside the region
    for i in range(warmup - late):
        trainer.step(dev_samples[i], dev_weights[i], modes[i])
    ns.check_status(dev)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for e in (t0, t1, *[x for row in ev for x in row]):
        e.record()  # create the CUDA events now, not inside the timed region
    barrier()
    for i in range(warmup - late, warmup):
        trainer.step(dev_samples[i], dev_weights[i], modes[i])
    if dist:
        torch.distributed.all_reduce(torch.zeros(1, device=dev))  # lines the DEVICE timelines up right before t0
    launches0 = _native.launches
    with clock_sampler as clocks:
        t0.record()
        for i in range(steps):
            trainer.hooks = ev[i]
            trainer.step(dev_samples[warmup + i], dev_weights[warmup + i], modes[warmup + i])
        t1.record()
        barrier()
    trainer.hooks = None
    launches = _native.launches - launches0
    ms_total = t0.elapsed_time(t1)
    # per-step durations (forward-start to forward-start) as this rank saw them: a transient shows up here
    per_step = [ev[i][0].elapsed_time(ev[i + 1][0]) for i in range(min(steps, 41) - 1)]
    fwd_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    bwd_ms = float(np.mean([e[2].elapsed_time(e[3]) for e in ev]))
    # the single-GPU / all-reduce / column-parallel flows also bracket the Adam launches
    adam_ms = (float(np.mean([e[4].elapsed_time(e[5]) for e in ev]))
               if trainer.mode in ("single", "allreduce", "colpar", "colshard") and not trainer.pooled_gemm else 0.0)
    final_loss = trainer.loss()
    ns.check_status(dev)

    # ---------------- end-to-end arm: the public API a user calls, host batches ----------------
    # compose.Pipeline.learn over a datasets.Dataset that lives in (pinned) HOST memory: every step copies
    # its sample/weight to the device, runs the step, and reads the loss back to the host.
    from mkb_b200 import compose, datasets, losses

    torch.manual_seed(42)
    model2 = getattr(models, mname)(hidden_dim=D, entities={i: i for i in range(N)},
                                    relations={i: i for i in range(R)}, gamma=gamma).to(dev)
    opt = optim.DenseAdam(filter(lambda p: p.requires_grad, model2.parameters()), lr=5e-5)
    ents, rels = {i: i for i in range(N)}, {i: i for i in range(R)}

    def host_dataset(batches):  # an epoch of Dataset = one head-batch + one tail-batch per GLOBAL batch of B * world
        # triples; under torch.distributed Pipeline.learn gives rank r the r-th block of B rows of every batch
        return datasets.Dataset(train=graph[batches.reshape(-1)], entities=ents, relations=rels, batch_size=B * world,
                                shuffle=False, seed=None, pin_memory=True)

    half = (steps + 1) // 2
    ds_warm = host_dataset(order_all[: max(warmup // 2, 2)])
    ds_time = host_dataset(order_all[warmup: warmup + half])
    e2e_steps = 2 * half
    pipe = compose.Pipeline(epochs=1, device=dev, trainer_options=topts)
    sys.stderr, _err = open(os.devnull, "w"), sys.stderr  # tqdm's bar
    e2e_error = None
    e2e_passes = []
    try:
        pipe.learn(model=model2, dataset=ds_warm, sampling=ns, optimizer=opt, loss=losses.Adversarial(0.5))
        for _ in range(3):  # three full passes over the same K-step dataset; the MEDIAN is reported, all are listed
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            if dist:
                torch.distributed.all_reduce(torch.zeros(1, device=dev))
            e0.record()
            pipe.learn(model=model2, dataset=ds_time, sampling=ns, optimizer=opt, loss=losses.Adversarial(0.5))
            e1.record()
            barrier()
            e2e_passes.append(e0.elapsed_time(e1) / e2e_steps)
    except Exception as e:  # single GPU only: the device-resident line is still reported, e2e carries the error
        if dist:
            raise  # a rank that leaves the collectives early would hang the others
        e2e_error = f"{type(e).__name__}: {e}"[:300]
        e2e_passes = [float("nan")] * 3
    finally:
        sys.stderr = _err
    e2e_ms = float(np.median(e2e_passes)) * steps  # per-step time of the median pass x `steps`
    e2e_loss = pipe.metric_loss.get()

    # ---------------- reduce over ranks ----------------
    e2e_passes_t = torch.tensor(e2e_passes, dtype=torch.float64, device=dev)
    if dist:
        torch.distributed.all_reduce(e2e_passes_t, op=torch.distributed.ReduceOp.MAX)
        e2e_ms = float(e2e_passes_t.median().item()) * steps  # median over passes of the max over ranks
    t = torch.tensor([ms_total, e2e_ms, fwd_ms, bwd_ms, adam_ms], dtype=torch.float64, device=dev)
    if dist:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_total, e2e_ms, fwd_ms, bwd_ms, adam_ms = t.tolist()
    triples_per_step = B * (1 + K) * world
    value = triples_per_step * steps / (ms_total * 1e-3)
    e2e_value = triples_per_step * steps / (e2e_ms * 1e-3)

    if rank == 0:
        hbm, peak_src = load_peaks()
        fwd_b, bwd_b = algorithmic_bytes(cfg)
        roof = kernel_roofline(cfg, "bwd", "score_bwd_kernel (fused backward, K3)" if trainer.backward == "scatter" else
                               "by-entity backward incl. the entity table's Adam (byent.cu; bytes still counted as K3's)",
                               bwd_b, bwd_ms, hbm, peak_src)
        roof["share_of_step"] = bwd_ms / (ms_total / steps)
        roof["also"] = kernel_roofline(cfg, "fwd", "score_neg_kernel<FUSED> (fused forward, K2)", fwd_b, fwd_ms, hbm, peak_src)
        if adam_ms and not dist:
            roof["adam"] = kernel_roofline(cfg, "adam", "adam_kernel x2 (dense Adam + gradient zeroing, both tables)",
                                           7 * (N * row_bytes(mname, D)[0] + R * row_bytes(mname, D)[1]),
                                           adam_ms, hbm, peak_src)
        # where the step goes (CUDA events, max over ranks): what is neither forward, backward nor optimizer is the
        # sampler plus, on several GPUs, the record push and the flag handshakes
        roof["per_step_ms_first_40"] = [round(x, 3) for x in per_step]
        roof["step_breakdown_ms"] = {"forward": fwd_ms, "backward": bwd_ms,
                                     "optimizer" + ("+table all-gather over NVLink" if dist else ""): adam_ms or None,
                                     "sampler, handshakes, launch gaps": ms_total / steps - fwd_ms - bwd_ms - (adam_ms or 0.0)}
        row_e, row_r = row_bytes(mname, D)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": workload_name(cfg), "global_batch": B * world, "negatives": K,
                "sampler": ("independent on-device Philox draws" if args.pool == "independent" else
                            "reference shared pool of 2K candidates per batch" +
                            (", scored as one tensor-core GEMM" if trainer.pooled_gemm else "")),
                "parallelism": (
                    "single GPU" if trainer.mode != "rowshard" else
                    f"single GPU, entity table in {trainer.n_shards} block-cyclic row shards (all local)")
                if not dist else (
                    f"dp{world}, replicated tables; batch-parallel forward, column-parallel backward over the "
                    f"global batch (step records exchanged by {'NVLink peer stores + flags, no NCCL on the step' if trainer.handshake == 'peer' else 'NCCL all-gather'}), "
                    f"fused Adam + all-gather through NVLink peer stores"
                    if trainer.mode == "colpar" else
                    f"dp{world}, entity table row-sharded block-cyclically over the GPUs (1/{world} of table, gradient "
                    f"and Adam state each); remote rows gathered by P2P loads, row gradients added into the owner's "
                    f"shard by system-scope vector REDs over NVLink; relation gradient all-reduced"
                    if trainer.mode == "rowshard" else
                    f"tp{world} over the hidden dim: each GPU holds 1/{world} of the COLUMNS of both tables (and of gradient "
                    f"and Adam state), scores the global batch over its columns, one all-reduce of the partial scores; "
                    f"no table or gradient row crosses NVLink"
                    if trainer.mode == "colshard" else
                    f"dp{world}, replicated tables; all-reduce of 3 loss sums + dense gradients" + (
                        f" [{trainer.mode_note}]" if trainer.mode_note else "")),
                "l2": "working set per step (tables+grads+Adam moments = "
                      f"{4 * (N * row_e + R * row_r) / 1e6:.0f} MB) exceeds the 126 MB L2; no explicit flush",
                "step": ("filter_pool + pooled_dot_fwd (gather P rows, GEMM S=Q.Pool^T, adv terms) + pooled_dot_bwd "
                         "(2 GEMMs, chain, row scatter) + adam(entity) + adam(relation)" if trainer.pooled_gemm else
                         "sample_negatives + fused_fwd + fused_bwd + adam(entity) + adam(relation)"
                         if trainer.backward == "scatter" else
                         "sample_negatives + fused_fwd + by-entity backward (CSR build, queries, dq pass, per-entity "
                         "gradient + Adam in place) + adam(relation)")
                        if trainer.mode in ("single", "allreduce") else
                        "sample_negatives + fused_fwd_sharded + fused_bwd_sharded + adam(own shard(s)) + adam(relation)"
                        if trainer.mode == "rowshard" else
                        "sample_negatives + push batch blocks + 2 x score_fwd(sub-table, global batch) + all-reduce(partial "
                        "scores) + adv_loss fwd/bwd + fused_bwd(sub-table) + 2 x adam(own columns)"
                        if trainer.mode == "colshard" else
                        (f"wait(slices) + sample_negatives + fused_fwd + peer_copy(record) + signal + wait(records) + "
                         f"1 multi-record fused_bwd_chunk + 2 x adam_slice_bcast + signal" if trainer.handshake == "peer" else
                         f"sample_negatives + fused_fwd + all-gather(step records) + {world} x fused_bwd_chunk + "
                         "2 x adam_slice_bcast"),
                "final_loss": final_loss,
            },
            "roofline": roof,
            "e2e": {"value": None if e2e_error else e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(B * 3 * 8 + B * 4) * world,
                    "d2h_bytes_per_step": 16, "ms_per_step": None if e2e_error else e2e_ms / steps,
                    "api": "compose.Pipeline.learn(models.*, datasets.Dataset(host, pinned), sampling.NegativeSampling, "
                           "optim.DenseAdam, losses.Adversarial): per step H2D sample+weight, D2H loss sums",
                    "rolling_loss": e2e_loss, "passes_ms_per_step": None if e2e_error else e2e_passes_t.tolist(),
                    "note": "median of three timed passes of Pipeline.learn over the same K-step dataset",
                    **({"error": e2e_error} if e2e_error else {})},
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                base = cpu_reference_run(cfg, steps=3, warmup=1, budget_s=20.0, graph=graph)
                line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as e:  # the GPU numbers above must not be lost to a host-side problem
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                        "sample": f"failed: {type(e).__name__}: {e}"[:300]}
            try:  # informative only: must never cost the bench line
                del trainer, model2, opt, pipe
                torch.cuda.empty_cache()
                line["cpu_baseline"]["same_operators_on_this_gpu"] = eager_gpu_run(cfg, graph, dev)
            except Exception as e:  # e.g. out of memory for the eager path's [B,K,2D] temporaries
                line["cpu_baseline"]["same_operators_on_this_gpu"] = {"error": f"{type(e).__name__}: {e}"[:200]}
            try:  # the unmodified reference with device='cuda', full batch, through its own Pipeline.learn
                torch.cuda.empty_cache()
                r = reference_run(cfg, steps=6, warmup=2, budget_s=0, graph=graph, device=str(dev), full_batch=True)
                if r is not None:
                    line["reference_on_this_gpu"] = {
                        **{k: r[k] for k in ("value", "unit", "ms_per_step", "sample")},
                        "ours_over_it": {"value": value / r["value"],
                                         "e2e": None if e2e_error else e2e_value / r["value"]}}
            except Exception as e:
                line["reference_on_this_gpu"] = {"error": f"{type(e).__name__}: {e}"[:200]}
        if world == 1 and cfg != "cfg4" and not args.no_hbm_config:
            try:  # the config whose table (493 MB) exceeds the 126 MB L2: the fractions there ARE HBM fractions
                torch.cuda.empty_cache()
                line["roofline_hbm_config"] = hbm_config_run(dev, hbm, peak_src)
            except Exception as e:
                line["roofline_hbm_config"] = {"error": f"{type(e).__name__}: {e}"[:200]}
        print(json.dumps(line))
    if dist:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hbm-config", action="store_true",
                    help="skip the short config-4 run that adds roofline_hbm_config to the line")
    ap.add_argument("--cpu-budget", type=float, default=150.0,
                    help="--impl reference: seconds of CPU work the bounded sample is sized for (all steps together)")
    ap.add_argument("--mode", default=None, choices=["colpar", "allreduce", "rowshard", "colshard"],
                    help="multi-GPU scheme of DeviceTrainer (default colpar: column-parallel backward + fused "
                         "Adam/all-gather over NVLink peer memory; allreduce: dense gradient all-reduce; rowshard: "
                         "entity table row-sharded over the GPUs, P2P row gathers + remote gradient REDs; colshard: tables "
                         "sharded by hidden-dim columns, partial scores all-reduced, nothing else moves)")
    ap.add_argument("--backward", default="scatter", choices=["scatter", "by_entity"],
                    help="single-GPU backward: scatter = K3 vector REDs + dense Adam (default, the measured path); "
                         "by_entity = atomics-free per-entity backward with Adam fused in (csrc/byent.cu)")
    ap.add_argument("--handshake", default=None, choices=["peer", "nccl"],
                    help="colpar only: peer = NVLink peer-memory flags + record push, no NCCL on the step (default); "
                         "nccl = round 1's three collectives per step")
    ap.add_argument("--packed-records", action="store_true",
                    help="colpar only: loss sums ride in the all-gathered step records, ONE multi-record backward launch "
                         "(98 %% efficiency on 2 GPUs; its 4/8-GPU runs timed out in round 1 and await diagnosis)")
    ap.add_argument("--merged-backward", action="store_true",
                    help="colpar only: one backward launch over all G gathered records, collectives unchanged")
    ap.add_argument("--pool", default="independent", choices=["independent", "reference"],
                    help="negative sampler: independent on-device draws (default, the headline) or the reference's "
                         "shared pool of 2K candidates per batch")
    ap.add_argument("--pooled-gemm", action="store_true",
                    help="with --pool reference and DistMult/ComplEx (cfg3): score the batch as S = Q.Pool^T on the "
                         "tensor cores (csrc/pooled.cu) instead of B*K row gathers")
    ap.add_argument("--virtual-shards", type=int, default=0,
                    help="single GPU only: run the row-sharded kernels with this many local shards")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
