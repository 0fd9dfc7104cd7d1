#!/usr/bin/env python
"""Headline benchmark: GNN-head episodes/sec (forward + backward), BASELINE.json's metric.

A "step" is one pass of the hot path over one synthetic episode: the n_query graphs of a
5-way 20-shot episode (B=16 graphs, N=105 nodes, F=133 features -- the configuration the
metric's target is quoted on) go through ``GNN_nl`` forward, a cross-entropy
on the query nodes, and the full backward (input + all 64 parameter gradients).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--shape 5w5s|5w20s|5w50c] [--precision auto|fp32|tf32]
                    [--mode train|eval] [--dp-bytes MB] [--no-share] [--no-graph]

Multi-GPU: launched under torchrun, one rank per GPU; every rank runs its own episodes
(weak scaling, episodes are independent) and the gradients are averaged with one NCCL
all-reduce per step (the episode-parallel meta-training path).  ``--dp-bytes 21.2``
adds a flat buffer standing for the backbone + fc gradients (5 307 706 floats in total with the GNN's)
to that all-reduce.  ``--mode eval``: the episode-sharded evaluation path (finetune.py:634-682): 600
forward-only episodes of B=15 graphs, episode e on rank e % G, one gather of the accuracies at the end.

``--impl reference`` times the reference's OWN head code (oracle/_ref: methods/gnn.py + gnnnet.py taken
verbatim from the reference by oracle/make_ref.py; the CPU port oracle/gnn_oracle.py when that is absent)
on the host cores, same workload, same metric.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

SHAPES = {
    # name: (n_way, n_shot, n_query, compress)
    "5w5s": (5, 5, 16, False),
    "5w20s": (5, 20, 16, False),
    "5w50c": (5, 50, 16, True),
}
NF = 96
METRIC = "gnn_head_episodes_per_sec_fwd_bwd"
BACKBONE_FC_FLOATS = 4905792 + 65920      # ResNet10 + fc of GnnNet; the GNN adds 335 994


def head_flops(bsz, n, backward=True):
    per_pair = sum(2 * (f * 192 + 192 * 192 + 192 * 96 + 96 * 96 + 96) for f in (133, 181, 229))
    return per_pair * (3 if backward else 1) * bsz * n * n


def gemm_traffic(shape, share, prec):
    """DRAM bytes per edge-MLP GEMM launch from the committed `ncu` capture of this workload
    (profiles/r02_gemm_traffic.json, else round 1's; written by tools/ncu_traffic.py), or None."""
    for name in ("r02_gemm_traffic.json", "r01_gemm_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        try:
            rec = json.load(open(path))
        except (OSError, ValueError):
            continue
        if rec.get("config") == {"shape": shape, "share_support": bool(share), "precision": prec}:
            return rec.get("bytes_per_launch")   # the capture describes ONE configuration
    return None


def hbm_view(shape, share, prec, ms_per_step, hbm_gbs):
    """The step seen from the memory side: DRAM bytes of one step (committed `ncu` capture, every kernel profiled
    alone) over the measured step time, against the measured copy bandwidth.  None without a matching capture."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "r02_gemm_traffic.json")))
    except (OSError, ValueError):
        return None
    step = rec.get("step")
    if not step or rec.get("config") != {"shape": shape, "share_support": bool(share), "precision": prec}:
        return None
    total = step["dram_read_bytes"] + step["dram_write_bytes"]
    gbs = total / (ms_per_step / 1e3) / 1e9
    return {"dram_bytes_per_step": total, "achieved_gbs_on_step": gbs, "hbm_peak_gbs": hbm_gbs,
            "frac_of_hbm_peak_on_step": gbs / hbm_gbs if hbm_gbs else None,
            "note": "upper bound (cold-L2, serialised capture); the step is bound by neither roof alone: see "
                    "DESIGN.md 4.4"}


def shape_dims(shape, n_query=None):
    n_way, n_shot, nq, compress = SHAPES[shape]
    k = round(n_shot / 2) if compress else n_shot
    return n_way, n_shot, (n_query or nq), compress, n_way * (k + 1)


def workload(shape, n_query=None, backward=True):
    """ONE sentence for both arms (the driver compares the arms' config)."""
    n_way, n_shot, nq, compress, n = shape_dims(shape, n_query)
    what = "fwd+bwd" if backward else "fwd"
    tail = "CE on the query nodes; input + 64 parameter gradients" if backward else "scores of the query nodes"
    return (f"GnnNet head {what}, {shape}: GNN_nl on B={nq} graphs x N={n} nodes, F=133, nf=96, n_way={n_way}; " + tail)


def synthetic_features(shape, seed, device="cpu", n_query=None):
    """Backbone features of one episode: [n_way, n_shot+n_query, 512] ~ N(0,1) (synthetic; the
    ResNet10 backbone is outside the hot path and stays on cuDNN)."""
    n_way, n_shot, nq, _, _ = shape_dims(shape, n_query)
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n_way, n_shot + nq, 512, generator=g).to(device)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "source": "MEASURED_PEAKS.json"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def measure_tf32_peak(seconds=2.0, n=8192):
    """Dense TF32 matmul peak of THIS GPU, measured the way MEASURED_PEAKS.json measured bf16 (SURVEY.md 8d:
    "measure with a TF32 8192^3 matmul"): torch.matmul on fp32 inputs with allow_tf32 (cuBLAS), best of 10
    (burst) and back to back for `seconds` (sustained)."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(n, n, device="cuda")
        b = torch.randn(n, n, device="cuda")
        c = torch.empty(n, n, device="cuda")
        fl = 2.0 * n ** 3
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b, out=c)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if seconds <= 0:
            return {"burst": fl / (best * 1e-3) / 1e12, "sustained": None}
        iters, t0 = 0, time.perf_counter()
        e0.record()
        while time.perf_counter() - t0 < seconds:
            for _ in range(20):
                torch.matmul(a, b, out=c)
            iters += 20
            torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        return {"burst": fl / (best * 1e-3) / 1e12, "sustained": fl * iters / (e0.elapsed_time(e1) * 1e-3) / 1e12}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the run (started before the warm-up so that short
    timed regions still see samples; stopped after the end-to-end region)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        rows, reasons = [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                rows.append((float(f[1]), float(f[2]), float(f[3])))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        pmax = max(r[2] for r in rows)
        load = [r for r in rows if r[2] >= 0.6 * pmax] or rows     # samples taken while the GPU was working
        return {"sm_mhz": statistics.median(r[0] for r in load), "sm_max_mhz": max(r[1] for r in rows),
                "reasons": sorted(reasons), "samples": len(rows), "samples_under_load": len(load),
                "power_w_max": pmax}


# ------------------------------------------------------------------------------------------
# the reference's own head code on the CPU (oracle/_ref), or the port when it is absent
# ------------------------------------------------------------------------------------------

class ReferenceHead:
    """fc -> graphs -> forward_gnn -> CE of the reference's GnnNet / gnnnet_copy.GnnNet on features, driven
    through the reference's own modules (the few glue lines of gnnnet.py:79-83 / gnnnet_copy.py:67-72 are
    restated because set_forward(is_feature=True) hard-codes 15 queries, gnnnet.py:73)."""

    def __init__(self, shape, device="cpu", variant="reference"):
        from oracle import ref_loader as R
        self.ns = R.load(variant)
        n_way, n_shot, n_query, compress, n = shape_dims(shape)
        mod = self.ns.gnnnet_copy if compress else self.ns.gnnnet
        torch.manual_seed(0)
        self.m = mod.GnnNet(self.ns.backbone.ResNet10, n_way, n_shot)
        self.m.n_query = n_query
        self.compress = compress
        self.n_shot = n_shot
        self.device = torch.device(device)
        self.m.fc.to(self.device)
        self.m.gnn.to(self.device)
        self.m.support_label = self.m.support_label.to(self.device)
        self.y = torch.from_numpy(np.repeat(range(n_way), n_query)).to(self.device)
        self.params = list(self.m.fc.parameters()) + list(self.m.gnn.parameters())

    def loss(self, feat):
        m = self.m
        z = m.fc(feat.reshape(-1, feat.size(-1)))
        z = z.view(m.n_way, -1, z.size(1))
        ns = m.n_support                                  # already halved by gnnnet_copy (:34)
        if self.compress:
            sup = z[:, :2 * ns].reshape(m.n_way, 2, ns, -1).mean(dim=1)
            q0 = 2 * ns
        else:
            sup, q0 = z[:, :ns], ns
        zs = [torch.cat([sup, z[:, q0 + i:q0 + i + 1]], dim=1).reshape(1, -1, z.size(2)) for i in range(m.n_query)]
        return m.loss_fn(m.forward_gnn(zs), self.y)

    def step(self, feat):
        for p in self.params:
            p.grad = None
        loss = self.loss(feat)
        loss.backward()
        return loss


def cpu_reference_episode_seconds(shape, steps, warmup, threads, budget_s=200.0):
    """(kind, [seconds per episode]) of fwd + bwd of the head on the host cores."""
    torch.set_num_threads(threads)
    kind = "port"
    try:
        from oracle import ref_loader as R
        if R.available():
            head = ReferenceHead(shape, "cpu")
            kind = "reference"
    except Exception as e:                                   # noqa: BLE001
        print(f"bench.py: oracle/_ref unusable ({e}); timing the CPU port instead", file=sys.stderr)
        kind = "port"
    if kind == "port":
        from oracle import gnn_oracle as O
        n_way, n_shot, n_query, compress, n = shape_dims(shape)
        p = O.random_params(128 + n_way, NF, n_way, seed=0, dtype=torch.float32, perturb_bn=False)
        params = {("gnn." + k): v.requires_grad_(True) for k, v in p.items()}
        g = torch.Generator().manual_seed(1)
        params["fc.0.weight"] = ((torch.rand(128, 512, generator=g) * 2 - 1) / 512 ** 0.5).requires_grad_(True)
        params["fc.0.bias"] = torch.zeros(128, requires_grad=True)
        params["fc.1.weight"] = torch.ones(128, requires_grad=True)
        params["fc.1.bias"] = torch.zeros(128, requires_grad=True)
    times, t_start = [], time.perf_counter()
    for it in range(warmup + steps):
        feat = synthetic_features(shape, 100 + it)
        t0 = time.perf_counter()
        if kind == "reference":
            head.step(feat)
        else:
            for v in params.values():
                v.grad = None
            O.head_loss(feat, params, n_way, n_shot, n_query, compress).backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        if time.perf_counter() - t_start > budget_s and times:     # bounded: never more than a few minutes
            break
    return kind, times


def cpu_baseline_record(kind, times, threads, warmup):
    sec = sum(times) / len(times)
    what = ("oracle/_ref: the reference's own methods/gnn.py + gnnnet.py head code (verbatim copies made by "
            "oracle/make_ref.py)") if kind == "reference" else "oracle/gnn_oracle.py (port of methods/gnn.py + gnnnet.py head)"
    return {"value": 1.0 / sec, "unit": "episodes/s", "cores": threads, "kind": kind,
            "sample": f"{len(times)} whole episodes of the same workload after {warmup} warm-up "
                      f"({sec:.2f} s each), torch CPU fp32, {what}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    kind, times = cpu_reference_episode_seconds(args.shape, args.steps, args.warmup, threads)
    sec = sum(times) / len(times)
    val = 1.0 / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "episodes/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload(args.shape), "shape": args.shape,
                   "arm": "host CPU, all cores; features -> fc -> graphs -> GNN_nl -> CE -> backward"},
        "cpu_baseline": cpu_baseline_record(kind, times, threads, args.warmup),
        "e2e": {"value": val, "unit": "episodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json(line)
    return 0


# ------------------------------------------------------------------------------------------
# the reference's head run with torch eager on the SAME GPU (cuDNN / cuBLAS): the unfused-GPU yardstick
# ------------------------------------------------------------------------------------------

def gpu_eager_reference(shape, dev, steps=10, warmup=3):
    from oracle import ref_loader as R
    if not R.available():
        return None
    out = {}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    try:
        for label, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            head = ReferenceHead(shape, dev)
            feats = [synthetic_features(shape, 300 + i, dev) for i in range(warmup + steps)]
            for i in range(warmup):
                head.step(feats[i])
            torch.cuda.synchronize()
            ms = 0.0
            for i in range(steps):
                flush.fill_(i & 0xff)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                head.step(feats[warmup + i])
                e1.record()
                torch.cuda.synchronize()
                ms += e0.elapsed_time(e1)
            out[label] = {"episodes_per_s": steps / (ms / 1e3), "ms_per_step": ms / steps}
            del head
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    out["what"] = ("the reference's own head (oracle/_ref: fc -> graphs -> methods/gnn.py GNN_nl -> CE -> backward) with "
                   "torch eager on this GPU: cuDNN 1x1 convs + cuBLAS, allow_tf32 off / on; same features, L2 flushed "
                   f"between steps, {steps} steps after {warmup} warm-up, CUDA events")
    return out


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------

def run_eval(args, world, rank, dev, lib, mft_b200):
    """Episode-sharded evaluation (finetune.py:634-682, configs C2/C5): forward only, B=15 graphs."""
    import torch.distributed as dist
    from mft_b200 import parallel
    n_episodes = args.episodes
    n_way, n_shot, n_query, compress, n = shape_dims(args.shape, 15)
    torch.manual_seed(0)
    head = mft_b200.GnnHead(n_way, n_shot, compress=compress, share_support=not args.no_share).to(dev)
    head.n_query = n_query
    parallel.broadcast_parameters(head)
    y = mft_b200.query_labels(n_way, n_query).to(dev)
    mine = parallel.owned_episodes(n_episodes, rank, world)
    feats = [synthetic_features(args.shape, 10 + e, "cpu", n_query).pin_memory() for e in mine[:64]]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    # the forward of one episode as a CUDA graph (features in a static buffer, accuracy out): the eager forward is
    # ~60 launches, which would make the sweep bound by the host's launch rate, not by the GPU
    static_f = feats[0].to(dev)
    use_graph = not args.no_graph
    if use_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(3):
                head.set_forward(static_f)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph), torch.no_grad():
            static_acc = (head.set_forward(static_f).argmax(1) == y).float().mean()

    def episode(f):
        with torch.no_grad():
            if use_graph:
                static_f.copy_(f, non_blocking=True)
                graph.replay()
                return static_acc.clone()
            scores = head.set_forward(f.to(dev, non_blocking=True))
            return (scores.argmax(1) == y).float().mean()

    for i in range(max(args.warmup, 3)):
        episode(feats[i % len(feats)])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(dev.index)
    sampler.start()
    l0 = lib.mft_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    accs = []
    for i, _ in enumerate(mine):
        if i % 8 == 0:
            flush.fill_(i & 0xff)
        accs.append(episode(feats[i % len(feats)]))
    acc_local = torch.stack(accs).double().cpu().tolist() if accs else []
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = lib.mft_launch_count() - l0
    if use_graph:                                # a replay does not pass through the library's launch counter
        with torch.no_grad():
            c0 = lib.mft_launch_count()
            head.set_forward(static_f)
            torch.cuda.synchronize()
            launches = (lib.mft_launch_count() - c0) * len(mine)
    clocks = sampler.stop()
    acc_all = parallel.gather_episode_results(acc_local, n_episodes, rank, world, device=dev)
    t = torch.tensor([ms, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t[:1], op=dist.ReduceOp.MAX)
        dist.all_reduce(t[1:], op=dist.ReduceOp.SUM)
    if rank == 0:
        mean, ci = parallel.accuracy_summary(acc_all.cpu())
        emit_json({
            "metric": "gnn_head_eval_episodes_per_sec_fwd", "value": n_episodes / (float(t[0]) / 1e3),
            "unit": "episodes/s", "n_gpus": world, "steps": n_episodes, "warmup": max(args.warmup, 3),
            "ms_per_step": float(t[0]) / max(1, len(mine)), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
            "config": {"workload": workload(args.shape, 15, backward=False), "shape": args.shape,
                       "episodes": n_episodes, "parallelism": f"episode-sharded eval, episode e on rank e % {world}, "
                       "no data-path collective; one all-reduce gather of the accuracies at the end",
                       "l2": "256 MiB fill every 8 episodes", "timing": "CUDA events around the rank's share, "
                       "host features -> H2D -> fc -> graphs -> GNN_nl -> argmax, accuracies read back once; max over ranks",
                       "launch": "CUDA graph replay of the forward" if use_graph else "eager launches"},
            "clocks": clocks, "gpu_launches": int(t[1]),
            "accuracy": {"mean": mean, "ci95": ci, "note": "synthetic features, random-init weights: chance level"},
        })
    return 0


def run_ours(args):
    import torch.distributed as dist
    import mft_b200
    from mft_b200 import _lib, parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load_library()
    mft_b200.set_precision(args.precision)
    if args.mode == "eval":
        rc = run_eval(args, world, rank, dev, lib, mft_b200)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return rc

    n_way, n_shot, n_query, compress, n = shape_dims(args.shape)
    torch.manual_seed(0)
    head = mft_b200.GnnHead(n_way, n_shot, compress=compress, share_support=not args.no_share).to(dev)
    head.n_query = n_query
    parallel.broadcast_parameters(head)
    gnn_params = list(head.gnn.parameters())
    from mft_b200.gnn import _resolve_precision
    prec = "tf32" if _resolve_precision([133, 181, 229], NF) == _lib.PREC_TF32 else "fp32"
    # data-parallel payload: the GNN gradients (and, end to end, the fc's) are the head's own; --dp-bytes adds
    # a flat buffer that stands for the rest of the model's gradients (backbone: outside the head)
    extra = None
    if world > 1 and args.dp_bytes > 0:
        extra_floats = max(0, int(args.dp_bytes * 1e6 / 4) - sum(p.numel() for p in gnn_params))
        extra = torch.zeros(extra_floats, dtype=torch.float32, device=dev)

    def allreduce(params):
        parallel.allreduce_mean_grads(params, world)
        if extra is not None:
            dist.all_reduce(extra, op=dist.ReduceOp.AVG)

    # dense TF32 peak of this GPU, once before the timed work (cold) and once after it: the larger one is the
    # roofline denominator (a peak measured on a chip the benchmark has just heated would flatter the fraction)
    tf32_before = measure_tf32_peak(seconds=0.0) if rank == 0 else None
    sampler = ClockSampler(local_rank)
    sampler.start()

    # Episode inputs.  Kernel-resident arm: the node tensors of every step already sit in HBM.
    total = args.warmup + args.steps
    nodes_dev = []
    with torch.no_grad():
        for it in range(total):
            feat = synthetic_features(args.shape, 1000 * rank + it, dev)
            nodes_dev.append(head.nodes(feat).contiguous())
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def fwd_bwd(nodes):
        return head.loss_from_nodes(nodes)      # GNN_nl + cross-entropy of the query nodes (mft_query_ce)

    def eager_step(nodes):
        for prm in gnn_params:
            prm.grad = None
        loss = fwd_bwd(nodes.detach().requires_grad_(True))
        loss.backward()
        return loss

    use_graph = not args.no_graph
    launches_per_step = None
    if use_graph:
        # launches of one step, counted on an eager run (a replay does not pass through the library)
        eager_step(nodes_dev[0]); torch.cuda.synchronize()
        l0 = lib.mft_launch_count(); eager_step(nodes_dev[0]); torch.cuda.synchronize()
        launches_per_step = lib.mft_launch_count() - l0
        gstep = mft_b200.GraphedStep(fwd_bwd, [nodes_dev[0]], gnn_params)

    def step(nodes):
        loss = gstep(nodes) if use_graph else eager_step(nodes)
        if world > 1:
            allreduce(gnn_params)
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: K steps, L2 flushed between steps, per-step CUDA events
    for it in range(args.warmup):
        step(nodes_dev[it])
    barrier()
    launches0 = lib.mft_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        flush.fill_(k & 0xff)
        ev[k][0].record()
        step(nodes_dev[args.warmup + k])
        ev[k][1].record()
    barrier()
    launches = lib.mft_launch_count() - launches0
    if use_graph:
        launches = launches_per_step * args.steps
    ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = sum(ms)

    # ---- end to end: host (pinned) features in, loss out, copies inside the timed region
    feats_host = [synthetic_features(args.shape, 5000 + 1000 * rank + it).pin_memory() for it in range(total)]
    all_params = list(head.parameters())
    if use_graph:
        g_e2e = mft_b200.GraphedStep(lambda f: head.set_forward_loss(f), [feats_host[0].to(dev)], all_params,
                                        inputs_require_grad=False)

    # The loss of every step is copied to pinned host memory inside the step (D2H, 4 bytes) and READ by
    # the host one step later, after that copy's event: the host never blocks the queue it is feeding.
    loss_host = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]
    losses = []

    # Input prefetch: the H2D copy of step k+1 runs on a copy stream into one of two staging buffers
    # while step k computes; the step itself starts with a device-side copy staging -> static input.
    copy_stream = torch.cuda.Stream()
    staging = [torch.empty_like(feats_host[0], device=dev) for _ in range(2)]
    staged_ev = [torch.cuda.Event() for _ in range(2)]
    consumed_ev = [torch.cuda.Event() for _ in range(2)]

    def prefetch(fh, k):
        slot = k & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed_ev[slot])        # the step that last read this buffer is done with it
            staging[slot].copy_(fh, non_blocking=True)
            staged_ev[slot].record(copy_stream)

    def e2e_step(fh, k, nxt=None):
        if k == 0 or not getattr(e2e_step, "primed", False):
            prefetch(fh, k)
            e2e_step.primed = True
        if nxt is not None:
            prefetch(nxt, k + 1)
        torch.cuda.current_stream().wait_event(staged_ev[k & 1])
        fdev = staging[k & 1]
        if use_graph:
            loss = g_e2e(fdev)                    # device copy into the graph's static input + replay
            consumed_ev[k & 1].record()
        else:
            for prm in all_params:
                prm.grad = None
            loss = head.set_forward_loss(fdev)
            loss.backward()
            consumed_ev[k & 1].record()
        if world > 1:
            allreduce(all_params)
        slot = k & 1
        if k >= 2:                                # the copy issued two steps ago into this slot has landed?
            loss_ev[slot].synchronize()
            losses.append(float(loss_host[slot][0]))
        loss_host[slot].copy_(loss.detach().reshape(1), non_blocking=True)   # device -> host read of the result
        loss_ev[slot].record()

    def e2e_drain(k_end):
        for k in range(max(0, k_end - 2), k_end):
            loss_ev[k & 1].synchronize()
            losses.append(float(loss_host[k & 1][0]))

    for it in range(args.warmup):
        e2e_step.primed = False
        e2e_step(feats_host[it], it)
    e2e_drain(args.warmup)
    losses.clear()
    barrier()
    e2e_step.primed = False
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        nxt = feats_host[args.warmup + k + 1] if k + 1 < args.steps else None
        e2e_step(feats_host[args.warmup + k], k, nxt)
    e2e_drain(args.steps)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    assert len(losses) == args.steps and all(l == l for l in losses), "end-to-end arm lost a loss value"
    clocks = sampler.stop()

    # ---- per-category device time of the library's kernels (same steps, events around each launch).
    # Every rank runs the steps (they contain the gradient all-reduce); only rank 0 records.
    prof = {}
    nprof = min(args.steps, 3)
    if rank == 0:
        lib.mft_prof_enable(1)
    for k in range(nprof):
        eager_step(nodes_dev[args.warmup + k])
        if world > 1:
            allreduce(gnn_params)
    if rank == 0:
        raw = _lib.profile_collect()
        lib.mft_prof_enable(0)
        prof = {k: (v[0] / nprof, v[1] // nprof) for k, v in raw.items()}
    barrier()

    # ---- max over ranks
    t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(t[0]), float(t[1])
    if world > 1:                               # the other ranks are done: rank 0 alone measures peaks below
        dist.barrier()
        dist.destroy_process_group()

    if rank == 0:
        peaks = measured_peaks()
        tf32 = measure_tf32_peak()
        tf32["burst_after"] = tf32["burst"]
        tf32["burst_before"] = tf32_before["burst"]
        tf32["burst"] = max(tf32["burst"], tf32_before["burst"])
        ms_per_step = total_ms / args.steps
        eps = world * args.steps / (total_ms / 1e3)
        e2e_eps = world * args.steps / (e2e_ms / 1e3)
        alg = head_flops(n_query, n, True)
        # dominant kernel family: the edge-MLP GEMM launches (fwd layers, dgrad, wgrad)
        # (wgrad and dgrad launches of a layer overlap on two streams, each on part of the SMs: their summed
        # durations would count that time twice, so the backward enters with the main-stream span of each
        # Wcompute's wgrad + dgrad region -- which also contains the dx gather -- the forward with its launches)
        spans = {k: round(prof.pop(k)[0], 4) for k in list(prof) if k.startswith("span_")}
        region = prof.pop("bwd_gemm_region", None)
        if region is not None:
            gemm_ms = sum(v[0] for k, v in prof.items() if k.startswith("fwd_gemm")) + region[0]
        else:
            gemm_ms = sum(v[0] for k, v in prof.items() if k.startswith(("fwd_gemm", "dgrad", "wgrad")))
        gemm_n = sum(v[1] for k, v in prof.items() if k.startswith(("fwd_gemm", "dgrad", "wgrad")))
        lib_ms = sum(v[0] for v in prof.values())
        achieved = alg / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None
        rows_exec = n_query * n * (n + 1) // 2           # unordered pairs of every graph
        n_sup = n - n_way                                 # support nodes: the same rows in every graph
        rows_w0 = rows_exec if args.no_share else (n_sup * (n_sup + 1) // 2
                                                   + n_query * (n * (n + 1) // 2 - n_sup * (n_sup + 1) // 2))
        pair_fl = [2 * (f * 192 + 192 * 192 + 192 * 96 + 96 * 96 + 96) for f in (133, 181, 229)]
        executed = 3.0 * (pair_fl[0] * rows_w0 + (pair_fl[1] + pair_fl[2]) * rows_exec)
        dp_floats = sum(p.numel() for p in gnn_params) + (extra.numel() if extra is not None else 0)
        line = {
            "metric": METRIC, "value": eps, "unit": "episodes/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32" if prec == "tf32" else "f32", "data": "synthetic",
            "config": {
                "workload": workload(args.shape), "shape": args.shape,
                "arm": "B200: device-resident nodes -> GNN_nl -> CE -> backward (value); host features -> ... (e2e)",
                "precision": prec,
                "share_support": (not args.no_share),
                "rows_layer_w0": rows_w0, "rows_other_layers": rows_exec,
                "tape": ("fp16 pre-BN activations with per-layer power-of-two scales, fp32 gradients (tensor-core path)"
                         if prec == "tf32" else "fp32"),
                "parallelism": f"episode-dp{world}" + (f" + nccl allreduce(avg) of {dp_floats} fp32 gradients "
                                                       f"({dp_floats * 4 / 1e6:.2f} MB) per step" if world > 1 else ""),
                "l2": "256 MiB fill between timed steps (L2 flushed); activation tape per step is 620 MB > L2",
                "timing": "CUDA events per step on torch's current stream, summed over K steps, max over ranks",
                "launch": "CUDA graph replay of fwd+bwd (captured once)" if use_graph else "eager launches",
            },
            "clocks": clocks,
            "e2e": {"value": e2e_eps, "unit": "episodes/s",
                    "h2d_bytes_per_step": int(feats_host[0].numel() * 4), "d2h_bytes_per_step": 4,
                    "what": "pinned host features -> H2D (prefetched one step ahead on a copy stream) -> fc(Linear+BN1d) -> "
                            "graphs -> GNN_nl -> CE -> backward "
                            "(all head parameters) -> loss copied to pinned host memory every step (read by the host one "
                            "step later, so the queue never drains)"},
            "gpu_launches": int(launches),
            "roofline": {
                "bound": "tensor", "achieved": achieved, "peak": tf32["burst"], "unit": "TFLOP/s",
                "frac": (achieved / tf32["burst"]) if achieved else None,
                "traffic": gemm_traffic(args.shape, not args.no_share, prec),
                "hbm_view": hbm_view(args.shape, not args.no_share, prec, ms_per_step, peaks.get("hbm_gbs")),
                "kernel": "edge-MLP GEMM launches (4 fwd + 4 dgrad + 4 wgrad per Wcompute, x3); time = forward launch "
                          "durations + span of each Wcompute's overlapped wgrad/dgrad region",
                "gemm_ms_per_step": gemm_ms,
                "launches_per_step": gemm_n, "avg_launch_ms": (gemm_ms / gemm_n) if gemm_n else None,
                "algorithmic_flops_per_step": alg,
                "executed_flops_per_step": executed,
                "executed_tflops": (executed / (gemm_ms / 1e3) / 1e12) if gemm_ms > 0 else None,
                "frac_executed": (executed / (gemm_ms / 1e3) / 1e12 / tf32["burst"]) if gemm_ms > 0 else None,
                "achieved_on_step": alg / (ms_per_step / 1e3) / 1e12,
                "frac_on_step": alg / (ms_per_step / 1e3) / 1e12 / tf32["burst"],
                "peak_sustained": tf32["sustained"],
                "peak_source": f"measured in this run: torch.matmul fp32 8192^3 with allow_tf32 (cuBLAS), best of 10, before the "
                               f"timed work {tf32['burst_before']:.1f} and after it {tf32['burst_after']:.1f} TF/s (the larger is "
                               f"the peak: burst, the GEMM launches are event-timed one by one), 2 s back to back = "
                               f"{tf32['sustained']:.1f} TF/s; {peaks['source']} has bf16 {peaks['bf16_burst']} / "
                               f"{peaks['bf16_sustained']} TF/s (half of it would be {peaks['bf16_burst'] / 2:.0f})",
                "note": "achieved counts the reference's dense B*N^2 pair FLOPs over the summed device time of the GEMM "
                        "launches (CUDA events around every launch, eager run of the same steps); the kernels execute "
                        "the N(N+1)/2 unordered pairs only and (share_support) the support-support pairs of layer_w0 once "
                        "for all graphs (executed_*); *_on_step divide by the whole graph-replayed step instead",
            },
            "head_tflops_algorithmic": alg * eps / world / 1e12,
            "kernel_ms_per_step": {k: round(v[0], 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])},
            "bwd_gemm_region_ms_per_step": round(region[0], 4) if region is not None else None,
            "main_stream_span_ms_per_step": spans,
            "library_kernel_ms_per_step": lib_ms,
        }
        if world == 1 and not args.no_gpu_reference:
            try:
                line["gpu_eager_reference"] = gpu_eager_reference(args.shape, dev)
            except Exception as e:                       # noqa: BLE001  (a yardstick, never fatal)
                line["gpu_eager_reference"] = {"error": str(e)[:200]}
        if world == 1 and not args.no_cpu_baseline:      # contract: rank 0 at N = 1 only
            threads = os.cpu_count() or 1
            n_s = 3 if args.shape != "5w5s" else 10
            kind, tms = cpu_reference_episode_seconds(args.shape, n_s, 1, threads, budget_s=60.0)
            line["cpu_baseline"] = cpu_baseline_record(kind, tms, threads, 1)
        emit_json(line)
    return 0


class _OnlyJsonOnStdout:
    """stdout carries exactly ONE JSON line.  Native libraries write to file descriptor 1 behind Python's back
    (the torch-bundled NCCL prints "NCCL version ..." there at communicator set-up, whatever NCCL_DEBUG and
    NCCL_DEBUG_FILE say), so descriptor 1 points at stderr for the whole run and the JSON line is written to
    the real stdout at the end."""

    def __enter__(self):
        sys.stdout.flush()
        self._real = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, line):
        sys.stdout.flush()
        os.write(self._real, (line + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self._real, 1)
        os.close(self._real)
        return False


_OUT = None


def emit_json(line):
    text = json.dumps(line)
    if _OUT is not None:
        _OUT.emit(text)
    else:
        print(text)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="5w20s", choices=sorted(SHAPES))
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "tf32"])
    ap.add_argument("--mode", default="train", choices=["train", "eval"])
    ap.add_argument("--episodes", type=int, default=600, help="--mode eval: episodes of the sweep (finetune.py: 600)")
    ap.add_argument("--dp-bytes", type=float, default=0.0,
                    help="MB of fp32 gradients all-reduced per step at N > 1 (default: the GNN's own 1.34 MB; 21.2 = "
                         "GNN + fc + ResNet10, SURVEY.md 8e)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--no-share", action="store_true",
                    help="evaluate the support-support pairs of layer_w0 in every graph (GnnHead(share_support=False))")
    ap.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    global _OUT
    with _OnlyJsonOnStdout() as _OUT:
        try:
            if args.impl == "reference":
                return run_reference(args)
            return run_ours(args)
        finally:
            _OUT = None


if __name__ == "__main__":
    sys.exit(main())
