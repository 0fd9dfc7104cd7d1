#!/usr/bin/env python
"""Headline benchmark: GNN-head episodes/sec (forward + backward), BASELINE.json's metric.

A "step" is one pass of the hot path over one synthetic episode: the n_query graphs of a
5-way 20-shot episode (B=16 graphs, N=105 nodes, F=133 features -- the configuration the
metric's target is quoted on) go through ``GNN_nl`` forward, a cross-entropy
on the query nodes, and the full backward (input + all 64 parameter gradients).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--shape 5w5s|5w20s|5w50c] [--precision auto|fp32|tf32]

Multi-GPU: launched under torchrun, one rank per GPU; every rank runs its own episodes
(weak scaling, episodes are independent) and the GNN gradients are averaged with one NCCL
all-reduce per step (the episode-parallel meta-training path).

``--impl reference`` times the CPU restatement of the reference (oracle/gnn_oracle.py --
the reference is Python/PyTorch and /root/reference does not exist on the GPU box) on the
host cores, same workload, same metric.
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


def head_flops(bsz, n, backward=True):
    per_pair = sum(2 * (f * 192 + 192 * 192 + 192 * 96 + 96 * 96 + 96) for f in (133, 181, 229))
    return per_pair * (3 if backward else 1) * bsz * n * n


def gemm_traffic(shape, share, prec):
    """DRAM bytes per edge-MLP GEMM launch from the committed `ncu` capture of this workload
    (profiles/r01_gemm_traffic.json, written by tools/ncu_traffic.py), or None."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_gemm_traffic.json")
    try:
        rec = json.load(open(path))
    except (OSError, ValueError):
        return None
    same = rec.get("config") == {"shape": shape, "share_support": bool(share), "precision": prec}
    return rec.get("bytes_per_launch") if same else None   # the capture describes ONE configuration


def shape_dims(shape):
    n_way, n_shot, n_query, compress = SHAPES[shape]
    k = round(n_shot / 2) if compress else n_shot
    return n_way, n_shot, n_query, compress, n_way * (k + 1)


def synthetic_features(shape, seed, device="cpu"):
    """Backbone features of one episode: [n_way, n_shot+n_query, 512] ~ N(0,1) (synthetic; the
    ResNet10 backbone is outside the hot path and stays on cuDNN)."""
    n_way, n_shot, n_query, _, _ = shape_dims(shape)
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n_way, n_shot + n_query, 512, generator=g).to(device)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "source": "MEASURED_PEAKS.json"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
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
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# reference arm: CPU restatement of the reference on the host cores
# ------------------------------------------------------------------------------------------

def cpu_reference_episode_seconds(shape, steps, warmup, threads):
    """fwd + bwd of the head on CPU (oracle port of methods/gnn.py + gnnnet.py glue)."""
    from oracle import gnn_oracle as O
    torch.set_num_threads(threads)
    n_way, n_shot, n_query, compress, n = shape_dims(shape)
    p = O.random_params(128 + n_way, NF, n_way, seed=0, dtype=torch.float32, perturb_bn=False)
    params = {("gnn." + k): v.requires_grad_(True) for k, v in p.items()}
    g = torch.Generator().manual_seed(1)
    params["fc.0.weight"] = ((torch.rand(128, 512, generator=g) * 2 - 1) / 512 ** 0.5).requires_grad_(True)
    params["fc.0.bias"] = torch.zeros(128, requires_grad=True)
    params["fc.1.weight"] = torch.ones(128, requires_grad=True)
    params["fc.1.bias"] = torch.zeros(128, requires_grad=True)
    times = []
    for it in range(warmup + steps):
        feat = synthetic_features(shape, 100 + it)
        for v in params.values():
            v.grad = None
        t0 = time.perf_counter()
        loss = O.head_loss(feat, params, n_way, n_shot, n_query, compress)
        loss.backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    # bounded sample: whole episodes of the same workload, few enough to end within minutes
    steps = max(1, min(args.steps, 3 if args.shape != "5w5s" else 20))
    warmup = 1
    times = cpu_reference_episode_seconds(args.shape, steps, warmup, threads)
    sec = sum(times) / len(times)
    n_way, n_shot, n_query, compress, n = shape_dims(args.shape)
    val = 1.0 / sec
    line = {
        "impl": "reference", "metric": "gnn_head_episodes_per_sec_fwd_bwd", "value": val, "unit": "episodes/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"GnnNet head fwd+bwd, {args.shape} (B={n_query} graphs, N={n} nodes, F=133, nf=96), "
                               f"features->fc->graphs->GNN_nl->CE->backward", "shape": args.shape},
        "cpu_baseline": {"value": val, "unit": "episodes/s", "cores": threads, "kind": "port",
                         "sample": f"{steps} whole episodes after {warmup} warm-up, torch CPU fp32, "
                                   f"oracle/gnn_oracle.py (port of methods/gnn.py + gnnnet.py head)"},
        "e2e": {"value": val, "unit": "episodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json(line)
    return 0


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------

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

    n_way, n_shot, n_query, compress, n = shape_dims(args.shape)
    torch.manual_seed(0)
    head = mft_b200.GnnHead(n_way, n_shot, compress=compress, share_support=not args.no_share).to(dev)
    head.n_query = n_query
    parallel.broadcast_parameters(head)
    gnn_params = list(head.gnn.parameters())
    y = mft_b200.query_labels(n_way, n_query).to(dev)
    from mft_b200.gnn import _resolve_precision
    prec = "tf32" if _resolve_precision([133, 181, 229], NF) == _lib.PREC_TF32 else "fp32"

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
        if use_graph:
            loss = gstep(nodes)
        else:
            loss = eager_step(nodes)
        if world > 1:
            parallel.allreduce_mean_grads(gnn_params, world)
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
    sampler = ClockSampler(local_rank)
    sampler.start()
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
    clocks = sampler.stop()
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
            parallel.allreduce_mean_grads(all_params, world)
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

    # ---- per-category device time of the library's kernels (same steps, events around each launch).
    # Every rank runs the steps (they contain the gradient all-reduce); only rank 0 records.
    prof = {}
    nprof = min(args.steps, 3)
    if rank == 0:
        lib.mft_prof_enable(1)
    for k in range(nprof):
        eager_step(nodes_dev[args.warmup + k])
        if world > 1:
            parallel.allreduce_mean_grads(gnn_params, world)
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

    if rank == 0:
        peaks = measured_peaks()
        ms_per_step = total_ms / args.steps
        eps = world * args.steps / (total_ms / 1e3)
        e2e_eps = world * args.steps / (e2e_ms / 1e3)
        alg = head_flops(n_query, n, True)
        # dominant kernel family: the edge-MLP GEMM launches (fwd layers, dgrad, wgrad)
        gemm_ms = sum(v[0] for k, v in prof.items() if k.startswith(("fwd_gemm", "dgrad", "wgrad")))
        gemm_n = sum(v[1] for k, v in prof.items() if k.startswith(("fwd_gemm", "dgrad", "wgrad")))
        lib_ms = sum(v[0] for v in prof.values())
        tf32_peak = peaks["bf16_sustained"] / 2.0
        achieved = alg / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None
        rows_exec = n_query * n * (n + 1) // 2           # unordered pairs of every graph
        n_sup = n - n_way                                 # support nodes: the same rows in every graph
        rows_w0 = rows_exec if args.no_share else (n_sup * (n_sup + 1) // 2
                                                   + n_query * (n * (n + 1) // 2 - n_sup * (n_sup + 1) // 2))
        pair_fl = [2 * (f * 192 + 192 * 192 + 192 * 96 + 96 * 96 + 96) for f in (133, 181, 229)]
        executed = 3.0 * (pair_fl[0] * rows_w0 + (pair_fl[1] + pair_fl[2]) * rows_exec)
        line = {
            "metric": "gnn_head_episodes_per_sec_fwd_bwd", "value": eps, "unit": "episodes/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32" if prec == "tf32" else "f32", "data": "synthetic",
            "config": {
                "workload": f"GnnNet head fwd+bwd, {args.shape}: GNN_nl on B={n_query} graphs x N={n} nodes, F=133, "
                            f"nf=96, n_way={n_way}; CE on the query nodes; input + 64 parameter gradients",
                "shape": args.shape, "precision": prec,
                "share_support": (not args.no_share),
                "rows_layer_w0": rows_w0, "rows_other_layers": rows_exec,
                "tape": "fp16 pre-BN activations, fp32 gradients (tensor-core path)" if prec == "tf32" else "fp32",
                "parallelism": f"episode-dp{world}" + (" + nccl allreduce(gnn grads, 1.34 MB)" if world > 1 else ""),
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
                "bound": "tensor", "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s",
                "frac": (achieved / tf32_peak) if achieved else None, "traffic": gemm_traffic(args.shape, not args.no_share, prec),
                "kernel": "edge-MLP GEMM launches (4 fwd + 4 dgrad + 4 wgrad per Wcompute, x3)",
                "launches_per_step": gemm_n, "avg_launch_ms": (gemm_ms / gemm_n) if gemm_n else None,
                "algorithmic_flops_per_step": alg,
                "executed_flops_per_step": executed,
                "peak_source": f"{peaks['source']}: bf16 sustained {peaks['bf16_sustained']} TF/s / 2 (dense TF32 "
                               f"is half the bf16 rate; no TF32 figure is driver-measured)",
                "note": "achieved counts the reference's dense B*N^2 pair FLOPs; the kernels execute the "
                        "N(N+1)/2 unordered pairs only and (share_support) the support-support pairs of "
                        "layer_w0 once for all graphs (executed_flops_per_step)",
            },
            "head_tflops_algorithmic": alg * eps / world / 1e12,
            "kernel_ms_per_step": {k: round(v[0], 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])},
            "library_kernel_ms_per_step": lib_ms,
        }
        if not args.no_cpu_baseline and world >= 1:
            threads = os.cpu_count() or 1
            n_s = 2 if args.shape != "5w5s" else 10
            tms = cpu_reference_episode_seconds(args.shape, n_s, 1, threads)
            sec = sum(tms) / len(tms)
            line["cpu_baseline"] = {"value": 1.0 / sec, "unit": "episodes/s", "cores": threads, "kind": "port",
                                    "sample": f"{n_s} whole episodes of the same workload after 1 warm-up "
                                              f"({sec:.2f} s each), torch CPU fp32, oracle/gnn_oracle.py"}
        emit_json(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
