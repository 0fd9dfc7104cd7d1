"""Kernel timeline INSIDE a CUDA-graph replay of the head step (forward + backward), from CUPTI kernel activity
records (torch.profiler): start, duration and stream of every kernel of one replay, so side-stream overlap and the
gaps between dependent kernels show as they are in the benchmarked schedule.  Timestamps taken under a profiler
are not bench values; the SHARES are what this is for.
usage: python tools/graph_timeline.py [shape] [out.json]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
import mft_b200

shape = sys.argv[1] if len(sys.argv) > 1 else "5w20s"
out = sys.argv[2] if len(sys.argv) > 2 else None
n_way, n_shot, n_query, compress, n = bench.shape_dims(shape)
torch.manual_seed(0)
dev = torch.device("cuda", 0)
head = mft_b200.GnnHead(n_way, n_shot, compress=compress).to(dev)
head.n_query = n_query
params = list(head.gnn.parameters())
with torch.no_grad():
    nodes = [head.nodes(bench.synthetic_features(shape, i, dev)).contiguous() for i in range(4)]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
step = mft_b200.GraphedStep(lambda x: head.loss_from_nodes(x), [nodes[0]], params)
for i in range(3):
    step(nodes[i])
torch.cuda.synchronize()

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(3):
        flush.fill_(i)
        torch.cuda.synchronize()
        step(nodes[i + 1])
        torch.cuda.synchronize()

ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
ker = sorted(((e.time_range.start, e.time_range.end, e.name, getattr(e, "device_index", 0)) for e in ev
              if "fill" not in e.name.lower() and "memset" not in e.name.lower() or True), key=lambda t: t[0])
# split into replays at the flush kernels (vectorized fill of 256 MB: the only > 20 us elementwise kernel)
groups, cur = [], []
for s, t, name, _ in ker:
    if "FillFunctor" in name and (t - s) > 15:
        if cur:
            groups.append(cur)
        cur = []
        continue
    cur.append((s, t, name))
if cur:
    groups.append(cur)
g = groups[-1]
t0 = min(s for s, _, _ in g)
t1 = max(t for _, t, _ in g)
print(f"replay span {1e-3 * (t1 - t0):.4f} ms, {len(g)} kernels/copies")


def short(name):
    name = name.split("(")[0]
    for pre in ("void ", "mft::", "(anonymous namespace)::"):
        name = name.replace(pre, "")
    return name[:70]


# union coverage and the time only ONE kernel was running
pts = sorted([(s, 1) for s, _, _ in g] + [(t, -1) for _, t, _ in g])
busy = alone = 0.0
depth, last = 0, pts[0][0]
for x, d in pts:
    if depth >= 1:
        busy += x - last
    if depth == 1:
        alone += x - last
    depth += d
    last = x
print(f"some kernel running {1e-3 * busy:.4f} ms, exactly one running {1e-3 * alone:.4f} ms, "
      f"idle {1e-3 * (t1 - t0 - busy):.4f} ms")
agg = {}
for s, t, name in g:
    a = agg.setdefault(short(name), [0.0, 0])
    a[0] += t - s
    a[1] += 1
print(f"{'kernel':72s} {'sum ms':>8s} {'n':>4s}")
for k, (d, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:72s} {1e-3 * d:8.4f} {c:4d}")
print("\n# chronological (start us, dur us, name)")
for s, t, name in g:
    print(f"{s - t0:9.2f} {t - s:8.2f}  {short(name)}")
if out:
    json.dump({"shape": shape, "span_us": t1 - t0, "busy_us": busy, "alone_us": alone,
               "kernels": [[s - t0, t - s, short(nm)] for s, t, nm in g]}, open(out, "w"))
