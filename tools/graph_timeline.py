"""Per-category device time INSIDE a CUDA-graph replay of the 5w20s head step: the library's event pairs are
captured into the graph as event-record nodes, so after a replay they hold the replay's own schedule (side
streams overlapping, no host launch gaps).  usage: python tools/graph_timeline.py [shape]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import mft_b200
from mft_b200 import _lib

shape = sys.argv[1] if len(sys.argv) > 1 else "5w20s"
lib = _lib.load_library()
n_way, n_shot, n_query, compress, n = bench.shape_dims(shape)
torch.manual_seed(0)
dev = torch.device("cuda", 0)
head = mft_b200.GnnHead(n_way, n_shot, compress=compress).to(dev)
head.n_query = n_query
params = list(head.gnn.parameters())
with torch.no_grad():
    nodes = [head.nodes(bench.synthetic_features(shape, i, dev)).contiguous() for i in range(4)]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for i in range(3):                                       # eager warm-up first: the profile must hold the graph only
    for p in params:
        p.grad = None
    head.loss_from_nodes(nodes[i].detach().requires_grad_(True)).backward()
torch.cuda.synchronize()
lib.mft_prof_enable(1)                                   # event pairs get captured with the launches
step = mft_b200.GraphedStep(lambda x: head.loss_from_nodes(x), [nodes[0]], params, warmup=0)
for i in range(3):
    flush.fill_(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); step(nodes[i + 1]); e1.record(); torch.cuda.synchronize()
    print("replay ms", e0.elapsed_time(e1))
raw = _lib.profile_collect()
lib.mft_prof_enable(0)
tot = 0.0
for k, v in sorted(raw.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:24s} {v[0]:8.4f} ms  {v[1]:4d} scopes")
