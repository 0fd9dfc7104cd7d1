"""Per-CTA phase timeline (clock64) of one tcgen05 rows-GEMM launch inside a real 5w20s forward /
backward (debug tool, GPU only).  usage: python tools/umma_timeline.py [launch index ...]
The three "epi cycles" phase counters are only filled by a library built with -DMFT_UMMA_TIMING
(make EXTRA=-DMFT_UMMA_TIMING); the default build leaves them at 0 to keep clock reads out of the
epilogue loop."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mft_b200
from mft_b200 import _lib
lib = _lib.load_library()
mft_b200.set_precision("tf32")
torch.manual_seed(0)
net = mft_b200.GNN_nl(133, 96, 5).cuda()
x = torch.randn(16, 105, 133, device="cuda", requires_grad=True)
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
names = {1: "weights", 2: "prod tile0", 3: "prod done", 4: "mma tile0", 5: "mma done", 12: "acc0 ready",
         8: "epi t0", 9: "epi t1", 10: "epi t2", 11: "epi t3", 6: "epi done", 7: "stats"}
phase = {13: "epi cycles in tmem wait", 14: "epi cycles TMEM->slab", 15: "epi cycles slab->global+stats"}
def run():
    out = net(x); out.sum().backward(); torch.cuda.synchronize()
for _ in range(2): run()
# forward launches per Wcompute: L1 (1-2 passes), L2, L3, L4 ; then backward dgrad launches
for idx in [int(a) for a in sys.argv[1:]] or [1, 2, 3]:
    tl = torch.zeros(148, 16, dtype=torch.int64, device="cuda")
    flush.fill_(1)
    lib.mft_debug_set_timeline(tl.data_ptr(), idx)
    run()
    t = tl.cpu()
    for b in (0, 73):
        base = int(t[b, 0])
        print("launch", idx, "CTA", b, {names[k]: int(t[b, k]) - base for k in names if int(t[b, k]) > 0},
              {phase[k]: int(t[b, k]) for k in phase})
