import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import util as U
import mft_b200
mft_b200.set_precision("tf32")
torch.manual_seed(0)
for F in (181, 133):
    m = mft_b200.Wcompute(F, 96).cuda()
    x = torch.randn(16, 105, F); up = torch.randn(16, 105, 105).cuda()
    res = []
    for rep in range(3):
        for p in m.parameters(): p.grad = None
        xg = x.cuda().requires_grad_(True)
        m.adjacency(xg, None).backward(up); torch.cuda.synchronize()
        res.append({k: v.grad.double().cpu().numpy().copy() for k, v in m.named_parameters()})
    print("F=%d" % F, " ".join("%s %.2e/%.2e" % (k[:8], U.rel(res[0][k], res[1][k]), U.rel(res[0][k], res[2][k])) for k in ("conv2d_1.weight", "conv2d_2.weight")), flush=True)
