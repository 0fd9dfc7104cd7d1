"""Print the per-tensor error table of the CUDA path against a golden fixture
(relative L2 vs the float64 reference run; the reference's own fp32 error beside it).

    python tools/parity_report.py [gnn_5w5s.npz] [fp32|tf32] [fused|modular]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import util as U  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "gnn_5w5s.npz"
    prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
    fused = (sys.argv[3] if len(sys.argv) > 3 else "fused") == "fused"
    rec = dict(np.load(os.path.join(ROOT, "tests", "golden", name)))
    params = {k[2:]: v for k, v in rec.items() if k.startswith("p.")}
    fin = params["layer_w0.conv2d_1.weight"].shape[1]
    nf = params["layer_w0.conv2d_4.weight"].shape[0]
    n_way = params["layer_last.fc.weight"].shape[0]
    out, dx, grads = U.run_cuda_gnn(rec["x"], params, rec["proj"], fin, nf, n_way, prec, fused)
    print(f"{name} precision={prec} fused={fused}")
    print(f"{'tensor':40s} {'ours':>10s} {'ref fp32':>10s}")
    print(f"{'out':40s} {U.rel(out, rec['out64']):10.2e} {U.rel(rec['out32'], rec['out64']):10.2e}")
    print(f"{'dx':40s} {U.rel(dx, rec['dx64']):10.2e} {U.rel(rec['dx32'], rec['dx64']):10.2e}")
    for k in params:
        g = grads[k].reshape(rec["g." + k].shape)
        if U.is_zero_grad(k):
            print(f"{k:40s} max|g|={np.abs(g).max():.1e} (analytically zero)")
        else:
            print(f"{k:40s} {U.rel(g, rec['g.' + k]):10.2e} {float(rec['e32.' + k]):10.2e}")


if __name__ == "__main__":
    main()
