"""Per-CTA phase timeline of one tcgen05 rows-GEMM launch (debug tool, GPU only)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mft_b200 import _lib
lib = _lib.load_library()
M, N, K = 89040, 192, 192
A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda")
out = torch.empty(M, N, device="cuda")
ws = torch.empty(lib.mft_debug_umma_gemm_workspace_bytes(N, K), dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def run():
    _lib.check(lib.mft_debug_umma_gemm(A.data_ptr(), K, W.data_ptr(), K, 0, out.data_ptr(), N, M, N, K, ws.data_ptr(), st), "gemm")
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
print("plain GEMM 89040x192x192 (image kernel + gemm): %.1f us" % (e0.elapsed_time(e1) * 1e3))
tl = torch.zeros(148, 16, dtype=torch.int64, device="cuda")
lib.mft_debug_set_timeline(tl.data_ptr()); run(); torch.cuda.synchronize(); lib.mft_debug_set_timeline(None)
t = tl.cpu()
names = {1: "weights resident", 2: "prod first tile", 3: "prod done", 4: "mma first tile issued", 5: "mma all issued",
         12: "epi first acc ready", 8: "epi tile0 done", 9: "epi tile1 done", 10: "epi tile2 done", 11: "epi tile3 done",
         6: "epi tiles done", 7: "stats committed"}
for b in (0, 1, 73, 147):
    base = int(t[b, 0])
    print("CTA", b, {names[k]: int(t[b, k]) - base for k in names if int(t[b, k]) > 0})
