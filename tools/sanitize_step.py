"""One small tensor-core-path forward + backward of the whole head (B=4 graphs, N=30 nodes: every tcgen05
kernel variant, the Gconv kernels, the fused pre-head and loss) for compute-sanitizer runs
(tools/r02_sanitize.sh).  Prints the loss so that a silent early exit cannot pass as a clean run."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mft_b200

mft_b200.set_precision(os.environ.get("MFT_PRECISION", "tf32"))
torch.manual_seed(0)
head = mft_b200.GnnHead(5, 5).cuda()
head.n_query = 4
feat = torch.randn(5, 9, 512, device="cuda")
for _ in range(2):
    loss = head.set_forward_loss(feat)
    loss.backward()
torch.cuda.synchronize()
print("loss", float(loss.detach()), "launches", mft_b200.load_library().mft_launch_count())
