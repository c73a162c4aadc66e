"""Runs the l0 spatial self-attention shape a few times (target of the ncu --set full capture)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mvoc_b200 import ops  # noqa: E402

B, H, N = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (80, 5, 4096)
dev = torch.device("cuda:0")
torch.manual_seed(0)
C = H * 64
qkv = torch.randn(B, N, 3 * C, device=dev).bfloat16()
q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
for _ in range(4):
    o = ops.attention(q, k, v, H)
torch.cuda.synchronize()
print("done", float(o.float().abs().mean()))
