"""Launch list helper: one bf16-mode and one exact-mode top-k on the clustered c3 corpus (run under ncu --metrics gpu__time_duration.sum)."""
import sys
import torch
sys.path.insert(0, ".")
from ah_rag_b200 import engine
dev = torch.device("cuda", 0)
n, dim, nq, k, n_cent = 10_000_000, 1024, 4096, 50, 4096
g = torch.Generator(device=dev).manual_seed(91011)
cent = torch.randn((n_cent, dim), generator=g, device=dev)
X = torch.empty((n, dim), dtype=torch.bfloat16, device=dev)
for c0 in range(0, n, 500_000):
    m = min(500_000, n - c0)
    x = cent[torch.randint(0, n_cent, (m,), generator=g, device=dev)] + 0.3 * torch.randn((m, dim), generator=g, device=dev)
    x /= x.norm(dim=1, keepdim=True)
    X[c0:c0 + m] = x.to(torch.bfloat16)
pick = torch.randint(0, n, (nq,), generator=g, device=dev)
q = X[pick].float() + 0.1 * torch.randn((nq, dim), generator=g, device=dev) / dim ** 0.5
Q = (q / q.norm(dim=1, keepdim=True)).to(torch.bfloat16).contiguous()
idx = engine.VectorIndex(X)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("bf16")
idx.topk(Q, k, mode="bf16")
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
torch.cuda.nvtx.range_push("exact")
idx.topk(Q, k, mode="exact")
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
