"""Does the bulk-tensor store clip ragged N per element?  (development probe)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["OFAB_GEMM_FORCE_TMA_STORE"] = "1"
from ofasys_b200 import ops
dev = torch.device("cuda:0")
M, N, K = 300, 773, 192
A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
o = torch.full((M + 64, N + 11), 7.0, dtype=torch.bfloat16, device=dev)
ops.gemm(M, N, K, A, K, 0, B, K, 0, o, N + 11)
torch.cuda.synchronize()
ref = A.float() @ B.float().t()
print("valid err", ((o[:M, :N].float() - ref).norm() / ref.norm()).item())
print("cols beyond N touched:", (o[:M, N:] != 7.0).sum().item(), "rows beyond M touched:", (o[M:] != 7.0).sum().item())
print(o[0, N - 2:N + 11])
