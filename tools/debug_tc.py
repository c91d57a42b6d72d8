import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vistaocr_b200 import ops
dev = torch.device("cuda:0")
torch.set_printoptions(linewidth=250, precision=1, sci_mode=False)

def run(a_mn, b_mn, M, N, K, A, B):
    lda, ldb = A.shape[1], B.shape[1]
    C = torch.zeros((M, N), device=dev)
    ops.tc_gemm(a_mn, b_mn, M, N, K, ops.split_tf32(A.to(dev)), lda, ops.split_tf32(B.to(dev)), ldb, C, N)
    torch.cuda.synchronize()
    return C.cpu()

# failing K-major case
g = torch.Generator().manual_seed(0)
M, N, K = 64, 96, 2304
A = torch.randn(M, K, generator=g); B = torch.randn(N, K, generator=g)
C = run(0, 0, M, N, K, A, B); W = A.double() @ B.double().t()
print("KK 64x96x2304 err", (C.double() - W).abs().max().item(), "max", W.abs().max().item())
M, N, K = 128, 128, 2304
A = torch.randn(M, K, generator=g); B = torch.randn(N, K, generator=g)
C = run(0, 0, M, N, K, A, B); W = A.double() @ B.double().t()
print("KK 128x128x2304 err", (C.double() - W).abs().max().item(), "max", W.abs().max().item())
M, N, K = 64, 128, 64
A = torch.randn(M, K, generator=g); B = torch.randn(N, K, generator=g)
C = run(0, 0, M, N, K, A, B); W = A.double() @ B.double().t()
print("KK 64x128x64 err", (C.double() - W).abs().max().item(), "max", W.abs().max().item())

# MN-major A probe: A stored [K][M], A[k][m] = m + 1000*k ; B K-major selects k = n (n < 8)
M, N, K = 128, 128, 32
A = torch.zeros(K, M)
for k in range(K):
    A[k] = torch.arange(M).float() + 1000 * k
B = torch.zeros(N, K)
for n in range(min(N, K)):
    B[n, n] = 1.0
C = run(1, 0, M, N, K, A, B)
print("A MN-major probe: C[m][n] should be m + 1000 n (n<32)")
print(C[:8, :10]); print(C[30:36, :4]); print(C[64:68, :4]); print(C[:4, 8:12], C[:4, 30:33])
# MN-major B probe: B stored [K][N], B[k][n] = n + 1000 k; A K-major selects k = m
A = torch.zeros(M, K)
for m in range(min(M, K)):
    A[m, m] = 1.0
B = torch.zeros(K, N)
for k in range(K):
    B[k] = torch.arange(N).float() + 1000 * k
C = run(0, 1, M, N, K, A, B)
print("B MN-major probe: C[m][n] should be n + 1000 m (m<32)")
print(C[:10, :8]); print(C[:4, 30:36]); print(C[:4, 64:68])
