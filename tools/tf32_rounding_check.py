"""Does skipping the cvt.rna pass (GB200_TC_TRUNCATE=1) change the GEMM error?  Prints rel-L2 and mean signed
relative bias of C = A B^T against fp64 for positive operands (bias shows truncation) and for N(0,1) operands."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from galerkin_transformer_b200 import functional as GF
torch.manual_seed(0)
M, N, K = 4096, 256, 512
for name, gen in (("uniform(1,2)", lambda *s: 1 + torch.rand(*s, device="cuda")), ("normal", lambda *s: torch.randn(*s, device="cuda"))):
    A, B = gen(M, K), gen(N, K)
    C = torch.empty(M, N, device="cuda")
    GF.set_precision("tf32")
    GF.gemm(A, B, C, M, N, K, lda=K, ldb=K, ldc=N, transB=True)
    ref = A.double() @ B.double().t()
    err = (C.double() - ref)
    print(f"TRUNCATE={os.environ.get('GB200_TC_TRUNCATE','0')} {name:13s} rel-L2 {err.norm()/ref.norm():.3e}  mean signed rel {(err/ref.abs().clamp_min(1e-9)).mean():+.3e}")
    torch.backends.cuda.matmul.allow_tf32 = True
    Cc = A @ B.t()
    e2 = Cc.double() - ref
    print(f"   cuBLAS tf32   {name:13s} rel-L2 {e2.norm()/ref.norm():.3e}  mean signed rel {(e2/ref.abs().clamp_min(1e-9)).mean():+.3e}")
