"""tcgen05 TF32 x 3 GEMM against numpy/torch float64 (correctness + time).  usage: tc_gemm_check.py [M N K tile128|persistent]"""
import sys, time
sys.path.insert(0, '/root/repo')
import torch
from pomdpy_b200 import vi
M, N, K = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (256, 256, 256)
tn = sys.argv[4] if len(sys.argv) > 4 else "tile128"
torch.manual_seed(1)
a = torch.rand(M, K, dtype=torch.float64, device="cuda") - 0.3
bt = torch.rand(N, K, dtype=torch.float64, device="cuda") - 0.6
sa, sb = vi.split_tf32(a), vi.split_tf32(bt)
torch.cuda.synchronize()
print("split ok: max |a - hi - lo| / |a| =", float(((a - sa[0].double() - sa[1].double()).abs() / a.abs().clamp_min(1e-30)).max()), flush=True)
c = vi.gemm_tf32x3(sa, sb, variant=tn)
torch.cuda.synchronize()
ref = a @ bt.T
scale = (a.abs() @ bt.abs().T)
err = ((c - ref).abs() / scale).max()
print("M N K variant", M, N, K, tn, "max |err| / (|A||B|) =", float(err), "max abs err", float((c - ref).abs().max()), flush=True)
bad = ((c - ref).abs() / scale) > 1e-5
print("bad entries", int(bad.sum()), "of", bad.numel(), flush=True)
if int(bad.sum()):
    idx = bad.nonzero()[:8]
    print("first bad (row, col):", idx.tolist(), [float(c[i, j]) for i, j in idx.tolist()], [float(ref[i, j]) for i, j in idx.tolist()])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(2): vi.gemm_tf32x3(sa, sb, variant=tn)
e0.record()
for _ in range(5): vi.gemm_tf32x3(sa, sb, variant=tn)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("%.3f ms per GEMM, %.1f TFLOP/s (fp64-equivalent 2MNK), %.1f TFLOP/s of tf32 MMAs" % (ms, 2.0 * M * N * K / ms / 1e9, 6.0 * M * N * K / ms / 1e9))
# VI-shaped data: the textbook projection at |S| = S2 against the fp64 DMMA kernel (all entries) and numpy (sampled rows)
import numpy as np
for (S2, A2, Z2, G2) in ((256, 3, 8, 16), (4096, 2, 64, 8)):
    rng = np.random.default_rng(123)
    T = rng.dirichlet(np.ones(S2), size=(A2, S2)); O = rng.dirichlet(np.ones(Z2), size=(A2, S2)); gamma = rng.uniform(-1, 1, size=(G2, S2))
    Td, Od, Gd = (torch.from_numpy(x).cuda() for x in (T, O, gamma))
    exact = vi.project_textbook(Td, Od, Gd)
    fast = vi.project_textbook(Td, Od, Gd, precision="tf32x3")
    torch.cuda.synchronize()
    print("projection S A Z G", S2, A2, Z2, G2, "max |fast - fp64| / max |fp64| =", float((fast - exact).abs().max() / exact.abs().max()),
          " max |fp64|", float(exact.abs().max()))
