"""Pure-write HBM bandwidth on this GPU (torch fill / memset), to place the H-build fill kernel
(a 12 B/nnz write stream) against what the memory system can absorb as writes alone."""
import torch
torch.cuda.set_device(0)
n = int(12.4e9 // 8)
a = torch.empty(n, dtype=torch.float64, device="cuda")
b = torch.empty(n // 2, dtype=torch.int32, device="cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: a.fill_(1.5)); print("fill f64 %.1f GB: %.3f ms -> %.0f GB/s" % (a.nbytes/1e9, ms, a.nbytes/ms/1e6))
ms = t(lambda: a.zero_()); print("memset %.1f GB: %.3f ms -> %.0f GB/s" % (a.nbytes/1e9, ms, a.nbytes/ms/1e6))
ms = t(lambda: (a.fill_(1.5), b.fill_(3))); tot = a.nbytes + b.nbytes; print("fill f64+i32 %.1f GB: %.3f ms -> %.0f GB/s" % (tot/1e9, ms, tot/ms/1e6))
c = torch.empty(n // 2, dtype=torch.float64, device="cuda")
ms = t(lambda: c.copy_(a[: n // 2])); print("copy %.1f GB moved: %.3f ms -> %.0f GB/s" % (2*c.nbytes/1e9, ms, 2*c.nbytes/ms/1e6))
