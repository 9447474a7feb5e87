import torch, time
n = 80 << 20
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); d = torch.empty(n, dtype=torch.uint8, device='cuda')
h2 = torch.empty(32 << 20, dtype=torch.uint8, pin_memory=True); d2 = torch.empty(32 << 20, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(f, reps=10):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
a = t(lambda: d.copy_(h, non_blocking=True)); print("H2D 80MB  %.3f ms  %.1f GB/s" % (a*1e3, n/a/1e9))
b = t(lambda: h2.copy_(d2, non_blocking=True)); print("D2H 32MB  %.3f ms  %.1f GB/s" % (b*1e3, (32<<20)/b/1e9))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both); print("both      %.3f ms" % (c*1e3))
# many small copies
hs = h.view(160, -1); ds = d.view(160, -1)
def small():
    for i in range(160): ds[i].copy_(hs[i], non_blocking=True)
e = t(small); print("160 x 512KB H2D %.3f ms" % (e*1e3))
