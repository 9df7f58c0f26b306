import torch, time
n = 4295753728
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, src, dst in (("D2H", d, h), ("H2D", h, d)):
    for _ in range(2):
        dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 3
    print(name, "%.1f ms  %.1f GB/s" % (dt * 1e3, n / dt / 1e9))
# both directions at once
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(2028992061, dtype=torch.uint8).pin_memory(); d2 = torch.empty(2028992061, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t = time.perf_counter()
with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print("D2H 4.3 GB + H2D 2.0 GB concurrently: %.1f ms" % (dt * 1e3))
