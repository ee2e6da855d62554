import torch, time
n = 407390232 // 8
xh = torch.empty(n, dtype=torch.float64).pin_memory(); yh = torch.empty(n, dtype=torch.float64).pin_memory()
xd = torch.empty(n, dtype=torch.float64, device="cuda"); yd = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
print("H2D ms", t(lambda: xd.copy_(xh, non_blocking=True)))
print("D2H ms", t(lambda: yh.copy_(yd, non_blocking=True)))
def both():
    with torch.cuda.stream(s1): xd.copy_(xh, non_blocking=True)
    with torch.cuda.stream(s2): yh.copy_(yd, non_blocking=True)
print("H2D || D2H ms", t(both))
def chunked(k):
    c = n // k
    def f():
        for i in range(k):
            with torch.cuda.stream(s1): xd[i*c:(i+1)*c].copy_(xh[i*c:(i+1)*c], non_blocking=True)
            with torch.cuda.stream(s2): yh[i*c:(i+1)*c].copy_(yd[i*c:(i+1)*c], non_blocking=True)
    return f
for k in (4, 8, 16, 32): print("chunks", k, t(chunked(k)))
