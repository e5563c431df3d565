import torch, time
n = 1 << 30  # bytes
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both(): h2d(); d2h()
print(f"PCIe pinned 1 GiB: H2D {n/t(h2d)/1e9:.1f} GB/s, D2H {n/t(d2h)/1e9:.1f} GB/s, both at once {n/t(both)/1e9:.1f} GB/s each way")
