import torch, time
n = 16789506
h = torch.empty(n, dtype=torch.float64, pin_memory=True); h.fill_(1.0)
d = torch.empty(n, dtype=torch.float64, device='cuda')
print('pinned?', h.is_pinned())
def t(f, reps=10):
    for _ in range(2): f()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/reps
a = t(lambda: d.copy_(h, non_blocking=True)); b = t(lambda: h.copy_(d, non_blocking=True))
print(f"H2D {a:.2f} ms ({n*8/a/1e6:.1f} GB/s)  D2H {b:.2f} ms ({n*8/b/1e6:.1f} GB/s)")
