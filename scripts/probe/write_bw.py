import torch
n = 1 << 30
t = torch.empty(n, device="cuda", dtype=torch.float32)
s = torch.empty(n, device="cuda", dtype=torch.float32).uniform_()
for name, fn, nbytes in (("fill_ (write only)", lambda: t.fill_(1.5), 4 * n), ("zero_ (memset)", lambda: t.zero_(), 4 * n), ("copy_ (read+write)", lambda: t.copy_(s), 8 * n),
                         ("sum (read only)", lambda: s.sum(), 4 * n)):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name}: {ms:.4f} ms  {nbytes / ms / 1e6:.0f} GB/s")
