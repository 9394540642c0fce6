import torch, time
dev = torch.device("cuda:0")
for mb in (16.384, 64):
    n = int(mb * 1e6 / 4)
    h = torch.empty(n, dtype=torch.float32).pin_memory()
    d = torch.empty(n, dtype=torch.float32, device=dev)
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"H2D pinned {mb} MB: {ms:.3f} ms = {mb / ms:.1f} GB/s")
    e0.record()
    for _ in range(20): h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"D2H pinned {mb} MB: {ms:.3f} ms = {mb / ms:.1f} GB/s")
