"""Scratch: pinned host <-> device copy bandwidth on the box, at the sizes bench.py's end-to-end step moves."""
import torch
dev = torch.device("cuda", 0)
for mb in (1, 42, 126):
    n = mb * 1024 * 1024 // 4
    h = torch.empty(n, dtype=torch.float32).pin_memory()
    d = torch.empty(n, dtype=torch.float32, device=dev)
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            fn()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 5
        print(f"{name} {mb:4d} MB: {ms:7.3f} ms  {mb / 1024 / (ms * 1e-3):6.1f} GB/s")
