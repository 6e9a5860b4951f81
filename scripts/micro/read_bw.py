"""Read-only streaming bandwidth on this GPU (reference point for the marching-cubes classify pass):
torch.sum / torch.max over a 512^3 float32 volume, L2 flushed in between."""
import torch
dev = torch.device("cuda", 0)
x = torch.rand(512, 512, 512, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
for name, fn in (("sum", lambda: x.sum()), ("max", lambda: x.max()), ("gt.any", lambda: (x > 0.5).any()), ("copy", lambda: x.clone())):
    fn(); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    tot = 0.0
    for _ in range(5):
        flush.fill_(0.0)
        ev[0].record(); fn(); ev[1].record(); torch.cuda.synchronize()
        tot += ev[0].elapsed_time(ev[1])
    ms = tot / 5
    print("%-7s %.3f ms  %.0f GB/s read" % (name, ms, x.numel() * 4 / ms / 1e6))
