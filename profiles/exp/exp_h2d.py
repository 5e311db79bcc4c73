import torch, time
x = torch.empty(130_000_000//4, dtype=torch.float32, pin_memory=True)
d = torch.empty_like(x, device='cuda')
s = torch.cuda.Stream()
for n in (1, 3):
    with torch.cuda.stream(s):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(10): d.copy_(x, non_blocking=True)
        e1.record(s)
    s.synchronize()
    print("H2D GB/s", 10*x.numel()*4/ (e0.elapsed_time(e1)*1e-3)/1e9)
