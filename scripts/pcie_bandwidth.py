import torch, time
dev=torch.device("cuda",0)
for mb in (8.6, 69, 276):
    n=int(mb*1e6/8)
    d=torch.empty(n,dtype=torch.float64,device=dev).normal_()
    h=torch.empty(n,dtype=torch.float64).pin_memory()
    for _ in range(3): h.copy_(d,non_blocking=True)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): h.copy_(d,non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    print(f"D2H {mb} MB pinned: {ms:.3f} ms  {mb/ms:.1f} GB/s")
    e0.record()
    for _ in range(10): d.copy_(h,non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    print(f"H2D {mb} MB pinned: {ms:.3f} ms  {mb/ms:.1f} GB/s")
