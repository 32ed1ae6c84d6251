import torch, time
d=torch.device('cuda:0')
for mb in [2,8,32,128,512]:
    n=mb*1024*1024
    src=torch.empty(n,dtype=torch.uint8,device=d); dst=torch.empty(n,dtype=torch.uint8).pin_memory()
    for _ in range(3): dst.copy_(src,non_blocking=True)
    torch.cuda.synchronize()
    reps=max(4,int(2048/mb))
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): dst.copy_(src,non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(mb,'MB D2H', round(n*reps/e0.elapsed_time(e1)/1e6,1),'GB/s')
    e0.record()
    for _ in range(reps): src.copy_(dst,non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(mb,'MB H2D', round(n*reps/e0.elapsed_time(e1)/1e6,1),'GB/s')
