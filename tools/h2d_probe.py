import torch, time
x = torch.empty(256*1024*1024//4, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
for n in (1,2):
    streams=[torch.cuda.Stream() for _ in range(n)]
    torch.cuda.synchronize()
    t0=time.perf_counter()
    for r in range(4):
        for i,s in enumerate(streams):
            with torch.cuda.stream(s):
                k = x.numel()//n
                d[i*k:(i+1)*k].copy_(x[i*k:(i+1)*k], non_blocking=True)
    torch.cuda.synchronize()
    el=time.perf_counter()-t0
    print("H2D streams=%d: %.1f GB/s"%(n, 4*x.numel()*4/el/1e9))
t0=time.perf_counter()
for r in range(4): x.copy_(d, non_blocking=True)
torch.cuda.synchronize(); print("D2H %.1f GB/s"%(4*x.numel()*4/(time.perf_counter()-t0)/1e9))
