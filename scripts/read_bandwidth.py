import torch, time
torch.cuda.init()
dev="cuda:0"
n=2_000_000_000
x=torch.randint(0,255,(n,),dtype=torch.uint8,device=dev)
xi=x.view(torch.int32)
xf=x.view(torch.float32)
def timeit(f,reps=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    ts=[]
    for _ in range(reps):
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts)//2], ts[0]
for name,f,bytes_ in [("sum int32 view (read 2 GB)", lambda: xi.sum(), n), ("max float view", lambda: xf.max(), n), ("sum u8", lambda: x.sum(dtype=torch.int64), n)]:
    med,best=timeit(f)
    print(f"{name}: median {med:.3f} ms -> {bytes_/med/1e6:.0f} GB/s, best {best:.3f} ms -> {bytes_/best/1e6:.0f} GB/s")
y=torch.empty_like(x)
med,best=timeit(lambda: y.copy_(x))
print(f"copy u8 2 GB (r+w 4 GB): median {med:.3f} ms -> {2*n/med/1e6:.0f} GB/s")
