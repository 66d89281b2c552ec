import torch, statistics, time
x = torch.zeros(32, device="cuda")
burn = torch.randn(4096, 4096, device="cuda")
t0=time.perf_counter()
while time.perf_counter()-t0<1.0: burn@burn
torch.cuda.synchronize()
for inner in (1, 10, 50):
    g = torch.cuda.CUDAGraph()
    x.add_(1); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(inner): x.add_(1)
    ts=[]
    for _ in range(30):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b)*1e3/inner)
    print("inner",inner,"us per tiny kernel", round(statistics.median(ts),2))
