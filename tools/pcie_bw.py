import torch, time
x = torch.empty(1<<30, dtype=torch.uint8).pin_memory(); y = torch.empty(1<<30, dtype=torch.uint8, device='cuda')
for sz in (1<<30, 100<<20, 8<<20):
    a = x[:sz]; b = y[:sz]
    b.copy_(a, non_blocking=True); torch.cuda.synchronize()
    t=time.perf_counter(); n=max(3, (2<<30)//sz)
    for _ in range(n): b.copy_(a, non_blocking=True)
    torch.cuda.synchronize(); dt=time.perf_counter()-t
    print('H2D', sz>>20, 'MB', sz*n/dt/1e9, 'GB/s')
    t=time.perf_counter()
    for _ in range(n): a.copy_(b, non_blocking=True)
    torch.cuda.synchronize(); dt=time.perf_counter()-t
    print('D2H', sz>>20, 'MB', sz*n/dt/1e9, 'GB/s')
# both directions at once
s1=torch.cuda.Stream(); s2=torch.cuda.Stream(); x2=torch.empty(1<<30, dtype=torch.uint8).pin_memory(); y2=torch.empty(1<<30, dtype=torch.uint8, device='cuda')
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(4):
    with torch.cuda.stream(s1): y.copy_(x, non_blocking=True)
    with torch.cuda.stream(s2): x2.copy_(y2, non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t
print('bidir each', 4*(1<<30)/dt/1e9, 'GB/s')
