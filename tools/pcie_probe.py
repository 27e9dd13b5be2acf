import torch, time
n=787251200
d=torch.empty(n,dtype=torch.uint8,device='cuda')
h=torch.empty(n,dtype=torch.uint8).pin_memory()
for _ in range(2): h.copy_(d,non_blocking=True); torch.cuda.synchronize()
t=time.perf_counter()
for _ in range(5): h.copy_(d,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
print("D2H %.2f ms %.1f GB/s"%(dt*1e3, n/dt/1e9))
t=time.perf_counter()
for _ in range(5): d.copy_(h,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
print("H2D %.2f ms %.1f GB/s"%(dt*1e3, n/dt/1e9))
