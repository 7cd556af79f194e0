import torch,time
x=torch.empty(8*1024**3,dtype=torch.uint8,device='cuda')
for _ in range(2): x.zero_()
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): x.zero_()
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/5
print("fill 8 GiB: %.3f ms -> %.2f TB/s"%(ms, 8*1024**3/ms/1e9))
y=torch.empty_like(x)
for _ in range(2): y.copy_(x)
e0.record()
for _ in range(5): y.copy_(x)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/5
print("copy 8 GiB: %.3f ms -> %.2f TB/s (r+w)"%(ms, 2*8*1024**3/ms/1e9))
s=torch.zeros(1,device='cuda')
e0.record()
for _ in range(5): s=s+x[:4*1024**3].view(torch.int32).sum()
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/5
print("read 4 GiB: %.3f ms -> %.2f TB/s"%(ms, 4*1024**3/ms/1e9))
