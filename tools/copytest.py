import torch
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for mb in (8, 25, 50, 100, 400):
    a = torch.empty(mb << 18, dtype=torch.float32, device=dev).normal_()
    b = torch.empty_like(a)
    for _ in range(2):
        flush.zero_()
        b.copy_(a)
torch.cuda.synchronize()
