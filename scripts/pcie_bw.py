import torch, time
for mb in (32, 256):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device='cuda')
    for name, fn in (('h2d', lambda: d.copy_(h, non_blocking=True)), ('d2h', lambda: h.copy_(d, non_blocking=True))):
        fn(); torch.cuda.synchronize(); t=time.perf_counter()
        for _ in range(10): fn()
        torch.cuda.synchronize(); dt=(time.perf_counter()-t)/10
        print(mb, 'MB', name, '%.1f GB/s' % (n/dt/1e9))
