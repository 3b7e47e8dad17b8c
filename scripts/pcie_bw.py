"""Host <-> device copy rates with pinned memory: one direction at a time, both at once on two streams (what the
host-buffer pipelines want), and both at once while host threads read memory (the packers)."""
import threading
import time

import numpy as np
import torch

for mb in (32, 256):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device='cuda')
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device='cuda')
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def h2d():
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
    def d2h():
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
    def both():
        h2d(); d2h()
    for name, fn, factor in (('h2d', h2d, 1), ('d2h', d2h, 1), ('h2d + d2h at once (each)', both, 1)):
        fn(); torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(10):
            fn()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
        print(mb, 'MB', name, '%.1f GB/s' % (factor * n / dt / 1e9), flush=True)
    # the same while 8 host threads stream through 1 GB of ordinary memory
    big = np.ones(1 << 27, dtype=np.uint64)
    stop = False
    def reader(k):
        while not stop:
            big[k::8][: 1 << 22].sum()
    threads = [threading.Thread(target=reader, args=(k,)) for k in range(8)]
    for t_ in threads:
        t_.start()
    both(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(10):
        both()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
    stop = True
    for t_ in threads:
        t_.join()
    print(mb, 'MB', 'h2d + d2h at once with 8 host threads reading (each)', '%.1f GB/s' % (n / dt / 1e9), flush=True)
