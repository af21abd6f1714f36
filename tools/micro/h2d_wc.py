#!/usr/bin/env python
"""H2D bandwidth from ordinary pinned memory against write-combined pinned memory (cudaHostAllocWriteCombined), 36 MB per copy."""
import ctypes
import torch

torch.cuda.init()
rt = ctypes.CDLL('libcudart.so.12')
rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
N = 36134912
dst = torch.empty(N, dtype=torch.uint8, device='cuda')

def wc_tensor(n, flags):
    p = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(p), n, flags) == 0
    return torch.frombuffer((ctypes.c_char * n).from_address(p.value), dtype=torch.uint8)

for name, src in (('pinned (torch)', torch.empty(N, dtype=torch.uint8).pin_memory()), ('cudaHostAlloc default', wc_tensor(N, 0)),
                  ('cudaHostAlloc write-combined', wc_tensor(N, 4))):
    src.fill_(1)
    print(name, 'is_pinned', src.is_pinned())
    for _ in range(5):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(f'  {ms:.4f} ms per 36.1 MB copy = {N / ms / 1e6:.1f} GB/s')
