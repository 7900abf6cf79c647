"""Micro-probe: does a freshly written buffer of `mb` MB stay in L2 until it is read back, with `noise` MB of
unrelated streaming reads in between?  Run under ncu and look at dram__bytes_read.sum of the reduce kernels."""
import sys
import torch
mb, noise = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device('cuda')
n = mb * (1 << 20) // 4
src = torch.randn(n, device=dev)
buf = torch.empty(n, device=dev)
other = torch.randn(max(noise, 1) * (1 << 20) // 4, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for it in range(3):
    flush.zero_()
    torch.cuda.synchronize()
    buf.copy_(src)                 # write `mb` MB (reads `mb` MB of src)
    if noise:
        other.max()                # unrelated streaming reads
    r = buf.sum()                  # read the buffer back: L2 hit or DRAM?
torch.cuda.synchronize()
print('done', float(r))
