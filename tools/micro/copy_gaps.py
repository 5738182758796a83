"""What costs PCIe time between two H2D copies of one stream?  1.8 MB pinned -> device copies, 400 in a row:
(a) back to back, (b) an event record after every copy, (c) a wait on an (already signalled) event of another
stream before every copy, (d) both, (e) like (d) on two alternating copy streams, (f) pairs of batches per copy."""
import torch
n, nbytes = 400, 1799582 // 4 * 4
host = [torch.empty(nbytes // 4, dtype=torch.float32).pin_memory() for _ in range(64)]
host2 = [torch.empty(nbytes // 2, dtype=torch.float32).pin_memory() for _ in range(32)]
dev = [torch.empty(nbytes // 4, dtype=torch.float32, device='cuda') for _ in range(8)]
dev2 = [torch.empty(nbytes // 2, dtype=torch.float32, device='cuda') for _ in range(4)]
cs, cs2, other = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
recs = [torch.cuda.Event() for _ in range(8)]
sig = [torch.cuda.Event() for _ in range(8)]
for e in sig:
    e.record(other)
torch.cuda.synchronize()


def run(mode):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record(cs)
    if mode == 'f':
        for i in range(n // 2):
            with torch.cuda.stream(cs):
                cs.wait_event(sig[i % 4])
                dev2[i % 4].copy_(host2[i % 32], non_blocking=True)
                recs[i % 4].record(cs)
    else:
        for i in range(n):
            st = cs2 if (mode == 'e' and i & 1) else cs
            with torch.cuda.stream(st):
                if mode in 'cde':
                    st.wait_event(sig[i % 8])
                dev[i % 8].copy_(host[i % 64], non_blocking=True)
                if mode in 'bde':
                    recs[i % 8].record(st)
    cs.wait_stream(cs2)
    e.record(cs)
    torch.cuda.synchronize()
    us = 1e3 * s.elapsed_time(e) / n
    print('(%s) %.1f us per 1.8 MB batch = %.1f GB/s' % (mode, us, nbytes / us / 1e3))


for m in 'abcdef':
    run(m)
    run(m)
