"""Probe of the symmetric-memory frame exchange (parallel.P2PFrameBuffer) under torchrun: correctness against an NCCL
all-gather, time of one exchange, and whether it overlaps a kernel that fills every SM."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gennerf_b200 import parallel  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
T, B, C, H, W = 32, 1, 32, 480, 640
fb = parallel.P2PFrameBuffer(T, B, C, H, W, dev)
ref = parallel.FrameBuffer(T, B, C, H, W, dev)
t0, t1 = fb.owned
ok = True
for k in (0, 1, 0, 1):
    g = torch.Generator(device=dev).manual_seed(100 * k + rank)
    mine = torch.randn((t1 - t0, B, H, W, C), device=dev, generator=g)
    fb.own(k).copy_(mine)
    ref.flat[t0:t1].copy_(mine)
    ref.all_gather()
    fb.exchange(k)
    fb.wait(k)
    torch.cuda.synchronize()
    ok &= bool(torch.equal(fb.slots[k]["flat"], ref.flat))
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
res = {}
for name, fn in (("p2p", lambda k: (fb.exchange(k), fb.wait(k))), ("nccl", lambda k: ref.all_gather())):
    for k in range(3):
        fn(k % 2)
    dist.barrier()
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for k in range(10):
        fn(k % 2)
    b.record()
    torch.cuda.synchronize()
    res[name] = a.elapsed_time(b) / 10
# overlap: a kernel that occupies every SM for ~10 ms (matmul chain) while the exchange runs on the copy stream
x = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
def busy():
    y = x
    for _ in range(12):
        y = y @ x
    return y
busy(); torch.cuda.synchronize()
a, b = ev(), ev()
a.record(); busy(); b.record(); torch.cuda.synchronize()
res["busy_alone"] = a.elapsed_time(b)
dist.barrier(); torch.cuda.synchronize()
a, b = ev(), ev()
a.record(); fb.exchange(0); busy(); fb.wait(0); b.record(); torch.cuda.synchronize()
res["busy_with_p2p_exchange"] = a.elapsed_time(b)
dist.barrier(); torch.cuda.synchronize()
a, b = ev(), ev()
a.record(); busy(); ref.all_gather(); b.record(); torch.cuda.synchronize()
res["busy_then_nccl"] = a.elapsed_time(b)
if rank == 0:
    gb = T * B * H * W * C * 4 * (world - 1) / world / 1e9
    print(f"world {world}: equal to NCCL all-gather: {ok}; per-exchange ms {res}; received {gb:.2f} GB per GPU -> "
          f"p2p {gb / res['p2p'] * 1e3:.0f} GB/s, nccl {gb / res['nccl'] * 1e3:.0f} GB/s")
dist.destroy_process_group()
