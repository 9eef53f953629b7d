"""parallel.P2PFrameBuffer on one GPU (a 1-rank NCCL group): construction in symmetric memory, the write / exchange / wait
protocol over both slots and the frame views the lift consumes.  The multi-rank behaviour (bit-equality with the NCCL
all-gather, bandwidth, overlap with a kernel that fills every SM) is what tools/p2p_probe.py checks under torchrun; the
bench's parity check runs on its frames at every N."""
import socket

import pytest
import torch
import torch.distributed as dist

from gennerf_b200 import synthetic as S

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_p2p_frame_buffer_single_rank():
    from gennerf_b200 import ops, parallel
    own_group = not dist.is_initialized()
    if own_group:
        try:
            dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{_port()}", world_size=1, rank=0,
                                    device_id=torch.device("cuda", 0))
        except Exception as e:                                             # noqa: BLE001
            pytest.skip(f"no 1-rank NCCL group on this box: {type(e).__name__}: {e}")
    try:
        T, B, C, H, W = 4, 1, 8, 12, 20
        try:
            fb = parallel.P2PFrameBuffer(T, B, C, H, W, torch.device("cuda", 0))
        except Exception as e:                                         # noqa: BLE001
            pytest.skip(f"symmetric memory is not available here: {type(e).__name__}: {e}")
        g = S.gen(3)
        for k in (0, 1, 0):
            frames = [torch.randn(B, C, H, W, generator=g).to(DEV) for _ in range(T)]
            ops.nchw_to_nhwc(frames, out=fb.own(k))
            fb.exchange(k)
            fb.wait(k)
            torch.cuda.synchronize()
            for t in range(T):
                assert torch.equal(fb.frames(k)[t], frames[t])
                assert fb.frames(k)[t].is_contiguous(memory_format=torch.channels_last)
        assert fb.owned == (0, T)
    finally:
        if own_group:
            dist.destroy_process_group()
