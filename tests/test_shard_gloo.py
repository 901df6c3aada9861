"""N > 1 path on CPU: world-size-2 gloo processes shard the channels, run their shard (the oracle stands in
for the GPU engine here: this test is about the partitioning and the gather, not the kernels) and gather to
rank 0; the result must equal the unsharded run bit for bit (channels are independent)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dsp_stuff_b200 import signals as S
from dsp_stuff_b200.shard import channel_range, gather_to_rank0


def test_channel_range_partitions_exactly():
    for total in (0, 1, 7, 8, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [channel_range(r, world, total) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        channel_range(2, 2, 8)


def _worker(rank, world, port, total, n, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle

    lo, hi = channel_range(rank, world, total)
    x = S.noise(hi - lo, n, channel_offset=lo)        # the same rows the unsharded run sees
    o = oracle.Oracle(hi - lo, threads=1)
    S.config3().apply(o)
    y = torch.from_numpy(o.process(x)[0])
    full = gather_to_rank0(y, total)
    if rank == 0:
        np.save(out_path, full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather_equals_single_run(tmp_path, oracle_mod):
    total, n = 7, 128 * 24                             # uneven split: 4 + 3 channels
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, port, total, n, out), nprocs=2, join=True)
    o = oracle_mod.Oracle(total, threads=1)
    S.config3().apply(o)
    ref = o.process(S.noise(total, n))[0]
    assert np.array_equal(np.load(out), ref)
