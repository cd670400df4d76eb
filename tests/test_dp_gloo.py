"""N > 1 host logic on CPU with the gloo backend, world_size 2: sharding, the IoU-stat all-reduce and the
bit-packed mask all-gather.  No forward runs here."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from anyref_b200 import dp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dp.init_from_env("gloo")
    g = torch.Generator().manual_seed(0)
    pred = [(torch.rand(32, 40, generator=g) > 0.5).to(torch.uint8) for _ in range(n_items)]
    gt = [(torch.rand(32, 40, generator=g) > 0.5).to(torch.uint8) for _ in range(n_items)]
    lo, hi = dp.shard_range(n_items, rank, world)
    stats = dp.all_reduce_stats(dp.iou_stats(pred[lo:hi], gt[lo:hi]))
    packed = dp.pack_bits(torch.stack(pred[lo:hi])) if hi > lo else torch.zeros(0, dtype=torch.uint8)
    gathered = dp.all_gather_packed(packed)
    if rank == 0:
        full = dp.iou_stats(pred, gt)
        masks = torch.cat([dp.unpack_bits(p, (dp.shard_range(n_items, r, world)[1] - dp.shard_range(n_items, r, world)[0]) * 32 * 40)
                           for r, p in enumerate(gathered)])
        q.put((stats.tolist(), full.tolist(), bool(torch.equal(masks, torch.stack(pred).reshape(-1)))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [5, 8])
def test_stats_and_mask_gather_world2(n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, want, masks_ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == pytest.approx(want, rel=1e-6)
    assert masks_ok


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 1024):
        for w in (1, 2, 4, 8):
            spans = [dp.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))


def test_pack_bits_matches_numpy():
    import numpy as np

    m = (torch.rand(3, 17, 13) > 0.4).to(torch.uint8)
    assert np.array_equal(dp.pack_bits(m).numpy(), np.packbits(m.numpy().reshape(-1)))
    assert torch.equal(dp.unpack_bits(dp.pack_bits(m), m.numel()), m.reshape(-1))


def test_intersection_and_union_matches_reference_semantics():
    pred = torch.tensor([[0, 1, 1], [1, 0, 0]])
    gt = torch.tensor([[0, 1, 0], [1, 1, 0]])
    i, u, t = dp.intersection_and_union(pred, gt, 2)
    assert i.tolist() == [2.0, 2.0] and u.tolist() == [4.0, 4.0] and t.tolist() == [3.0, 3.0]


def test_gather_checksum_is_order_and_length_sensitive():
    """The checksum all_gather_packed verifies transfers with must notice what a corrupted or shifted piece looks like:
    a swapped pair of bytes, a dropped byte, a shift by one; and must not depend on alignment or length padding."""
    import torch

    from anyref_b200.dp import _checksum

    g = torch.Generator().manual_seed(0)
    a = torch.randint(0, 256, (100003,), dtype=torch.uint8, generator=g)
    base = int(_checksum(a))
    b = a.clone()
    b[[10, 11]] = b[[11, 10]]
    assert int(_checksum(b)) != base or int(a[10]) == int(a[11])
    assert int(_checksum(a[:-1])) != base and int(_checksum(a[1:])) != base
    c = a.clone()
    c[77777] ^= 1
    assert int(_checksum(c)) != base
    # a non-contiguous / unaligned view of the same bytes gives the same value
    padded = torch.zeros(a.numel() + 3, dtype=torch.uint8)
    padded[3:] = a
    assert int(_checksum(padded[3:])) == base
    assert int(_checksum(torch.zeros(0, dtype=torch.uint8))) == 0

