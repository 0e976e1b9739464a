"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: contiguous frame sharding and the compact BEV-token
gather.  The kernels are not involved: each rank fabricates its per-rank outputs and the test checks that the
destination rank reassembles the whole batch with global frame numbering."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lidar_vision_vqa_b200 import sharding


def test_shard_bounds_cover_and_balance():
    for n, w in ((16, 1), (16, 2), (16, 8), (17, 8), (3, 8), (0, 4), (256, 8)):
        b = sharding.shard_bounds(n, w)
        assert len(b) == w and b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1


def test_shard_points_rebases_offsets():
    offs = torch.tensor([0, 5, 5, 9, 20, 21], dtype=torch.int32)  # frame 1 is empty
    pts = torch.arange(21 * 5, dtype=torch.float32).view(21, 5)
    seen = []
    for r in range(2):
        p, lo_offs, base = sharding.shard_points(pts, offs, r, 2)
        assert lo_offs[0] == 0 and lo_offs[-1] == p.shape[0]
        seen.append(p)
        assert base == (0 if r == 0 else 3)
    assert torch.equal(torch.cat(seen), pts)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_frames, result_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharding.shard_bounds(n_frames, world)[rank]
        g = torch.Generator().manual_seed(100 + rank)
        per_frame = [int(torch.randint(0, 7, (1,), generator=g)) for _ in range(hi - lo)]
        if rank == 1:
            per_frame[-1] = 0  # a trailing empty frame
        m = sum(per_frame)
        feats = torch.randn(m, 64, generator=g)
        coords = torch.zeros(m, 4, dtype=torch.int32)
        coords[:, 0] = torch.repeat_interleave(torch.arange(hi - lo), torch.tensor(per_frame)).int()
        coords[:, 2] = torch.randint(0, 32, (m,), generator=g).int()
        coords[:, 3] = torch.randint(0, 32, (m,), generator=g).int()
        tok = sharding.gather_bev_tokens(feats, coords, frame_base=lo, n_frames_total=n_frames, dst=0)
        # every rank also publishes what it sent, for the check on rank 0
        sent = [None] * world
        dist.all_gather_object(sent, (feats.numpy(), coords.numpy(), lo))
        if rank == 0:
            assert tok is not None and tok.n_frames == n_frames
            exp_f = np.concatenate([s[0] for s in sent])
            exp_c = np.concatenate([s[1] + np.array([s[2], 0, 0, 0], np.int32) for s in sent])
            ok = (np.array_equal(tok.pillar_features.numpy(), exp_f) and np.array_equal(tok.voxel_coords.numpy(), exp_c)
                  and tok.pillars_per_rank == [len(s[0]) for s in sent])
            with open(result_path, "w") as f:
                f.write("ok" if ok else "mismatch")
        else:
            assert tok is None
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gather_bev_tokens_world2_gloo(tmp_path):
    result = tmp_path / "result.txt"
    mp.spawn(_worker, args=(2, _free_port(), 7, str(result)), nprocs=2, join=True)
    assert result.read_text() == "ok"


def test_gather_without_process_group_is_identity():
    feats = torch.randn(5, 64)
    coords = torch.zeros(5, 4, dtype=torch.int32)
    tok = sharding.gather_bev_tokens(feats, coords, frame_base=3, n_frames_total=4)
    assert torch.equal(tok.pillar_features, feats) and int(tok.voxel_coords[0, 0]) == 3


def test_numa_binding_is_best_effort():
    """No NVML / no GPU here: the helper must report None (or a CPU list) and never raise or change the affinity to nothing."""
    import os

    from lidar_vision_vqa_b200 import sharding

    before = os.sched_getaffinity(0)
    got = sharding.bind_to_gpu_numa_node(0)
    after = os.sched_getaffinity(0)
    assert got is None or (len(got) > 0 and set(got) <= before)
    assert len(after) > 0
    os.sched_setaffinity(0, before)


def _gatherer_worker(rank, world, port, result_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rows, f, frames = 40, 8, 3
        gat = sharding.TokenGatherer(rows, f, frames, "cpu", dst=0, slots=2)
        a, b, c = sharding.wire_layout(rows, f, frames)
        ok = True
        for step in range(3):  # three steps through two slots, different counts every step
            g = torch.Generator().manual_seed(10 * step + rank)
            per_frame = torch.randint(0, 9, (frames,), generator=g).int()
            m = int(per_frame.sum())
            # the encoder's slab: features | coords | counts, rows beyond the count hold stale garbage
            wire = torch.zeros(a + b + c, dtype=torch.uint8)
            feats = wire[:a].view(torch.float32).view(rows, f)
            coords = wire[a:a + b].view(torch.int32).view(rows, 4)
            count = wire[a + b:a + b + 4 * (frames + 1)].view(torch.int32)
            feats.copy_(torch.randn(rows, f, generator=g))
            coords.copy_(torch.randint(0, 50, (rows, 4), generator=g).int())
            coords[:m, 0] = torch.repeat_interleave(torch.arange(frames), per_frame.long()).int()
            count.copy_(torch.cat([per_frame, per_frame.sum(dim=0, keepdim=True)]).int())
            slot = gat.exchange(wire)
            sent = [None] * world
            dist.all_gather_object(sent, (feats.numpy().copy(), coords.numpy().copy(), count.numpy().copy()))
            if rank == 0:
                for r in range(world):
                    ok &= np.array_equal(slot["feats"][r].numpy(), sent[r][0])
                    ok &= np.array_equal(slot["coords"][r].numpy(), sent[r][1])  # (rebasing is a device kernel)
                    ok &= np.array_equal(slot["counts"][r].numpy(), sent[r][2])
        flags = [None] * world
        dist.all_gather_object(flags, bool(ok))
        if rank == 0:
            with open(result_path, "w") as fh:
                fh.write("ok" if all(flags) else "mismatch")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_token_gatherer_world2_gloo(tmp_path):
    """The steady-state gatherer on two gloo ranks: one slab per rank and step into fixed segments, counts travelling with
    the payload, slots reused across steps (rebasing and the overflow flag are a device kernel: GPU tests)."""
    result = tmp_path / "result.txt"
    mp.spawn(_gatherer_worker, args=(2, _free_port(), str(result)), nprocs=2, join=True)
    assert result.read_text() == "ok"
