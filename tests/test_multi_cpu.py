"""world_size-2 gloo test of the N>1 host logic: screen-tile partition, compact buffers, gather,
assembly (hashdag_b200/partition.py mirrors hdt::PixelMap).  The per-pixel work is done by the CPU
oracle here; on the GPU box the same layout is exercised by tests/test_gpu_parity.py."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import get_scene, scene_cameras
from hashdag_b200 import camera, partition

W, H, TILE = 160, 96, 5


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, frame_path, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = np.load(frame_path)
        mine = partition.pack_compact(full, rank, world, TILE)      # what this rank's tracer would have rendered
        # pixels of other ranks' tiles never enter this rank's buffer
        owned = partition.owned_tiles(rank, world, W, H, TILE)
        assert len(owned) <= partition.max_tiles_per_rank(world, W, H, TILE)
        t = torch.from_numpy(mine.view(np.int32).copy())
        gathered = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, gathered, dst=0)
        if rank == 0:
            frame = partition.assemble([g.numpy().view(np.uint32) for g in gathered], W, H, TILE)
            np.save(out_path, frame)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_tile_partition_gather_reassembles_frame(tmp_path):
    from oracle import hdo
    s = get_scene(12, 10)
    cam = scene_cameras(s, 1, 10)[0]
    prm = camera.trace_params(cam, camera.DAGInfo(s.bounds_min, s.bounds_max), s.levels, W, H)
    d = hdo.make_dag(s, hdo.DAG_BASIC)
    p, _ = hdo.trace_paths(d, W, H, prm)
    c, _ = hdo.trace_colors(d, hdo.make_colors(s, hdo.COLORS_COMPRESSED), p)
    frame_path, out_path = str(tmp_path / "frame.npy"), str(tmp_path / "out.npy")
    np.save(frame_path, c)
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), frame_path, out_path), nprocs=world, join=True)
    assert np.array_equal(np.load(out_path), c)


def test_partition_covers_every_tile_once():
    for world in (1, 2, 3, 4, 8):
        seen = []
        for r in range(world):
            seen += partition.owned_tiles(r, world, 1920, 1080, 6)
        tx, ty = partition.tile_grid(1920, 1080, 6)
        assert sorted(seen) == list(range(tx * ty))
