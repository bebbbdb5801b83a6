"""The N > 1 path on CPU: world_size 2 over gloo.  Each rank packs the tiles it owns (as the device
render would), ONE all-gather collects them, un-tiling rebuilds the image."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as td
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, w, h, q):
    sys.path.insert(0, ROOT)
    from nrays_b200 import dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(123)          # same "frame" on every rank (the scene is replicated)
    img = rng.uniform(size=(h, w, 3)).astype(np.float32)
    local = torch.from_numpy(dist.pack_tiles_host(img, rank, world))
    gathered = dist.all_gather_tiles(local, world)
    full = dist.untile_host(gathered.numpy(), world, w, h)
    ok = np.array_equal(full, img)
    # every rank holds the complete frame after the single collective
    q.put((rank, bool(ok), tuple(gathered.shape)))
    td.destroy_process_group()


def test_two_rank_tile_gather_gloo():
    world, w, h = 2, 70, 45
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, w, h, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=60) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from nrays_b200 import dist

    assert res == [(0, True, (2, dist.tiles_per_rank(w, h, 2), 16, 16, 3)), (1, True, (2, dist.tiles_per_rank(w, h, 2), 16, 16, 3))]
