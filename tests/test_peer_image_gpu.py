"""PeerImage (the fused multi-GPU exchange) across real process boundaries: two ranks, CUDA IPC, two DIFFERENT frames
back to back through the fallback end-to-end recipe (owner copies the device image to the host after every frame).

The driver's GPU test box has one GPU, so both ranks share cuda:0 (CUDA IPC works between processes on one device) and
the process group is gloo; PeerImage then runs its ordering collectives on host tensors after draining the stream.
What is checked: writers -> reader ordering (sync) and reader -> writers ordering (release): the host
copy of frame i must be frame i even though every rank goes straight on to frame i+1.
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C

    import torch
    import torch.distributed as td

    from nrays_b200 import _abi as A, _lib, configs, dist, make_camera

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    lib = _lib.load()
    scene, camd, _cfg = configs.build("C3", target_tris=30000, lod=4)
    w, h = 208, 112
    peer = dist.PeerImage(w, h, rank, world, 0)
    host = torch.empty(h * w * 3, dtype=torch.float32).pin_memory() if rank == 0 else None
    frames = []
    for step, bg in enumerate(((1.0, 1.0, 1.0), (0.0, 0.2, 0.9), (0.5, 0.0, 0.0))):
        scene.set_background(bg)
        cam = make_camera(w, h, 2, 1.0, camd.eye, camd.projection((w, h)), seed=step)
        dist.render_tiles_to_image(scene, cam, rank, world, peer.ptr)
        peer.sync()                      # all ranks' stores are in the owner's image
        if rank == 0:
            host.copy_(peer.tensor(), non_blocking=True)   # the owner's read of frame `step` ...
        peer.release()                   # ... is ordered before anyone's stores of frame `step + 1`
        if rank == 0:
            torch.cuda.synchronize()
            frames.append(host.numpy().copy())
    if rank == 0:
        # the same three frames rendered unsharded by this process
        want = []
        for step, bg in enumerate(((1.0, 1.0, 1.0), (0.0, 0.2, 0.9), (0.5, 0.0, 0.0))):
            scene.set_background(bg)
            cam = make_camera(w, h, 2, 1.0, camd.eye, camd.projection((w, h)), seed=step)
            full = np.empty(w * h * 3, np.float32)
            _lib.check(lib.nrb_render(scene.handle, C.byref(cam), full.ctypes.data_as(C.POINTER(C.c_float)), None))
            want.append(full)
        np.save(os.path.join(out_dir, "got.npy"), np.stack(frames))
        np.save(os.path.join(out_dir, "want.npy"), np.stack(want))
    td.barrier()
    peer.close()
    scene.close()
    td.destroy_process_group()


def test_peer_image_two_processes_back_to_back_frames(gpu, tmp_path):
    import torch.multiprocessing as mp

    port = 29600 + (os.getpid() % 300)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got, want = np.load(tmp_path / "got.npy"), np.load(tmp_path / "want.npy")
    assert got.shape == want.shape == (3, 208 * 112 * 3)
    for k in range(3):
        np.testing.assert_allclose(got[k], want[k], rtol=0, atol=3e-5, err_msg="frame %d" % k)
    assert np.abs(want[0] - want[1]).max() > 0.1   # the frames really differ
