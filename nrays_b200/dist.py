"""Tile sharding across GPUs and the single gather of finished tiles.

One process per GPU.  The scene is replicated; the image is cut into 16x16-pixel tiles and tile t
belongs to rank t mod world (the reference's contiguous pixel ranges, src/scene.rs:61-63, would be
badly balanced between sky rows and geometry rows).  Each rank renders its tiles into a packed
buffer [tiles_per_rank][16][16][3]; ONE all-gather (NCCL over NVLink on GPUs, gloo in the CPU tests)
collects them and an un-tile pass rebuilds the row-major image.  The RNG is keyed by the global
pixel index, so the image does not depend on the sharding.
"""
import ctypes as C

import numpy as np

from . import _abi as A

TILE = A.NRB_TILE


def tile_grid(width, height):
    return (width + TILE - 1) // TILE, (height + TILE - 1) // TILE


def tile_count(width, height):
    tx, ty = tile_grid(width, height)
    return tx * ty


def tiles_per_rank(width, height, world):
    """Packed-buffer length every rank allocates (equal sizes so one all-gather suffices)."""
    return (tile_count(width, height) + world - 1) // world


def local_tiles(width, height, rank, world):
    return list(range(rank, tile_count(width, height), world))


def pack_tiles_host(image, rank, world):
    """Host reference of the packed layout: image (H,W,3) -> (tiles_per_rank,16,16,3), zero padded."""
    h, w, _ = image.shape
    tx, _ty = tile_grid(w, h)
    out = np.zeros((tiles_per_rank(w, h, world), TILE, TILE, 3), dtype=image.dtype)
    for lt, t in enumerate(local_tiles(w, h, rank, world)):
        y0, x0 = (t // tx) * TILE, (t % tx) * TILE
        blk = image[y0:y0 + TILE, x0:x0 + TILE]
        out[lt, :blk.shape[0], :blk.shape[1]] = blk
    return out


def untile_host(gathered, world, width, height):
    """Host reference of the un-tile pass: (world, tiles_per_rank,16,16,3) -> (H,W,3)."""
    tx, _ty = tile_grid(width, height)
    g = np.asarray(gathered).reshape(world, -1, TILE, TILE, 3)
    img = np.zeros((height, width, 3), dtype=g.dtype)
    for t in range(tile_count(width, height)):
        r, lt = t % world, t // world
        y0, x0 = (t // tx) * TILE, (t % tx) * TILE
        hh, ww = min(TILE, height - y0), min(TILE, width - x0)
        img[y0:y0 + hh, x0:x0 + ww] = g[r, lt, :hh, :ww]
    return img


def all_gather_tiles(local, world):
    """The one collective of the path.  `local`: torch tensor (tiles_per_rank,16,16,3) on this rank."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return local.reshape(1, *local.shape)
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out.reshape((world,) + tuple(local.shape))


# ---- device side --------------------------------------------------------------------------------
def render_device(scene, cam, out):
    """nrb_render_device into a torch CUDA tensor `out` (H*W*3 floats)."""
    from . import _lib

    stats = A.NrbStats()
    _lib.check(_lib.load().nrb_render_device(scene.handle, C.byref(cam), C.c_void_p(out.data_ptr()), C.byref(stats)))
    return stats


def render_tiles_device(scene, cam, rank, world, out):
    """Render this rank's tiles into the packed CUDA tensor `out` (tiles_per_rank*16*16*3 floats)."""
    from . import _lib

    ts = A.NrbTileSet(rank, world)
    n_local = C.c_uint32()
    stats = A.NrbStats()
    _lib.check(_lib.load().nrb_render_tiles_device(scene.handle, C.byref(cam), C.byref(ts), C.c_void_p(out.data_ptr()),
                                                   C.byref(n_local), C.byref(stats)))
    return stats, n_local.value


def untile_device(gathered, world, width, height, out, device=0, stream=None):
    from . import _lib

    tpr = tiles_per_rank(width, height, world)
    _lib.check(_lib.load().nrb_untile_device(int(device), C.c_void_p(stream or 0), C.c_void_p(gathered.data_ptr()), world,
                                             tpr, width, height, C.c_void_p(out.data_ptr())))


def use_stream(scene, stream_ptr):
    from . import _lib

    _lib.check(_lib.load().nrb_scene_set_stream(scene.handle, C.c_void_p(stream_ptr or 0)))
