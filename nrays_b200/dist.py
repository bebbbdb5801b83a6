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


# ---- fused exchange: every rank resolves its tiles straight into the owner's image over NVLink --------------
class _DevArray:
    """Minimal __cuda_array_interface__ carrier so torch can view library-owned device memory."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class PeerImage:
    """Row-major W*H*3 float image in the memory of rank `owner`, mapped into every rank of the node with CUDA IPC.

    Replaces gather + un-tile: `render_tiles_to_image` makes each rank's resolve kernel store its finished pixels
    into this image directly (peer stores over NVLink / NVSwitch), so each pixel crosses the fabric once and only
    the owner receives anything; `sync()` (one 4-byte all-reduce on the render stream) orders the ranks before the
    owner reads.  Collective constructor: call it on every rank of the default process group.
    """

    def __init__(self, width, height, rank, world, device, owner=0):
        import torch
        import torch.distributed as td

        from . import _lib

        self.lib, self.rank, self.world, self.device, self.owner = _lib.load(), rank, world, int(device), owner
        self.n_floats = width * height * 3
        self.ptr = C.c_void_p()
        # NCCL: collectives on CUDA tensors, ordered on the render stream.  Any other backend (gloo in the tests): the
        # stream is drained first and the collective runs on host tensors — same ordering, paid with a host sync.
        self._on_device = td.get_backend() == "nccl"
        self._cdev = ("cuda:%d" % self.device) if self._on_device else "cpu"
        handle = A.NrbIpcHandle()
        ok = True
        if rank == owner:
            ok = self.lib.nrb_ipc_alloc(self.device, self.n_floats * 4, C.byref(self.ptr), C.byref(handle)) == A.NRB_OK
        hb = torch.tensor(list(bytes(handle.bytes)) + [1 if ok else 0], dtype=torch.uint8, device=self._cdev)
        td.broadcast(hb, src=owner)
        raw = hb.cpu().numpy().tobytes()
        if not raw[64]:
            raise RuntimeError("PeerImage: the owner could not allocate / export the image")
        if rank != owner:
            C.memmove(handle.bytes, raw[:64], 64)
            ok = self.lib.nrb_ipc_open(self.device, C.byref(handle), C.byref(self.ptr)) == A.NRB_OK
        # every rank must know whether every rank has the mapping
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self._cdev)
        td.all_reduce(flag, op=td.ReduceOp.MIN)
        self._flag = torch.zeros(1, dtype=torch.int32, device=self._cdev)
        if int(flag.item()) == 0:
            self.close()
            raise RuntimeError("PeerImage: CUDA IPC mapping failed on some rank: " + self.lib.nrb_last_error().decode("utf-8", "replace"))

    def tensor(self):
        """torch view of the image (owner only)."""
        import torch

        assert self.rank == self.owner
        return torch.as_tensor(_DevArray(self.ptr.value, self.n_floats), device="cuda:%d" % self.device)

    def _fence(self):
        import torch
        import torch.distributed as td

        if not self._on_device:
            torch.cuda.synchronize(self.device)
        td.all_reduce(self._flag)

    def sync(self):
        """WRITERS -> READER: orders all ranks' peer stores before whatever the owner enqueues next on its current stream."""
        self._fence()

    def release(self):
        """READER -> WRITERS: call on every rank after the owner has enqueued its reads of this frame (copy to the host,
        encode, ...).  The next frame's peer stores of every rank are ordered behind those reads; without it rank r may
        be storing frame i+1 into the image while the owner still copies frame i (write-after-read race)."""
        self._fence()

    def close(self):
        if self.ptr:
            if self.rank == self.owner:
                self.lib.nrb_ipc_free(self.device, self.ptr)
            else:
                self.lib.nrb_ipc_close(self.device, self.ptr)
            self.ptr = C.c_void_p()


def render_tiles_to_image(scene, cam, rank, world, image_ptr):
    """Render this rank's tiles and resolve them into the row-major image at `image_ptr` (local or peer memory)."""
    from . import _lib

    ts = A.NrbTileSet(rank, world)
    stats = A.NrbStats()
    ptr = image_ptr if isinstance(image_ptr, C.c_void_p) else C.c_void_p(int(image_ptr))
    _lib.check(_lib.load().nrb_render_tiles_to_image(scene.handle, C.byref(cam), C.byref(ts), ptr, C.byref(stats)))
    return stats



def render_tiles_to_image_rgb8(scene, cam, rank, world, image_ptr):
    """RGB8 form of render_tiles_to_image: W*H*3 BYTES at `image_ptr`, quantised like Image::to_png (src/image.rs:64-77)."""
    from . import _lib

    ts = A.NrbTileSet(rank, world)
    stats = A.NrbStats()
    ptr = image_ptr if isinstance(image_ptr, C.c_void_p) else C.c_void_p(int(image_ptr))
    _lib.check(_lib.load().nrb_render_tiles_to_image_rgb8(scene.handle, C.byref(cam), C.byref(ts), ptr, C.byref(stats)))
    return stats


def render_tiles_to_host_rgb8(scene, cam, rank, world, host_addr):
    """RGB8 form of render_tiles_to_host: the rank's tile columns as 48-byte segments, one strided 2-D DMA."""
    from . import _lib

    ts = A.NrbTileSet(rank, world)
    stats = A.NrbStats()
    _lib.check(_lib.load().nrb_render_tiles_to_host_rgb8(scene.handle, C.byref(cam), C.byref(ts), C.c_void_p(int(host_addr)), C.byref(stats)))
    return stats


# ---- end-to-end exchange: N ranks fill ONE shared pinned host image through their own PCIe links ------------------
class SharedHostImage:
    """Row-major W*H*3 float image in POSIX shared memory, mapped by every rank of the node and registered with CUDA in
    each process (nrb_host_register).  `render_tiles_to_host` drops each rank's tile columns into it with one strided 2-D
    DMA, so the host image — the thing scene::render returns — is filled through N PCIe links in parallel instead of
    25 MB funnelling through rank 0's link.  Collective constructor (default process group); after `sync()` every rank can
    read `array`."""

    def __init__(self, width, height, rank, world, device):
        import torch
        import torch.distributed as td
        from multiprocessing import shared_memory

        from . import _lib

        self.lib, self.rank, self.world, self.device = _lib.load(), rank, world, int(device)
        self.nbytes = width * height * 3 * 4
        self.shm, self.array, self._cbuf, self._registered, self.addr = None, None, None, False, 0
        self._on_device = td.get_backend() == "nccl"
        cdev = ("cuda:%d" % self.device) if self._on_device else "cpu"
        name = [None]
        if rank == 0:
            try:  # a failure here (e.g. /dev/shm full) must still reach the broadcast, or the other ranks wait forever
                self.shm = shared_memory.SharedMemory(create=True, size=self.nbytes)
                name[0] = self.shm.name
            except Exception:
                name[0] = None
        td.broadcast_object_list(name, src=0)
        ok = name[0] is not None
        try:
            if ok and rank != 0:
                self.shm = shared_memory.SharedMemory(name=name[0])
                try:  # only the owner unlinks the segment; keep this process's resource tracker out of it
                    from multiprocessing import resource_tracker

                    resource_tracker.unregister(self.shm._name, "shared_memory")
                except Exception:
                    pass
            if ok:
                self.array = np.ndarray((width * height * 3,), dtype=np.float32, buffer=self.shm.buf)
                self._cbuf = C.c_char.from_buffer(self.shm.buf)
                self.addr = C.addressof(self._cbuf)
                dptr = C.c_void_p()
                ok = self.lib.nrb_host_register(self.device, C.c_void_p(self.addr), self.nbytes, C.byref(dptr)) == A.NRB_OK
                self._registered = ok
        except Exception:
            ok = False
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=cdev)
        td.all_reduce(flag, op=td.ReduceOp.MIN)
        self._flag = torch.zeros(1, dtype=torch.int32, device=cdev)
        if int(flag.item()) == 0:
            self.close()
            raise RuntimeError("SharedHostImage: shared memory / cudaHostRegister failed on some rank")

    def sync(self):
        """All ranks' DMAs have landed when this all-reduce completes (each rank's render call returns after its own)."""
        import torch
        import torch.distributed as td

        if not self._on_device:
            torch.cuda.synchronize(self.device)
        td.all_reduce(self._flag)

    def close(self):
        if self._registered:
            self.lib.nrb_host_unregister(self.device, C.c_void_p(self.addr))
            self._registered = False
        self.array = None
        self._cbuf = None
        if self.shm is not None:
            try:
                self.shm.close()  # raises BufferError while a caller still holds a view of `array`
            except Exception:
                pass
            if self.rank == 0:
                try:  # on its own: the segment must go away whatever close() did
                    self.shm.unlink()
                except Exception:
                    pass
            self.shm = None


def render_tiles_to_host(scene, cam, rank, world, host_addr):
    """Render this rank's tile columns and DMA them into the row-major HOST image at `host_addr` (registered memory).
    Raises NraysError(NRB_ERR_UNSUPPORTED) when the frame geometry does not give every rank whole tile columns."""
    from . import _lib

    ts = A.NrbTileSet(rank, world)
    stats = A.NrbStats()
    _lib.check(_lib.load().nrb_render_tiles_to_host(scene.handle, C.byref(cam), C.byref(ts), C.c_void_p(int(host_addr)), C.byref(stats)))
    return stats
