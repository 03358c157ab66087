"""ctypes binding of libh264r.so (the C ABI of include/h264r.h).

There is no CPU fallback anywhere in this package: if the CUDA library is missing or no sm_100
device is visible, creating an Engine raises.
"""
import ctypes
import os

import numpy as np

from . import abi, build

class PictureBufStruct(ctypes.Structure):
    _fields_ = [("hdr", ctypes.c_void_p), ("mv", ctypes.c_void_p), ("mv_end", ctypes.c_void_p), ("blk_mask", ctypes.c_void_p), ("blocks", ctypes.c_void_p),
                ("max_blocks", ctypes.c_size_t), ("n_intra", ctypes.c_int32), ("level_bytes", ctypes.c_int32), ("base", ctypes.c_void_p),
                ("hdr_bytes", ctypes.c_int32), ("ring_index", ctypes.c_int32), ("hdr8", ctypes.c_void_p),
                ("mb_words", ctypes.c_void_p), ("stream", ctypes.c_void_p), ("max_stream_words", ctypes.c_size_t)]


_lib = None


class H264RError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("h264r error %d: %s" % (code, msg))
        self.code = code


def lib_path():
    # H264R_LIB: development switch to compare builds of the same library (tuning runs); the product is lib/libh264r.so
    return os.environ.get("H264R_LIB") or os.path.join(build.LIB, "libh264r.so")


def load_library():
    """Loads (building if needed and possible) libh264r.so and declares the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.environ.get("H264R_LIB"):
        path = build.build_cuda()      # rebuilds when a source is newer than the library (and nvcc exists); else the prebuilt one
    L = ctypes.CDLL(path)
    vp, ci = ctypes.c_void_p, ctypes.c_int
    L.h264r_abi_version.restype = ci
    L.h264r_last_error.restype = ctypes.c_char_p
    L.h264r_last_error.argtypes = [vp]
    L.h264r_create.argtypes = [ctypes.POINTER(vp), ci, ci, ci, ci, ci, ctypes.c_uint32]
    L.h264r_destroy.argtypes = [vp]
    L.h264r_destroy.restype = None
    L.h264r_host_alloc.argtypes = [vp, ctypes.c_size_t]
    L.h264r_host_alloc.restype = vp
    L.h264r_host_free.argtypes = [vp, vp]
    L.h264r_host_free.restype = None
    L.h264r_frame_begin.argtypes = [vp, ci, ctypes.POINTER(abi.PicParams)]
    L.h264r_submit_mbs.argtypes = [vp, ci, ci, ci, vp, vp, vp]
    L.h264r_submit_mbs_packed.argtypes = [vp, ci, ci, ci, vp, vp, vp, vp, ctypes.c_size_t]
    L.h264r_wait_uploads.argtypes = [vp]
    L.h264r_set_lanes.argtypes = [vp, ci]
    L.h264r_picture_buf_alloc.argtypes = [vp, ctypes.c_size_t, ctypes.POINTER(PictureBufStruct)]
    L.h264r_picture_buf_free.argtypes = [vp, ctypes.POINTER(PictureBufStruct)]
    L.h264r_picture_buf_format.argtypes = [vp, ctypes.POINTER(PictureBufStruct), ctypes.c_int]
    L.h264r_picture_ring_alloc.argtypes = [vp, ci, ctypes.c_size_t, ctypes.POINTER(PictureBufStruct)]
    L.h264r_submit_pictures.argtypes = [vp, ci, ctypes.POINTER(ci), ctypes.POINTER(PictureBufStruct), ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_long)]
    L.h264r_picture_buf_free.restype = None
    L.h264r_submit_picture.argtypes = [vp, ci, ctypes.POINTER(PictureBufStruct), ctypes.c_size_t, ctypes.c_long]
    L.h264r_frame_end.argtypes = [vp, ci]
    L.h264r_frame_abort.argtypes = [vp, ci]
    L.h264r_device_status.argtypes = [vp, ctypes.POINTER(ctypes.c_uint32)]
    L.h264r_flush.argtypes = [vp]
    L.h264r_sync.argtypes = [vp]
    L.h264r_read_planes.argtypes = [vp, ci, vp, ci, vp, vp, ci]
    L.h264r_frame_layout.argtypes = [vp, ctypes.POINTER(abi.FrameLayout)]
    L.h264r_read_frame_async.argtypes = [vp, ci, vp]
    L.h264r_wait_readbacks.argtypes = [vp]
    L.h264r_frame_digest.argtypes = [vp, ci, vp]
    L.h264r_frame_digests.argtypes = [vp, vp, ci, vp]
    L.h264r_range_excursions.argtypes = [vp, ctypes.POINTER(ctypes.c_uint64)]
    L.h264r_excursion_map.argtypes = [vp, ci, vp]
    L.h264r_residual_mbs.argtypes = [vp, ci, vp, vp, vp]
    L.h264r_release.argtypes = [vp, ci]
    L.h264r_last_flush_stats.argtypes = [vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ci)]
    L.h264r_set_profiling.argtypes = [vp, ci]
    L.h264r_last_flush_kernel_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_float * 5), ctypes.POINTER(ci * 5)]
    L.h264r_last_flush_launch_ms.argtypes = [vp, vp, vp, ci]
    L.h264r_last_flush_launch_ms.restype = ci
    L.h264r_mark.argtypes = [vp, ci]
    L.h264r_elapsed_ms.argtypes = [vp, ci, ci, ctypes.POINTER(ctypes.c_float)]
    if L.h264r_abi_version() != abi.ABI_VERSION:
        raise RuntimeError("libh264r.so ABI version mismatch")
    _lib = L
    return L


EXPORTS = [
    "h264r_abi_version", "h264r_last_error", "h264r_create", "h264r_destroy", "h264r_host_alloc", "h264r_host_free",
    "h264r_frame_begin", "h264r_submit_mbs", "h264r_submit_mbs_packed", "h264r_wait_uploads", "h264r_set_lanes", "h264r_picture_buf_alloc", "h264r_picture_buf_free", "h264r_picture_buf_format", "h264r_picture_ring_alloc", "h264r_submit_picture", "h264r_submit_pictures", "h264r_frame_end", "h264r_frame_abort", "h264r_device_status", "h264r_flush", "h264r_sync", "h264r_read_planes",
    "h264r_frame_layout", "h264r_read_frame_async", "h264r_wait_readbacks",
    "h264r_frame_digest", "h264r_frame_digests", "h264r_range_excursions", "h264r_excursion_map",
    "h264r_residual_mbs", "h264r_release", "h264r_last_flush_stats", "h264r_set_profiling",
    "h264r_last_flush_kernel_ms", "h264r_last_flush_launch_ms", "h264r_mark", "h264r_elapsed_ms",
]


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return ctypes.c_void_p(a)
    return a.ctypes.data_as(ctypes.c_void_p)


class PinnedBuffer:
    """numpy view over memory from h264r_host_alloc (page-locked)."""

    def __init__(self, engine, dtype, shape):
        self.engine = engine
        dtype = np.dtype(dtype)
        n = int(np.prod(shape))
        self.nbytes = n * dtype.itemsize
        self.ptr = engine.L.h264r_host_alloc(engine.ctx, self.nbytes)
        if not self.ptr:
            raise H264RError(abi.E_NOMEM, "pinned allocation of %d bytes failed" % self.nbytes)
        buf = (ctypes.c_uint8 * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)

    def free(self):
        if self.ptr:
            self.engine.L.h264r_host_free(self.engine.ctx, self.ptr)
            self.ptr = None
            self.array = None


class PictureBuf:
    """h264r_picture_buf: pinned staging of one picture in the engine's upload layout (numpy views)."""

    def __init__(self, engine, max_blocks=0, struct=None):
        """struct: a member of a PictureRing's array (allocated there), else a buffer of its own"""
        self.engine = engine
        if struct is None:
            self.c = PictureBufStruct()
            engine._check(engine.L.h264r_picture_buf_alloc(engine.ctx, max_blocks, ctypes.byref(self.c)))
        else:
            self.c = struct
        n = engine.n_mbs

        def view(ptr, dtype, count):
            dt = np.dtype(dtype)
            buf = (ctypes.c_uint8 * (count * dt.itemsize)).from_address(ptr)
            return np.frombuffer(buf, dtype=dt, count=count)
        self._view = view
        self.mv = view(self.c.mv, np.int16, n * 32).reshape(n, 32)
        self.mv_tail = view(self.c.mv_end - n * 64, np.int16, n * 32).reshape(-1, 2)     # ends at mv_end: partition-order stream
        self._map()
        self.n_blocks = 0
        self.n_mvs = -1

    def _map(self):
        n = self.engine.n_mbs
        self.hdr = self._view(self.c.hdr, abi.MB_HDR_DTYPE, n)
        self.hdr8 = self._view(self.c.hdr8, np.uint32, n * 2).reshape(n, 2) if self.c.hdr8 else None
        self.mb_words = self._view(self.c.mb_words, np.uint8, n) if self.c.mb_words else None
        self.stream = self._view(self.c.stream, np.uint32, int(self.c.max_stream_words)) if self.c.stream else None
        self.blk_mask = self._view(self.c.blk_mask, np.uint32, n)
        self.blocks = self._view(self.c.blocks, np.int16, int(self.c.max_blocks) * 16).reshape(-1, 16)

    def set_header_bytes(self, hdr_bytes):
        """16: h264r_mb_hdr records; 8: h264r_mb_hdr8 (masks and blocks move up behind them; compact levels only); abi.HDR_STREAM: the
        stream form (8-byte headers + word counts + one stream per picture)"""
        if self.c.hdr_bytes != hdr_bytes:
            self.engine._check(self.engine.L.h264r_picture_buf_format(self.engine.ctx, ctypes.byref(self.c), hdr_bytes))
            self._map()

    def fill_stream(self, hdr, mv, coeff):
        """stream form (abi.HDR_STREAM) from the dense per-macroblock arrays, packed by the C writer (include/h264r_writer.h through
        libh264host.so: h264w_pack_arrays).  Returns False when a level leaves int8 (the picture then needs another form)."""
        from . import hoststream
        self.set_header_bytes(abi.HDR_STREAM)
        hdr = np.ascontiguousarray(hdr).view(abi.MB_HDR_DTYPE).reshape(-1)
        coeff = np.ascontiguousarray(coeff, dtype=np.int16)
        mvp = np.ascontiguousarray(mv, dtype=np.int16) if mv is not None else None
        nb, nm = ctypes.c_size_t(0), ctypes.c_long(0)
        L = hoststream.lib()
        rc = L.h264w_pack_arrays(ctypes.byref(self.c), hdr.size, hdr.ctypes.data_as(ctypes.c_void_p), mvp.ctypes.data_as(ctypes.c_void_p) if mvp is not None else None,
                                 coeff.ctypes.data_as(ctypes.c_void_p), abi.LEVELS_COMPACT, ctypes.byref(nb), ctypes.byref(nm))
        if rc == abi.E_INVAL:
            return False
        if rc:
            raise H264RError(rc, "h264w_pack_arrays (-6: the buffer's stream area is too small for the picture)")
        self.n_blocks, self.n_mvs = int(nb.value), int(nm.value)
        return True

    def fill(self, hdr, mv, blk_mask, blocks, mv_partition=False, level8=None, compact=None, short_hdr=False):
        """mv: per-4x4 vectors (n, 32); mv_partition: store them in partition order (pack.pack_mvs_partition).
        level8 / compact: level storage (None: the narrowest that holds the picture -- H264R_LEVELS_COMPACT when every
        level fits int8, else int16).  short_hdr: 8-byte headers when the levels go compact (else the 16-byte ones)"""
        from . import pack
        blocks = np.asarray(blocks)
        if level8 is None:
            level8 = blocks.shape[0] > 0 and int(blocks.min()) >= -128 and int(blocks.max()) <= 127
        if compact is None:
            compact = level8
        short_hdr = bool(short_hdr and level8 and compact)
        self.set_header_bytes(8 if short_hdr else 16)
        if short_hdr:
            self.hdr8[:] = pack.pack_hdr8(hdr)
        else:
            self.hdr[:] = np.ascontiguousarray(hdr).view(abi.MB_HDR_DTYPE).reshape(-1)
        self.n_mvs = -1
        if mv is not None:
            if mv_partition:
                pm = pack.pack_mvs_partition(hdr, mv)
                self.n_mvs = int(pm.shape[0])
                if self.n_mvs:
                    self.mv_tail[-self.n_mvs:] = pm[::-1]     # vector i at mv_end[-2(i+1)]: the stream grows backward
            else:
                self.mv[:] = mv
        self.blk_mask[:] = blk_mask
        self.n_blocks = int(blocks.shape[0])
        if level8 and compact:      # 8-byte units: significance map + levels where a macroblock's blocks are sparse
            mask, units = pack.pack_compact(blk_mask, blocks)
            if short_hdr:
                units = pack.insert_intra_units(hdr, mask, units)
            self.blk_mask[:] = mask
            self.n_blocks = int(units.shape[0])       # what h264r_submit_picture is told: units
            self.blocks.view(np.uint8).reshape(-1, 8)[:self.n_blocks] = units
            self.c.level_bytes = abi.LEVELS_COMPACT
        elif level8:      # int8 transport: 16 bytes per block through the same area
            self.blocks.view(np.int8).reshape(-1, 16)[:self.n_blocks] = blocks.astype(np.int8)
            self.c.level_bytes = abi.LEVELS_INT8
        else:
            self.blocks[:self.n_blocks] = blocks
            self.c.level_bytes = abi.LEVELS_INT16
        self.c.n_intra = int((np.ascontiguousarray(hdr).view(abi.MB_HDR_DTYPE).reshape(-1)["mb_class"] <= abi.MB_I16x16).sum())       # a parser counts this as it goes

    def free(self):
        if self.c.base:
            self.hdr = self.hdr8 = self.mv = self.mv_tail = self.blk_mask = self.blocks = self.mb_words = self.stream = None
            self.engine.L.h264r_picture_buf_free(self.engine.ctx, ctypes.byref(self.c))


class PictureRing:
    """n picture buffers in one pinned allocation at a constant stride (h264r_picture_ring_alloc): filled like single buffers,
    shipped together by Engine.submit_ring with one strided transfer."""

    def __init__(self, engine, n, max_blocks=0):
        self.engine = engine
        self.n = n
        self.array = (PictureBufStruct * n)()
        engine._check(engine.L.h264r_picture_ring_alloc(engine.ctx, n, max_blocks, self.array))
        self.bufs = [PictureBuf(engine, struct=self.array[i]) for i in range(n)]
        self._ids = (ctypes.c_int * n)()
        self._nb = (ctypes.c_size_t * n)()
        self._nm = (ctypes.c_long * n)()

    def __getitem__(self, i):
        return self.bufs[i]

    def free(self):
        if self.bufs:
            for b in self.bufs[::-1]:     # member 0 last: it owns the allocation
                b.free()
            self.bufs = []


class Engine:
    """One reconstruction context on one GPU (h264r_ctx)."""

    def __init__(self, w_mbs, h_mbs, max_frames, max_pending=None, flags=abi.REF_COMPAT, device=0):
        self.L = load_library()
        self.w_mbs, self.h_mbs, self.n_mbs = w_mbs, h_mbs, w_mbs * h_mbs
        self.flags = flags
        self.max_frames = max_frames
        self.ctx = ctypes.c_void_p()
        rc = self.L.h264r_create(ctypes.byref(self.ctx), device, w_mbs, h_mbs, max_frames, max_pending or max_frames, flags)
        if rc != 0:
            raise H264RError(rc, self.L.h264r_last_error(None).decode())

    def close(self):
        if self.ctx:
            self.L.h264r_destroy(self.ctx)
            self.ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise H264RError(rc, self.L.h264r_last_error(self.ctx).decode())

    def pinned(self, dtype, shape):
        return PinnedBuffer(self, dtype, shape)

    def frame_begin(self, frame_id, pp):
        self._check(self.L.h264r_frame_begin(self.ctx, frame_id, ctypes.byref(pp)))

    def submit_mbs(self, frame_id, hdr, mv, coeff, first_mb=0, n_mbs=None):
        n = n_mbs if n_mbs is not None else (hdr.shape[0] if hasattr(hdr, "shape") else self.n_mbs)
        self._check(self.L.h264r_submit_mbs(self.ctx, frame_id, first_mb, n, _ptr(hdr), _ptr(mv), _ptr(coeff)))

    def submit_mbs_packed(self, frame_id, hdr, mv, blk_mask, blocks, first_mb=0, n_mbs=None):
        n = n_mbs if n_mbs is not None else hdr.shape[0]
        n_blocks = 0 if blocks is None else int(blocks.size // 16)
        self._check(self.L.h264r_submit_mbs_packed(self.ctx, frame_id, first_mb, n, _ptr(hdr), _ptr(mv), _ptr(blk_mask), _ptr(blocks),
                                                   n_blocks))

    def set_lanes(self, lanes):
        self._check(self.L.h264r_set_lanes(self.ctx, lanes))

    def wait_uploads(self):
        self._check(self.L.h264r_wait_uploads(self.ctx))

    def frame_end(self, frame_id):
        self._check(self.L.h264r_frame_end(self.ctx, frame_id))

    def submit_picture(self, frame_id, pp, hdr, mv, coeff):
        hdr = np.ascontiguousarray(hdr)
        coeff = np.ascontiguousarray(coeff, dtype=np.int16)
        mv = np.ascontiguousarray(mv, dtype=np.int16) if mv is not None else None
        self.frame_begin(frame_id, pp)
        self.submit_mbs(frame_id, hdr, mv, coeff)
        self.frame_end(frame_id)

    def submit_picture_packed(self, frame_id, pp, hdr, mv, blk_mask, blocks):
        """blk_mask, blocks as produced by pack.pack_blocks()"""
        self.frame_begin(frame_id, pp)
        self.submit_mbs_packed(frame_id, hdr, mv, blk_mask, blocks)
        self.frame_end(frame_id)

    def picture_buf(self, max_blocks=0):
        return PictureBuf(self, max_blocks)

    def submit_picture_buf(self, frame_id, pp, buf):
        """whole picture from a PictureBuf: one DMA, queued at once"""
        self.frame_begin(frame_id, pp)
        self._check(self.L.h264r_submit_picture(self.ctx, frame_id, ctypes.byref(buf.c), buf.n_blocks, buf.n_mvs))

    def picture_ring(self, n, max_blocks=0):
        return PictureRing(self, n, max_blocks)

    def submit_ring(self, ring, frame_ids, pps, count=None):
        """pictures frame_ids[k] <- ring[k], begun back to back and shipped with one strided transfer where the library can"""
        n = ring.n if count is None else count
        for k in range(n):
            self.frame_begin(frame_ids[k], pps[k])
            ring._ids[k] = frame_ids[k]
            ring._nb[k] = ring.bufs[k].n_blocks
            ring._nm[k] = ring.bufs[k].n_mvs
        self._check(self.L.h264r_submit_pictures(self.ctx, n, ring._ids, ring.array, ring._nb, ring._nm))

    def frame_abort(self, frame_id):
        self._check(self.L.h264r_frame_abort(self.ctx, frame_id))

    def device_status(self):
        f = ctypes.c_uint32()
        self._check(self.L.h264r_device_status(self.ctx, ctypes.byref(f)))
        return int(f.value)

    def flush(self):
        self._check(self.L.h264r_flush(self.ctx))

    def sync(self):
        self._check(self.L.h264r_sync(self.ctx))

    def read_planes(self, frame_id):
        W, H = self.w_mbs * 16, self.h_mbs * 16
        y = np.empty((H, W), dtype=np.uint8)
        u = np.empty((H // 2, W // 2), dtype=np.uint8)
        v = np.empty((H // 2, W // 2), dtype=np.uint8)
        self._check(self.L.h264r_read_planes(self.ctx, frame_id, _ptr(y), W, _ptr(u), _ptr(v), W // 2))
        return y, u, v

    def frame_layout(self):
        lay = abi.FrameLayout()
        self._check(self.L.h264r_frame_layout(self.ctx, ctypes.byref(lay)))
        return lay

    def read_frame_async(self, frame_id, dst_ptr):
        """dst_ptr: address of frame_layout().frame_bytes bytes (pinned: PinnedBuffer.ptr + offset)"""
        self._check(self.L.h264r_read_frame_async(self.ctx, frame_id, ctypes.c_void_p(dst_ptr)))

    def wait_readbacks(self):
        self._check(self.L.h264r_wait_readbacks(self.ctx))

    @staticmethod
    def planes_of(lay, buf):
        """views of the three planes inside a frame image read by read_frame_async (buf: uint8 array of frame_bytes)"""
        W, H = lay.width, lay.height
        y = np.lib.stride_tricks.as_strided(buf[lay.y_offset:], shape=(H, W), strides=(lay.y_pitch, 1))
        u = np.lib.stride_tricks.as_strided(buf[lay.u_offset:], shape=(H // 2, W // 2), strides=(lay.c_pitch, 1))
        v = np.lib.stride_tricks.as_strided(buf[lay.v_offset:], shape=(H // 2, W // 2), strides=(lay.c_pitch, 1))
        return y, u, v

    def digests(self, frame_ids):
        ids = np.ascontiguousarray(frame_ids, dtype=np.int32)
        out = np.empty((ids.size, 16), dtype=np.uint8)
        self._check(self.L.h264r_frame_digests(self.ctx, _ptr(ids), int(ids.size), _ptr(out)))
        return out

    def range_excursions(self):
        c = ctypes.c_uint64()
        self._check(self.L.h264r_range_excursions(self.ctx, ctypes.byref(c)))
        return int(c.value)

    def excursion_map(self, frame_id):
        out = np.empty(self.n_mbs, dtype=np.uint8)
        self._check(self.L.h264r_excursion_map(self.ctx, frame_id, _ptr(out)))
        return out

    def residual_mbs(self, hdr, coeff):
        hdr = np.ascontiguousarray(hdr)
        coeff = np.ascontiguousarray(coeff, dtype=np.int16)
        n = hdr.shape[0]
        out = np.empty((n, 384), dtype=np.int16)
        self._check(self.L.h264r_residual_mbs(self.ctx, n, _ptr(hdr), _ptr(coeff), _ptr(out)))
        return out[:, :256].reshape(n, 16, 16), out[:, 256:320].reshape(n, 8, 8), out[:, 320:].reshape(n, 8, 8)

    def release(self, frame_id):
        self._check(self.L.h264r_release(self.ctx, frame_id))

    def last_flush_stats(self):
        ms, n = ctypes.c_float(), ctypes.c_int()
        self._check(self.L.h264r_last_flush_stats(self.ctx, ctypes.byref(ms), ctypes.byref(n)))
        return float(ms.value), int(n.value)

    def set_profiling(self, on):
        self._check(self.L.h264r_set_profiling(self.ctx, 1 if on else 0))

    def last_flush_launch_ms(self, cap=4096):
        """per-launch (ms, family) of the last flush, issue order"""
        ms, fam = (ctypes.c_float * cap)(), (ctypes.c_int * cap)()
        n = self.L.h264r_last_flush_launch_ms(self.ctx, ms, fam, cap)
        if n < 0:
            self._check(n)
        return [(ms[i], fam[i]) for i in range(n)]

    def last_flush_kernel_ms(self):
        ms, n = (ctypes.c_float * 5)(), (ctypes.c_int * 5)()
        self._check(self.L.h264r_last_flush_kernel_ms(self.ctx, ctypes.byref(ms), ctypes.byref(n)))
        return list(ms), list(n)

    def mark(self, slot):
        self._check(self.L.h264r_mark(self.ctx, slot))

    def elapsed_ms(self, a, b):
        ms = ctypes.c_float()
        self._check(self.L.h264r_elapsed_ms(self.ctx, a, b, ctypes.byref(ms)))
        return float(ms.value)
