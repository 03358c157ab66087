"""ctypes front of libh264host.so (include/h264_host.h): the CABAC stream writer (h264e_*) and the host-side decoder (h264d_*)
that turns an Annex-B stream into the per-picture buffers of include/h264r.h.

The decoder here is the product's host side (what NAL::decode / Slice::PraseSliceDataer / macroblock::Parse are in the reference,
h264/NAL.cpp:14-316, h264/slice.cpp:278-389, h264/macroblock.cpp:143-475); the writer is tooling -- the image has no H.264 encoder,
so every genuine bitstream the tests, the libavcodec pin and the reference arm use is written by it.
"""
import ctypes

import numpy as np

from . import abi, synth

REF_QUIRKS, KEEP_SYNTAX = 0x1, 0x2          # h264d_open flags
BENCH_PACK = 0x100                          # h264d_benchmark: + packing of every picture into a stream-form picture buffer
E_ABSOLUTE, E_REF_QUIRKS = 0x1, 0x2         # h264e_encode_picture flags


class FrameRef(ctypes.Structure):
    _fields_ = [("frame_id", ctypes.c_int32), ("poc", ctypes.c_int32), ("decode_index", ctypes.c_int32)]


class Picture(ctypes.Structure):
    _fields_ = [
        ("w_mbs", ctypes.c_int32), ("h_mbs", ctypes.c_int32), ("frame_id", ctypes.c_int32), ("decode_index", ctypes.c_int32),
        ("is_idr", ctypes.c_int32), ("nal_ref_idc", ctypes.c_int32), ("frame_num", ctypes.c_int32), ("poc", ctypes.c_int32),
        ("slice_qp", ctypes.c_int32),
        ("pp", abi.PicParams),
        ("hdr", ctypes.c_void_p), ("mv", ctypes.c_void_p), ("coeff", ctypes.c_void_p), ("syntax", ctypes.c_void_p),
        ("n_intra", ctypes.c_int32), ("n_output", ctypes.c_int32), ("output", FrameRef * 17),
        ("n_release", ctypes.c_int32), ("release", ctypes.c_int32 * 17),
        ("slice_data_bytes", ctypes.c_int64),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = synth.host_lib()
        _lib.h264e_write_sps_pps.restype = ctypes.c_long
        _lib.h264e_encode_picture.restype = ctypes.c_long
        _lib.h264d_open.restype = ctypes.c_void_p
        _lib.h264d_open.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_int, ctypes.c_int]
        _lib.h264d_close.argtypes = [ctypes.c_void_p]
        _lib.h264d_next.argtypes = [ctypes.c_void_p, ctypes.POINTER(Picture)]
        _lib.h264d_error.restype = ctypes.c_char_p
        _lib.h264d_error.argtypes = [ctypes.c_void_p]
        _lib.h264d_fill_picture_buf.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_size_t),
                                                ctypes.POINTER(ctypes.c_long)]
        _lib.h264w_pack_arrays.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                           ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_long)]
        _lib.h264d_benchmark.restype = ctypes.c_long
        _lib.h264d_benchmark.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    return _lib


def encode_stream(seq, pictures, flags=0, per_picture=False):
    """pictures: [(Pic, mbs ndarray[MB_DTYPE])] -> Annex-B bytes with real CABAC slice data.  mbs arrays are normalised in place
    (and, with E_ABSOLUTE, rewritten from final vectors / modes to mvd / prev / rem).  per_picture: also return the access units."""
    L = lib()
    n_mbs = seq.w_mbs * seq.h_mbs
    cap = 4096 + n_mbs * 2048
    buf = (ctypes.c_uint8 * cap)()
    n = L.h264e_write_sps_pps(ctypes.byref(seq), buf, ctypes.c_long(cap))
    assert n > 0
    units = [bytes(buf[:n])]
    for pic, mbs in pictures:
        assert mbs.dtype == synth.MB_DTYPE and mbs.size == n_mbs and mbs.flags["C_CONTIGUOUS"]
        n = L.h264e_encode_picture(ctypes.byref(seq), ctypes.byref(pic), mbs.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(n_mbs),
                                   ctypes.c_uint32(flags), buf, ctypes.c_long(cap))
        assert n > 0, "h264e_encode_picture failed: %d" % n
        units.append(bytes(buf[:n]))
    stream = b"".join(units)
    if per_picture:
        return stream, [units[0] + units[1]] + units[2:]
    return stream


class HostDecoder:
    """Annex-B stream -> pictures as dicts(pp, hdr, mv, coeff, frame_id, poc, ...), arrays copied out of the decoder."""

    def __init__(self, stream, flags=0, max_frames=0, reorder_depth=0):
        self._buf = np.frombuffer(stream, dtype=np.uint8).copy()
        self._flags = flags
        self._h = lib().h264d_open(self._buf.ctypes.data_as(ctypes.c_void_p), self._buf.size, flags, max_frames, reorder_depth)
        self.tail = None

    def close(self):
        if self._h:
            lib().h264d_close(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def __iter__(self):
        return self

    def __next__(self):
        p = Picture()
        rc = lib().h264d_next(self._h, ctypes.byref(p))
        if rc < 0:
            raise RuntimeError("h264d_next: %d (%s)" % (rc, lib().h264d_error(self._h).decode()))
        outs = [(p.output[i].frame_id, p.output[i].poc, p.output[i].decode_index) for i in range(p.n_output)]
        rel = [p.release[i] for i in range(p.n_release)]
        if rc == 0:
            self.tail = {"output": outs, "release": rel}
            raise StopIteration
        n = p.w_mbs * p.h_mbs
        hdr = np.ctypeslib.as_array(ctypes.cast(p.hdr, ctypes.POINTER(ctypes.c_uint8)), shape=(n * 16,)).copy().view(abi.MB_HDR_DTYPE)
        coeff = np.ctypeslib.as_array(ctypes.cast(p.coeff, ctypes.POINTER(ctypes.c_int16)), shape=(n, 384)).copy()
        mv = None
        if p.mv:
            mv = np.ctypeslib.as_array(ctypes.cast(p.mv, ctypes.POINTER(ctypes.c_int16)), shape=(n, 32)).copy()
        pp = abi.PicParams()
        ctypes.memmove(ctypes.byref(pp), ctypes.byref(p.pp), ctypes.sizeof(pp))
        out = {"w_mbs": p.w_mbs, "h_mbs": p.h_mbs, "frame_id": p.frame_id, "decode_index": p.decode_index, "is_idr": p.is_idr,
               "frame_num": p.frame_num, "poc": p.poc, "slice_qp": p.slice_qp, "pp": pp, "hdr": hdr, "mv": mv, "coeff": coeff,
               "n_intra": p.n_intra, "output": outs, "release": rel, "slice_data_bytes": p.slice_data_bytes}
        if p.syntax:
            out["syntax"] = np.ctypeslib.as_array(ctypes.cast(p.syntax, ctypes.POINTER(ctypes.c_uint8)),
                                                  shape=(n * synth.MB_DTYPE.itemsize,)).copy().view(synth.MB_DTYPE)
        return out

    def fill_picture_buf(self, buf, level_fmt=abi.LEVELS_COMPACT):
        """Packs the picture returned last into an engine picture buffer (engine.PictureBuf.raw)."""
        nb, nm = ctypes.c_size_t(0), ctypes.c_long(0)
        rc = lib().h264d_fill_picture_buf(self._h, ctypes.byref(buf), level_fmt, ctypes.byref(nb), ctypes.byref(nm))
        return rc, nb.value, nm.value


def decode_stream(stream, flags=0, max_frames=0):
    d = HostDecoder(stream, flags, max_frames)
    pics = list(d)
    tail = d.tail
    d.close()
    return pics, tail


def benchmark(stream, flags=0, repeat=1):
    """(pictures parsed, seconds) of the host side alone."""
    buf = np.frombuffer(stream, dtype=np.uint8)
    sec = ctypes.c_double(0)
    n = lib().h264d_benchmark(buf.ctypes.data_as(ctypes.c_void_p), buf.size, flags, repeat, ctypes.byref(sec))
    if n < 0:
        raise RuntimeError("h264d_benchmark: %d" % n)
    return n, sec.value
