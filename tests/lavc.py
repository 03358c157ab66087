"""TEST INFRASTRUCTURE: libavcodec's H.264 decoder as an independent, spec-conformant second oracle.

The image carries no ffmpeg CLI and no PyAV, but the opencv wheel bundles libavcodec / libavutil (decode side); they are loaded
here with ctypes and driven through the public send_packet / receive_frame API, which hands back the raw Y, U, V planes
(cv2.VideoCapture would convert to BGR, which is lossy).  Used to pin the spec-mode pieces the reference decoder lacks
(Intra8x8, chroma prediction / MC, 8x8 transform, deblocking: SURVEY a13, a14, a19, a20, a21) and to generate the fixtures under
tests/golden/spec_*.npz (tests/golden/make_spec_golden.py).  Never imported by the product.
"""
import ctypes
import glob
import os

import numpy as np

AV_CODEC_ID_H264 = 27
AVERROR_EAGAIN = -11
AVERROR_EOF = -541478725          # FFERRTAG('E','O','F',' ')

_libs = None


def available():
    try:
        _load()
        return True
    except Exception:
        return False


def _load():
    global _libs
    if _libs is not None:
        return _libs
    import cv2  # noqa: F401  (locates the wheel)
    base = os.path.join(os.path.dirname(os.path.dirname(cv2.__file__)), "opencv_python_headless.libs")
    if not os.path.isdir(base):
        base = os.path.join(os.path.dirname(os.path.dirname(cv2.__file__)), "opencv_python.libs")

    def one(pat):
        m = sorted(glob.glob(os.path.join(base, pat)))
        if not m:
            raise OSError("no %s under %s" % (pat, base))
        return m[0]
    avu = ctypes.CDLL(one("libavutil-*.so*"), mode=ctypes.RTLD_GLOBAL)
    try:
        ctypes.CDLL(one("libswresample-*.so*"), mode=ctypes.RTLD_GLOBAL)
    except OSError:
        pass
    avc = ctypes.CDLL(one("libavcodec-*.so*"), mode=ctypes.RTLD_GLOBAL)
    avc.avcodec_find_decoder.restype = ctypes.c_void_p
    avc.avcodec_find_decoder.argtypes = [ctypes.c_int]
    avc.avcodec_alloc_context3.restype = ctypes.c_void_p
    avc.avcodec_alloc_context3.argtypes = [ctypes.c_void_p]
    avc.avcodec_open2.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    avc.avcodec_free_context.argtypes = [ctypes.c_void_p]
    avc.av_packet_alloc.restype = ctypes.c_void_p
    avc.av_new_packet.argtypes = [ctypes.c_void_p, ctypes.c_int]
    avc.av_packet_unref.argtypes = [ctypes.c_void_p]
    avc.av_packet_free.argtypes = [ctypes.c_void_p]
    avc.avcodec_send_packet.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    avc.avcodec_receive_frame.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    avu.av_frame_alloc.restype = ctypes.c_void_p
    avu.av_frame_unref.argtypes = [ctypes.c_void_p]
    avu.av_frame_free.argtypes = [ctypes.c_void_p]
    avu.av_opt_set.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
    avu.av_log_set_level.argtypes = [ctypes.c_int]
    _libs = (avu, avc)
    return _libs


class _Packet(ctypes.Structure):        # leading fields of AVPacket (stable public ABI)
    _fields_ = [("buf", ctypes.c_void_p), ("pts", ctypes.c_int64), ("dts", ctypes.c_int64), ("data", ctypes.c_void_p),
                ("size", ctypes.c_int)]


class _Frame(ctypes.Structure):         # leading fields of AVFrame (stable public ABI)
    _fields_ = [("data", ctypes.c_void_p * 8), ("linesize", ctypes.c_int * 8), ("extended_data", ctypes.c_void_p),
                ("width", ctypes.c_int), ("height", ctypes.c_int), ("nb_samples", ctypes.c_int), ("format", ctypes.c_int)]


def _plane(ptr, stride, w, h):
    buf = (ctypes.c_uint8 * (stride * h)).from_address(ptr)
    return np.frombuffer(buf, dtype=np.uint8).reshape(h, stride)[:, :w].copy()


def decode(access_units, quiet=True):
    """access_units: list of bytes (Annex B, one picture each; parameter sets may ride along).  Returns [(y, u, v)] in output
    order (uint8 planes, 4:2:0)."""
    avu, avc = _load()
    if quiet:
        avu.av_log_set_level(8)          # AV_LOG_FATAL
    codec = avc.avcodec_find_decoder(AV_CODEC_ID_H264)
    assert codec, "libavcodec has no H.264 decoder"
    ctx = avc.avcodec_alloc_context3(codec)
    avu.av_opt_set(ctx, b"threads", b"1", 0)
    rc = avc.avcodec_open2(ctx, codec, None)
    assert rc == 0, "avcodec_open2: %d" % rc
    pkt = avc.av_packet_alloc()
    frame = avu.av_frame_alloc()
    out = []

    def drain():
        while True:
            rc = avc.avcodec_receive_frame(ctx, frame)
            if rc in (AVERROR_EAGAIN, AVERROR_EOF):
                return
            assert rc == 0, "avcodec_receive_frame: %d" % rc
            f = _Frame.from_address(frame)
            assert f.format in (0, 12), "pixel format %d is not yuv420p" % f.format       # AV_PIX_FMT_YUV420P / YUVJ420P
            w, h = f.width, f.height
            out.append((_plane(f.data[0], f.linesize[0], w, h), _plane(f.data[1], f.linesize[1], w // 2, h // 2),
                        _plane(f.data[2], f.linesize[2], w // 2, h // 2)))
            avu.av_frame_unref(frame)

    for au in access_units:
        rc = avc.av_new_packet(pkt, len(au))
        assert rc == 0
        p = _Packet.from_address(pkt)
        ctypes.memmove(p.data, au, len(au))
        rc = avc.avcodec_send_packet(ctx, pkt)
        avc.av_packet_unref(pkt)
        assert rc == 0, "avcodec_send_packet: %d" % rc
        drain()
    avc.avcodec_send_packet(ctx, None)
    drain()
    fp = ctypes.c_void_p(frame)
    avu.av_frame_free(ctypes.byref(fp))
    pp = ctypes.c_void_p(pkt)
    avc.av_packet_free(ctypes.byref(pp))
    cp = ctypes.c_void_p(ctx)
    avc.avcodec_free_context(ctypes.byref(cp))
    return out
