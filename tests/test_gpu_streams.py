"""GPU parity on genuine bitstreams, end to end through both C ABIs: Annex-B bytes -> libh264host.so (h264d_*: NAL / CABAC /
derivations / DPB) -> libh264r.so (sm_100a kernels) -> planes, compared with
  * what LIBAVCODEC decoded from the same bytes (tests/golden/spec_*.npz): spec mode -- Intra8x8, intra chroma prediction, chroma
    motion compensation, 8x8 transform, deblocking, 1080p High included (BASELINE config 3);
  * what the UNMODIFIED reference decoder reconstructed from the same bytes with its own CABAC (tests/golden/ref_cif_ip_stream.npz):
    REF_COMPAT mode with the reference's host-side derivations (H264D_REF_QUIRKS).
"""
import ctypes

import numpy as np
import pytest

from common import *  # noqa: F401,F403
from videodecoder_b200 import engine, hoststream

pytestmark = pytest.mark.gpu


def engine_on_stream(stream, flags=0, dflags=0, transport="dense", pool=0):
    """Decodes `stream` on the host, reconstructs on the GPU.  pool > 0: the DPB reuses frame slots and pictures are read back when
    the decoder says they are output; else one slot per picture.  Returns planes per picture in output order."""
    d = hoststream.HostDecoder(stream, flags=dflags, max_frames=pool)
    eng, bufs, outs, n = None, [], {}, 0
    for p in d:
        if eng is None:
            n_slots = pool if pool else 64
            eng = engine.Engine(p["w_mbs"], p["h_mbs"], max_frames=n_slots, max_pending=64, flags=flags)
        if transport == "dense":
            eng.submit_picture(p["frame_id"], p["pp"], p["hdr"], p["mv"], p["coeff"])
        else:
            b = eng.picture_buf(0)
            fmt = abi.LEVELS_COMPACT if transport == "compact" else abi.LEVELS_INT16
            rc, nb, nm = d.fill_picture_buf(b.c, fmt)
            if rc == abi.E_INVAL and fmt == abi.LEVELS_COMPACT:      # a level outside int8: the writer says so, the caller falls back
                rc, nb, nm = d.fill_picture_buf(b.c, abi.LEVELS_INT16)
            assert rc == 0, rc
            b.n_blocks, b.n_mvs = nb, nm
            eng.submit_picture_buf(p["frame_id"], p["pp"], b)
            bufs.append(b)
        n += 1
        if pool:
            for slot, poc, idx in p["output"]:
                outs[idx] = dict(zip("yuv", eng.read_planes(slot)))
            for slot in p["release"]:
                eng.release(slot)
    if pool:
        for slot, poc, idx in d.tail["output"]:
            outs[idx] = dict(zip("yuv", eng.read_planes(slot)))
    else:
        eng.sync()
        for i in range(n):
            outs[i] = dict(zip("yuv", eng.read_planes(i)))
    exc = eng.range_excursions()
    status = eng.device_status()
    for b in bufs:
        b.free()
    eng.close()
    d.close()
    return [outs[i] for i in range(n)], exc, status


@pytest.mark.parametrize("transport", ["dense", "compact", "int16"])
@pytest.mark.parametrize("name", ["spec_qcif_i", "spec_cif_ip", "spec_qcif_wp", "spec_dpb", "spec_hd1080_ip"])
def test_spec_mode_equals_libavcodec(built, name, transport):
    g = load_stream_golden(name)
    outs, _, status = engine_on_stream(g["stream"], flags=0, transport=transport)
    assert status == 0
    check_against_stream_golden(g, outs, "engine (spec mode) vs libavcodec")


@pytest.mark.parametrize("transport", ["dense", "compact"])
def test_ref_compat_equals_the_unmodified_reference(built, transport):
    g = load_stream_golden("ref_cif_ip_stream")
    outs, exc, status = engine_on_stream(g["stream"], flags=abi.REF_COMPAT, dflags=hoststream.REF_QUIRKS, transport=transport)
    assert exc == 0 and status == 0
    check_against_stream_golden(g, outs, "engine (REF_COMPAT) vs reference with its own CABAC")


def test_dpb_slot_reuse_on_the_engine(built):
    """pool of 6 frame slots for a 30-picture stream with list modification, MMCO and long-term pictures: slots are overwritten as
    the DPB frees them (h264r_release + frame_begin on a used slot), pictures read back at output time."""
    g = load_stream_golden("spec_dpb")
    outs, _, status = engine_on_stream(g["stream"], flags=0, transport="compact", pool=6)
    assert status == 0
    check_against_stream_golden(g, outs, "engine with a pool of 6 vs libavcodec")
