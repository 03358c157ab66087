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
            fmt = abi.LEVELS_COMPACT if transport in ("compact", "compact8") else abi.LEVELS_INT16
            if transport == "compact8":          # 8-byte headers: the host decoder writes through h264r_writer.h, which follows buf.hdr_bytes
                b.set_header_bytes(8)
            rc, nb, nm = d.fill_picture_buf(b.c, fmt)
            if rc == abi.E_INVAL and fmt == abi.LEVELS_COMPACT:      # a level outside int8: the writer says so, the caller falls back
                b.set_header_bytes(16)
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


@pytest.mark.parametrize("transport", ["dense", "compact", "compact8", "int16"])
@pytest.mark.parametrize("name", ["spec_qcif_i", "spec_cif_ip", "spec_qcif_wp", "spec_dpb", "spec_hd1080_ip", "base_qcif_i", "base_cif_ip"])
def test_spec_mode_equals_libavcodec(built, name, transport):
    g = load_stream_golden(name)
    outs, _, status = engine_on_stream(g["stream"], flags=0, transport=transport)
    assert status == 0
    check_against_stream_golden(g, outs, "engine (spec mode) vs libavcodec")


@pytest.mark.parametrize("transport", ["dense", "compact", "compact8"])
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


@pytest.mark.parametrize("name", ["kat_intra_modes", "kat_deblock_strengths"])
def test_known_answer_streams(built, name):
    """the per-mode / per-boundary-strength streams (pinned against libavcodec by the CPU suite): engine == oracle"""
    import streams
    s = streams.KAT_RECIPES[name]()
    pics, _ = hoststream.decode_stream(s["stream"])
    want = run_sequence(oracle.recon_picture, pics[0]["w_mbs"], pics[0]["h_mbs"], 0, pics)
    outs, _, status = engine_on_stream(s["stream"], flags=0, transport="compact")
    assert status == 0
    for i, (w, o) in enumerate(zip(want, outs)):
        for k in "yuv":
            assert np.array_equal(w[k], o[k]), "%s picture %d plane %s" % (name, i, k)


def test_more_pictures_than_one_launch_list_holds(built):
    """a flush with more pictures per wave than a kernel's by-value picture list (PIC_LIST_MAX = 940) and than the intra work list of
    one k_intra_list launch (SPARSE_PICS = 64) holds: the scheduler splits the waves into several launches"""
    from videodecoder_b200 import workload
    W, H, N = 3, 3, 960
    rng = np.random.default_rng(77)
    cfg = workload.WorkloadConfig(p_intra_in_p=0.1, p_skip=0.2)
    eng = engine.Engine(W, H, max_frames=2 * N, max_pending=2 * N, flags=abi.REF_COMPAT)
    gops = [workload.make_gop(rng, W, H, 2, cfg) for _ in range(8)]
    for g in range(N):
        i_pic, p_pic = gops[g % 8]
        eng.submit_picture(2 * g, i_pic["pp"], i_pic["hdr"], None, i_pic["coeff"])
        pp = abi.PicParams.from_buffer_copy(p_pic["pp"])
        pp.ref_list0[0] = 2 * g
        mask, blocks = pack.pack_blocks(p_pic["coeff"])
        eng.submit_picture_packed(2 * g + 1, pp, np.ascontiguousarray(p_pic["hdr"]), p_pic["mv"], mask, blocks)
    eng.sync()
    want = [run_sequence(oracle.recon_picture, W, H, abi.REF_COMPAT, gop) for gop in gops]
    for g in list(range(0, N, 37)) + [N - 1]:
        for k in range(2):
            y, u, v = eng.read_planes(2 * g + k)
            assert np.array_equal(y, want[g % 8][k]["y"]) and np.array_equal(u, want[g % 8][k]["u"]), "GOP %d picture %d" % (g, k)
    eng.close()
