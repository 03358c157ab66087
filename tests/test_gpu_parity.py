"""GPU parity tests: the sm_100a engine, called through its C ABI (libh264r.so), against
(a) the golden fixtures produced by the compiled reference decoder, (b) the CPU oracle."""
import numpy as np
import pytest

from common import *  # noqa: F401,F403
from videodecoder_b200 import engine

pytestmark = pytest.mark.gpu


def engine_sequence(w_mbs, h_mbs, flags, pics, batch=True, packed=False):
    eng = engine.Engine(w_mbs, h_mbs, max_frames=len(pics), flags=flags)
    bufs = []
    for i, p in enumerate(pics):
        if packed in ("buf", "bufmv", "bufmvh8", "bufstream"):
            mask, blocks = pack.pack_blocks(p["coeff"])
            b = eng.picture_buf(0 if packed == "bufstream" else max(1, blocks.shape[0]))
            # bufmvh8: 8-byte headers; bufstream: the stream form (a picture with a level beyond int8 falls back to the 16-byte form)
            if not (packed == "bufstream" and b.fill_stream(p["hdr"], p["mv"], p["coeff"])):
                b.fill(p["hdr"], p["mv"], mask, blocks, mv_partition=packed != "buf", short_hdr=packed == "bufmvh8")
            bufs.append(b)
            eng.submit_picture_buf(i, p["pp"], b)
        elif packed:
            mask, blocks = pack.pack_blocks(p["coeff"])
            eng.submit_picture_packed(i, p["pp"], np.ascontiguousarray(p["hdr"]), p["mv"], mask, blocks)
        else:
            eng.submit_picture(i, p["pp"], p["hdr"], p["mv"], p["coeff"])
        if not batch:
            eng.sync()
    eng.sync()
    outs = [dict(zip("yuv", eng.read_planes(i))) for i in range(len(pics))]
    exc = eng.range_excursions()
    dig = eng.digests(list(range(len(pics))))
    for b in bufs:
        b.free()
    eng.close()
    return outs, exc, dig


@pytest.mark.parametrize("name", ["qcif_i", "cif_ip", "qcif_wp", "hd1080_ip"])
@pytest.mark.parametrize("batch,packed", [(True, False), (False, False), (True, True), (False, True), (True, "buf"), (False, "buf"), (True, "bufmv"), (False, "bufmv"), (True, "bufmvh8"), (False, "bufmvh8"), (True, "bufstream"), (False, "bufstream")])
def test_golden_reference_parity(built, name, batch, packed):
    """bit-exact Y/U/V against what the reference decoder reconstructed (MD5 + planes); dense and packed submits,
    one flush for the whole sequence and one flush per picture."""
    g = load_golden(name)
    outs, exc, dig = engine_sequence(g["w_mbs"], g["h_mbs"], abi.REF_COMPAT, g["pics"], batch, packed)
    assert exc == 0
    for i, (p, o) in enumerate(zip(g["pics"], outs)):
        if "y" in p:
            for k in "yuv":
                assert np.array_equal(p[k], o[k]), "picture %d plane %s differs from the reference" % (i, k)
        assert planes_md5(o["y"], o["u"], o["v"]) == p["md5"], "picture %d MD5" % i
        assert bytes(dig[i]) == oracle.digest(o["y"], o["u"], o["v"])


@pytest.mark.parametrize("packed", [False, True, "buf", "bufmv", "bufmvh8", "bufstream"])
@pytest.mark.parametrize("name", ["qcif_i", "cif_ip", "qcif_wp"])
def test_spec_mode_vs_oracle(built, name, packed):
    """same syntax in spec mode (Clip1, chroma prediction/MC, deblocking): engine == oracle."""
    g = load_golden(name)
    for p in g["pics"]:
        p["pp"].disable_deblocking_filter_idc = 0
        p["hdr"]["qp_cb"] = abi.chroma_qp(p["hdr"]["qp_y"])
        p["hdr"]["qp_cr"] = abi.chroma_qp(p["hdr"]["qp_y"], 2)
    want = run_sequence(oracle.recon_picture, g["w_mbs"], g["h_mbs"], 0, g["pics"])
    outs, _, _ = engine_sequence(g["w_mbs"], g["h_mbs"], 0, g["pics"], packed=packed)
    for i, (w, o) in enumerate(zip(want, outs)):
        for k in "yuv":
            assert np.array_equal(w[k], o[k]), "picture %d plane %s" % (i, k)


@pytest.mark.parametrize("flags", [abi.REF_COMPAT, 0])
def test_residual_kernel(built, flags):
    g = load_golden("cif_ip")
    eng = engine.Engine(g["w_mbs"], g["h_mbs"], max_frames=1, flags=flags)
    for p in g["pics"][:3]:
        got = eng.residual_mbs(p["hdr"], p["coeff"])
        want = oracle.residual(g["w_mbs"], g["h_mbs"], flags, p["hdr"], p["coeff"])
        for a, b in zip(got, want):
            assert np.array_equal(a.astype(np.int32), b)
    eng.close()
