"""The reference-side binding of INTEGRATION.md, compiled against the reference's own classes
(oracle/shims/adapter_main.cpp -> oracle/_ref/h264ref_adapter, built where /root/reference exists; the binary travels).

CPU: what the adapter packs out of the reference's macroblock / residual objects equals what the Python twin (pack.py, fed
with the synthesizer's abstract syntax) packs -- the two host-side packers pin each other.
GPU: reference parser -> adapter -> C ABI -> engine, planes compared inside the adapter with the reference decoder's own
CPU reconstruction of the same stream: the drop-in, end to end, in the reference's language."""
import numpy as np
import pytest

from common import *  # noqa: F401,F403

pytestmark = pytest.mark.skipif(not refharness.have_adapter(), reason="reference adapter not built")


def _stream(seed, w, h, n_pics, **cfg):
    rng = np.random.default_rng(seed)
    seq = synth.make_seq(w, h, max_num_ref_frames=cfg.get("num_ref", 1))
    pics = synth.random_gop(rng, seq, n_pics, synth.ContentConfig(**cfg))
    stream, queue, dump = sanitize_with_reference(seq, pics)
    return seq, pics, stream, queue, dump


@pytest.mark.parametrize("seed,num_ref", [(61, 1), (62, 2)])
def test_adapter_packs_what_pack_py_packs(built, seed, num_ref):
    seq, pics, stream, queue, dump = _stream(seed, 7, 5, 4, num_ref=num_ref, p_intra_in_p=0.15, mvd_range=20)
    packed = refharness.run_adapter(stream, queue, n_mbs=seq.w_mbs * seq.h_mbs, pack_only=True)
    assert len(packed) == len(pics)
    for i, ((pic, mbs), d, (hdr_c, mv_c, coeff_c)) in enumerate(zip(pics, dump, packed)):
        pp, hdr, mv, coeff = buffers_from_dump(seq, pic, mbs, d)
        assert np.array_equal(hdr.view(np.uint8).reshape(-1, 16), hdr_c), "picture %d headers" % i
        assert np.array_equal(coeff, coeff_c), "picture %d levels" % i
        if mv is not None:
            inter = hdr["mb_class"] >= abi.MB_P16x16
            assert np.array_equal(mv[inter], mv_c[inter]), "picture %d vectors" % i


@pytest.mark.parametrize("seed,big_levels", [(63, False), (64, True)])
def test_c_writer_fills_picture_buffers_like_pack_py(built, seed, big_levels):
    """include/h264r_writer.h (the C helper a parser calls per macroblock) against the Python twin: masks with the compact
    flag, 8-byte units / int16 blocks, partition-order vectors, n_intra -- byte for byte"""
    seq, pics, stream, queue, dump = _stream(seed, 8, 6, 4, num_ref=2, p_intra_in_p=0.15, mvd_range=20)
    if big_levels:      # a level outside int8 somewhere in picture 1: that picture must fall back to int16 blocks
        mbs = pics[1][1]
        a = int(np.flatnonzero(mbs["cbp_luma"] != 0)[0])
        b = int(np.flatnonzero((mbs["luma"][a] != 0).any(axis=1))[0])
        k = int(np.flatnonzero(mbs["luma"][a][b])[0])
        mbs["luma"][a][b][k] = 300
        stream, queue = synth.serialize_stream(seq, pics)
    n = seq.w_mbs * seq.h_mbs
    bufs = refharness.run_adapter(stream, queue, n_mbs=n, pack_only=True, transport="buf")
    dump, _ = refharness.run_reference(stream, queue, level=2)
    saw_int16 = False
    for i, ((pic, mbs), d, b) in enumerate(zip(pics, dump, bufs)):
        pp, hdr, mv, coeff = buffers_from_dump(seq, pic, mbs, d)
        assert np.array_equal(hdr.view(np.uint8).reshape(-1, 16), b["hdr"]), "picture %d headers" % i
        mask, blocks = pack.pack_blocks(coeff)
        if blocks.size and (blocks.min() < -128 or blocks.max() > 127):
            saw_int16 = True
            assert b["level_bytes"] == abi.LEVELS_INT16
            assert np.array_equal(mask, b["blk_mask"]) and np.array_equal(blocks.view(np.uint8).reshape(-1, 32), b["units"])
        else:
            cmask, units = pack.pack_compact(mask, blocks)
            assert b["level_bytes"] == abi.LEVELS_COMPACT
            assert np.array_equal(cmask, b["blk_mask"]), "picture %d masks" % i
            assert np.array_equal(units, b["units"]), "picture %d units" % i
        want_mvs = pack.pack_mvs_partition(hdr, mv) if mv is not None else np.zeros((0, 2), dtype=np.int16)
        assert np.array_equal(want_mvs, b["mvs"]), "picture %d vectors" % i
        assert b["n_intra"] == int((hdr["mb_class"] <= abi.MB_I16x16).sum())
    assert saw_int16 == big_levels


@pytest.mark.gpu
@pytest.mark.parametrize("transport", [None, "buf"])
@pytest.mark.parametrize("seed,w,h,n_pics,num_ref", [(71, 11, 9, 3, 1), (72, 22, 18, 6, 2)])
def test_reference_host_side_drives_the_engine(built, seed, w, h, n_pics, num_ref, transport):
    """QCIF / CIF streams: the engine, fed by the reference's own parser through the adapter, reproduces the reference
    decoder's reconstruction bit for bit"""
    seq, pics, stream, queue, dump = _stream(seed, w, h, n_pics, num_ref=num_ref, p_intra_in_p=0.1, mvd_range=24)
    verdict = refharness.run_adapter(stream, queue, transport=transport)      # None: dense arrays; "buf": h264r_writer + h264r_submit_picture
    assert verdict["pictures"] == n_pics
    assert verdict["luma_mismatches"] == 0 and verdict["chroma_mismatches"] == 0 and verdict["returncode"] == 0
    assert verdict["luma_range_excursions"] == 0
