"""CPU tests of the host side (libh264host.so: CABAC stream writer + decoder, include/h264_host.h) and of the spec-mode oracle
against LIBAVCODEC (tests/lavc.py) -- the independent pin of the pieces the reference decoder lacks
-- and against the UNMODIFIED reference running its own arithmetic decoder on the same genuine bitstreams."""
import numpy as np
import pytest

from common import *  # noqa: F401,F403
import lavc
import streams
from videodecoder_b200 import hoststream

SPEC_SMALL = ["spec_qcif_i", "spec_cif_ip", "spec_qcif_wp", "spec_dpb", "base_qcif_i", "base_cif_ip"]
need_lavc = pytest.mark.skipif(not lavc.available(), reason="no libavcodec in this image")
need_ref_cabac = pytest.mark.skipif(not refharness.have_cabac_harness(), reason="reference harness (own CABAC) not built")


def oracle_on_stream(stream, flags=0, dflags=0):
    pics, tail = hoststream.decode_stream(stream, flags=dflags)
    g = pics[0]
    return pics, run_sequence(oracle.recon_picture, g["w_mbs"], g["h_mbs"], flags, pics)


@pytest.mark.parametrize("high", [False, True])
def test_writer_reader_round_trip(built, high):
    """what the CABAC writer emits, the decoder reads back: syntax-element records identical, macroblock by macroblock."""
    rng = np.random.default_rng(11)
    if high:
        seq = synth.make_seq(13, 7, max_num_ref_frames=3, num_ref_idx_default=2, profile_idc=100, transform_8x8_mode=1, deblocking_control=1)
        cfg = synth.ContentConfig(num_ref=3, sub_types=(0, 1, 2, 3), p_intra_in_p=0.15, p_i8x8=0.4, p_t8x8=0.5, p_qp_delta=0.5, qp_delta_range=6)
    else:
        seq = synth.make_seq(13, 7, max_num_ref_frames=2, num_ref_idx_default=1)
        cfg = synth.ContentConfig(num_ref=2, sub_types=(0, 1, 2, 3), p_intra_in_p=0.15)
    pics = synth.random_gop(rng, seq, 5, cfg)
    # a few large values: Exp-Golomb suffixes of coeff_abs_level_minus1 and mvd
    pics[0][1]["luma"][3][2][0] = 2000
    pics[0][1]["cbp_luma"][3] |= 1
    pics[1][1]["mvd"][:, 0, :] = rng.integers(-900, 900, size=(pics[1][1].size, 2))
    stream = hoststream.encode_stream(seq, pics)
    dec, tail = hoststream.decode_stream(stream, flags=hoststream.KEEP_SYNTAX)
    assert len(dec) == len(pics) and tail == {"output": [], "release": []}
    for i, ((pic, mbs), d) in enumerate(zip(pics, dec)):
        for name in mbs.dtype.names:
            assert np.array_equal(mbs[name], d["syntax"][name]), "picture %d field %s" % (i, name)
        assert d["is_idr"] == pic.is_idr and d["frame_num"] == pic.frame_num and d["poc"] == pic.poc_lsb


@pytest.mark.parametrize("name", SPEC_SMALL + ["spec_hd1080_ip"])
def test_spec_oracle_equals_libavcodec_golden(built, name):
    """host decoder + spec-mode oracle on the committed bitstream == the planes libavcodec decoded from it (fixture)."""
    g = load_stream_golden(name)
    _, outs = oracle_on_stream(g["stream"])
    check_against_stream_golden(g, outs, "oracle (spec mode) vs libavcodec")


@need_lavc
@pytest.mark.parametrize("name", SPEC_SMALL)
def test_recipes_live_against_libavcodec(built, name):
    """the recipe reproduces the committed stream byte for byte, and libavcodec run now gives the committed planes."""
    g = load_stream_golden(name)
    s = streams.SPEC_RECIPES[name]()
    assert s["stream"] == g["stream"]
    frames = lavc.decode(s["aus"])
    check_against_stream_golden(g, [dict(zip("yuv", f)) for f in frames], "libavcodec now vs fixture")


@need_lavc
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_high_profile_streams_against_libavcodec(built, seed):
    """fresh random content every seed: constrained intra prediction, chroma QP offsets, cabac_init_idc, odd geometries."""
    rng = np.random.default_rng(500 + seed)
    w, h = [(5, 3), (9, 4), (3, 8)][seed - 1]
    cip = seed == 2
    seq = synth.make_seq(w, h, max_num_ref_frames=2, num_ref_idx_default=1, profile_idc=100, transform_8x8_mode=1, deblocking_control=1,
                         constrained_intra_pred=int(cip), chroma_qp_index_offset=int(rng.integers(-6, 7)),
                         second_chroma_qp_index_offset=int(rng.integers(-6, 7)), pic_init_qp=int(rng.integers(18, 40)))
    cfg = synth.ContentConfig(num_ref=2, sub_types=(0, 1, 2, 3), p_intra_in_p=0.25, p_i8x8=0.4, p_t8x8=0.5, constrained_intra=cip,
                              p_qp_delta=0.5, qp_delta_range=5, mv_range=120)
    pics = synth.random_gop(rng, seq, 5, cfg, absolute=True)
    for i, (p, _) in enumerate(pics):
        p.disable_deblocking_filter_idc = (0, 0, 1, 0, 0)[i]
        p.cabac_init_idc = i % 3
        p.slice_qp_delta = int(rng.integers(-3, 4))
        p.slice_alpha_c0_offset_div2 = int(rng.integers(-6, 7))
        p.slice_beta_offset_div2 = int(rng.integers(-6, 7))
    stream, aus = hoststream.encode_stream(seq, pics, flags=hoststream.E_ABSOLUTE, per_picture=True)
    frames = lavc.decode(aus)
    _, outs = oracle_on_stream(stream)
    assert len(frames) == len(outs) == len(pics)
    for i, (f, o) in enumerate(zip(frames, outs)):
        for k, ref in zip("yuv", f):
            assert np.array_equal(o[k], ref), "picture %d plane %s" % (i, k)


def test_absolute_mode_gives_the_requested_vectors_and_modes(built):
    """E_ABSOLUTE: the writer derives mvd / prev / rem so that the decoder's final vectors and intra modes are the requested ones."""
    rng = np.random.default_rng(12)
    seq = synth.make_seq(9, 6, profile_idc=100, transform_8x8_mode=1)
    cfg = synth.ContentConfig(sub_types=(0, 1, 2, 3), p_intra_in_p=0.3, p_i8x8=0.5, p_skip=0.1)
    pics = synth.random_gop(rng, seq, 3, cfg, absolute=True)
    want = [(m["kind"].copy(), m["mvd"].copy(), m["rem_mode"].copy(), m["sub_type"].copy()) for _, m in pics]
    stream = hoststream.encode_stream(seq, pics, flags=hoststream.E_ABSOLUTE)
    dec, _ = hoststream.decode_stream(stream)
    for (kind, mv, modes, sub), d in zip(want, dec):
        for a in range(kind.size):
            h = d["hdr"][a]
            if kind[a] == synth.I4x4:
                got = [(h["u"][b >> 1] >> (4 * (b & 1))) & 15 for b in range(16)]
                assert got == [int(x) for x in modes[a]]
            elif kind[a] == synth.I8x8:
                got = [(h["u"][b >> 1] >> (4 * (b & 1))) & 15 for b in range(4)]
                assert got == [int(x) for x in modes[a][:4]]
            elif kind[a] in (synth.P16x16, synth.P16x8, synth.P8x16, synth.P8x8):
                part_mv = mv[a].reshape(4, 4, 2)
                assert np.array_equal(d["mv"][a].reshape(16, 2), pack.expand_mv(int(kind[a]), sub[a], part_mv)), "MB %d" % a


def test_ref_quirks_reproduce_the_reference_host_side(built):
    """H264D_REF_QUIRKS: vectors, intra modes, QPs and reference lists equal what the reference decoder's own parser derived
    (picture::get_MvNeighbour with its neighbour-lookup quirks, D2, D11, D12), on every partition shape."""
    rng = np.random.default_rng(13)
    seq = synth.make_seq(13, 8, max_num_ref_frames=2, num_ref_idx_default=1)
    pics = synth.random_gop(rng, seq, 5, synth.ContentConfig(num_ref=2, p_intra_in_p=0.2))
    stream_q, queue = synth.serialize_stream(seq, pics)
    dump, _ = refharness.run_reference(stream_q, queue, level=2)
    dec, _ = hoststream.decode_stream(hoststream.encode_stream(seq, pics), flags=hoststream.REF_QUIRKS)
    for i, ((pic, mbs), d, h) in enumerate(zip(pics, dump, dec)):
        pp, hdr, mv, coeff = buffers_from_dump(seq, pic, mbs, d)
        assert np.array_equal(hdr.view(np.uint8), h["hdr"].view(np.uint8)), "picture %d headers" % i
        assert np.array_equal(coeff, h["coeff"]), "picture %d levels" % i
        if mv is not None:
            assert np.array_equal(mv, h["mv"]), "picture %d vectors" % i
        assert [pp.ref_list0[k] for k in range(pp.num_ref_idx_l0)] == [h["pp"].ref_list0[k] for k in range(h["pp"].num_ref_idx_l0)]


@need_ref_cabac
def test_unmodified_reference_reads_the_writers_stream(built):
    """the reference decoder with its OWN arithmetic decoder (h264/cabac.cpp) on a genuine CABAC stream: its planes are the
    fixture's, and host decoder (REF_QUIRKS) + oracle (REF_COMPAT) reconstruct the same planes from the same bytes."""
    g = load_stream_golden("ref_cif_ip_stream")
    dump, stats = refharness.run_reference_cabac(g["stream"], level=1)
    assert stats["pictures"] == g["n_pics"]
    check_against_stream_golden(g, dump, "reference (own CABAC) now vs fixture")
    _, outs = oracle_on_stream(g["stream"], flags=abi.REF_COMPAT, dflags=hoststream.REF_QUIRKS)
    check_against_stream_golden(g, outs, "host decoder + oracle vs reference")
    assert all(o["excursions"] == 0 for o in outs)


def test_reference_stream_fixture_through_host_decoder(built):
    """the same without the reference binary (what the GPU box can run): fixture planes == host decoder + oracle."""
    g = load_stream_golden("ref_cif_ip_stream")
    _, outs = oracle_on_stream(g["stream"], flags=abi.REF_COMPAT, dflags=hoststream.REF_QUIRKS)
    check_against_stream_golden(g, outs, "host decoder + oracle vs reference fixture")


def test_dpb_with_a_small_frame_pool(built):
    """frame slots are reused as the DPB frees them: with a pool of 6 (4 references + current + 1) the stream of spec_dpb decodes to
    the same pictures as with one slot per picture; a slot is released only when no later picture lists it."""
    g = load_stream_golden("spec_dpb")
    d = hoststream.HostDecoder(g["stream"], max_frames=6)
    store, outs, live = {}, [], set()
    for p in d:
        assert 0 <= p["frame_id"] < 6
        pp = p["pp"]
        refs = []
        for i in range(pp.num_ref_idx_l0):
            assert pp.ref_list0[i] in live, "picture %d lists slot %d, which was released" % (p["decode_index"], pp.ref_list0[i])
            refs.append(store[pp.ref_list0[i]])
        o = oracle.recon_picture(p["w_mbs"], p["h_mbs"], 0, pp, p["hdr"], p["mv"], p["coeff"], refs=refs, ref_ids=[pp.ref_list0[i] for i in range(pp.num_ref_idx_l0)])
        store[p["frame_id"]] = (o["y"], o["u"], o["v"])
        live.add(p["frame_id"])
        outs.append(o)
        assert [x[2] for x in p["output"]] == [p["decode_index"]]       # I / P: output order is decoding order
        for s in p["release"]:
            live.discard(s)
    assert d.tail["output"] == []
    d.close()
    check_against_stream_golden(g, outs, "small pool")


def test_decoder_rejects_damaged_and_unsupported_streams(built):
    g = load_stream_golden("spec_qcif_i")
    s = g["stream"]
    # truncated inside the first picture
    with pytest.raises(RuntimeError, match="-1"):
        hoststream.decode_stream(s[:len(g["aus"][0]) - 200])
    # bytes flipped in the slice data: must end with an error or a picture, never a crash
    rng = np.random.default_rng(5)
    for k in range(20):
        b = bytearray(s)
        for pos in rng.integers(60, len(b), size=8):
            b[pos] ^= 1 << int(rng.integers(0, 8))
        try:
            hoststream.decode_stream(bytes(b))
        except RuntimeError:
            pass
    # the same for a CAVLC stream (mb_skip_run, coeff_token, level escapes, run_before all read from damaged bits)
    s2 = load_stream_golden("base_cif_ip")["stream"]
    for k in range(40):
        b = bytearray(s2)
        for pos in rng.integers(60, len(b), size=8):
            b[pos] ^= 1 << int(rng.integers(0, 8))
        try:
            hoststream.decode_stream(bytes(b))
        except RuntimeError:
            pass
    # a CABAC stream whose parameter set claims CAVLC: garbage in, an error or garbage out, no crash
    seq = synth.make_seq(3, 3)
    pics = synth.random_gop(np.random.default_rng(1), seq, 1, synth.ContentConfig())
    good = hoststream.encode_stream(seq, pics)
    i = good.index(b"\x00\x00\x00\x01\x68")
    bad = bytearray(good)
    bad[i + 5] &= 0xDF          # pps: ue(0) ue(0) then entropy_coding_mode_flag = bit 2 of the first payload byte ('1 1 1' -> '1 1 0')
    try:
        hoststream.decode_stream(bytes(bad))
    except RuntimeError:
        pass


def test_host_decoder_throughput_is_reported(built):
    g = load_stream_golden("spec_cif_ip")
    n, sec = hoststream.benchmark(g["stream"], repeat=2)
    assert n == 2 * g["n_pics"] and sec > 0
    # ... and with every picture packed into the stream form of a picture buffer (what bench.py's host block reports beside it)
    n2, sec2 = hoststream.benchmark(g["stream"], flags=hoststream.BENCH_PACK, repeat=2)
    assert n2 == n and sec2 > 0


@need_lavc
@pytest.mark.parametrize("name", sorted(streams.KAT_RECIPES))
def test_known_answer_streams_against_libavcodec(built, name):
    """one tool at a time (every Intra4x4 / Intra8x8 / Intra16x16 / chroma mode; every boundary-strength class of the deblocking
    filter): libavcodec's planes == host decoder + spec-mode oracle"""
    s = streams.KAT_RECIPES[name]()
    frames = lavc.decode(s["aus"])
    pics, outs = oracle_on_stream(s["stream"])
    assert len(frames) == len(outs) == len(s["pics"])
    for i, (f, o) in enumerate(zip(frames, outs)):
        for k, ref in zip("yuv", f):
            assert np.array_equal(o[k], ref), "%s picture %d plane %s" % (name, i, k)
    if name == "kat_intra_modes":       # the stream really carries every mode
        seen4, seen8, seen16 = set(), set(), set()
        for p in pics:
            for h in p["hdr"]:
                nib = [(int(h["u"][b >> 1]) >> (4 * (b & 1))) & 15 for b in range(16)]
                if h["mb_class"] == abi.MB_I4x4:
                    seen4.update(nib)
                elif h["mb_class"] == abi.MB_I8x8:
                    seen8.update(nib[:4])
                elif h["mb_class"] == abi.MB_I16x16:
                    seen16.add(int(h["intra_modes"]) & 3)
        assert seen4 == set(range(9)) and seen8 == set(range(9)) and seen16 == set(range(4))


# ---- CAVLC (entropy_coding_mode_flag = 0; profile_idc 66 selects it in the writer) ---------------------------------------------
def _cavlc_table(name):
    import os
    import re
    src = open(os.path.join(ROOT, "videodecoder_b200", "csrc", "host", "cavlc_tables.h")).read()
    m = re.search(r"%s(\[[^=]*\])+\s*=\s*(\{.*?\});" % name, src, re.S)
    rows = re.findall(r"\{([^{}]*)\}", m.group(2))
    return [[int(x) for x in r.replace("\n", " ").split(",") if x.strip()] for r in rows]


def test_cavlc_tables_are_prefix_codes():
    """every VLC table of cavlc_tables.h (Tables 9-5, 9-7, 9-8, 9-9a, 9-10 of the standard) is prefix-free, and complete up to the one
    all-zero codeword the standard leaves unused in its longest-code tables; the me(v) maps are permutations of 0..47"""
    from fractions import Fraction

    def kraft(lens, bits):
        codes = [format(b, "0%db" % l) for l, b in zip(lens, bits) if l]
        assert len(set(codes)) == len(codes)
        for a in codes:
            assert not any(b != a and b.startswith(a) for b in codes), a
        return sum(Fraction(1, 2 ** len(c)) for c in codes), max(len(c) for c in codes), len(codes)

    for lens, bits, entries in zip(_cavlc_table("kCoeffTokenLen"), _cavlc_table("kCoeffTokenBits"), (62, 62, 62, 62)):
        k, longest, n = kraft(lens, bits)
        assert n == entries and (1 - k).numerator == 1 and (1 - k).denominator >= 32        # one unused (all-zero) codeword
    assert kraft(_cavlc_table("kChromaDcTokenLen")[0], _cavlc_table("kChromaDcTokenBits")[0])[0] == 1
    for tc, (lens, bits) in enumerate(zip(_cavlc_table("kTotalZerosLen"), _cavlc_table("kTotalZerosBits")), 1):
        k, longest, n = kraft(lens, bits)
        assert n == 17 - tc and (k == 1 if tc > 1 else (1 - k).numerator == 1)
    for lens, bits in zip(_cavlc_table("kChromaDcTotalZerosLen"), _cavlc_table("kChromaDcTotalZerosBits")):
        assert kraft(lens, bits)[0] == 1
    for zl, (lens, bits) in enumerate(zip(_cavlc_table("kRunLen"), _cavlc_table("kRunBits")), 1):
        lens, bits = lens + [0] * (16 - len(lens)), bits + [0] * (16 - len(bits))
        k, longest, n = kraft(lens, bits)
        assert n == (15 if zl == 7 else zl + 1) and (k == 1 if zl < 7 else (1 - k).numerator == 1)
    assert sorted(_cavlc_table("kCbpIntra")[0]) == list(range(48)) and sorted(_cavlc_table("kCbpInter")[0]) == list(range(48))


@pytest.mark.parametrize("profile", [66, 100])
def test_cavlc_writer_reader_round_trip(built, profile):
    """CAVLC slices: what the writer emits the decoder reads back, record for record -- Baseline (profile_idc 66), and the 8x8
    transform travelling as four interleaved 4x4 blocks (checked through the codec templates with a High-profile PPS is not
    possible from the writer, which ties CAVLC to profile 66: the High case round-trips through CABAC above)"""
    if profile != 66:
        pytest.skip("the writer selects CAVLC through profile_idc 66 only")
    rng = np.random.default_rng(21)
    seq = synth.make_seq(13, 7, max_num_ref_frames=3, num_ref_idx_default=2, profile_idc=66, deblocking_control=1)
    cfg = synth.ContentConfig(num_ref=3, sub_types=(0, 1, 2, 3), p_intra_in_p=0.15, p_qp_delta=0.5, qp_delta_range=6)
    pics = synth.random_gop(rng, seq, 5, cfg)
    # large values: level_prefix 14 / 15 escapes and beyond, long mvd codes
    pics[0][1]["luma"][3][2][0] = 2000
    pics[0][1]["luma"][3][2][1] = -17
    pics[0][1]["luma"][3][2][5] = 40
    pics[0][1]["cbp_luma"][3] |= 1
    pics[1][1]["mvd"][:, 0, :] = rng.integers(-900, 900, size=(pics[1][1].size, 2))
    stream = hoststream.encode_stream(seq, pics)
    dec, tail = hoststream.decode_stream(stream, flags=hoststream.KEEP_SYNTAX)
    assert len(dec) == len(pics)
    for i, ((pic, mbs), d) in enumerate(zip(pics, dec)):
        for name in mbs.dtype.names:
            assert np.array_equal(mbs[name], d["syntax"][name]), "picture %d field %s" % (i, name)


@need_lavc
@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_baseline_cavlc_streams_against_libavcodec(built, seed):
    """Baseline-profile (CAVLC) streams of random content -- BASELINE config 1's "baseline-profile stream" made literal: QCIF I-only
    (seed 1) and I+P with every partition shape, several references, QP deltas, deblocking on: libavcodec's planes == host decoder +
    spec-mode oracle"""
    rng = np.random.default_rng(700 + seed)
    w, h = [(11, 9), (9, 4), (5, 6), (22, 18)][seed - 1]
    seq = synth.make_seq(w, h, max_num_ref_frames=2, num_ref_idx_default=1, profile_idc=66, deblocking_control=1,
                         chroma_qp_index_offset=int(rng.integers(-4, 5)), pic_init_qp=int(rng.integers(20, 38)))
    cfg = synth.ContentConfig(num_ref=2, sub_types=(0, 1, 2, 3), p_intra_in_p=0.25, p_qp_delta=0.5, qp_delta_range=5, mv_range=100)
    n = 3 if seed == 1 else 5
    pics = synth.random_gop(rng, seq, n, cfg, absolute=True, intra_only=(seed == 1))
    for i, (p, _) in enumerate(pics):
        p.slice_qp_delta = int(rng.integers(-3, 4))
    stream, aus = hoststream.encode_stream(seq, pics, flags=hoststream.E_ABSOLUTE, per_picture=True)
    frames = lavc.decode(aus)
    _, outs = oracle_on_stream(stream)
    assert len(frames) == len(outs) == len(pics)
    for i, (f, o) in enumerate(zip(frames, outs)):
        for k, ref in zip("yuv", f):
            assert np.array_equal(o[k], ref), "picture %d plane %s" % (i, k)
