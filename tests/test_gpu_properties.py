"""GPU tests at BASELINE.json's full sizes and at the edges of the input space, through properties that do not need a
full-size oracle run for every picture: one flush vs one flush per picture, dense vs packed vs picture-buffer
transport, any number of lanes, determinism -- all must give identical digests -- plus direct oracle comparison on
ragged / tiny / large geometries and on the content edge cases of the domain (all-skip pictures, vectors far
outside the picture, every quarter-sample position, every intra mode, QP extremes, empty residual)."""
import os

import numpy as np
import pytest

from common import *  # noqa: F401,F403
from videodecoder_b200 import engine, workload

pytestmark = pytest.mark.gpu


def run_engine(w, h, flags, gops, transport="buf", flush="all", max_frames=None):
    """gops: list of GOPs (lists of pictures); frame ids are assigned GOP-major.  Returns (digests, planes of the last
    picture of every GOP, excursions)."""
    n = sum(len(g) for g in gops)
    eng = engine.Engine(w, h, max_frames=max_frames or n, flags=flags | abi.DIGESTS)
    bufs, fid0 = [], 0
    order = []
    for g in gops:
        for i, p in enumerate(g):
            order.append((fid0 + i, p, fid0))
        fid0 += len(g)
    if flush == "round":        # picture i of every GOP, then picture i + 1, ... flushed per round
        order.sort(key=lambda t: (t[0] - t[2], t[2]))
    last_round = None
    for fid, p, base in order:
        if flush == "round" and last_round is not None and fid - base != last_round:
            eng.flush()
        last_round = fid - base
        pp = abi.PicParams.from_buffer_copy(p["pp"])
        for k in range(pp.num_ref_idx_l0):
            pp.ref_list0[k] = base + p["pp"].ref_list0[k]
        if transport == "dense":
            eng.submit_picture(fid, pp, p["hdr"], p["mv"], p["coeff"])
        else:
            mask, blocks = pack.pack_blocks(p["coeff"])
            if transport == "packed":
                eng.submit_picture_packed(fid, pp, np.ascontiguousarray(p["hdr"]), p["mv"], mask, blocks)
            else:
                b = eng.picture_buf(0 if transport == "bufstream" else max(1, blocks.shape[0]))
                # bufmv: the narrowest storage (compact units); bufmv8: int8 blocks; bufmv16: int16 blocks
                # bufmvh8: compact units + 8-byte headers; bufstream: the stream form
                if not (transport == "bufstream" and b.fill_stream(p["hdr"], p["mv"], p["coeff"])):
                    b.fill(p["hdr"], p["mv"], mask, blocks, mv_partition=transport in ("bufmv", "bufmv8", "bufmv16", "bufmvh8", "bufstream"),
                           level8=False if transport == "bufmv16" else None, compact=False if transport == "bufmv8" else None,
                           short_hdr=transport == "bufmvh8")
                bufs.append(b)
                eng.submit_picture_buf(fid, pp, b)
        if flush == "each":
            eng.sync()
    eng.sync()
    dig = eng.digests(list(range(n)))
    lasts, k = [], 0
    for g in gops:
        k += len(g)
        lasts.append(eng.read_planes(k - 1))
    exc = eng.range_excursions()
    for b in bufs:
        b.free()
    eng.close()
    return dig, lasts, exc


def oracle_gop(w, h, flags, gop):
    return run_sequence(oracle.recon_picture, w, h, flags, gop)


def check_vs_oracle(w, h, flags, gop, **kw):
    dig, lasts, _ = run_engine(w, h, flags, [gop], **kw)
    want = oracle_gop(w, h, flags, gop)
    for i, o in enumerate(want):
        assert bytes(dig[i]) == oracle.digest(o["y"], o["u"], o["v"]), "picture %d digest" % i
    for k, pl in zip("yuv", lasts[0]):
        assert np.array_equal(pl, want[-1][k]), "last picture plane %s" % k


# ---- geometries ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("compat", [True, False])
@pytest.mark.parametrize("w,h", [(1, 1), (1, 5), (7, 1), (2, 2), (33, 3), (45, 36)])
def test_ragged_geometries(built, compat, w, h):
    """one macroblock, single rows / columns, widths that are not multiples of the 32-macroblock scan words"""
    rng = np.random.default_rng(w * 100 + h)
    cfg = workload.WorkloadConfig(compat=compat, p_intra_in_p=0.15, mv_range=80, num_ref=2, p_t8x8=0.0 if compat else 0.4,
                                  p_i8x8=0.0 if compat else 0.4, p_sub=(0.4, 0, 0.3, 0.3) if compat else (0.25, 0.25, 0.25, 0.25))
    gop = workload.make_gop(rng, w, h, 4, cfg)
    flags = abi.REF_COMPAT if compat else 0
    if compat:
        sanitize_with_oracle(w, h, gop)
    check_vs_oracle(w, h, flags, gop, transport="bufmv")


@pytest.mark.gpu
@pytest.mark.parametrize("transport", ["bufmv", "bufmv8", "bufmvh8", "bufstream", "packed"])
def test_compact_transport_dense_and_extreme_levels(built, transport):
    """H264R_LEVELS_COMPACT with macroblocks of every kind: all blocks sparse (one unit per block), blocks with 7..16 levels
    (plain int8 blocks, two units), levels at -128 / 127; Intra16x16 and chroma DC levels read through the compact form.
    Spec mode (Clip1 keeps any residual legal), against the oracle."""
    from common import densify_levels
    w, h = 13, 7
    rng = np.random.default_rng(77)
    cfg = workload.WorkloadConfig(compat=False, p_intra_in_p=0.2, num_ref=2, p_sub=(0.25, 0.25, 0.25, 0.25))
    gop = densify_levels(rng, workload.make_gop(rng, w, h, 3, cfg))
    if transport != "packed":
        mask, blocks = pack.pack_blocks(gop[1]["coeff"])
        cmask, units = pack.pack_compact(mask, blocks)
        compact = (cmask & abi.BLK_COMPACT) != 0
        assert compact.any() and (~compact & (mask != 0)).any()        # both kinds of macroblock are on the wire
    check_vs_oracle(w, h, 0, gop, transport=transport)


def test_4k_pictures_vs_oracle(built):
    """BASELINE config 5 geometry (3840x2160 = 240x135 MBs): I + 2 P against the oracle"""
    w, h = 240, 135
    rng = np.random.default_rng(5)
    cfg = workload.WorkloadConfig(compat=True)
    gop = workload.make_gop(rng, w, h, 3, cfg)
    sanitize_with_oracle(w, h, gop)
    check_vs_oracle(w, h, abi.REF_COMPAT, gop, transport="bufmv")


# ---- full-size invariants ----------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def hd_gops():
    """BASELINE config 4 shape, shortened: 1080p, 4 GOPs x 6 pictures (content model of bench.py)"""
    w, h = 120, 68
    rng = np.random.default_rng(44)
    cfg = workload.WorkloadConfig(compat=True)
    gops = [workload.make_gop(rng, w, h, 6, cfg) for _ in range(2)]
    for g in gops:
        workload.sanitize(lambda n: engine.Engine(w, h, max_frames=n, flags=abi.REF_COMPAT), w, h, g)
    return w, h, [gops[0], gops[1], gops[0], gops[1]]


def test_1080p_transport_and_flush_invariance(built, hd_gops):
    """dense / packed / picture-buffer transports, one flush / per picture / per round: identical digests, and the
    first GOP equals the oracle"""
    w, h, gops = hd_gops
    ref, lasts, exc = run_engine(w, h, abi.REF_COMPAT, gops, transport="dense", flush="all")
    assert exc == 0
    for transport, flush in (("packed", "all"), ("buf", "each"), ("bufmv", "round"), ("bufmv", "all"), ("bufmvh8", "round"), ("bufmvh8", "all"), ("bufstream", "round"), ("bufstream", "all"), ("bufmv8", "all"), ("bufmv16", "all")):
        dig, _, exc = run_engine(w, h, abi.REF_COMPAT, gops, transport=transport, flush=flush)
        assert exc == 0
        assert np.array_equal(dig, ref), (transport, flush)
    want = oracle_gop(w, h, abi.REF_COMPAT, gops[0])
    for i, o in enumerate(want):
        assert bytes(ref[i]) == oracle.digest(o["y"], o["u"], o["v"])
    assert np.array_equal(ref[:6], ref[12:18]) and np.array_equal(ref[6:12], ref[18:24])      # equal GOPs, equal digests
    for k, pl in zip("yuv", lasts[0]):
        assert np.array_equal(pl, want[-1][k])


@pytest.mark.parametrize("lanes", ["1", "3", "8"])
def test_1080p_lane_count_invariance(built, hd_gops, lanes):
    w, h, gops = hd_gops
    old = os.environ.get("H264R_LANES")
    try:
        os.environ["H264R_LANES"] = "4"
        ref, _, _ = run_engine(w, h, abi.REF_COMPAT, gops, transport="bufmv", flush="round")
        os.environ["H264R_LANES"] = lanes
        dig, _, _ = run_engine(w, h, abi.REF_COMPAT, gops, transport="bufmv", flush="round")
    finally:
        if old is None:
            os.environ.pop("H264R_LANES", None)
        else:
            os.environ["H264R_LANES"] = old
    assert np.array_equal(dig, ref)


def test_1080p_tma_window_staging(built, hd_gops):
    """k_inter<.., TMA>: reference windows by cp.async.bulk.tensor (16-byte aligned boxes + consumer-side shift) give the
    same pictures as load staging"""
    w, h, gops = hd_gops
    old = os.environ.get("H264R_TMA")
    try:
        os.environ["H264R_TMA"] = "0"
        ref, _, _ = run_engine(w, h, abi.REF_COMPAT, gops, transport="bufmv")
        os.environ["H264R_TMA"] = "1"
        dig, _, exc = run_engine(w, h, abi.REF_COMPAT, gops, transport="bufmv")
        assert exc == 0 and np.array_equal(dig, ref)
        rng = np.random.default_rng(77)         # spec mode, small picture, vectors far outside, against the oracle
        cfg = workload.WorkloadConfig(compat=False, mv_range=200, num_ref=2, p_t8x8=0.3)
        gop = workload.make_gop(rng, 9, 5, 4, cfg)
        check_vs_oracle(9, 5, 0, gop, transport="bufmv")
    finally:
        if old is None:
            os.environ.pop("H264R_TMA", None)
        else:
            os.environ["H264R_TMA"] = old


def test_group_kernel_tensor_copy_ring(built, hd_gops):
    """k_inter_grp<.., TMA> (H264R_GROUP_TMA=1): the windows of 16x16 macroblocks by cp.async.bulk.tensor into a two-deep ring, every
    other window as aligned words, shifted consumer: the same pictures as the default load staging -- 1080p GOPs, and a small
    picture whose vectors point far outside (boxes at every border of the padded plane) against the oracle"""
    w, h, gops = hd_gops
    old = os.environ.get("H264R_GROUP_TMA")
    try:
        os.environ["H264R_GROUP_TMA"] = "0"
        ref, _, _ = run_engine(w, h, abi.REF_COMPAT, gops, transport="bufmv")
        os.environ["H264R_GROUP_TMA"] = "1"
        dig, _, exc = run_engine(w, h, abi.REF_COMPAT, gops, transport="bufmvh8", flush="round")
        assert exc == 0 and np.array_equal(dig, ref)
        for seed, (pw, ph) in ((78, (9, 5)), (79, (4, 3)), (80, (1, 2))):
            rng = np.random.default_rng(seed)
            cfg = workload.WorkloadConfig(compat=True, mv_range=300, num_ref=2, p_intra_in_p=0.1, p_skip=0.5)
            gop = workload.make_gop(rng, pw, ph, 4, cfg)
            sanitize_with_oracle(pw, ph, gop)
            check_vs_oracle(pw, ph, abi.REF_COMPAT, gop, transport="bufmv")
    finally:
        if old is None:
            os.environ.pop("H264R_GROUP_TMA", None)
        else:
            os.environ["H264R_GROUP_TMA"] = old


def test_frame_slot_reuse_across_gops(built, hd_gops):
    """a frame pool smaller than the job: slots are released and reused GOP after GOP (Decoder::ctrl_Memory's role)"""
    w, h, gops = hd_gops
    want, _, _ = run_engine(w, h, abi.REF_COMPAT, gops[:2], transport="bufmv")
    eng = engine.Engine(w, h, max_frames=6, max_pending=6, flags=abi.REF_COMPAT | abi.DIGESTS)
    got = []
    for g in gops[:2]:
        for i, p in enumerate(g):
            eng.submit_picture(i, p["pp"], p["hdr"], p["mv"], p["coeff"])
        eng.flush()
        got.append(eng.digests(list(range(len(g)))))
        for i in range(len(g)):
            eng.release(i)
    eng.close()
    assert np.array_equal(np.concatenate(got), want)


# ---- content edge cases ------------------------------------------------------------------------------------------
def test_all_skip_picture_is_a_copy(built):
    """P picture of P_Skip macroblocks with zero vectors: every plane equals the reference picture (COMPAT: luma)"""
    w, h = 12, 9
    rng = np.random.default_rng(8)
    cfg = workload.WorkloadConfig(compat=True)
    gop = workload.make_gop(rng, w, h, 2, cfg)
    sanitize_with_oracle(w, h, gop[:1])
    p = gop[1]
    p["hdr"]["mb_class"] = abi.MB_P16x16
    p["hdr"]["flags"] |= abi.SKIP
    p["hdr"]["cbp"] = 0
    p["hdr"]["sub_mb_type"] = 0
    p["hdr"]["u"][:] = 0
    p["mv"][:] = 0
    p["coeff"][:] = 0
    dig, lasts, exc = run_engine(w, h, abi.REF_COMPAT, [gop], transport="bufmv")
    want = oracle_gop(w, h, abi.REF_COMPAT, gop)
    assert exc == 0
    assert np.array_equal(lasts[0][0], want[0]["y"]) and np.array_equal(lasts[0][0], want[1]["y"])


@pytest.mark.parametrize("compat", [True, False])
def test_every_fractional_position_and_far_vectors(built, compat):
    """16x16 macroblocks sweeping all 16 quarter-sample positions, with vectors up to 300 samples outside the picture
    (coordinate clamp of h264/macroblock.cpp:828-830 = replicated border + clamped block position)"""
    w, h = 16, 4
    rng = np.random.default_rng(12)
    cfg = workload.WorkloadConfig(compat=compat, p_intra_in_p=0.0, p_skip=0.0)
    gop = workload.make_gop(rng, w, h, 2, cfg)
    if compat:
        sanitize_with_oracle(w, h, gop[:1])
    p = gop[1]
    n = w * h
    p["hdr"]["mb_class"] = abi.MB_P16x16
    p["hdr"]["cbp"] = 0
    p["hdr"]["sub_mb_type"] = 0
    p["hdr"]["flags"] &= ~np.uint8(abi.T8x8)
    p["hdr"]["u"][:] = 0
    p["coeff"][:] = 0
    far = np.array([0, 1200, -1200, 37, -41, 600, -600, 5], dtype=np.int16)
    mv = np.zeros((n, 16, 2), dtype=np.int16)
    for a in range(n):
        mv[a, :, 0] = (a % 16) % 4 + 4 * far[(a // 4) % 8]
        mv[a, :, 1] = (a % 16) // 4 + 4 * far[(a // 7) % 8]
    p["mv"][:] = mv.reshape(n, 32)
    check_vs_oracle(w, h, abi.REF_COMPAT if compat else 0, gop, transport="bufmv")


@pytest.mark.parametrize("qp", [0, 1, 23, 24, 51])
def test_qp_extremes(built, qp):
    """both branches of the dequantisation shift (qP < 24 rounds, qP >= 24 shifts left) and the ends of the tables"""
    w, h = 6, 4
    rng = np.random.default_rng(qp)
    cfg = workload.WorkloadConfig(compat=False, qp=qp, max_amp=30.0)
    gop = workload.make_gop(rng, w, h, 3, cfg)
    for p in gop:
        p["hdr"]["qp_y"] = qp
        p["hdr"]["qp_cb"] = abi.chroma_qp(p["hdr"]["qp_y"])
        p["hdr"]["qp_cr"] = abi.chroma_qp(p["hdr"]["qp_y"])
    check_vs_oracle(w, h, 0, gop, transport="bufmv")


def test_empty_residual_i_picture(built):
    """an I picture with no coefficient at all (every block absent in the packed transport)"""
    w, h = 5, 4
    rng = np.random.default_rng(2)
    gop = workload.make_gop(rng, w, h, 1, workload.WorkloadConfig(compat=True))
    gop[0]["coeff"][:] = 0
    gop[0]["hdr"]["cbp"] = 0
    check_vs_oracle(w, h, abi.REF_COMPAT, gop, transport="bufmv")


def test_api_errors_on_device(built):
    """error behaviour of the calls added for the packed transport"""
    eng = engine.Engine(4, 3, max_frames=2, flags=abi.REF_COMPAT)
    rng = np.random.default_rng(1)
    gop = workload.make_gop(rng, 4, 3, 2, workload.WorkloadConfig(compat=True))
    p = gop[0]
    mask, blocks = pack.pack_blocks(p["coeff"])
    eng.frame_begin(0, p["pp"])
    with pytest.raises(engine.H264RError):      # dense after packed within one picture
        eng.submit_mbs_packed(0, np.ascontiguousarray(p["hdr"][:6]), None, mask[:6], pack.pack_blocks(p["coeff"][:6])[1], n_mbs=6)
        eng.submit_mbs(0, np.ascontiguousarray(p["hdr"][6:]), None, np.ascontiguousarray(p["coeff"][6:]), first_mb=6, n_mbs=6)
    eng.close()
    eng = engine.Engine(4, 3, max_frames=2, flags=abi.REF_COMPAT)
    eng.frame_begin(0, p["pp"])
    with pytest.raises(engine.H264RError):      # packed chunks out of macroblock order
        eng.submit_mbs_packed(0, np.ascontiguousarray(p["hdr"][6:]), None, mask[6:], pack.pack_blocks(p["coeff"][6:])[1], first_mb=6, n_mbs=6)
    eng.close()


# ---- scheduler hazards and API edges (round-1 advisor findings) -------------------------------------------------------
@pytest.mark.parametrize("lanes", [1, 4])
def test_frame_slot_overwritten_while_queued_readers_exist(built, lanes, monkeypatch):
    """Write after read on a frame slot: two independent P chains share ONE pool, and each picture overwrites the slot the
    other chain's queued picture still reads.  All in one flush, 1 and 4 kernel pipelines.  Result must equal the oracle
    of each chain (h264r_api.cu: reader_wave / frame_reader_lanes)."""
    monkeypatch.setenv("H264R_LANES", str(lanes))
    w, h = 24, 10
    rng = np.random.default_rng(321 + lanes)
    cfg = workload.WorkloadConfig(compat=False, p_intra_in_p=0.1, mv_range=40)
    chains = [workload.make_gop(rng, w, h, 5, cfg) for _ in range(2)]
    want = [oracle_gop(w, h, 0, c) for c in chains]
    # slots: chain c, picture i lives in slot (2*i + c) % 3 -- three slots for two chains: a slot is reused as soon as its
    # picture has been referenced by the NEXT picture of its chain, while that picture is still queued
    eng = engine.Engine(w, h, max_frames=3, max_pending=16, flags=0)
    eng.set_lanes(lanes)
    slot_of = {}
    got = {}

    def submit(c, i):
        p = chains[c][i]
        pp = abi.PicParams.from_buffer_copy(p["pp"])
        for k in range(pp.num_ref_idx_l0):
            pp.ref_list0[k] = slot_of[(c, p["pp"].ref_list0[k])]
        used = {slot_of[(c, i - 1)]} if i > 0 else set()
        live = {slot_of[(1 - c, j)] for j in range(5) if (1 - c, j) in slot_of and j == max(k for (cc, k) in slot_of if cc == 1 - c)}
        free = [s for s in range(3) if s not in used and s not in live]
        s = free[0]
        slot_of[(c, i)] = s
        eng.submit_picture(s, pp, p["hdr"], p["mv"], p["coeff"])

    # one flush per round of two pictures (a slot whose picture is still QUEUED cannot be begun again; a DONE one can, and
    # that is the hazard: round i's second picture overwrites the slot round i's first picture reads).  The check is on
    # the LAST picture of each chain: it depends on every earlier one through prediction.
    for i in range(5):
        for c in range(2):
            submit(c, i)
        eng.flush()
    eng.sync()
    for c in range(2):
        y, u, v = eng.read_planes(slot_of[(c, 4)])
        assert np.array_equal(y, want[c][4]["y"]) and np.array_equal(u, want[c][4]["u"]) and np.array_equal(v, want[c][4]["v"]), "chain %d" % c
    eng.close()


def test_frame_abort_and_open_reader_pins_slot(built):
    w, h = 6, 4
    rng = np.random.default_rng(9)
    cfg = workload.WorkloadConfig(compat=False)
    gop = workload.make_gop(rng, w, h, 2, cfg)
    eng = engine.Engine(w, h, max_frames=3, max_pending=2, flags=0)
    eng.submit_picture(0, gop[0]["pp"], gop[0]["hdr"], None, gop[0]["coeff"])
    pp = abi.PicParams.from_buffer_copy(gop[1]["pp"])
    eng.frame_begin(1, pp)                       # open reader of slot 0
    with pytest.raises(engine.H264RError) as e:
        eng.frame_begin(0, gop[0]["pp"])         # overwriting its reference: refused
    assert e.value.code == abi.E_STATE
    # a failed submit (chunk out of order) can be retried or the picture given up
    with pytest.raises(engine.H264RError):
        eng.submit_mbs(1, gop[1]["hdr"][4:], gop[1]["mv"][4:], gop[1]["coeff"][4:], first_mb=4)
    eng.frame_abort(1)
    with pytest.raises(engine.H264RError):
        eng.frame_abort(1)
    # both the frame slot and the syntax slot are usable again
    eng.submit_picture(1, pp, gop[1]["hdr"], gop[1]["mv"], gop[1]["coeff"])
    eng.sync()
    want = oracle_gop(w, h, 0, gop)
    y, u, v = eng.read_planes(1)
    assert np.array_equal(y, want[1]["y"]) and np.array_equal(u, want[1]["u"]) and np.array_equal(v, want[1]["v"])
    eng.close()


def test_device_status_reports_bad_picture_buffers(built):
    """picture buffers with a parser-supplied n_intra skip the host pass over the headers; the device checks instead"""
    w, h = 8, 5
    rng = np.random.default_rng(10)
    cfg = workload.WorkloadConfig(compat=False, p_intra_in_p=0.3)
    gop = workload.make_gop(rng, w, h, 2, cfg)
    eng = engine.Engine(w, h, max_frames=4, flags=0)

    def send(fid, p, n_intra=None, patch=None):
        mask, blocks = pack.pack_blocks(p["coeff"])
        b = eng.picture_buf(max(1, blocks.shape[0]))
        hdr = p["hdr"].copy()
        if patch:
            patch(hdr)
        b.fill(hdr, p["mv"], mask, blocks, mv_partition=True)
        if n_intra is not None:
            b.c.n_intra = n_intra
        eng.submit_picture_buf(fid, p["pp"], b)
        return b

    bufs = [send(0, gop[0]), send(1, gop[1])]
    eng.sync()
    assert eng.device_status() == 0
    true_intra = int((gop[1]["hdr"]["mb_class"] <= abi.MB_I16x16).sum())
    assert true_intra >= 2
    bufs.append(send(2, gop[1], n_intra=true_intra - 1))
    eng.sync()
    assert eng.device_status() == abi.STATUS_N_INTRA
    assert eng.device_status() == 0              # reset on read

    def bad_class(hdr):
        hdr["mb_class"][3] = 3
    bufs.append(send(3, gop[1], patch=bad_class))
    eng.sync()
    assert eng.device_status() & abi.STATUS_BAD_CLASS
    # stream form: word counts that do not add up, or do not match what a macroblock's header and mask announce
    for damage in (None, "count", "total"):
        p = gop[1]
        b = eng.picture_buf(0)
        assert b.fill_stream(p["hdr"], p["mv"], p["coeff"])
        if damage == "count":
            a = int(np.flatnonzero(b.mb_words)[0])
            b.mb_words[a] -= 1; b.mb_words[a + 1] += 1          # the total still agrees
        elif damage == "total":
            b.n_blocks += 1
        eng.submit_picture_buf(2, p["pp"], b)
        eng.sync()
        assert eng.device_status() == (abi.STATUS_STREAM if damage else 0), damage
        bufs.append(b)
    for b in bufs:
        b.free()
    eng.close()


def test_create_rejects_unsupported_width(built):
    with pytest.raises(engine.H264RError) as e:
        engine.Engine(513, 2, max_frames=1, flags=0)
    assert e.value.code == abi.E_INVAL


@pytest.mark.gpu
@pytest.mark.parametrize("short_hdr", [False, True])
def test_ring_submit_equals_single_submits(built, short_hdr):
    """h264r_picture_ring_alloc + h264r_submit_pictures: a round of pictures (one per GOP) shipped with one strided transfer gives
    the digests of per-picture submits; members begun in another order than they lie in the ring (slots not consecutive) fall
    back to single transfers and give the same"""
    import ctypes
    w, h, n_gops, n_pics = 9, 6, 5, 4
    rng = np.random.default_rng(91)
    cfg = workload.WorkloadConfig(compat=True, p_intra_in_p=0.1, num_ref=1)
    gops = [workload.make_gop(rng, w, h, n_pics, cfg) for _ in range(n_gops)]
    for g in gops:
        sanitize_with_oracle(w, h, g)
    ref, _, exc = run_engine(w, h, abi.REF_COMPAT, gops, transport="bufmv", flush="round")
    assert exc == 0
    for order in ("ring", "reversed"):
        eng = engine.Engine(w, h, max_frames=n_gops * n_pics, flags=abi.REF_COMPAT)
        rings = []
        for i in range(n_pics):
            packed = [pack.pack_blocks(g[i]["coeff"]) for g in gops]
            ring = eng.picture_ring(n_gops, max(1, max(b.shape[0] for _, b in packed)))
            fids, pps = [], []
            for k, g in enumerate(gops):
                p = g[i]
                ring[k].fill(p["hdr"], p["mv"], packed[k][0], packed[k][1], mv_partition=True, short_hdr=short_hdr)
                pp = abi.PicParams.from_buffer_copy(p["pp"])
                for r in range(pp.num_ref_idx_l0):
                    pp.ref_list0[r] = k * n_pics + p["pp"].ref_list0[r]
                fids.append(k * n_pics + i); pps.append(pp)
            if order == "ring":
                eng.submit_ring(ring, fids, pps)
            else:
                for k in reversed(range(n_gops)):        # slots are handed out in begin order: the ring's members get descending slots
                    eng.frame_begin(fids[k], pps[k])
                ids = (ctypes.c_int * n_gops)(*fids)
                nb = (ctypes.c_size_t * n_gops)(*[ring[k].n_blocks for k in range(n_gops)])
                nm = (ctypes.c_long * n_gops)(*[ring[k].n_mvs for k in range(n_gops)])
                eng._check(eng.L.h264r_submit_pictures(eng.ctx, n_gops, ids, ring.array, nb, nm))
            rings.append(ring)
            eng.flush()
        eng.sync()
        dig = eng.digests(list(range(n_gops * n_pics)))
        assert eng.range_excursions() == 0
        assert np.array_equal(dig, ref), order
        for r in rings:
            r.free()
        eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("compat", [True, False])
def test_maximum_picture_width(built, compat):
    """the widest picture the engine takes (512 macroblocks = 8192 samples; 513 is rejected, test_create_rejects_unsupported_width):
    the row masks of the wavefront kernels are full, the group kernel's last group of a row straddles picture rows"""
    w, h = 512, 2
    rng = np.random.default_rng(512)
    cfg = workload.WorkloadConfig(compat=compat, p_intra_in_p=0.05, num_ref=2, mv_range=120, p_t8x8=0.0 if compat else 0.4, p_i8x8=0.0 if compat else 0.3)
    gop = workload.make_gop(rng, w, h, 3, cfg)
    if compat:
        sanitize_with_oracle(w, h, gop)
    check_vs_oracle(w, h, abi.REF_COMPAT if compat else 0, gop, transport="bufmvh8")
    check_vs_oracle(w, h, abi.REF_COMPAT if compat else 0, gop, transport="bufstream")
