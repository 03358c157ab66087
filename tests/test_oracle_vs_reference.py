"""CPU: the oracle (oracle/restate) and the host emulation of the kernel code against the COMPILED
REFERENCE DECODER run live (oracle/_ref/h264ref_harness), on freshly synthesized streams.  This is
what pins the oracle; it needs the harness binary (built in the container that has /root/reference)."""
import numpy as np
import pytest

from common import *  # noqa: F401,F403
import emul

pytestmark = pytest.mark.skipif(not refharness.have_harness(), reason="reference harness not built")


def _check(seq, pics):
    stream, queue, dump = sanitize_with_reference(seq, pics)
    outs = oracle_sequence(seq, pics, dump)
    for i, (d, o) in enumerate(zip(dump, outs)):
        assert o["excursions"] == 0
        assert np.array_equal(d["raw"], o["raw"]), "picture %d: int luma differs from the reference" % i
        for k in "yuv":
            assert np.array_equal(d[k], o[k]), "picture %d plane %s" % (i, k)
    return dump


@pytest.mark.parametrize("seed", [11, 12])
def test_intra_pictures(built, seed):
    rng = np.random.default_rng(seed)
    seq = synth.make_seq(7, 5)
    _check(seq, synth.random_gop(rng, seq, 2, synth.ContentConfig(p_i16=0.4), intra_only=True))


@pytest.mark.parametrize("seed,num_ref", [(21, 1), (22, 3)])
def test_inter_pictures(built, seed, num_ref):
    rng = np.random.default_rng(seed)
    seq = synth.make_seq(9, 6, max_num_ref_frames=num_ref)
    _check(seq, synth.random_gop(rng, seq, 5, synth.ContentConfig(num_ref=num_ref, p_intra_in_p=0.1, mvd_range=12)))


def test_far_motion_vectors_clamp(built):
    rng = np.random.default_rng(31)
    seq = synth.make_seq(6, 4)
    _check(seq, synth.random_gop(rng, seq, 4, synth.ContentConfig(mvd_range=400, p_skip=0.1)))


def test_residual_against_reference_dump(built):
    """residual::Decode output per macroblock (a1-a7) straight from the reference's own matrices."""
    rng = np.random.default_rng(41)
    seq = synth.make_seq(8, 5)
    pics = synth.random_gop(rng, seq, 2, synth.ContentConfig(p_qp_delta=0.6, qp_delta_range=6))
    stream, queue = synth.serialize_stream(seq, pics)
    dump, _ = refharness.run_reference(stream, queue, level=2)
    for (pic, mbs), d in zip(pics, dump):
        pp, hdr, mv, coeff = buffers_from_dump(seq, pic, mbs, d)
        for impl in (oracle, emul):
            ry, ru, rv = impl.residual(seq.w_mbs, seq.h_mbs, abi.REF_COMPAT, hdr, coeff)
            assert np.array_equal(ry.astype(np.int32), d["mb"]["res_y"])
            assert np.array_equal(ru.astype(np.int32), d["mb"]["res_u"])
            assert np.array_equal(rv.astype(np.int32), d["mb"]["res_v"])


def test_queue_is_consumed_exactly(built):
    rng = np.random.default_rng(51)
    seq = synth.make_seq(5, 4)
    pics = synth.random_gop(rng, seq, 3, synth.ContentConfig())
    stream, queue = synth.serialize_stream(seq, pics)
    _, stats = refharness.run_reference(stream, queue, level=0)
    assert stats["pictures"] == 3 and stats["queue_used"] == stats["queue_total"] == queue.size // 2
