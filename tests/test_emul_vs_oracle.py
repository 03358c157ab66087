"""CPU: the kernel code (h264r_core.cuh / h264r_seq.cuh, run by the host SIMT emulator) against the
oracle on directly generated syntax, in both modes, including what the reference lacks
(8x8 transform, Intra8x8, 8x4 sub-partitions, chroma prediction / MC, deblocking)."""
import numpy as np
import pytest

from common import *  # noqa: F401,F403
import emul
from videodecoder_b200 import workload


@pytest.mark.parametrize("compat", [True, False])
@pytest.mark.parametrize("seed,w,h", [(7, 8, 6), (8, 5, 3), (9, 1, 1), (10, 2, 7), (11, 1, 4), (12, 9, 1)])
def test_gop(built, compat, seed, w, h):
    rng = np.random.default_rng(seed)
    cfg = workload.WorkloadConfig(compat=compat, num_ref=2, p_intra_in_p=0.12, mv_range=90,
                                  p_t8x8=0.0 if compat else 0.5, p_i8x8=0.0 if compat else 0.5,
                                  p_sub=(0.4, 0, 0.3, 0.3) if compat else (0.25, 0.25, 0.25, 0.25), qp=int(rng.integers(14, 44)))
    gop = workload.make_gop(rng, w, h, 4, cfg)
    flags = abi.REF_COMPAT if compat else 0
    if compat:
        sanitize_with_oracle(w, h, gop)     # outside [0,255] the reference's int samples have no 8-bit equivalent (D3)
    want = run_sequence(oracle.recon_picture, w, h, flags, gop)
    got = run_sequence(emul.recon_picture, w, h, flags, gop)
    for i, (a, b) in enumerate(zip(want, got)):
        for k in "yuv":
            assert np.array_equal(a[k], b[k]), "picture %d plane %s" % (i, k)
        assert a["excursions"] == b["excursions"]
        assert oracle.digest(a["y"], a["u"], a["v"]) == emul.digest(b["y"], b["u"], b["v"])


@pytest.mark.parametrize("compat", [True, False])
@pytest.mark.parametrize("variant", [emul.PACKED, emul.PACKED | emul.COMPACT, emul.PACKED | emul.COMPACT | emul.GENERIC_INTRA, emul.GENERIC_INTRA,
                                     emul.PACKED | emul.COMPACT | emul.PARTS, emul.PARTS])
def test_gop_variants(built, compat, variant):
    """packed coefficient storage (what h264r_submit_mbs_packed delivers) and the generic intra phases give the
    same pictures as the default paths"""
    rng = np.random.default_rng(21)
    w, h = 7, 5
    cfg = workload.WorkloadConfig(compat=compat, num_ref=2, p_intra_in_p=0.2, mv_range=70, p_t8x8=0.0 if compat else 0.4,
                                  p_i8x8=0.0 if compat else 0.4, p_sub=(0.4, 0, 0.3, 0.3) if compat else (0.25, 0.25, 0.25, 0.25))
    gop = workload.make_gop(rng, w, h, 3, cfg)
    flags = abi.REF_COMPAT if compat else 0
    if compat:
        sanitize_with_oracle(w, h, gop)
    want = run_sequence(oracle.recon_picture, w, h, flags, gop)
    got = run_sequence(emul.recon_picture, w, h, flags | variant, gop)
    for i, (a, b) in enumerate(zip(want, got)):
        for k in "yuv":
            assert np.array_equal(a[k], b[k]), "picture %d plane %s" % (i, k)
        assert a["excursions"] == b["excursions"]


@pytest.mark.parametrize("raw", [False, True])
@pytest.mark.parametrize("seed,w,h,dense", [(31, 7, 5, False), (32, 11, 3, False), (33, 3, 3, True), (34, 1, 1, False), (35, 9, 2, True), (36, 17, 1, False)])
def test_group_sequence(built, seed, w, h, dense, raw):
    """the group sequence of k_inter_grp (h264r_mcgrp.cuh: residual of the coded blocks of eight macroblocks in dense passes, then
    staging + interpolation per macroblock) gives the oracle's pictures; picture sizes that leave partial groups, content dense
    enough that a group's coded blocks exceed the residual slots (the group is then taken in parts), intra macroblocks between
    the inter ones, weighted prediction off (the lean path); raw: the staging of the tensor-copy variant (aligned words, consumer shifts)"""
    rng = np.random.default_rng(seed)
    cfg = workload.WorkloadConfig(compat=True, num_ref=2, p_intra_in_p=0.15, mv_range=80, p_sub=(0.4, 0, 0.3, 0.3), qp=int(rng.integers(16, 40)))
    if dense:       # nearly every block of every macroblock coded, small amplitudes
        cfg = workload.WorkloadConfig(compat=True, num_ref=2, p_intra_in_p=0.1, mv_range=80, p_sub=(0.4, 0, 0.3, 0.3), qp=30, p_skip=0.05,
                                      p_cbp8=0.95, p_block=0.95, p_chroma=(0.1, 0.2, 0.7), max_amp=3.0)
    gop = workload.make_gop(rng, w, h, 4, cfg)
    sanitize_with_oracle(w, h, gop)
    if dense:
        assert max(int((p["coeff"].reshape(-1, 24, 16) != 0).any(axis=2).reshape(-1, 24)[:8].sum()) for p in gop[1:]) > 64
    want = run_sequence(oracle.recon_picture, w, h, abi.REF_COMPAT, gop)
    got = run_sequence(emul.recon_picture, w, h, abi.REF_COMPAT | emul.PACKED | emul.COMPACT | emul.GROUP | (emul.GROUP_RAW if raw else 0), gop)
    for i, (a, b) in enumerate(zip(want, got)):
        for k in "yuv":
            assert np.array_equal(a[k], b[k]), "picture %d plane %s" % (i, k)
        assert a["excursions"] == b["excursions"]


def test_compact_storage_dense_and_extreme_levels(built):
    """the kernels' block readers on H264R_LEVELS_COMPACT storage with compact, plain and mixed macroblocks (7..16 levels
    per block, levels at the int8 limits), fast and generic phases"""
    from common import densify_levels
    rng = np.random.default_rng(78)
    w, h = 6, 4
    cfg = workload.WorkloadConfig(compat=False, p_intra_in_p=0.25, num_ref=2, p_t8x8=0.3, p_i8x8=0.3, p_sub=(0.25, 0.25, 0.25, 0.25))
    gop = densify_levels(rng, workload.make_gop(rng, w, h, 3, cfg))
    want = run_sequence(oracle.recon_picture, w, h, 0, gop)
    for variant in (emul.PACKED | emul.COMPACT, emul.PACKED | emul.COMPACT | emul.GENERIC_INTRA):
        got = run_sequence(emul.recon_picture, w, h, variant, gop)
        for i, (a, b) in enumerate(zip(want, got)):
            for k in "yuv":
                assert np.array_equal(a[k], b[k]), "variant %x picture %d plane %s" % (variant, i, k)


def test_deblock_offsets_and_qp_extremes(built):
    """alpha/beta offsets and QPs at both ends of the tables"""
    for qp, offs in ((2, (0, 6, 6)), (50, (0, -6, -6)), (30, (0, 3, -3))):
        rng = np.random.default_rng(qp)
        cfg = workload.WorkloadConfig(compat=False, qp=qp, p_intra_in_p=0.2, max_amp=40.0)
        gop = workload.make_gop(rng, 6, 4, 3, cfg)
        for p in gop:
            p["pp"].disable_deblocking_filter_idc, p["pp"].slice_alpha_c0_offset_div2, p["pp"].slice_beta_offset_div2 = offs
        want = run_sequence(oracle.recon_picture, 6, 4, 0, gop)
        got = run_sequence(emul.recon_picture, 6, 4, 0, gop)
        for a, b in zip(want, got):
            for k in "yuv":
                assert np.array_equal(a[k], b[k])


def test_deblocking_disabled_equals_plain_reconstruction(built):
    rng = np.random.default_rng(5)
    cfg = workload.WorkloadConfig(compat=False)
    gop = workload.make_gop(rng, 4, 4, 2, cfg)
    for p in gop:
        p["pp"].disable_deblocking_filter_idc = 1
    a = run_sequence(emul.recon_picture, 4, 4, 0, gop)
    for p in gop:
        p["pp"].disable_deblocking_filter_idc = 0
    b = run_sequence(emul.recon_picture, 4, 4, 0, gop)
    assert not np.array_equal(a[0]["y"], b[0]["y"])      # the filter does something on this content


def test_residual_linearity_property(built):
    """with QP >= 24 dequantisation is linear: residual(2c) - 2 residual(c) stays within rounding."""
    rng = np.random.default_rng(3)
    cfg = workload.WorkloadConfig(compat=True, qp=30)
    pic = workload.make_picture(rng, 6, 4, False, cfg)
    pic["hdr"]["mb_class"] = abi.MB_I4x4
    r1 = emul.residual(6, 4, abi.REF_COMPAT, pic["hdr"], pic["coeff"])[0].astype(np.int32)
    r2 = emul.residual(6, 4, abi.REF_COMPAT, pic["hdr"], pic["coeff"] * 2)[0].astype(np.int32)
    assert np.abs(r2 - 2 * r1).max() <= 1
