"""TEST INFRASTRUCTURE shared by the test modules: turning (abstract syntax + what the compiled
reference derived) into h264r buffers, sanitising content, running sequences through the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from videodecoder_b200 import abi, pack, synth  # noqa: E402
import oracle  # noqa: E402
import refharness  # noqa: E402


def buffers_from_dump(seq, pic, mbs, dpic):
    """h264r buffers of one picture from its abstract syntax plus the values the reference's host
    side derived (final MVs, Intra4x4PredMode, QPs: macroblock::Parse, h264/macroblock.cpp:143-475)."""
    m = dpic["mb"]
    n = mbs.shape[0]
    hdr = pack.pack_headers(seq.w_mbs, seq.h_mbs, mbs, m["qpy"], m["qpc"], i4_modes=m["i4_modes"], part_ref=m["ref_idx"])
    coeff = pack.pack_coeff(mbs)
    is_p = pic.slice_type == synth.SLICE_P
    mv = pack.pack_mvs(mbs, m["mv"]) if is_p else None
    pp = abi.default_pic_params(abi.SLICE_P if is_p else abi.SLICE_I, ref_list=dpic["ref_list"] if is_p else ())
    if is_p and seq.weighted_pred_flag:
        pp.weighted_pred = 1
        pp.luma_log2_weight_denom = pic.luma_log2_weight_denom
        for i in range(pic.num_ref_idx_l0_active_minus1 + 1):
            if pic.luma_weight_flag[i]:
                pp.luma_weight[i], pp.luma_offset[i] = pic.luma_weight[i], pic.luma_offset[i]
            else:
                pp.luma_weight[i], pp.luma_offset[i] = 1 << pic.luma_log2_weight_denom, 0
    return pp, hdr, mv, coeff


def excursion_mbs(dpic):
    """addresses of MBs whose reference int samples left [0,255]."""
    wm, hm = dpic["w_mbs"], dpic["h_mbs"]
    r = dpic["raw"].reshape(hm, 16, wm, 16)
    bad = (r.min(axis=(1, 3)) < 0) | (r.max(axis=(1, 3)) > 255)
    return np.flatnonzero(bad.reshape(-1))


def sanitize_with_reference(seq, pics, max_iter=40):
    """Zeroes the residual of macroblocks that make the reference's samples leave [0,255] until the
    whole sequence is clean (content restriction that removes the no-Clip1 deviation D3).
    Returns (stream, queue, dump)."""
    for it in range(max_iter):
        stream, queue = synth.serialize_stream(seq, pics)
        dump, _ = refharness.run_reference(stream, queue, level=2)
        dirty = False
        for (pic, mbs), d in zip(pics, dump):
            bad = excursion_mbs(d)
            if bad.size:
                synth.zero_residual(mbs, bad)
                dirty = True
                break       # later pictures depend on this one: fix in decode order
        if not dirty:
            return stream, queue, dump
    raise RuntimeError("content did not converge to the [0,255] range")


def oracle_sequence(seq, pics, dump, flags=abi.REF_COMPAT):
    """Runs a whole sequence through the oracle using buffers derived from `dump`.
    Returns list of oracle outputs (dict y,u,v,...)."""
    out = []
    for (pic, mbs), d in zip(pics, dump):
        pp, hdr, mv, coeff = buffers_from_dump(seq, pic, mbs, d)
        refs = [(out[i]["y"], out[i]["u"], out[i]["v"]) for i in d["ref_list"]] if pic.slice_type == synth.SLICE_P else []
        o = oracle.recon_picture(seq.w_mbs, seq.h_mbs, flags, pp, hdr, mv, coeff, refs=refs, want_raw=True)
        out.append(o)
    return out


GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    """Fixture written by tests/golden/make_golden.py -> dict(w_mbs, h_mbs, pics=[dict(pp, hdr, mv, coeff, md5, y,u,v)])."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    out = {"w_mbs": int(z["w_mbs"]), "h_mbs": int(z["h_mbs"]), "pics": []}
    for i in range(int(z["n_pics"])):
        v = z["pp%d" % i]
        st, nref, wp, ld = (int(x) for x in v[:4])
        pp = abi.default_pic_params(st, ref_list=[int(x) for x in v[4:4 + nref]])
        pp.weighted_pred, pp.luma_log2_weight_denom = wp, ld
        for k in range(abi.MAX_REFS):
            pp.luma_weight[k] = int(v[20 + k])
            pp.luma_offset[k] = int(v[36 + k])
        pic = {"pp": pp, "hdr": z["hdr%d" % i].copy().view(abi.MB_HDR_DTYPE).reshape(-1), "coeff": z["coeff%d" % i],
               "mv": z["mv%d" % i] if ("mv%d" % i) in z.files else None, "md5": bytes(z["md5_%d" % i]).hex()}
        if ("y%d" % i) in z.files:
            pic["y"], pic["u"], pic["v"] = z["y%d" % i], z["u%d" % i], z["v%d" % i]
        out["pics"].append(pic)
    return out


def planes_md5(y, u, v):
    import hashlib
    return hashlib.md5(np.ascontiguousarray(y).tobytes() + np.ascontiguousarray(u).tobytes() + np.ascontiguousarray(v).tobytes()).hexdigest()


def run_sequence(recon, w_mbs, h_mbs, flags, pics):
    """Runs pictures [(pp, hdr, mv, coeff)...] through `recon` (oracle.recon_picture or
    emul.recon_picture), resolving ref_list0 frame ids as indices into the sequence."""
    outs = []
    for p in pics:
        pp = p["pp"]
        refs = [(outs[pp.ref_list0[i]]["y"], outs[pp.ref_list0[i]]["u"], outs[pp.ref_list0[i]]["v"]) for i in range(pp.num_ref_idx_l0)]
        outs.append(recon(w_mbs, h_mbs, flags, pp, p["hdr"], p["mv"], p["coeff"], refs=refs))
    return outs


def sanitize_with_oracle(w_mbs, h_mbs, gop, max_iter=40):
    """REF_COMPAT content restriction for directly generated syntax: zero the residual of macroblocks
    whose luma leaves [0,255] (oracle's excursion map), picture by picture in decode order."""
    from videodecoder_b200 import workload
    outs = []
    for p in gop:
        pp = p["pp"]
        refs = [(outs[pp.ref_list0[i]]["y"], outs[pp.ref_list0[i]]["u"], outs[pp.ref_list0[i]]["v"]) for i in range(pp.num_ref_idx_l0)]
        for _ in range(max_iter):
            o = oracle.recon_picture(w_mbs, h_mbs, abi.REF_COMPAT, pp, p["hdr"], p["mv"], p["coeff"], refs=refs)
            if o["excursions"] == 0:
                break
            workload.zero_mbs(p, np.flatnonzero(o["excursion_map"]))
        else:
            raise RuntimeError("no convergence")
        outs.append(o)
    return gop


def densify_levels(rng, gop, p_block=0.4, big=True):
    """Makes the coefficient blocks of a GOP (spec-mode content: Clip1 keeps any value legal) dense and varied: a share of
    the blocks that carry AC levels gets 7..15 of them, a few levels sit at the int8 limits, so that compact, plain-int8
    and mixed macroblocks all occur in the compact transport.  Only positions 1..15 of blocks that already have an AC level
    are touched: those are legal whatever the macroblock type and coded_block_pattern say."""
    for p in gop:
        c = p["coeff"].reshape(-1, 24, 16)
        has_ac = (c[:, :, 1:] != 0).any(axis=2)
        for a, b in zip(*np.nonzero(has_ac)):
            if rng.random() < p_block:
                n = int(rng.integers(7, 16))
                pos = 1 + rng.choice(15, size=n, replace=False)
                c[a, b, pos] = rng.choice([-3, -2, -1, 1, 2, 3], size=n)
            elif big and rng.random() < 0.1:
                c[a, b, 1 + int(rng.integers(0, 15))] = int(rng.choice([-128, 127]))
    return gop


def load_stream_golden(name):
    """Fixture written by tests/golden/make_spec_golden.py -> dict(w_mbs, h_mbs, stream bytes, aus [bytes], md5 [hex], planes [(y,u,v)] or None)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    stream = z["stream"].tobytes()
    aus, off = [], 0
    for n in z["au_sizes"]:
        aus.append(stream[off:off + int(n)])
        off += int(n)
    n_pics = int(z["n_pics"])
    out = {"w_mbs": int(z["w_mbs"]), "h_mbs": int(z["h_mbs"]), "stream": stream, "aus": aus, "n_pics": n_pics,
           "md5": [bytes(z["md5_%d" % i]).hex() for i in range(n_pics)],
           "planes": [(z["y%d" % i], z["u%d" % i], z["v%d" % i]) for i in range(n_pics)] if "y0" in z.files else None}
    return out


def check_against_stream_golden(g, outs, what):
    """outs: per picture dict(y,u,v) in decoding order (= output order for these streams)."""
    assert len(outs) == g["n_pics"]
    for i, o in enumerate(outs):
        if g["planes"] is not None:
            for k, ref in zip("yuv", g["planes"][i]):
                assert np.array_equal(o[k], ref), "%s: picture %d plane %s differs (%d samples)" % (what, i, k, int((o[k] != ref).sum()))
        assert planes_md5(o["y"], o["u"], o["v"]) == g["md5"][i], "%s: picture %d MD5" % (what, i)
