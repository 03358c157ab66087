"""Generates the golden fixtures of tests/golden/ by running the COMPILED REFERENCE DECODER
(oracle/_ref/h264ref_harness, built from /root/reference/h264 by oracle/Makefile) on synthesized
streams.  Run in the build container (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Each fixture holds the h264r input buffers of every picture (derived from the abstract syntax and
from what the reference's own host side computed: final motion vectors, intra modes, QPs) and the
planes the reference reconstructed (picture::get_SampleXY semantics), plus their MD5.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from common import *  # noqa: E402,F403


def pp_to_dict(pp):
    return {
        "slice_type": pp.slice_type, "num_ref_idx_l0": pp.num_ref_idx_l0, "ref_list0": list(pp.ref_list0),
        "weighted_pred": pp.weighted_pred, "luma_log2_weight_denom": pp.luma_log2_weight_denom,
        "luma_weight": list(pp.luma_weight), "luma_offset": list(pp.luma_offset),
    }


def make(name, seq, pics, keep_planes=True):
    stream, queue, dump = sanitize_with_reference(seq, pics)
    out = {"w_mbs": seq.w_mbs, "h_mbs": seq.h_mbs, "n_pics": len(pics)}
    for i, ((pic, mbs), d) in enumerate(zip(pics, dump)):
        pp, hdr, mv, coeff = buffers_from_dump(seq, pic, mbs, d)
        out["hdr%d" % i] = hdr.view(np.uint8).reshape(-1, 16)
        out["coeff%d" % i] = coeff
        if mv is not None:
            out["mv%d" % i] = mv
        ppd = pp_to_dict(pp)
        out["pp%d" % i] = np.array([ppd["slice_type"], ppd["num_ref_idx_l0"], ppd["weighted_pred"], ppd["luma_log2_weight_denom"]]
                                   + ppd["ref_list0"] + ppd["luma_weight"] + ppd["luma_offset"], dtype=np.int32)
        md5 = hashlib.md5(d["y"].tobytes() + d["u"].tobytes() + d["v"].tobytes()).hexdigest()
        out["md5_%d" % i] = np.frombuffer(bytes.fromhex(md5), dtype=np.uint8)
        if keep_planes:
            out["y%d" % i], out["u%d" % i], out["v%d" % i] = d["y"], d["u"], d["v"]
        assert d["raw"].min() >= 0 and d["raw"].max() <= 255
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    # the stream + queue themselves for the smallest case: lets the CPU suite re-run the reference
    print(name, "pictures", len(pics), "stream bytes", len(stream), "queue ints", queue.size,
          "file KiB", os.path.getsize(os.path.join(HERE, name + ".npz")) // 1024)


def main():
    # config 1: QCIF I-only (Intra4x4 all 9 modes, Intra16x16 all 4 modes)
    rng = np.random.default_rng(1)
    seq = synth.make_seq(11, 9)
    make("qcif_i", seq, synth.random_gop(rng, seq, 3, synth.ContentConfig(), intra_only=True))
    # config 2: CIF I+P: all partition shapes, 16 fractional positions, 2 references, intra MBs in P
    rng = np.random.default_rng(2)
    seq = synth.make_seq(22, 18, max_num_ref_frames=2)
    make("cif_ip", seq, synth.random_gop(rng, seq, 8, synth.ContentConfig(num_ref=2, p_intra_in_p=0.08)))
    # explicit weighted prediction, 3 references, vectors far outside the picture
    rng = np.random.default_rng(3)
    seq = synth.make_seq(11, 9, max_num_ref_frames=3, weighted_pred=1)
    pics = synth.random_gop(rng, seq, 6, synth.ContentConfig(num_ref=3, p_intra_in_p=0.1, mvd_range=90, p_skip=0.2))
    for i, (pic, mbs) in enumerate(pics):
        if pic.slice_type == synth.SLICE_P:
            pic.luma_log2_weight_denom = 5 if i % 2 else 0
            for r in range(pic.num_ref_idx_l0_active_minus1 + 1):
                pic.luma_weight_flag[r] = 1 if (r + i) % 2 == 0 else 0
                pic.luma_weight[r] = int(rng.integers(20, 44)) if i % 2 else 1
                pic.luma_offset[r] = int(rng.integers(-6, 7))
    make("qcif_wp", seq, pics)
    # 1080p (120x68 MBs) I + P pair, MD5 only
    rng = np.random.default_rng(4)
    seq = synth.make_seq(120, 68)
    make("hd1080_ip", seq, synth.random_gop(rng, seq, 2, synth.ContentConfig(p_intra_in_p=0.03)), keep_planes=False)


if __name__ == "__main__":
    main()
