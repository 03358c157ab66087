"""Generates the spec-mode fixtures tests/golden/spec_*.npz and ref_*_stream.npz.  Run in the build container:

    python tests/golden/make_spec_golden.py

spec_* / base_*: a genuine High-profile CABAC / Baseline-profile CAVLC bitstream written by the repository's stream writer (recipes in tests/streams.py) and the
planes LIBAVCODEC decoded from it (tests/lavc.py: the libavcodec 62 bundled with the opencv wheel, driven through ctypes) -- an
implementation that shares no code and no author with this repository.  Every spec-mode parity test compares against these planes.
ref_cif_ip_stream: a Main-profile CABAC bitstream and the planes the UNMODIFIED reference decoder reconstructed from it with its own
arithmetic decoder (oracle/_ref/h264ref_cabac).
Fixture layout: stream (uint8), au_sizes, w_mbs, h_mbs, n_pics, md5_<i> (16 bytes of MD5(Y|U|V)), and y<i>/u<i>/v<i> for the small ones.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from common import *  # noqa: E402,F403
import lavc  # noqa: E402
import streams  # noqa: E402


def save(name, seq, stream, aus, frames, keep_planes):
    out = {"w_mbs": seq.w_mbs, "h_mbs": seq.h_mbs, "n_pics": len(frames), "stream": np.frombuffer(stream, dtype=np.uint8),
           "au_sizes": np.array([len(a) for a in aus], dtype=np.int64)}
    for i, (y, u, v) in enumerate(frames):
        md5 = hashlib.md5(y.tobytes() + u.tobytes() + v.tobytes()).hexdigest()
        out["md5_%d" % i] = np.frombuffer(bytes.fromhex(md5), dtype=np.uint8)
        if keep_planes:
            out["y%d" % i], out["u%d" % i], out["v%d" % i] = y, u, v
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "pictures", len(frames), "stream bytes", len(stream), "file KiB", os.path.getsize(path) // 1024)


def main():
    only = sys.argv[1:]          # names given on the command line: regenerate just those
    for name, recipe in streams.SPEC_RECIPES.items():
        if only and name not in only:
            continue
        s = recipe()
        frames = lavc.decode(s["aus"])
        assert len(frames) == len(s["pics"])
        save(name, s["seq"], s["stream"], s["aus"], frames, keep_planes=name != "spec_hd1080_ip")
    if only and "ref_cif_ip_stream" not in only:
        return
    # Main-profile stream for the unmodified reference (own CABAC); content kept inside [0,255]
    s = streams.ref_cif_ip(sanitize=lambda seq, pics: sanitize_with_reference(seq, pics))
    dump, _ = refharness.run_reference_cabac(s["stream"], level=1)
    assert all(d["raw"].min() >= 0 and d["raw"].max() <= 255 for d in dump)
    save("ref_cif_ip_stream", s["seq"], s["stream"], s["aus"], [(d["y"], d["u"], d["v"]) for d in dump], keep_planes=True)


if __name__ == "__main__":
    main()
