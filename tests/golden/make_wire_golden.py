"""Generates tests/golden/stream_form_md5.json: MD5 of the stream form (H264R_HDR_STREAM) of every picture of the committed
REF_COMPAT fixtures, as include/h264r_writer.h writes it (through libh264host's h264w_pack_arrays) -- a known answer for the WIRE
format itself.  A change of the writer that alters a byte on the wire fails tests/test_pack.py::test_wire_format_known_answer until
the ABI version is raised and this script is run again:

    python tests/golden/make_wire_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from common import load_golden  # noqa: E402
from test_pack import wire_md5  # noqa: E402
from videodecoder_b200 import abi  # noqa: E402

out = {"abi_version": abi.ABI_VERSION, "pictures": {}}
for name in ("qcif_i", "cif_ip", "qcif_wp"):
    g = load_golden(name)
    out["pictures"][name] = [wire_md5(g["w_mbs"] * g["h_mbs"], p) for p in g["pics"]]
json.dump(out, open(os.path.join(HERE, "stream_form_md5.json"), "w"), indent=1)
print({k: len(v) for k, v in out["pictures"].items()})
