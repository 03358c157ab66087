"""CPU: the C-ABI library loads, exports every symbol declared in include/h264r.h, the ctypes /
numpy mirrors have the C layouts, and the engine refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from common import ROOT
from videodecoder_b200 import abi, build, engine, synth


def declared_functions(header):
    # through the preprocessor: comments and development-only blocks (#ifdef H264R_TRACE) are not part of the ABI
    text = subprocess.run(["gcc", "-E", "-P", os.path.join(ROOT, "include", header)], capture_output=True, text=True, check=True).stdout
    return sorted(set(re.findall(r"\b(h264r_[a-z_0-9]+|h264s_[a-z_0-9]+)\s*\(", text)))


def test_cuda_library_exports_header_symbols():
    path = build.build_cuda()
    lib = ctypes.CDLL(path)
    names = declared_functions("h264r.h")
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libh264r.so does not export %s" % n
    assert sorted(engine.EXPORTS) == names, "engine.EXPORTS out of sync with include/h264r.h"
    assert lib.h264r_abi_version() == abi.ABI_VERSION


def test_host_library_exports_header_symbols(built):
    lib = ctypes.CDLL(build.build_host())
    for n in declared_functions("h264_syntax.h"):
        assert hasattr(lib, n), "libh264host.so does not export %s" % n


def test_struct_layouts_match_c():
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "h264r.h"
#include "h264_syntax.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(h264r_mb_hdr), sizeof(h264r_pic_params),
         offsetof(h264r_pic_params, weighted_pred), offsetof(h264r_pic_params, luma_weight),
         offsetof(h264r_pic_params, chroma_offset), offsetof(h264r_pic_params, disable_deblocking_filter_idc),
         sizeof(h264s_mb), sizeof(h264s_pic));
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        got = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    P = abi.PicParams
    want = [abi.MB_HDR_DTYPE.itemsize, ctypes.sizeof(P), P.weighted_pred.offset, P.luma_weight.offset, P.chroma_offset.offset,
            P.disable_deblocking_filter_idc.offset, synth.MB_DTYPE.itemsize, ctypes.sizeof(synth.Pic)]
    assert got == want


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(engine.H264RError) as e:
        engine.Engine(11, 9, max_frames=2)
    assert e.value.code == abi.E_NODEVICE


def test_avail_flags():
    f = abi.avail_flags(3, 2)
    assert list(f) == [0, abi.AVAIL_A, abi.AVAIL_A,
                       abi.AVAIL_B | abi.AVAIL_C, abi.AVAIL_A | abi.AVAIL_B | abi.AVAIL_C | abi.AVAIL_D, abi.AVAIL_A | abi.AVAIL_B | abi.AVAIL_D]


def test_writer_header_is_plain_c_and_cxx(tmp_path):
    """include/h264r_writer.h is header-only C99 that also compiles as C++ (hosts of either language include it)"""
    import subprocess
    inc = os.path.join(ROOT, "include")
    src = '#include "h264r_writer.h"\nint main(void) { h264r_writer w; h264r_picture_buf b; h264r_writer_begin(&w, &b, 0, H264R_LEVELS_COMPACT); return w.next_mb; }\n'
    for name, cc, std in (("t.c", "gcc", "-std=c99"), ("t.cpp", "g++", "-std=c++11")):
        p = tmp_path / name
        p.write_text(src)
        r = subprocess.run([cc, std, "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", inc, str(p), "-o", str(tmp_path / (name + ".out"))],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
