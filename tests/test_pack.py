"""Host-side packing helpers (what a parser would write into a picture buffer): round trips on CPU."""
import numpy as np

from videodecoder_b200 import abi, pack
import emul


def unpack_compact(mask, units):
    """plain restatement of H264R_LEVELS_COMPACT (include/h264r.h): units -> [n_blocks, 16] int16"""
    out = []
    u = 0
    for m in mask:
        n = bin(int(m) & 0xFFFFFF).count("1")
        for _ in range(n):
            blk = np.zeros(16, dtype=np.int16)
            if int(m) & abi.BLK_COMPACT:
                sig = int(units[u, 0]) | (int(units[u, 1]) << 8)
                k = 0
                for pos in range(16):
                    if (sig >> pos) & 1:
                        blk[pos] = np.int8(units[u, 2 + k].astype(np.int8))
                        k += 1
                assert k <= 6
                u += 1
            else:
                blk[:8] = units[u].view(np.int8)
                blk[8:] = units[u + 1].view(np.int8)
                u += 2
            out.append(blk)
    assert u == units.shape[0]
    return np.array(out, dtype=np.int16).reshape(-1, 16)


def test_pack_compact_round_trip():
    rng = np.random.default_rng(5)
    n_mbs = 300
    coeff = np.zeros((n_mbs, 24, 16), dtype=np.int16)
    for a in range(n_mbs):
        for b in rng.choice(24, size=rng.integers(0, 9), replace=False):
            nnz = int(rng.choice([1, 1, 2, 3, 6, 7, 16], p=[0.3, 0.2, 0.2, 0.15, 0.1, 0.03, 0.02]))
            pos = rng.choice(16, size=nnz, replace=False)
            coeff[a, b, pos] = rng.choice([-128, -3, -1, 1, 2, 127], size=nnz)
    mask, blocks = pack.pack_blocks(coeff.reshape(n_mbs, 384))
    cmask, units = pack.pack_compact(mask, blocks)
    assert np.array_equal(cmask & 0xFFFFFF, mask)
    compact = (cmask & abi.BLK_COMPACT) != 0
    assert compact.any() and (~compact & (mask != 0)).any()       # both kinds of macroblock occur
    assert not (compact & (mask == 0)).any()
    assert np.array_equal(unpack_compact(cmask, units), blocks)
    # fewer bytes than int8 blocks whenever a macroblock is compact
    assert units.shape[0] * 8 < blocks.shape[0] * 16


def test_pack_compact_empty_picture():
    mask = np.zeros(10, dtype=np.uint32)
    cmask, units = pack.pack_compact(mask, np.zeros((0, 16), dtype=np.int16))
    assert units.shape == (0, 8) and not cmask.any()


_WRITER_PROG = r'''
#include <stdio.h>
#include <stdlib.h>
#include "h264r_writer.h"
/* reads n, then hdr[n] (16 B), mv[n][32] int16, coeff[n][384] int16 from stdin-named file; runs the writer with 8-byte headers on a
   plain host buffer; writes n_units, n_mvs, n_intra, hdr8[n], mask[n], units, mvs */
int main(int argc, char** argv)
{
    FILE* f = fopen(argv[1], "rb"); FILE* o = fopen(argv[2], "wb");
    int n; if (fread(&n, 4, 1, f) != 1) return 1;
    h264r_mb_hdr* hdr = malloc(sizeof(h264r_mb_hdr) * n); int16_t* mv = malloc(64 * n); int16_t* coeff = malloc(768 * n);
    if (fread(hdr, 16, n, f) != (size_t)n || fread(mv, 64, n, f) != (size_t)n || fread(coeff, 768, n, f) != (size_t)n) return 1;
    h264r_picture_buf b; memset(&b, 0, sizeof b);
    b.hdr_bytes = 8; b.hdr8 = malloc(8 * n); b.blk_mask = malloc(4 * n); b.blocks = malloc(768 * n + 8 * n); b.max_blocks = 24 * (size_t)n + n;
    int16_t* mvbuf = malloc(64 * n); b.mv = mvbuf; b.mv_end = mvbuf + 32 * n;
    h264r_writer w; h264r_writer_begin(&w, &b, n, H264R_LEVELS_COMPACT);
    for (int a = 0; a < n; a++) { int rc = h264r_writer_mb(&w, &hdr[a], hdr[a].mb_class >= H264R_MB_P16x16 ? mv + 32 * a : NULL, coeff + 384 * a); if (rc) return 10 - rc; }
    size_t nu; long nm; if (h264r_writer_end(&w, &nu, &nm)) return 2;
    int32_t head[3] = {(int32_t)nu, (int32_t)nm, b.n_intra};
    fwrite(head, 4, 3, o); fwrite(b.hdr8, 8, n, o); fwrite(b.blk_mask, 4, n, o); fwrite(b.blocks, 8, nu, o); fwrite(b.mv_end - 2 * nm, 4, nm, o);
    fclose(o); return 0;
}
'''


def unpack_hdr8(h8):
    """plain restatement of h264r_mb_hdr8 (include/h264r.h) -> 16-byte records (u.pred4x4 of intra macroblocks left zero)"""
    out = np.zeros(h8.shape[0], dtype=abi.MB_HDR_DTYPE)
    w0, w1 = h8[:, 0].astype(np.uint32), h8[:, 1].astype(np.uint32)
    out["mb_class"] = w0 & 7
    out["flags"] = (w0 >> 3) & 63
    out["intra_modes"] = (w0 >> 9) & 15
    out["qp_y"] = (w0 >> 13) & 63
    out["qp_cb"] = (w0 >> 19) & 63
    out["qp_cr"] = (w0 >> 25) & 63
    out["cbp"] = w1 & 63
    out["sub_mb_type"] = (w1 >> 6) & 255
    u = out["u"].reshape(-1, 8)
    inter = out["mb_class"] > abi.MB_I16x16
    for q in range(4):
        u[inter, q] = ((w1 >> (14 + 4 * q)) & 15)[inter]
    return out


def test_short_headers_c_writer_and_python_twin(tmp_path):
    """8-byte headers: include/h264r_writer.h and pack.py (pack_hdr8, insert_intra_units) write the same bytes; the headers expand
    back to the 16-byte records and the intra macroblocks' prediction modes sit in front of their blocks"""
    import os
    import subprocess
    from videodecoder_b200 import workload
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    (tmp_path / "w.c").write_text(_WRITER_PROG)
    r = subprocess.run(["gcc", "-std=c99", "-O1", "-Wall", "-Werror", "-I", os.path.join(root, "include"), str(tmp_path / "w.c"), "-o", str(tmp_path / "w")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rng = np.random.default_rng(77)
    wm, hm = 9, 7
    cfg = workload.WorkloadConfig(compat=False, num_ref=3, p_intra_in_p=0.2, p_i8x8=0.3, p_t8x8=0.3, p_sub=(0.25, 0.25, 0.25, 0.25), chroma_qp_offset=-3)
    for p in workload.make_gop(rng, wm, hm, 3, cfg):
        n = wm * hm
        hdr = np.ascontiguousarray(p["hdr"]).view(abi.MB_HDR_DTYPE).reshape(-1)
        mv = p["mv"] if p["mv"] is not None else np.zeros((n, 32), dtype=np.int16)
        with open(tmp_path / "in.bin", "wb") as f:
            f.write(np.int32(n).tobytes()); f.write(hdr.tobytes()); f.write(np.ascontiguousarray(mv, dtype=np.int16).tobytes())
            f.write(np.ascontiguousarray(p["coeff"], dtype=np.int16).tobytes())
        r = subprocess.run([str(tmp_path / "w"), str(tmp_path / "in.bin"), str(tmp_path / "out.bin")])
        assert r.returncode == 0
        raw = open(tmp_path / "out.bin", "rb").read()
        nu, nm, ni = np.frombuffer(raw, dtype=np.int32, count=3)
        off = 12
        h8 = np.frombuffer(raw, dtype=np.uint32, count=2 * n, offset=off).reshape(n, 2); off += 8 * n
        mask_c = np.frombuffer(raw, dtype=np.uint32, count=n, offset=off); off += 4 * n
        units_c = np.frombuffer(raw, dtype=np.uint8, count=8 * nu, offset=off).reshape(-1, 8); off += 8 * nu
        mvs_c = np.frombuffer(raw, dtype=np.int16, count=2 * nm, offset=off).reshape(-1, 2)
        mask, blocks = pack.pack_blocks(p["coeff"])
        cmask, units = pack.pack_compact(mask, blocks)
        units = pack.insert_intra_units(hdr, cmask, units)
        assert np.array_equal(pack.pack_hdr8(hdr), h8)
        assert np.array_equal(cmask, mask_c) and np.array_equal(units, units_c)
        assert ni == int((hdr["mb_class"] <= abi.MB_I16x16).sum()) and nu == units.shape[0]
        if p["mv"] is not None:
            assert np.array_equal(pack.pack_mvs_partition(hdr, p["mv"])[::-1], mvs_c)
        # expansion: everything but the intra prediction modes comes back from the 8 bytes
        back = unpack_hdr8(h8)
        intra = hdr["mb_class"] <= abi.MB_I16x16
        for k in ("mb_class", "flags", "intra_modes", "qp_y", "qp_cb", "qp_cr", "cbp"):
            assert np.array_equal(back[k], hdr[k]), k
        assert np.array_equal(back["sub_mb_type"][~intra], hdr["sub_mb_type"][~intra])
        assert np.array_equal(back["u"].reshape(-1, 8)[~intra][:, :4], hdr["u"].reshape(-1, 8)[~intra][:, :4])
        # the first unit of every intra macroblock = its u.pred4x4 bytes
        per_mb = np.array([bin(int(m) & 0xFFFFFF).count("1") * (1 if int(m) & abi.BLK_COMPACT else 2) for m in cmask]) + intra
        first = np.concatenate(([0], np.cumsum(per_mb)))[:-1]
        assert np.array_equal(units_c[first[intra]], hdr["u"].reshape(-1, 8)[intra].view(np.uint8))


def unpack_stream(hdr8, mb_words, stream):
    """plain restatement of the stream form (H264R_HDR_STREAM, include/h264r.h) -- what k_blk_scan's expansion produces on the device:
    16-byte headers, dense block masks, 8-byte units, the vectors in partition order"""
    n = hdr8.shape[0]
    hdr = unpack_hdr8(hdr8)
    mask = np.zeros(n, dtype=np.uint32)
    units, mvs = [], []
    wo = 0
    for a in range(n):
        w0, w1 = int(hdr8[a, 0]), int(hdr8[a, 1])
        cls = w0 & 7
        intra = cls <= abi.MB_I16x16
        inline = (not intra) and bool(w0 >> 31)
        p = wo
        if (w1 & 63) or cls == abi.MB_I16x16:
            mask[a] = stream[p]; p += 1
        if intra:
            hdr["u"].reshape(-1, 8)[a] = stream[p:p + 2].view(np.uint8); p += 2
        # behind the words the macroblock is a byte run: short vectors, compact blocks; zero-padded to a word in front of int8 blocks
        # and at the end
        run = stream[p:].view(np.uint8)
        q = 0
        if intra:
            pass
        elif inline:
            x = ((w1 >> 6) & 255) | (((w0 >> 9) & 15) << 8) | (((w1 >> 30) & 1) << 12)
            y = ((w1 >> 18) & 0xFFF) | ((w1 >> 31) << 12)
            sx = lambda v: v - 8192 if v & 4096 else v
            mvs.append((sx(x), sx(y)))
            hdr["intra_modes"][a] = 0; hdr["sub_mb_type"][a] = 0
            hdr["u"].reshape(-1, 8)[a, 1:4] = hdr["u"].reshape(-1, 8)[a, 0]       # the reference index of a 16x16 macroblock, in all four quadrant fields
        else:
            sub = (w1 >> 6) & 255
            cnt = {abi.MB_P16x16: 1, abi.MB_P16x8: 2, abi.MB_P8x16: 2}.get(cls, sum((1, 2, 2, 4)[(sub >> (2 * k)) & 3] for k in range(4)))
            if (w1 >> 30) & 1:                                  # H264R_HDR8_MV_SHORT: int8 pairs
                for i in range(cnt):
                    mvs.append((int(run[q:q + 1].view(np.int8)[0]), int(run[q + 1:q + 2].view(np.int8)[0]))); q += 2
            else:
                for i in range(cnt):
                    mvs.append(tuple(int(v) for v in run[q:q + 4].view(np.int16))); q += 4
        m = int(mask[a])
        if m & abi.BLK_COMPACT:
            for b in range(24):
                if not (m >> b) & 1:
                    continue
                n = bin(int(run[q]) | (int(run[q + 1]) << 8)).count("1")
                assert 1 <= n <= 6
                unit = np.zeros(8, dtype=np.uint8)
                unit[:2 + n] = run[q:q + 2 + n]
                units.append(unit); q += 2 + n
        elif m & 0xFFFFFF:
            assert not run[q:(q + 3) // 4 * 4].any(), "padding in front of the int8 blocks of macroblock %d is not zero" % a
            q = (q + 3) // 4 * 4
            for b in range(24):
                if (m >> b) & 1:
                    units.append(run[q:q + 8].copy()); units.append(run[q + 8:q + 16].copy()); q += 16
        assert not run[q:(q + 3) // 4 * 4].any(), "padding of macroblock %d is not zero" % a
        p += (q + 3) // 4
        assert p - wo == int(mb_words[a]), "macroblock %d: %d words walked, %d announced" % (a, p - wo, mb_words[a])
        wo = p
    return hdr, mask, (np.array(units, dtype=np.uint8).reshape(-1, 8)), np.array(mvs, dtype=np.int16).reshape(-1, 2), wo


def test_stream_form_c_writer_against_restatement(built):
    """the stream form written by include/h264r_writer.h (through libh264host's h264w_pack_arrays) expands -- by the plain restatement
    above of what k_blk_scan does -- to exactly what the other picture-buffer forms carry: headers, masks, compact units, vectors in
    partition order; P_L0_16x16 vectors ride in the header when they fit 13 bits and in the stream when they do not"""
    import ctypes
    from videodecoder_b200 import engine, hoststream, workload
    rng = np.random.default_rng(99)
    wm, hm = 10, 6
    n = wm * hm
    cfg = workload.WorkloadConfig(compat=False, num_ref=3, p_intra_in_p=0.15, p_i8x8=0.3, p_t8x8=0.3, p_sub=(0.25, 0.25, 0.25, 0.25), mv_range=200)
    total_saved = 0
    for pi, p in enumerate(workload.make_gop(rng, wm, hm, 4, cfg)):
        hdr = np.ascontiguousarray(p["hdr"]).view(abi.MB_HDR_DTYPE).reshape(-1).copy()
        coeff = np.ascontiguousarray(p["coeff"], dtype=np.int16).copy()
        mv = np.ascontiguousarray(p["mv"], dtype=np.int16).copy() if p["mv"] is not None else None
        if mv is not None:      # a few vectors beyond the inline range
            big = np.flatnonzero(hdr["mb_class"] == abi.MB_P16x16)[:3]
            mv[big] = np.tile(np.array([5000, -4500], dtype=np.int16), 16)
            small = np.flatnonzero(hdr["mb_class"] > abi.MB_P16x16)[::2]       # ... and every other partitioned macroblock within int8
            mv[small] = np.clip(mv[small], -128, 127)
        c3 = coeff.reshape(n, 24, 16)
        if pi == 2:             # a macroblock with a dense block: plain int8 storage, four words per block
            a = int(np.flatnonzero(hdr["cbp"] & 1)[0])
            c3[a, 0, :9] = np.arange(1, 10)
        hdr8 = np.zeros((n, 2), dtype=np.uint32); words = np.zeros(n, dtype=np.uint8); stream = np.zeros(n * 120, dtype=np.uint32)
        buf = engine.PictureBufStruct()
        buf.hdr8 = hdr8.ctypes.data; buf.mb_words = words.ctypes.data; buf.stream = stream.ctypes.data
        buf.max_stream_words = stream.size; buf.hdr_bytes = abi.HDR_STREAM
        nb, nm = ctypes.c_size_t(0), ctypes.c_long(0)
        rc = hoststream.lib().h264w_pack_arrays(ctypes.byref(buf), n, hdr.ctypes.data_as(ctypes.c_void_p), mv.ctypes.data_as(ctypes.c_void_p) if mv is not None else None,
                                                coeff.ctypes.data_as(ctypes.c_void_p), abi.LEVELS_COMPACT, ctypes.byref(nb), ctypes.byref(nm))
        assert rc == 0
        got_hdr, got_mask, got_units, got_mvs, used = unpack_stream(hdr8, words, stream)
        assert used == nb.value == int(words.sum())
        mask, blocks = pack.pack_blocks(coeff)
        cmask, units = pack.pack_compact(mask, blocks)
        assert np.array_equal(got_mask, cmask) and np.array_equal(got_units, units)
        # the device's own expansion functions (h264r_core.cuh: stream_mb_counts / stream_expand_mb, run by the host emulator in
        # macroblock order) give the same headers, masks, units, vectors -- and the scans and the intra list the kernels use
        dev = emul.stream_expand(hdr8, words, stream, units.shape[0], 16 * n)
        assert dev["words"] == used
        assert np.array_equal(dev["hdr"].view(np.uint8), got_hdr.view(np.uint8))
        assert np.array_equal(dev["mask"], cmask) and np.array_equal(dev["units"], units) and np.array_equal(dev["mvs"], got_mvs)
        per_mb = np.array([bin(int(m) & 0xFFFFFF).count("1") * (1 if int(m) & abi.BLK_COMPACT else 2) for m in cmask])
        assert np.array_equal(dev["blk_off"], np.concatenate([[0], np.cumsum(per_mb)[:-1]]))
        assert np.array_equal(dev["intra_list"], np.flatnonzero(hdr["mb_class"] <= abi.MB_I16x16))
        # a word count that does not match what the macroblock holds is noticed at that macroblock (H264R_STATUS_STREAM on the device)
        bad = words.copy(); bad[7] += 1
        assert emul.stream_expand(hdr8, bad, stream, units.shape[0] + 48, 16 * n)["words"] == -1 - 7
        intra = hdr["mb_class"] <= abi.MB_I16x16
        for k in ("mb_class", "flags", "qp_y", "qp_cb", "qp_cr", "cbp"):
            assert np.array_equal(got_hdr[k], hdr[k]), k
        assert np.array_equal(got_hdr["intra_modes"][intra], hdr["intra_modes"][intra])
        assert np.array_equal(got_hdr["u"].reshape(-1, 8)[intra], hdr["u"].reshape(-1, 8)[intra])
        p8 = hdr["mb_class"] == abi.MB_P8x8
        assert np.array_equal(got_hdr["sub_mb_type"][p8], hdr["sub_mb_type"][p8])
        inter = ~intra
        assert np.array_equal(got_hdr["u"].reshape(-1, 8)[inter][:, 0], hdr["u"].reshape(-1, 8)[inter][:, 0])
        assert np.array_equal(got_hdr["u"].reshape(-1, 8)[inter][:, :4], hdr["u"].reshape(-1, 8)[inter][:, :4])
        if mv is not None:
            want = pack.pack_mvs_partition(hdr, mv)
            assert np.array_equal(got_mvs, want) and nm.value == want.shape[0]
            inl = (hdr8[:, 0] >> 31).astype(bool)
            assert inl.sum() == (hdr["mb_class"] == abi.MB_P16x16).sum() - 3
            short = ((hdr8[:, 1] >> 30) & 1).astype(bool) & ~inl
            multi = (hdr["mb_class"] > abi.MB_P16x16) & ~inl
            assert short.any() and (multi & ~short).any(), "both vector widths must occur: %d short, %d long of %d" % (short.sum(), (multi & ~short).sum(), multi.sum())
        short_bytes = 8 * n + 4 * n + units.shape[0] * 8 + (0 if mv is None else 4 * pack.pack_mvs_partition(hdr, mv).shape[0]) + 8 * int(intra.sum())
        total_saved += short_bytes - (9 * n + 4 * used)
    assert total_saved > 0


def test_stream_expansion_of_garbage_stays_inside_its_areas(built):
    """whatever a damaged stream-form buffer holds, the expansion the kernels run (stream_mb_counts / stream_expand_mb) writes at most
    48 units and 16 vectors per macroblock -- the sizes of the slot's areas -- and reads at most 116 words past a macroblock's first
    word (the device keeps the stream in front of areas of the same slot that are longer than that)"""
    rng = np.random.default_rng(5)
    n = 200
    for trial in range(20):
        hdr8 = rng.integers(0, 2 ** 32, size=(n, 2), dtype=np.uint64).astype(np.uint32)
        if trial % 2:
            hdr8[:, 0] &= ~np.uint32(4)              # intra / inter mix of legal classes now and then
        words = rng.integers(0, 256, size=n, dtype=np.uint8)
        stream = rng.integers(0, 2 ** 32, size=n * 255 + 128, dtype=np.uint64).astype(np.uint32)
        dev = emul.stream_expand(hdr8, words, stream, 48 * n, 16 * n, strict=False, canary=0xC0FFEE)
        assert dev["words"] < 0          # and it is reported


def wire_md5(n, p):
    """MD5 over what travels for picture p in the stream form: 8-byte headers | word counts | stream words (+ the two counts the
    submit call takes)"""
    import ctypes, hashlib
    from videodecoder_b200 import engine, hoststream
    hdr = np.ascontiguousarray(p["hdr"]).view(abi.MB_HDR_DTYPE).reshape(-1).copy()
    coeff = np.ascontiguousarray(p["coeff"], dtype=np.int16).copy()
    mv = np.ascontiguousarray(p["mv"], dtype=np.int16).copy() if p["mv"] is not None else None
    hdr8 = np.zeros((n, 2), dtype=np.uint32); words = np.zeros(n, dtype=np.uint8); stream = np.zeros(n * 120, dtype=np.uint32)
    buf = engine.PictureBufStruct()
    buf.hdr8 = hdr8.ctypes.data; buf.mb_words = words.ctypes.data; buf.stream = stream.ctypes.data
    buf.max_stream_words = stream.size; buf.hdr_bytes = abi.HDR_STREAM
    nb, nm = ctypes.c_size_t(0), ctypes.c_long(0)
    rc = hoststream.lib().h264w_pack_arrays(ctypes.byref(buf), n, hdr.ctypes.data_as(ctypes.c_void_p), mv.ctypes.data_as(ctypes.c_void_p) if mv is not None else None,
                                            coeff.ctypes.data_as(ctypes.c_void_p), abi.LEVELS_COMPACT, ctypes.byref(nb), ctypes.byref(nm))
    if rc:
        return "rc %d" % rc           # a picture with a level beyond int8: not in this form
    h = hashlib.md5()
    h.update(hdr8.tobytes()); h.update(words.tobytes()); h.update(stream[:nb.value].tobytes())
    h.update(np.array([nb.value, nm.value], dtype=np.int64).tobytes())
    return h.hexdigest()


def test_wire_format_known_answer(built):
    """the bytes of the stream form are a contract between a host that fills picture buffers and the device that expands them:
    committed MD5s of the committed fixtures' pictures (tests/golden/make_wire_golden.py), tied to the ABI version"""
    import json, os
    from common import load_golden
    kat = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stream_form_md5.json")))
    assert kat["abi_version"] == abi.ABI_VERSION, "ABI version changed: regenerate tests/golden/stream_form_md5.json"
    for name, want in kat["pictures"].items():
        g = load_golden(name)
        got = [wire_md5(g["w_mbs"] * g["h_mbs"], p) for p in g["pics"]]
        assert got == want, name


def test_stream_form_of_empty_macroblocks(built):
    """edge cases of the stream form: an all-skip P picture is its headers and nothing else (every vector inline, zero stream words),
    an Intra16x16 picture without levels still carries a mask word and its prediction modes per macroblock -- through the C writer
    and the expansion functions the kernels call"""
    import ctypes
    from videodecoder_b200 import engine, hoststream
    n = 12
    for intra in (False, True):
        hdr = np.zeros(n, dtype=abi.MB_HDR_DTYPE)
        hdr["mb_class"] = abi.MB_I16x16 if intra else abi.MB_P16x16
        hdr["qp_y"] = 30; hdr["qp_cb"] = 29; hdr["qp_cr"] = 29
        hdr["intra_modes"] = 2 if intra else 0
        coeff = np.zeros((n, 384), dtype=np.int16)
        mv = None if intra else np.tile(np.array([3, -7], dtype=np.int16), (n, 16))
        hdr8 = np.zeros((n, 2), dtype=np.uint32); words = np.zeros(n, dtype=np.uint8); stream = np.zeros(n * 120, dtype=np.uint32)
        buf = engine.PictureBufStruct()
        buf.hdr8 = hdr8.ctypes.data; buf.mb_words = words.ctypes.data; buf.stream = stream.ctypes.data
        buf.max_stream_words = stream.size; buf.hdr_bytes = abi.HDR_STREAM
        nb, nm = ctypes.c_size_t(0), ctypes.c_long(0)
        rc = hoststream.lib().h264w_pack_arrays(ctypes.byref(buf), n, hdr.ctypes.data_as(ctypes.c_void_p), mv.ctypes.data_as(ctypes.c_void_p) if mv is not None else None,
                                                coeff.ctypes.data_as(ctypes.c_void_p), abi.LEVELS_COMPACT, ctypes.byref(nb), ctypes.byref(nm))
        assert rc == 0
        assert nb.value == (3 * n if intra else 0) and nm.value == (0 if intra else n)
        dev = emul.stream_expand(hdr8, words, stream, 4, 16 * n, canary=0xABCD)
        assert dev["words"] == nb.value and dev["units"].shape[0] == 0 and not dev["mask"].any()
        assert np.array_equal(dev["hdr"]["mb_class"], hdr["mb_class"]) and np.array_equal(dev["hdr"]["qp_y"], hdr["qp_y"])
        if intra:
            assert np.array_equal(dev["intra_list"], np.arange(n)) and np.array_equal(dev["hdr"]["intra_modes"], hdr["intra_modes"])
        else:
            assert dev["intra_list"].size == 0 and np.array_equal(dev["mvs"], np.tile(np.array([3, -7], dtype=np.int16), (n, 1)))
            assert ((hdr8[:, 0] >> 31) == 1).all()
