"""Host-side packing helpers (what a parser would write into a picture buffer): round trips on CPU."""
import numpy as np

from videodecoder_b200 import abi, pack


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
