"""ctypes / numpy mirror of include/h264r.h (struct layouts and constants only)."""
import ctypes

import numpy as np

ABI_VERSION = 7
HDR_STREAM = 9            # h264r_picture_buf.hdr_bytes: the stream form (include/h264r.h)
REF_COMPAT = 0x1
DIGESTS = 0x2      # digest of every picture as part of its reconstruction

MB_I4x4, MB_I8x8, MB_I16x16 = 0, 1, 2
MB_P16x16, MB_P16x8, MB_P8x16, MB_P8x8 = 4, 5, 6, 7
AVAIL_A, AVAIL_B, AVAIL_C, AVAIL_D, T8x8, SKIP = 0x01, 0x02, 0x04, 0x08, 0x10, 0x20
SUB_8x8, SUB_8x4, SUB_4x8, SUB_4x4 = 0, 1, 2, 3
SLICE_P, SLICE_I = 0, 2
COEFF_PER_MB, MV_PER_MB, MAX_REFS = 384, 32, 16
LEVELS_INT16, LEVELS_INT8, LEVELS_COMPACT = 2, 1, 3      # h264r_picture_buf.level_bytes
BLK_COMPACT = 1 << 24

E_INVAL, E_CUDA, E_NOMEM, E_STATE, E_NODEVICE, E_CAPACITY = -1, -2, -3, -4, -5, -6
STATUS_BAD_CLASS, STATUS_INTER_IN_I, STATUS_N_INTRA, STATUS_STREAM = 1, 2, 4, 8      # h264r_device_status

MB_HDR_DTYPE = np.dtype([
    ("mb_class", "u1"), ("flags", "u1"), ("qp_y", "u1"), ("qp_cb", "u1"), ("qp_cr", "u1"), ("cbp", "u1"),
    ("intra_modes", "u1"), ("sub_mb_type", "u1"), ("u", "u1", (8,)),
])
assert MB_HDR_DTYPE.itemsize == 16


class PicParams(ctypes.Structure):
    _fields_ = [
        ("slice_type", ctypes.c_int32),
        ("num_ref_idx_l0", ctypes.c_int32),
        ("ref_list0", ctypes.c_int32 * MAX_REFS),
        ("weighted_pred", ctypes.c_int32),
        ("luma_log2_weight_denom", ctypes.c_int32),
        ("chroma_log2_weight_denom", ctypes.c_int32),
        ("luma_weight", ctypes.c_int16 * MAX_REFS),
        ("luma_offset", ctypes.c_int16 * MAX_REFS),
        ("chroma_weight", (ctypes.c_int16 * 2) * MAX_REFS),
        ("chroma_offset", (ctypes.c_int16 * 2) * MAX_REFS),
        ("disable_deblocking_filter_idc", ctypes.c_int32),
        ("slice_alpha_c0_offset_div2", ctypes.c_int32),
        ("slice_beta_offset_div2", ctypes.c_int32),
    ]


class FrameLayout(ctypes.Structure):
    _fields_ = [("frame_bytes", ctypes.c_size_t), ("y_offset", ctypes.c_size_t), ("u_offset", ctypes.c_size_t), ("v_offset", ctypes.c_size_t),
                ("y_pitch", ctypes.c_int32), ("c_pitch", ctypes.c_int32), ("width", ctypes.c_int32), ("height", ctypes.c_int32)]


def default_pic_params(slice_type=SLICE_I, ref_list=(), deblock=(1, 0, 0)):
    pp = PicParams()
    pp.slice_type = slice_type
    pp.num_ref_idx_l0 = len(ref_list)
    for i in range(MAX_REFS):
        pp.ref_list0[i] = ref_list[i] if i < len(ref_list) else -1
        pp.luma_weight[i] = 1
        pp.chroma_weight[i][0] = pp.chroma_weight[i][1] = 1
    pp.disable_deblocking_filter_idc, pp.slice_alpha_c0_offset_div2, pp.slice_beta_offset_div2 = deblock
    return pp


def avail_flags(w_mbs, h_mbs):
    """AVAIL_* bits of every macroblock of a one-slice picture, raster order
    (macroblock::is_avaiable, reference h264/macroblock.cpp:1446-1452)."""
    mbx = np.tile(np.arange(w_mbs), h_mbs)
    mby = np.repeat(np.arange(h_mbs), w_mbs)
    f = np.zeros(w_mbs * h_mbs, dtype=np.uint8)
    f |= np.where(mbx > 0, AVAIL_A, 0).astype(np.uint8)
    f |= np.where(mby > 0, AVAIL_B, 0).astype(np.uint8)
    f |= np.where((mby > 0) & (mbx < w_mbs - 1), AVAIL_C, 0).astype(np.uint8)
    f |= np.where((mby > 0) & (mbx > 0), AVAIL_D, 0).astype(np.uint8)
    return f


# Table 8-15: QPc as a function of qPI (spec mode; the reference uses QPY'-2 instead)
_QPC = list(range(30)) + [29, 30, 31, 32, 32, 33, 34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39]


def chroma_qp(qp_y, offset=0, compat=False):
    qp_y = np.asarray(qp_y, dtype=np.int32)
    if compat:
        return np.where(qp_y > 2, qp_y - 2, 0).astype(np.uint8)
    return np.asarray(_QPC, dtype=np.uint8)[np.clip(qp_y + offset, 0, 51)]
