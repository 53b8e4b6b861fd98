"""H.264 constant tables the host side hands to the kernels (quantiser scales, scans, coefficient
costs).  Values are the H.264/AVC standard's; JM holds the same numbers in
lencod/src/q_matrix.c:20-170 (quant_coef, dequant_coef, quant_coef8, dequant_coef8),
lencod/src/block.c:72-77,170-176 (COEFF_COST4x4, SNGL_SCAN), lencod/src/transform8x8.c (SNGL_SCAN8x8,
SNGL_SCAN8x8_CAVLC, COEFF_COST8x8) and lencod/src/q_offsets.c:60-207 (default rounding offsets, Q11).
tests/test_tables_vs_reference.py re-reads those files (when /root/reference is mounted) and checks
every number below.
"""
import numpy as np

# per qp%6: (a: positions with both coords even, b: both odd, c: the rest)
_Q4 = [(13107, 5243, 8066), (11916, 4660, 7490), (10082, 4194, 6554),
       (9362, 3647, 5825), (8192, 3355, 5243), (7282, 2893, 4559)]
_D4 = [(10, 16, 13), (11, 18, 14), (13, 20, 16), (14, 23, 18), (16, 25, 20), (18, 29, 23)]
# per qp%6, 8x8 classes 0..5
_Q8 = [(13107, 11428, 20972, 12222, 16777, 15481), (11916, 10826, 19174, 11058, 14980, 14290),
       (10082, 8943, 15978, 9675, 12710, 11985), (9362, 8228, 14913, 8931, 11984, 11259),
       (8192, 7346, 13159, 7740, 10486, 9777), (7282, 6428, 11570, 6830, 9118, 8640)]
_D8 = [(20, 18, 32, 19, 25, 24), (22, 19, 35, 21, 28, 26), (26, 23, 42, 24, 33, 31),
       (28, 25, 45, 26, 35, 33), (32, 28, 51, 30, 40, 38), (36, 32, 58, 34, 46, 43)]


def _class4(j, i):
    if j % 2 == 0 and i % 2 == 0:
        return 0
    if j % 2 == 1 and i % 2 == 1:
        return 1
    return 2


def _class8(j, i):
    j, i = j % 4, i % 4
    if j == 0 and i == 0:
        return 0
    if j % 2 == 1 and i % 2 == 1:
        return 1
    if j == 2 and i == 2:
        return 2
    if (j == 0 and i % 2 == 1) or (i == 0 and j % 2 == 1):
        return 3
    if (j == 0 and i == 2) or (j == 2 and i == 0):
        return 4
    return 5


QUANT_COEF4 = np.array([[[_Q4[k][_class4(j, i)] for i in range(4)] for j in range(4)] for k in range(6)], np.int32)
DEQUANT_COEF4 = np.array([[[_D4[k][_class4(j, i)] for i in range(4)] for j in range(4)] for k in range(6)], np.int32)
QUANT_COEF8 = np.array([[[_Q8[k][_class8(j, i)] for i in range(8)] for j in range(8)] for k in range(6)], np.int32)
DEQUANT_COEF8 = np.array([[[_D8[k][_class8(j, i)] for i in range(8)] for j in range(8)] for k in range(6)], np.int32)

# zig-zag scans as {i (horizontal), j (vertical)} pairs
SNGL_SCAN = np.array([(0, 0), (1, 0), (0, 1), (0, 2), (1, 1), (2, 0), (3, 0), (2, 1),
                      (1, 2), (0, 3), (1, 3), (2, 2), (3, 1), (3, 2), (2, 3), (3, 3)], np.uint8)


def _zigzag8():
    out, i, j, up = [], 0, 0, True
    for _ in range(64):
        out.append((i, j))
        if up:
            if i == 7:
                j += 1; up = False
            elif j == 0:
                i += 1; up = False
            else:
                i += 1; j -= 1
        else:
            if j == 7:
                i += 1; up = True
            elif i == 0:
                j += 1; up = True
            else:
                i -= 1; j += 1
    return np.array(out, np.uint8)


SNGL_SCAN8x8 = _zigzag8()
# CAVLC 8x8: four interleaved 4x4-sized lists; list k takes zig-zag entries 4n+k
SNGL_SCAN8x8_CAVLC = np.array([SNGL_SCAN8x8[4 * n + k] for k in range(4) for n in range(16)], np.uint8)

COEFF_COST4x4 = np.array([[3, 2, 2, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
                          [9] * 16,
                          [3, 2, 2, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]], np.uint8)
COEFF_COST8x8 = np.array([[3] * 4 + [2] * 8 + [1] * 12 + [0] * 40,
                          [9] * 64,
                          [3] * 4 + [2] * 8 + [1] * 12 + [0] * 40], np.uint8)

OFFSET_BITS = 11
OFFSET_INTRA, OFFSET_INTER = 682, 342   # default Q11 rounding offsets (1/3, 1/6)


def q_params(qp, intra, n=4, offset=None):
    """LevelQuantParams table as JM builds it with no scaling matrix: [n][n][3] =
    {OffsetComp, ScaleComp, InvScaleComp} (lcommon/inc/quant_params.h:17-21; q_matrix.c:566-577,
    q_offsets.c:238-262)."""
    per, rem = qp // 6, qp % 6
    off = (OFFSET_INTRA if intra else OFFSET_INTER) if offset is None else offset
    out = np.zeros((n, n, 3), np.int32)
    if n == 4:
        out[..., 0] = off << (15 + per - OFFSET_BITS)
        out[..., 1] = QUANT_COEF4[rem]
        out[..., 2] = DEQUANT_COEF4[rem] << 4
    else:
        out[..., 0] = off << (16 + per - OFFSET_BITS)
        out[..., 1] = QUANT_COEF8[rem]
        out[..., 2] = DEQUANT_COEF8[rem] << 4
    return out


def lambda_me(qp):
    """Integer ME lambda JM derives for a P slice without RDO scaling tricks:
    lambda_md = 0.85 * 2^((qp-12)/3), lambda_mf = (int)(32*sqrt(lambda_md)+0.5)
    (lencod/src/lambda.c:20-32, LAMBDA_FACTOR lencod/inc/defines.h:130-131)."""
    import math
    return int(32.0 * math.sqrt(0.85 * 2.0 ** ((qp - 12) / 3.0)) + 0.5)
