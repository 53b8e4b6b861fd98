"""jm_b200/h264_tables.py against the numbers in the JM sources.  Reads the reference tree, so it only runs where
/root/reference is mounted (the authoring container); skipped elsewhere."""
import os
import re

import numpy as np
import pytest

from jm_b200 import h264_tables as T

JM = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(JM, "lencod", "src")), reason="reference tree not mounted")


def _array(path, name):
    src = open(os.path.join(JM, path), errors="ignore").read()
    m = re.search(r"\b" + re.escape(name) + r"\s*(\[[^\]]*\])+\s*=\s*\{(.*?)\};", src, re.S)
    assert m, f"{name} not found in {path}"
    body = re.sub(r"//[^\n]*", "", m.group(2))
    return np.array([int(x) for x in re.findall(r"-?\d+", body)])


def test_quant_tables():
    assert np.array_equal(_array("lencod/src/q_matrix.c", "quant_coef").reshape(6, 4, 4), T.QUANT_COEF4)
    assert np.array_equal(_array("lencod/src/q_matrix.c", "dequant_coef").reshape(6, 4, 4), T.DEQUANT_COEF4)
    assert np.array_equal(_array("lencod/src/q_matrix.c", "quant_coef8").reshape(6, 8, 8), T.QUANT_COEF8)
    assert np.array_equal(_array("lencod/src/q_matrix.c", "dequant_coef8").reshape(6, 8, 8), T.DEQUANT_COEF8)


def test_scans_and_costs():
    assert np.array_equal(_array("lencod/src/block.c", "SNGL_SCAN").reshape(16, 2), T.SNGL_SCAN)
    assert np.array_equal(_array("lencod/src/block.c", "COEFF_COST4x4").reshape(3, 16), T.COEFF_COST4x4)
    assert np.array_equal(_array("lencod/src/transform8x8.c", "SNGL_SCAN8x8").reshape(64, 2), T.SNGL_SCAN8x8)
    assert np.array_equal(_array("lencod/src/transform8x8.c", "SNGL_SCAN8x8_CAVLC").reshape(64, 2), T.SNGL_SCAN8x8_CAVLC)
    assert np.array_equal(_array("lencod/src/transform8x8.c", "COEFF_COST8x8").reshape(2, 64), T.COEFF_COST8x8[:2])   # JM holds 2 rows
