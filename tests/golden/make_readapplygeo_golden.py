"""Writes tests/golden/readapplygeo_test2.npz from the reference's own fixture images
(src/xmipp/resources/test/image/test2.spi, test2_wrap_false.spi, test2_wrap_true.spi: input and expected outputs of
ImageTest.readApplyGeo, applications/tests/function_tests/test_image_main.cpp:80-98 — anglePsi = 45, BSPLINE3, wrap off / on).
Needs /root/reference; run from the repository root."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from xmipp3_b200 import io          # noqa: E402

R = "/root/reference/src/xmipp/resources/test/image/"
arr = {k: io.read_spider(R + f).astype(np.float32).reshape(128, 128)
       for k, f in (("input", "test2.spi"), ("wrap_false", "test2_wrap_false.spi"), ("wrap_true", "test2_wrap_true.spi"))}
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "readapplygeo_test2.npz"), **arr)
print("written", {k: v.shape for k, v in arr.items()})
