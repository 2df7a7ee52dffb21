#!/usr/bin/env python
"""Records byte strings produced by the reference's own wire codec
(rvap/common/util.py, imported unmodified in the build container) as a small
fixture for tests/test_wire.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
import rvap.common.util as ref  # noqa: E402

rng = np.random.default_rng(5)
x1 = rng.standard_normal(160) * 0.1
x2 = rng.standard_normal(160) * 0.1
pkt = ref.conv_2floatarray_2_bytearray(list(x1), list(x2))
res = {"t": 1718000000.123456, "x1": list(rng.standard_normal(800) * 0.1), "x2": list(rng.standard_normal(800) * 0.1),
       "p_now": [0.25, 0.75], "p_future": [0.4, 0.6], "vad": [0.9, 0.1]}
res_b = ref.conv_vapresult_2_bytearray(res)
bc = {"t": 12.5, "x1": res["x1"], "x2": res["x2"], "p_bc_react": [0.125], "p_bc_emo": [0.5]}
bc_b = ref.conv_vapresult_2_bytearray_bc(bc)
assert ref.conv_bytearray_2_vapresult(res_b)["p_now"] == res["p_now"]
np.savez_compressed(os.path.join(ROOT, "tests/golden/ref_wire.npz"), x1=x1, x2=x2, pkt=np.frombuffer(pkt, np.uint8),
                    res_t=res["t"], res_x1=res["x1"], res_x2=res["x2"], res_bytes=np.frombuffer(res_b, np.uint8),
                    bc_bytes=np.frombuffer(bc_b, np.uint8))
print(len(pkt), len(res_b), len(bc_b))
