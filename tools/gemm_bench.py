#!/usr/bin/env python
"""Times the tcgen05 GEMM building block on the shapes of the step (GPU box only)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vap_realtime_b200.engine import selftest_gemm

for case in (8, 9, 5, 12, 13):
    for tsel in (1, 2):
        try:
            err, rep = selftest_gemm(case + 16 * tsel)
            print(rep)
        except Exception as e:
            print("case", case, "tile", tsel, "->", e)
