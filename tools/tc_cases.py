#!/usr/bin/env python
"""Ad-hoc shape sweep of the tensor-core PHMLinear against the fp64 oracle (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from tc_check import case
for spec in sys.argv[1:]:
    n, fin, fout, M = (int(v) for v in spec.split(","))
    case(n, fin, fout, M, 1, True)
