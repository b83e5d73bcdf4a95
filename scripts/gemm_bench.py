#!/usr/bin/env python
"""Times the tcgen05 GEMM variants on the Nn layer shapes (device-resident operands)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rasr_b200 import capi

L = capi.lib()
shapes = [(18944, 2048, 2048), (18944, 2048, 448), (18944, 12000, 2048), (4096, 2048, 2048)]
for M, N, K in shapes:
    for variant in (0,):
        ms = C.c_float()
        capi.check(L.rb_test_gemm_bench(M, N, K, variant, 10, C.byref(ms), 0))
        tf = 2.0 * M * N * K / (ms.value * 1e-3) / 1e12
        print(json.dumps(dict(M=M, N=N, K=K, variant=variant, ms=ms.value, tflops=tf)), flush=True)
