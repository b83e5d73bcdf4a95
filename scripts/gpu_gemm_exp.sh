#!/bin/bash
for g in 148 111 74 37; do echo "grid $g"; RB_GEMM_GRID=$g python scripts/gemm_bench.py 2>&1 | grep '"variant": 1' | head -3; done
