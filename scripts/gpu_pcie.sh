#!/bin/bash
cat > /tmp/pc.py <<'PY'
import torch, time
n = 128 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=10):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t) / reps
for _ in range(2): run(True, True)
a, b, c = run(True, False), run(False, True), run(True, True)
print("H2D alone %.1f GB/s, D2H alone %.1f GB/s, both at once: %.1f GB/s each direction (%.1f combined)" % (n / a / 1e9, n / b / 1e9, n / c / 1e9, 2 * n / c / 1e9))
PY
python /tmp/pc.py
