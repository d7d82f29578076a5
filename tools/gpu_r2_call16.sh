#!/bin/bash
# round 2, GPU call 16: full GPU suite with the complex64 multiplicity-aware walk and the 128x3 K1x launch shape; timings
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu16.log 2>&1; echo "pytest exit=$?" >> gpurun_out/r2_pytest_gpu16.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2_pytest_gpu16.log | tail -10
python tools/k1x_quick.py base 2>&1 | tee gpurun_out/r2_k1x_quick.log
python - <<'PY' 2>&1 | tee gpurun_out/r2_c64_timing.log
import sys, torch
sys.path.insert(0, ".")
from lightworks_b200 import _lib as L
from oracle import oracle as O
ctx = L.Context(0); s = torch.cuda.Stream(); ctx.set_stream(s.cuda_stream)
with torch.cuda.stream(s):
    for (m, n, seed) in [(24, 10, 3), (20, 12, 9)]:
        U = torch.from_numpy(O.random_unitary(m, seed)).cuda()
        ins = torch.tensor([1] * n + [0] * (m - n), dtype=torch.uint8, device="cuda")
        S = min(L.fock_count(m, n), 100_000_000)
        p = torch.empty(S, dtype=torch.float64, device="cuda")
        for dt in (0, 1):
            best = 1e9
            for rep in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s); ctx.perm_basis_probs_dev(U, m, ins, n, 0, S, p, dtype=dt); e1.record(s); e1.synchronize()
                best = min(best, e0.elapsed_time(e1))
            print(f"m={m} n={n} S={S} dtype={'c64' if dt else 'c128'}: {best:.3f} ms sum={float(p.sum()):.9f}")
PY
