"""FP64 pipe issue-rate probes: how fast do DFMA/DMUL/DADD issue with register (not constant) operands?"""
import sys

sys.path.insert(0, ".")
from lightworks_b200 import _lib as L  # noqa: E402

ctx = L.Context(0)
peak = ctx.fp64_peak()
print(f"DFMA peak (1 register operand): {peak/1e12:.2f} TFLOP/s-equivalent")
names = {1: "DFMA 3 distinct register operands", 2: "DFMA shared A operand (reuse)", 3: "DMUL 2 registers",
         4: "DADD 2 registers", 5: "complex multiply chain (2 DMUL + 2 DFMA)"}
for mode in range(1, 6):
    r = ctx.fp64_probe(mode)
    print(f"mode {mode} {names[mode]:45s}: {r/1e12:6.2f}  ratio to peak {r/peak:.3f}  -> {2*peak/r:.2f} cycles/instr/scheduler")
