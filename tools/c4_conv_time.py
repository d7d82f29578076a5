"""C4 pdist_calc once (after a warm-up): target for `ncu -k regex:convolve --metrics gpu__time_duration.sum`."""
import sys
import time

sys.path.insert(0, ".")
import lightworks_b200 as lb  # noqa: E402
from lightworks_b200.convolution import SourceClasses, pdist_calc  # noqa: E402
from oracle import oracle as O  # noqa: E402, N812

circ16 = lb.CompiledCircuitView(O.random_unitary(16, 4), 16)
be = lb.PermanentBackend()
classes = SourceClasses.from_source([1] * 8 + [0] * 8, 0.99, 0.9, 0.95, 1e-9)
for _ in range(2):
    t0 = time.perf_counter()
    acc = pdist_calc(circ16, classes, be, keep_on_device=True)
    be.ctx.synchronize()
    print("pdist_calc", time.perf_counter() - t0, acc.n_convolutions, flush=True)
