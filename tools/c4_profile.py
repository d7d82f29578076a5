import sys, time, cProfile, pstats
sys.path.insert(0, ".")
import lightworks_b200 as lb
from oracle import oracle as O
circ16 = lb.CompiledCircuitView(O.random_unitary(16, 4), 16)
be = lb.PermanentBackend()
from lightworks_b200.convolution import SourceClasses, pdist_calc
t0=time.perf_counter(); classes = SourceClasses.from_source([1]*8+[0]*8, 0.99, 0.9, 0.95, 1e-9); print("source classes", time.perf_counter()-t0, len(classes))
acc = pdist_calc(circ16, classes, be, keep_on_device=True)  # warm
be.ctx.synchronize()
t0=time.perf_counter(); acc = pdist_calc(circ16, classes, be, keep_on_device=True); be.ctx.synchronize(); print("pdist_calc", time.perf_counter()-t0, acc.n_convolutions)
pr=cProfile.Profile(); pr.enable(); acc = pdist_calc(circ16, classes, be, keep_on_device=True); be.ctx.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
# the whole Sampler chain (bench.py extra c4_noisy_sampler), stage by stage
import numpy as np
from lightworks_b200 import _lib
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter()
counts = lb.sample_source_outputs(circ16, [1] * 8 + [0] * 8, 1_000_000, be, purity=0.99, brightness=0.9, indistinguishability=0.95,
                                  detector_efficiency=0.9, seed=7)
print("sample_source_outputs", time.perf_counter() - t0, len(counts))
pr.disable()
pstats.Stats(pr).sort_stats("cumtime").print_stats(22)
