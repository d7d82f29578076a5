import sys
import numpy as np
sys.path.insert(0, ".")
import ctypes as C
import lightworks_b200 as lb
from lightworks_b200 import _lib
from oracle import oracle as O


def pend(tag):
    buf = C.create_string_buffer(200)
    rc = _lib.cdll().lwb_debug_pending_cuda_error(buf, 200)
    print(f"[{tag}] pending cuda error: {rc} {buf.value.decode()}", flush=True)


ctx = lb.default_context(0)
U = O.random_unitary(9, 4)
ins = [1] * 7 + [0, 0]
print("single ctx ok", ctx.perm_basis_probs(U, ins).sum()); pend("after first call")
comm = lb.Comm.local(2); pend("after Comm.local(2)")
print(comm.info())
try:
    print("default ctx again", ctx.perm_basis_probs(U, ins).sum())
except Exception as e:
    print("default ctx FAILED:", e)
pend("after default ctx call")
try:
    c1 = comm.context(1)
    print("device-1 ctx", c1.perm_basis_probs(U, ins).sum())
except Exception as e:
    print("device-1 ctx FAILED:", e)
pend("after device-1 ctx call")
try:
    got, tot = comm.perm_basis_probs(U, ins)
    print("comm ok", tot, np.array_equal(got, ctx.perm_basis_probs(U, ins)))
except Exception as e:
    print("comm FAILED:", e)
pend("after comm call")
