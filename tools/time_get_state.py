"""Times bsfg_get_state / bsfg_set_state field by field on the C4 shape (diagnostic for bench.py's e2e leg)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sparsefactormixedmodel_b200 import BSFGSweep, _lib
from sparsefactormixedmodel_b200._lib import ptr, check
n, p, k = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "C4"]
data, rv, pri, rp, cs = bench.make_problem(n, p, k)
sw = BSFGSweep(data, rv, pri, rp, k_init=k, k_max=k, seed=1)
sw.set_state(cs); sw.sweep(2)
shapes = dict(Lambda=(p, k), Lambda_prec=(p, k), Plam=(p, k), F=(n, k), F_a=(n, k), F_h2=(k,), B=(1, p), resid_Y_prec=(p,),
              E_a_prec=(p,), delta=(k,), tauh=(k,))
for rep in range(2):
    for name, shp in shapes.items():
        buf = np.zeros(shp, order="F")
        st = _lib.bsfg_state(); setattr(st, name, ptr(buf))
        t0 = time.perf_counter(); check(sw.lib.bsfg_get_state(sw._ctx, C.byref(st))); dt = time.perf_counter() - t0
        print(f"rep{rep} get {name:14s} {buf.nbytes/1e6:8.2f} MB {dt*1e3:8.2f} ms", flush=True)
for rep in range(2):
    t0 = time.perf_counter(); out = sw.get_state(with_E_a=False); print("get_state total ms", (time.perf_counter() - t0) * 1e3)
    t0 = time.perf_counter(); sw.set_state(cs); print("set_state total ms", (time.perf_counter() - t0) * 1e3)
    t0 = time.perf_counter(); sw.sweep(1); print("sweep(1) ms", (time.perf_counter() - t0) * 1e3)
