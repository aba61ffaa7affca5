"""The checker the `-m gpu` parity tests compare the CUDA path with -- TEST INFRASTRUCTURE ONLY.

Every Rcpp export (`BSFG_functions_c.cpp:26-339`) is served by the compiled reference itself
(`oracle/_ref/libbsfg_reference.so`, the unmodified source file; `oracle/ref_oracle.py`) whenever that library is
present -- it is built in the development container and travels to the GPU box with the snapshot.  The R-only steps
of the loop body (`fast_BSFG_sampler.R:176-270`, `BSFG_functions.R:272-372`; R cannot run here) come from the NumPy
restatement `oracle/bsfg_oracle.py`, which `tests/test_oracle.py` pins to the compiled reference on CPU.
`KIND` says which of the two answered, and `tests/test_golden_gpu.py::test_checker_is_the_compiled_reference` fails
when the GPU suite silently fell back to the restatement.
"""
from . import bsfg_oracle as _O
from . import ref_oracle as _R
from .bsfg_oracle import *  # noqa: F401,F403  (R-only steps, helpers)

for _name in dir(_O):                      # also the underscore-free helpers `import *` skips when __all__ is absent
    if not _name.startswith("__") and _name not in globals():
        globals()[_name] = getattr(_O, _name)

if _R.available():
    KIND = "reference"
    for _name in _R.C_FUNCTIONS:
        globals()[_name] = getattr(_R, _name)
    gibbs_iteration = _R.gibbs_iteration
    gibbs_iteration_px = _R.gibbs_iteration_px

    def gibbs_iteration_fixedlambda(data, rv, priors, rp, st, Lambda, v):
        return _O.gibbs_iteration_fixedlambda(data, rv, priors, rp, st, Lambda, v, fns=_R)
else:                                      # pragma: no cover - only on a box that received no oracle/_ref
    KIND = "port"
