"""`tools/build_operators.py` under its own name: only `build_kernel` (:327-351) is live code upstream (the sparse
operator builders of that file have no caller); it is host arithmetic on a (2k-1) x (2k-1) matrix."""
from ..stencils import build_kernel  # noqa: F401
