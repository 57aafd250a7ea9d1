"""`import pyhispmv` -- the drop-in name of the reference's pybind11 module (pyhispmv/setup.py:15-38).

The compiled extension lives inside the package (hispmv_b200/pyhispmv.*.so, built in-tree by `make`);
this shim only re-exports it under the reference's top-level name.  No fallback: if the extension is not
built, or no B200 is visible when FpgaHandle is constructed, the error propagates.
"""
from hispmv_b200.pyhispmv import FpgaHandle  # noqa: F401
from hispmv_b200.pyhispmv import __doc__  # noqa: F401
