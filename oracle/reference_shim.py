"""
Import the UNMODIFIED reference `repet.py` from /root/reference (build container only).

TEST INFRASTRUCTURE.  The reference cannot be imported as-is on this image
(SURVEY.md section 8(c)): matplotlib is absent and `scipy.signal.hamming/triang` moved to
`scipy.signal.windows`.  This shim stubs/aliases those names *before* import and loads the
reference under the module name `repet_reference` without touching the reference file.
`/root/reference` does not exist on the GPU box: `load()` returns None there and callers
(only `oracle/make_golden.py` and the optional `-m "not gpu"` cross-check) skip.
"""

import importlib.util
import os
import sys
import types

REFERENCE_PATH = "/root/reference/repet.py"


def available():
    return os.path.isfile(REFERENCE_PATH)


def load():
    if not available():
        return None
    if "repet_reference" in sys.modules:
        return sys.modules["repet_reference"]
    import scipy.signal
    import scipy.signal.windows

    if "matplotlib" not in sys.modules:
        stub = types.ModuleType("matplotlib")
        stub.pyplot = types.ModuleType("matplotlib.pyplot")
        sys.modules["matplotlib"] = stub
        sys.modules["matplotlib.pyplot"] = stub.pyplot
    if not hasattr(scipy.signal, "hamming"):
        scipy.signal.hamming = scipy.signal.windows.hamming
    if not hasattr(scipy.signal, "triang"):
        scipy.signal.triang = scipy.signal.windows.triang
    spec = importlib.util.spec_from_file_location("repet_reference", REFERENCE_PATH)
    module = importlib.util.module_from_spec(spec)
    sys.modules["repet_reference"] = module
    spec.loader.exec_module(module)
    return module
