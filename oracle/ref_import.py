"""Import the UNMODIFIED reference package from /root/reference/src (build container only).

Test infrastructure: used by oracle/gen_golden.py to pin oracle/ctrm_oracle.py against the
reference itself and to write tests/golden/*.  /root/reference does not exist on the GPU box;
nothing under tests/ -m gpu, smoke() or bench.py may call this.
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CTRM_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "src", "ctrm")) and os.path.isdir(os.path.join(HERE, "_ref"))


def import_reference():
    """returns the reference `ctrm` module, pinned to CPU and one torch thread."""
    if not available():
        raise RuntimeError("reference not available (needs /root/reference and oracle/_ref; run `make -C oracle ref`)")
    for p in (os.path.join(REF, "src"), os.path.join(HERE, "_ref"), os.path.join(HERE, "ref_stubs")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch

    torch.set_num_threads(1)
    import ctrm  # noqa: F401  (the reference)
    import ctrm.utils

    ctrm.utils.set_device_name("cpu")
    return ctrm
