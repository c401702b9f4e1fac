"""ctypes view of the lfvio C-ABI (include/lfvio.h) — test/bench harness only.

`import lfvio` loads lf-vio_b200/liblfvio.so and declares its prototypes (lfvio/_api.py): there is no compute and no fallback
here, a missing library fails the import loudly.

With LFVIO_TYPES_ONLY=1 in the environment only the library-free parts are importable — the interface structs
(lfvio.backend, lfvio.prestep), the synthetic workload generators (lfvio.synth, lfvio.synthwin) and the configuration
(lfvio.config).  bench.py's `--impl reference` arm uses that mode, so that the process that times the CPU path never maps the
product library.
"""
import os

TYPES_ONLY = os.environ.get("LFVIO_TYPES_ONLY") == "1"

from .config import PAL_UP, scaled_camera, tracker_defaults  # noqa: E402,F401

if not TYPES_ONLY:
    from ._api import *  # noqa: E402,F401,F403
    from ._api import _as, _img  # noqa: E402,F401  (helpers some tests use)
