"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container).

    python -m tests.golden.make_golden            # from the repo root: every case
    python -m tests.golden.make_golden NAME ...   # only the named cases (new ones: the others stay byte-identical)

For every case in tests/cases.GOLDEN_CASES this imports the reference class (tests/refload.py),
loads the case's weights, runs forward + backward on CPU in fp32 and stores inputs, weights and
the reference's outputs.  The fixtures travel to the GPU box, where /root/reference does not
exist.  Never generate fixtures from our own implementation.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests import cases as K  # noqa: E402
from tests import refload, runners  # noqa: E402


def main():
    if not refload.available():
        raise SystemExit("reference not available; fixtures can only be generated in the build container")
    import torch
    torch.set_num_threads(1)  # fixed reduction order
    total = 0
    only = set(sys.argv[1:])
    unknown = only - {c.name for c in K.GOLDEN_CASES}
    if unknown:
        raise SystemExit(f"not in tests/cases.GOLDEN_CASES: {sorted(unknown)}")
    for c in K.GOLDEN_CASES:
        if only and c.name not in only:
            continue
        params = K.make_params(c)
        u, g = K.make_io(c)
        ref = runners.run_reference(c, params=params, io=(u, g))
        blob = {"u": u, "gout": g}
        blob.update({"p_" + k: v for k, v in params.items()})
        blob.update({"ref_" + k: v.astype(np.float32) for k, v in ref.items()})
        path = os.path.join(HERE, c.name + ".npz")
        np.savez(path, **blob)
        total += os.path.getsize(path)
        print(f"{c.name:28s} {os.path.getsize(path)/1024:8.1f} KiB  keys={sorted(ref)}")
    print(f"total {total/1e6:.2f} MB; torch {torch.__version__}")


if __name__ == "__main__":
    main()
