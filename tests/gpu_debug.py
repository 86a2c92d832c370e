"""Not a test: prints the per-key errors of every case (run on the GPU box while debugging)."""
import sys
import traceback

import numpy as np

from . import cases as K
from . import golden_io, runners


def main():
    which = sys.argv[1:] or ["golden", "config"]
    worst = 0.0
    if "golden" in which:
        for c in K.GOLDEN_CASES:
            try:
                params, io, ref = golden_io.load(c)
                got = runners.run_cuda(c, params=params, io=io)
                e = runners.compare(got, ref)
                worst = max(worst, max(e.values()))
                print(f"{c.name:26s} max {max(e.values()):.2e}  " + " ".join(f"{k}={v:.1e}" for k, v in e.items()), flush=True)
            except Exception:
                print(f"{c.name:26s} FAILED"); traceback.print_exc()
    if "config" in which:
        for c in K.CONFIG_CASES:
            try:
                params, io = K.make_params(c), K.make_io(c)
                got = runners.run_cuda(c, params=params, io=io)
                o32 = runners.run_oracle(c, params=params, io=io, dtype=np.float32)
                e = runners.compare(got, o32)
                worst = max(worst, max(e.values()))
                print(f"{c.name:26s} max {max(e.values()):.2e}  " + " ".join(f"{k}={v:.1e}" for k, v in e.items()), flush=True)
            except Exception:
                print(f"{c.name:26s} FAILED"); traceback.print_exc()
    print("WORST", worst)


if __name__ == "__main__":
    main()
