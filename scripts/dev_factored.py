"""developer timing (not part of the bench contract): dense vs factored sweep
tables on configs #3 (41x61), #5 (2000x500) and a reduced SEAREV grid.

    python scripts/dev_factored.py [ar1] [large] [searev]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import stodynprog_b200 as sdp  # noqa: E402
import workloads as wl  # noqa: E402
from stodynprog_b200 import _cabi  # noqa: E402
from dev_timing import time_sweeps, opt  # noqa: E402


def make(which, compress, layout="auto"):
    if which == "ar1":
        prob = wl.storage_ar1(sdp)
    elif which == "large":
        prob = wl.storage_ar1_large(sdp, n_E=int(os.environ.get("N_E", "2000")), n_P=500)
    else:
        prob = wl.searev(sdp, n_E=int(os.environ.get("SEAREV_N_E", "31")))
    sv = prob.solver
    sv.table_compress = compress
    sv.table_layout = layout
    return sv


if __name__ == "__main__":
    todo = sys.argv[1:] or ["ar1", "large"]
    for which in todo:
        sums = {}
        for compress in os.environ.get("COMPRESS", "on,off").split(","):
            layouts = ["auto"] if which != "large" else ["state_minor"]
            for layout in layouts:
                sv = make(which, compress, layout)
                t0 = time.perf_counter()
                T = sv.sweep_tables()
                print("%s compress=%s layout=%s setup %.1fs items %d table %.3f GB (%.2f B/backup)"
                      % (which, compress, T.layout_name, time.perf_counter() - t0, T.n_items,
                         T.device_bytes / 1e9, T.streamed_bytes_per_backup), flush=True)
                lib = sv.engine.lib
                if T.tiled:
                    variants = [dict()]
                elif T.factored:
                    variants = [dict(hoist=1, hoist_upl=2), dict(hoist=1, hoist_upl=4), dict(hoist=0, upl=4)]
                else:
                    variants = [dict(upl=4), dict(upl=2)]
                for v in variants:
                    opt(lib, **v)
                    r = time_sweeps(sv, T)
                    sums.setdefault(r["checksum"], []).append((compress, v))
                    print("   ", v, r, flush=True)
                del sv, T
                torch.cuda.empty_cache()
        print(which, "distinct checksums:", len(sums), flush=True)
