"""problem factories shared by make_golden.py (which needs /root/reference) and the tests
(which only read the committed fixtures)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import workloads as wl  # noqa: E402


def column_cases(api, **kw):
    """the two problems of column_cases.npz: grids with >= 32 rows of state axis 0, which
    the column-shared layout (CF) takes"""
    ar1 = wl.storage_ar1(api, n_E=70, n_P=5, n_w=9, steps=(0.5, 0.1), **kw)
    sea = wl.searev(api, n_E=33, n_S=4, n_A=3, **kw)
    sea.solver.control_steps = (.05,)
    return ar1, sea
