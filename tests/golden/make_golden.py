"""Generate the committed golden fixtures: seeded inputs and the oracle's residual for small
instances of every BASELINE.json configuration.  Run from the repo root:

    python tests/golden/make_golden.py

The reference itself cannot run here (Julia), so these vectors come from the oracle, which is
pinned to the reference's own golden L2 errors by tests/test_oracle_golden.py.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import cases  # noqa: E402
import sse_oracle as oc  # noqa: E402
from bridge import oracle_problem  # noqa: E402

FIXTURES = {
    "cfg1_adv2d_tri_p4": ("advection_tri_case", dict(p=4, M=2)),
    "cfg2_euler2d_tri_p4_lf": ("euler_tri_case", dict(p=4, M=2)),
    "cfg3_adv3d_tet_p4": ("advection_tet_case", dict(p=4, M=2)),
    "cfg4_euler3d_tet_p4_lf": ("euler_tet_case", dict(p=4, M=2, warp=True, ic="periodic")),
    "cfg4_euler3d_tet_p3_ec": ("euler_tet_case", dict(p=3, M=2, warp=True, interface="ec",
                                                      ic="periodic")),
    "cfg5_advdiff1d_p4": ("advection_diffusion_case", dict(d=1, p=4, M=4)),
    "cfg5_advdiff2d_p3": ("advection_diffusion_case", dict(d=2, p=3, M=2)),
}


def build(name):
    fn, kw = FIXTURES[name]
    solver, u0 = getattr(cases, fn)(lazy=True, **kw)
    u = cases.rough_state(solver, u0, seed=11)
    return solver, u


if __name__ == "__main__":
    for name in FIXTURES:
        solver, u = build(name)
        dudt = oc.semi_discrete_residual(oracle_problem(solver), u)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), u=u, dudt=dudt)
        print(name, u.shape, float(np.abs(dudt).max()))
