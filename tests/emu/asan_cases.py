"""Run under LD_PRELOAD=libasan with the AddressSanitizer build of the emulated library
(tests/test_kernels_host_emulation.py::test_kernels_address_sanitizer): one residual of every
kernel family, plus device geometry and a sharded BR1 flow."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), HERE):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

import cases  # noqa: E402
from sse_b200 import device as dev  # noqa: E402

lib = dev.load_library(sys.argv[1], allow_emulation=True)
dev._LIB = lib
import test_kernels_host_emulation as te  # noqa: E402

for name in ("euler3d_tet_p4_warp_lf", "adv3d_tet_p4", "euler2d_tri_p4_lf", "advdiff2d_p3",
             "euler3d_hex_nodal_p3_ec", "burgers2d_tri_p3_ec"):
    solver, u0 = te.CASES[name][0]()
    u = cases.rough_state(solver, u0, seed=1)
    d = dev.DeviceResidual(solver)
    dudt = np.full_like(u, np.nan)
    d.residual_host(u, dudt)
    assert np.all(np.isfinite(dudt)), name
    d.close()
    print("ok", name, flush=True)
# (the functionals are left out: their emulated warp shuffles cost two fiber switches per thread
# and step, which AddressSanitizer's swapcontext interception makes ~100x slower)
import test_gpu_geometry as tg  # noqa: E402
for case in tg._meshes():
    if case[0] in ("tet_exact_straight", "hex_curl_warped", "line_exact"):
        tg.test_device_geometry_matches_host(case)
print("ok geometry", flush=True)
import test_gpu_sharded_emulation as ts  # noqa: E402
ts.test_sharded_flow_matches_single_domain("advdiff2d_p3_br1", 3, shard_cls=te._emu_shard_class())
print("ok sharded", flush=True)
print("ASAN-CASES-DONE")
