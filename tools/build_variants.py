"""Build tuning variants of libsse_b200.so for an A/B sweep on the GPU box (tools/gpu_variants.sh):

    python tools/build_variants.py fd_minb3:SSE_FD_MINB=3 std_nb2:SSE_STD_NB=2,SSE_NODAL_NB=4 ...

Each argument is name:DEFINE[=v][,DEFINE...]; the libraries go to build/variants/<name>.so
(git-ignored, shipped to the box by gpurun), compiled in parallel.  Check a variant's results
on the CPU first with tests/emu/build_emu.build(defines=...) -- see
tests/test_kernels_host_emulation.py::test_tuning_knob_variants_in_emulation."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "stablespectralelements.jl_b200", "csrc")
OUT = os.path.join(ROOT, "build", "variants")


def build(spec: str) -> str:
    name, defs = spec.split(":", 1)
    out = os.path.join(OUT, name + ".so")
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    return ge.compile_library(out, defines=[d for d in defs.split(",") if d],
                              objdir=os.path.join(ROOT, "build", "obj_" + name))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    with ThreadPoolExecutor(max_workers=2) as pool:
        for path in pool.map(build, sys.argv[1:]):
            print("built", path)
