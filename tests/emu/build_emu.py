"""Build tests/emu/_build/libsse_b200_emu.so: the CUDA sources of the product compiled with g++
against tests/emu/cuda_emu.h (-DSSE_HOST_EMU), every CUDA thread a fiber.  Test infrastructure
only -- see the header of cuda_emu.h."""
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "stablespectralelements.jl_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libsse_b200_emu.so")
CMD = ["g++", "-O1", "-std=c++17", "-DSSE_HOST_EMU", "-x", "c++", "-I", HERE, "-fPIC", "-shared"]


def _digest():
    h = hashlib.sha256(" ".join(CMD).encode())
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))]
    files += [os.path.join(HERE, "cuda_emu.h"), os.path.join(ROOT, "include", "sse_b200.h")]
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(asan: bool = False, defines=()) -> str:
    """``asan``: an AddressSanitizer build (every "device" buffer and the dynamic shared memory
    of each launch are heap blocks, so out-of-bounds accesses of the kernels trap).
    ``defines``: extra -D tuning knobs (e.g. ("SSE_STD_NB=2",)) to check a kernel variant."""
    tag = ("_asan" if asan else "") + "".join("_" + d.replace("=", "") for d in defines)
    out = OUT.replace(".so", tag + ".so")
    extra = ["-g", "-fsanitize=address", "-fno-omit-frame-pointer"] if asan else []
    extra += ["-D" + d for d in defines]
    os.makedirs(os.path.dirname(out), exist_ok=True)
    stamp, dig = out + ".stamp", _digest() + tag
    if os.path.exists(out) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return out
    subprocess.run(CMD + extra + ["-o", out, os.path.join(CSRC, "sse_b200.cu"), "-ldl"],
                   check=True, cwd=CSRC)
    with open(stamp, "w") as f:
        f.write(dig)
    return out


if __name__ == "__main__":
    print(build())
