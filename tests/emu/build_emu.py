"""Build tests/emu/libwbk_emu.so: the csrc/*.cu kernels compiled by g++ against the SIMT
emulation shim (TEST INFRASTRUCTURE ONLY -- never loaded by the product package)."""

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "wavebreaking_b200", "csrc")
LIB = os.path.join(HERE, "libwbk_emu.so")


def build(force=False):
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    deps += [os.path.join(HERE, "simt_emu.h"), os.path.join(HERE, "simt_emu.cpp"),
             os.path.join(ROOT, "include", "wbk.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    objs = []
    for s in srcs + [os.path.join(HERE, "simt_emu.cpp")]:
        o = os.path.join(HERE, "_obj_" + os.path.basename(s) + ".o")
        cmd = ["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-DWBK_EMU", "-ffp-contract=off", "-Wno-unused-result",
               "-x", "c++", "-include", os.path.join(HERE, "simt_emu.h"), "-I", CSRC, "-c", s, "-o", o]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("emu build failed:\n" + " ".join(cmd) + "\n" + res.stderr[-6000:])
        objs.append(o)
    res = subprocess.run(["g++", "-shared", "-o", LIB] + objs, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("emu link failed:\n" + res.stderr[-4000:])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
