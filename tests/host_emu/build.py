"""TEST INFRASTRUCTURE: builds tests/host_emu/_build/libamr_emu.so = csrc/amr_ops.cu (kernel launches rewritten into
serial loops, CUDA runtime calls mapped to malloc/memcpy by cuda_host_shim.h) + csrc/amr_plan.cpp, with g++.
The product library is NOT involved and nothing here ships."""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "cup2d_b200", "csrc")
OUT = os.path.join(HERE, "_build")

GLUE = r'''
#include <string>
namespace cup2d {
static std::string g_err;
void set_error(const std::string &m) { g_err = m; }
int dim_of(int f) { return (f == CUP2D_VEL || f == CUP2D_VOLD || f == CUP2D_TMPV) ? 2 : 1; }
}
extern "C" const char *cup2d_last_error(void) { return cup2d::g_err.c_str(); }
'''


def build():
    os.makedirs(OUT, exist_ok=True)
    src = open(os.path.join(CSRC, "amr_ops.cu")).read()
    src = src.replace('#include "sim.h"', '#include "cuda_host_shim.h"')
    launch = re.compile(r"(\w+(?:<[^<>;]*>)?)<<<(.*?),\s*(\d+|\w+),\s*0,\s*a->stream>>>\((.*?)\);", re.S)
    src, n = launch.subn(lambda m: f"emu_launch({m.group(2)}, {m.group(3)}, [&] {{ {m.group(1)}({m.group(4)}); }});", src)
    assert n >= 10 and "<<<" not in src, f"launch rewrite incomplete ({n})"
    emu = os.path.join(OUT, "amr_ops_emu.cpp")
    open(emu, "w").write(src + GLUE)
    lib = os.path.join(OUT, "libamr_emu.so")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++20", "-fPIC", "-shared", "-pthread", "-DCUP2D_AMR_EMU", "-I", HERE, "-o", lib, emu,
                    os.path.join(CSRC, "amr_plan.cpp")], check=True)
    return lib


if __name__ == "__main__":
    print(build())
