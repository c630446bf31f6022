"""TEST INFRASTRUCTURE: builds tests/host_emu/_build/libamr_emu.so = csrc/amr_ops.cu + csrc/amr_fast.cu (kernel launches
rewritten into emulated launches, CUDA runtime calls mapped to malloc/memcpy by cuda_host_shim.h, the one PTX instruction of
weno.cuh replaced by a single-precision reciprocal seed) + csrc/amr_plan.cpp, with g++.  The product library is NOT involved
and nothing here ships."""
import os
import re
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "cup2d_b200", "csrc")
OUT = os.path.join(HERE, "_build")
COOP = {"amr_advect_fast_kernel", "amr_div_fast_kernel", "amr_scalar_fast_kernel"}   # kernels that use shared memory / warp barriers: threads must really run together

GLUE = r'''
#include <string>
namespace cup2d {
static std::string g_err;
void set_error(const std::string &m) { g_err = m; }
int dim_of(int f) { return (f == CUP2D_VEL || f == CUP2D_VOLD || f == CUP2D_TMPV) ? 2 : 1; }
}
extern "C" const char *cup2d_last_error(void) { return cup2d::g_err.c_str(); }
'''

def _fresh(out, inputs):
    """True if `out` exists and is newer than every input file (the builds below are skipped then)"""
    if not os.path.exists(out):
        return False
    t = os.path.getmtime(out)
    return all(os.path.getmtime(i) <= t for i in inputs)


def _inputs():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h", ".cpp"))]
    files += [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".h", ".py", ".cpp"))]
    return files + [os.path.join(ROOT, "include", "cup2d_b200.h")]


LAUNCH = re.compile(r"(\w+)((?:<[^<>;]*>)?)<<<(.*?),\s*(\w+),\s*(\w+),\s*a->stream>>>\((.*?)\);", re.S)


def rewrite(src):
    def sub(m):
        name, targs, grid, block, _smem, args = m.groups()
        fn = "emu_launch_coop" if name in COOP else "emu_launch"
        return f"{fn}({grid}, {block}, [&] {{ {name}{targs}({args}); }});"
    src, n = LAUNCH.subn(sub, src)
    assert "<<<" not in src, "launch rewrite incomplete"
    src = src.replace("extern __shared__ __align__(16) double af_smem[];", "static double af_smem[AF_SMEM / 8];")
    return src, n


def build(defines=(), tag=""):
    """defines: extra -D flags (e.g. the arithmetic variants of weno.cuh); tag names the resulting library"""
    os.makedirs(OUT, exist_ok=True)
    if _fresh(os.path.join(OUT, f"libamr_emu{tag}.so"), _inputs()):
        return os.path.join(OUT, f"libamr_emu{tag}.so")
    open(os.path.join(OUT, "sim.h"), "w").write('#pragma once\n#include "cuda_host_shim.h"\n')
    open(os.path.join(OUT, "common.cuh"), "w").write('#pragma once\n#include "cuda_host_shim.h"\n')
    shutil.copy(os.path.join(CSRC, "amr.h"), OUT)
    weno = open(os.path.join(CSRC, "weno.cuh")).read()
    weno, k = re.subn(r'asm\("rcp\.approx\.ftz\.f64 %0, %1;" : "=d"\(r\) : "d"\(x\)\);', "r = (double)(1.0f / (float)x);", weno)
    assert k == 1, "reciprocal seed not found in weno.cuh"
    open(os.path.join(OUT, "weno.cuh"), "w").write(weno)
    total = 0
    srcs = []
    for name in ("amr_ops.cu", "amr_fast.cu"):
        src, n = rewrite(open(os.path.join(CSRC, name)).read())
        total += n
        dst = os.path.join(OUT, name.replace(".cu", "_emu.cpp"))
        open(dst, "w").write(src + (GLUE if name == "amr_ops.cu" else ""))
        srcs.append(dst)
    assert total >= 16, f"only {total} launches rewritten"
    lib = os.path.join(OUT, f"libamr_emu{tag}.so")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++20", "-fPIC", "-shared", "-pthread", "-DCUP2D_AMR_EMU", *[f"-D{d}" for d in defines], "-I", OUT, "-I", HERE,
                    "-o", lib, *srcs, os.path.join(CSRC, "amr_plan.cpp")], check=True)
    return lib


# ---- the WHOLE product library under emulation -------------------------------------------------------------------------
FULL = os.path.join(HERE, "_build_full")
ASM_REWRITES = [  # (regex over the source text, replacement): the PTX of common.cuh / halo.cu / weno.cuh in host terms
    (r'asm volatile\("mbarrier\.init.*?\);', "(void)bar; (void)count;"),
    (r'asm volatile\("fence\.mbarrier_init.*?\);', ";"),
    (r'asm volatile\("mbarrier\.arrive\.expect_tx.*?: "memory"\);', "(void)bar; (void)bytes;"),
    (r'asm volatile\("\{\\n".*?: "memory"\);', "(void)bar; (void)parity; /* copies are synchronous here */"),
    (r'asm volatile\("cp\.async\.bulk\.shared.*?: "memory"\);', "emu_bulk_copy(smem_dst, gmem_src, bytes); (void)bar;"),
    (r'asm\("rcp\.approx\.ftz\.f64 %0, %1;" : "=d"\(r\) : "d"\(x\)\);', "r = (double)(1.0f / (float)x);"),
    (r'asm volatile\("cp\.async\.cg\.shared\.global.*?: "memory"\);', "emu_bulk_copy(smem_dst, gmem_src, 16);"),
    (r'asm volatile\("cp\.async\.wait_all;" ::: "memory"\);', ";"),
    (r'asm volatile\("st\.release\.sys.*?: "memory"\);', "__threadfence(); *(volatile unsigned long long *)p = v;"),
    (r'asm volatile\("st\.relaxed\.sys.*?: "memory"\);', "*(volatile unsigned long long *)p = v;"),
    (r'asm volatile\("ld\.acquire\.sys.*?: "memory"\);', "v = *(volatile const unsigned long long *)p; __threadfence();"),
    (r'asm volatile\("ld\.relaxed\.sys\.global\.u64.*?: "memory"\);', "v = *(volatile const unsigned long long *)p;"),
    (r'asm volatile\("ld\.relaxed\.sys\.global\.v2\.f64.*?: "memory"\);', "v = *p;"),
    (r'asm volatile\("ld\.relaxed\.sys\.global\.f64.*?: "memory"\);', "v = *(volatile const double *)p;"),
    (r'asm volatile\("mov\.u64 %0, %globaltimer;".*?"memory"\);', "t = emu_now_ns();"),
]
LAUNCH_ANY = re.compile(r"(\b\w+)((?:<[^<>;]*>)?)<<<(.*?)>>>\((.*?)\);", re.S)


def split_top(s):
    """split a launch configuration at top-level commas"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


SYNC_PAT = re.compile(r"__sync|__shfl|atomic|__shared__|mbar_|tma_load|__threadfence|volatile|asm\b|emu_bulk_copy")


def _functions(text, qualifier):
    """(name, body) of every function defined with `qualifier` in `text` (brace matching; declarations are skipped)"""
    out = []
    for m in re.finditer(re.escape(qualifier), text):
        i, depth = m.end(), 0
        while i < len(text) and not (text[i] in "{;" and depth == 0):
            depth += text[i] == "("
            depth -= text[i] == ")"
            i += 1
        if i >= len(text) or text[i] == ";":
            continue
        header = re.sub(r"__launch_bounds__\([^)]*\)", "", text[m.end():i])
        nm = re.search(r"(\w+)\s*\(", header)
        if not nm:
            continue
        j, depth = i, 0
        while j < len(text):
            depth += text[j] == "{"
            depth -= text[j] == "}"
            j += 1
            if depth == 0:
                break
        out.append((nm.group(1), text[i:j]))
    return out


def serial_safe_kernels(texts):
    """kernels whose threads never interact (no shared memory, barrier, shuffle, atomic, fence, volatile or PTX — directly or
    through a device function): they can be emulated one thread after the other, which is far cheaper than one OS thread
    per CUDA thread.  Everything else keeps real concurrent threads."""
    text = "\n".join(re.sub(r"//.*", "", t) for t in texts)
    dev = dict(_functions(text, "__device__"))
    syncing = {n for n, b in dev.items() if SYNC_PAT.search(b)}
    changed = True
    while changed:
        changed = False
        for n, b in dev.items():
            if n not in syncing and any(re.search(r"\b%s\s*[(<]" % re.escape(s), b) for s in syncing):
                syncing.add(n)
                changed = True
    safe = set()
    for n, b in _functions(text, "__global__"):
        if not SYNC_PAT.search(b) and not any(re.search(r"\b%s\s*[(<]" % re.escape(s), b) for s in syncing):
            safe.add(n)
    return safe


def _compile_link(flags, srcs, out, objdir, link_flags=()):
    """g++ every source to an object in parallel (one process per core), then link: the emulated library is ~12 translation
    units and rebuilt whenever a product source changes"""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(objdir, exist_ok=True)
    objs = [os.path.join(objdir, os.path.basename(f) + ".o") for f in srcs]

    def cc(pair):
        subprocess.run(["/usr/bin/g++", *flags, "-c", pair[0], "-o", pair[1]], check=True)
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        list(ex.map(cc, zip(srcs, objs)))
    subprocess.run(["/usr/bin/g++", *[f for f in flags if f.startswith(("-fsanitize", "-pthread", "-shared", "-fPIC"))], *link_flags, "-o", out, *objs],
                   check=True)


def build_full(defines=(), tag=""):
    """libcup2d_emu<tag>.so: every .cu / .cpp of cup2d_b200/csrc compiled with g++, one OS thread per CUDA thread;
    defines: extra -D flags (the measurement variants of advect.cu / weno.cuh)"""
    os.makedirs(FULL, exist_ok=True)
    if _fresh(os.path.join(FULL, f"libcup2d_emu{tag}.so"), _inputs()):
        return os.path.join(FULL, f"libcup2d_emu{tag}.so")
    srcs = []
    names = [n for n in sorted(os.listdir(CSRC)) if n.endswith((".cu", ".cuh", ".h", ".cpp"))]
    safe = serial_safe_kernels([open(os.path.join(CSRC, n)).read() for n in names])
    assert "amr_gather_kernel" in safe and "k_spmv" not in safe and "advect_stage_kernel" not in safe and "amr_advect_fast_kernel" not in safe, safe
    for name in names:
        text = open(os.path.join(CSRC, name)).read()
        for pat, rep in ASM_REWRITES:
            text = re.sub(pat, rep, text, flags=re.S)
        assert "asm" not in re.sub(r"//.*", "", text).replace("__asm", ""), f"PTX left in {name}"

        def sub(m):
            cfg = split_top(m.group(3))
            fn = "emu_launch_auto" if m.group(1) in safe else "emu_launch_coop"   # auto: serial unless EMU_ALL_COOP (sanitizer builds)
            return f"{fn}({cfg[0]}, {cfg[1]}, [&] {{ {m.group(1)}{m.group(2)}({m.group(4)}); }});"
        text = LAUNCH_ANY.sub(sub, text)
        assert "<<<" not in text, f"launch left in {name}"
        # shared memory becomes block-local storage of the emulated block (ranks emulated in one process run kernels concurrently)
        text = text.replace("extern __shared__ __align__(128) unsigned char smem[];",
                            "unsigned char *smem = (unsigned char *)emu_shared(__COUNTER__, ADV_SMEM);")
        text = text.replace("extern __shared__ __align__(16) double af_smem[];", "double *af_smem = (double *)emu_shared(__COUNTER__, AF_SMEM);")
        text = re.sub(r"__shared__\s+(?:__align__\(\d+\)\s+)?(\w+)\s+(\w+)\[([^\]]+)\];",
                      lambda m: f"{m.group(1)} *{m.group(2)} = ({m.group(1)} *)emu_shared(__COUNTER__, sizeof({m.group(1)}) * ({m.group(3)}));", text)
        text = re.sub(r"__shared__\s+(\w+)\s+(\w+);",
                      lambda m: f"{m.group(1)} &{m.group(2)} = *({m.group(1)} *)emu_shared(__COUNTER__, sizeof({m.group(1)}));", text)
        assert "__shared__" not in re.sub(r"//.*", "", text), f"__shared__ left in {name}"
        out = name.replace(".cu", "_emu.cpp") if name.endswith(".cu") else name
        open(os.path.join(FULL, out), "w").write(text)
        if out.endswith(".cpp"):
            srcs.append(os.path.join(FULL, out))
    lib = os.path.join(FULL, f"libcup2d_emu{tag}.so")
    _compile_link(["-O2", "-std=c++20", "-fPIC", "-shared", "-pthread", "-w", *[f"-D{d}" for d in defines], "-I", FULL, "-I", HERE],
                  srcs, lib, os.path.join(FULL, f"obj{tag}"))
    return lib


def build_tsan(defines=(), tag="", sanitize="thread"):
    """the emulated product sources + tsan_driver.cpp with -fsanitize=thread -> an executable that hunts for data races;
    sanitize="address,alignment,bounds" -> the same driver hunting for out-of-bounds accesses of device buffers and shared
    arrays (each its own exact-size allocation) and for vector loads/stores that are not aligned to their size"""
    build_full()
    srcs = [os.path.join(FULL, f) for f in sorted(os.listdir(FULL)) if f.endswith(".cpp")]
    exe = os.path.join(FULL, f"{'tsan' if sanitize == 'thread' else 'asan'}_driver{tag}")
    if _fresh(exe, _inputs()):
        return exe
    _compile_link(["-O1", "-g", "-std=c++20", "-pthread", f"-fsanitize={sanitize}",
                   *([] if sanitize == "thread" else ["-fno-sanitize-recover=all"]), "-w", "-DEMU_ALL_COOP",
                   *[f"-D{d}" for d in defines], "-I", FULL, "-I", HERE],
                  [os.path.join(HERE, "tsan_driver.cpp"), *srcs], exe, os.path.join(FULL, f"obj_{'tsan' if sanitize == 'thread' else 'asan'}{tag}"))
    return exe


if __name__ == "__main__":
    print(build())
