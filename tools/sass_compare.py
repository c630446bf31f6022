"""Compare the SASS of every kernel two builds of the library have in common:
     python tools/sass_compare.py old/libcup2d_b200.so cup2d_b200/libcup2d_b200.so
Used to show that host-side changes and added kernels left the kernels measured on the GPU instruction-identical (the build a
profile under profiles/ was taken from: `git archive <commit> cup2d_b200/csrc include | tar -x -C /tmp/old && make -C ...`)."""
import re
import subprocess
import sys


def kernels(so):
    out = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True, check=True).stdout
    res, cur, name = {}, [], None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                res[name] = cur
            name, cur = m.group(1), []
        elif name and re.match(r"\s*/\*[0-9a-f]{4}\*/", line):
            cur.append(re.sub(r"/\*[0-9a-f]{4}\*/", "", line).strip())
    if name:
        res[name] = cur
    return res


def main():
    old, new = kernels(sys.argv[1]), kernels(sys.argv[2])
    diff = [k for k in sorted(old) if k in new and old[k] != new[k]]
    gone = [k for k in sorted(old) if k not in new]
    for k in diff:
        print("DIFFERENT", k, len(old[k]), "->", len(new[k]), "instructions")
    for k in gone:
        print("MISSING  ", k)
    print(f"{len(old)} kernels in {sys.argv[1]}: {len(old) - len(diff) - len(gone)} identical in {sys.argv[2]}, {len(diff)} different, "
          f"{len(gone)} missing; {len(set(new) - set(old))} kernels only in the second")
    return 1 if diff or gone else 0


if __name__ == "__main__":
    sys.exit(main())
