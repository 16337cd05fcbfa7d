"""Build libepgx.so in-tree with nvcc for sm_100a:  python -m epgpy_b200.build"""

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libepgx.so")


def sources():
    return sorted(os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith(".cu"))


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(SRC, f) for f in os.listdir(SRC)] + [os.path.join(ROOT, "include", "epgx.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None, only=None):
    """defines / out: experimental builds (e.g. defines=["-DEPGX_WF=0"], out="libepgx_nowf.so"; run them with
    EPGX_LIB=<path>); only: substrings of the sources to recompile (the other objects are reused)"""
    out = os.path.join(HERE, out) if out else OUT
    if not force and not defines and out == OUT and not stale():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    tag = ("." + "".join(c for c in "_".join(defines) if c.isalnum() or c == "_")) if defines else ""
    for src in sources():
        obj = src[:-3] + tag + ".o"
        objs.append(obj)
        if only and not any(o in os.path.basename(src) for o in only) and os.path.exists(obj):
            continue
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", *defines,
               "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-c", src, "-o", obj]
        if os.path.basename(src) == "epgx_host.cu":  # host marshalling: AVX2 streaming stores (every x86 host of a B200 has them)
            cmd[cmd.index("-Xcompiler") + 1] = "-fPIC,-mavx2"
        if verbose:
            cmd += ["-Xptxas", "-v"]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        log, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{log}")
        if verbose:
            print(log)
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    defs = [a for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    only = [a[7:] for a in sys.argv[1:] if a.startswith("--only=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else None, only=only or None))
