"""Build libfuxi_b200.so in-tree with nvcc for sm_100a (no torch extension machinery needed:
the library is a plain C ABI and is loaded with ctypes)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libfuxi_b200.so")
SOURCES = ["api.cu", "project.cu", "inflate.cu", "edt.cu", "search.cu", "paths.cu", "band.cu", "small.cu", "field.cu", "assemble.cu", "post.cu", "replan.cu", "cloud.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "--expt-relaxed-constexpr"]


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "fuxi_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra_flags=(), out=None):
    """extra_flags / out: tuning experiments only (scripts/build_variants.sh builds libfuxi_b200_<tag>.so beside the
    default library; FUXI_B200_SO selects one at load time)."""
    global SO
    if out is not None:
        force = True
    if not force and not needs_build():
        return SO
    so = out or SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for f in SOURCES:
        o = os.path.join(HERE, "build", f.replace(".cu", ".o"))
        if out is not None:
            o = o.replace(".o", "." + os.path.basename(out) + ".o")
        cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, f), "-o", o]
        procs.append((f, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    failed = False
    for f, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            failed = True
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", so] + objs + ["-lcudart"], check=True)
    return so


if __name__ == "__main__":
    extra = [a for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[len("--out="):] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force=True, verbose="-v" in sys.argv, extra_flags=extra, out=os.path.join(HERE, outs[0]) if outs else None))
