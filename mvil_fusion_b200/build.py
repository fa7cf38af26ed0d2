"""Builds libvils_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libvils_b200.so")
HOST_OUT = os.path.join(HERE, "libvils_host.so")
SOURCES = ["common.cu", "vils_ba.cu", "vils_lidar.cu", "vils_preint.cu", "vils_klt.cu", "vils_frontend.cu", "vils_assoc.cu", "vils_vgicp.cu"]
# per-file extra flags: the FP32 LiDAR path must not contract a*b+c into FMA (bit parity with the reference arithmetic)
EXTRA = {"vils_lidar.cu": ["--fmad=false"], "vils_klt.cu": ["--fmad=false"], "vils_frontend.cu": ["--fmad=false"], "vils_assoc.cu": ["--fmad=false"]}
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if os.path.isfile(os.path.join(CSRC, f))] + [os.path.join(CSRC, "host", f) for f in os.listdir(os.path.join(CSRC, "host"))] + [os.path.join(HERE, "..", "include", "vils_cabi.h")]
    if not os.path.exists(HOST_OUT):
        return True
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        if not os.path.exists(src):
            continue
        obj = os.path.join(HERE, "build", s + ".o")
        cmd = [nvcc] + NVCC_FLAGS + EXTRA.get(s, []) + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    ok = True
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {s}\n{out}\n")
        ok = ok and p.returncode == 0
    if not ok:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([nvcc, "-shared", "-o", OUT] + objs + ["-lcudart", "-ldl"])
    # C++ host mirror of the reference's class API (Estimator / FeatureTracker), plain g++ on top of the C-ABI
    host = [os.path.join(CSRC, "host", f) for f in ("vils_host.cpp", "vils_initial.cpp")]
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DVILS_HOST_HAS_INITIAL", "-o", HOST_OUT] + host + ["-L" + HERE, "-lvils_b200", "-Wl,-rpath,$ORIGIN"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
