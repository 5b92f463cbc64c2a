"""Build libdetex_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch dependency)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdetex_b200.so")
SOURCES = ["kernels.cu", "convert.cu", "aux.cu", "ingest.cu", "host_tables.cu", "api.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr", "-diag-suppress", "177", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared",
]


def _newest_source_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def find_nvcc():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    return None


def build_id():
    """Hash of everything the library is compiled from (csrc/, include/, the flags): goes into
    detexb200Version() so that measurements can be tied to the build that produced them."""
    import hashlib
    h = hashlib.sha1(" ".join(NVCC_FLAGS).encode())
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            h.update(f.encode())
            h.update(open(os.path.join(root, f), "rb").read())
    return h.hexdigest()[:12]


def build(force=False, verbose=False):
    """Compile the CUDA library if it is missing or older than its sources."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_source_mtime():
        return LIB
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found; cannot build libdetex_b200.so")
    cmd = [nvcc] + NVCC_FLAGS + ['-DDETEXB200_BUILD_ID="%s"' % build_id(), "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
