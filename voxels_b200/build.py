"""Builds every native artefact of the repo, in-tree (nothing is JIT-cached under ~/.cache):

  build/gen/vxb_tables_data.h     Transvoxel tables re-packed from the reference checkout (tools/gen_tables.cpp)
  voxels_b200/lib/libvxb200.so    CUDA kernels + C ABI (include/vxb200.h), sm_100a only
  voxels_b200/lib/libvoxels_b200.so, build/libvxh_b200.so   C++ drop-in Polygonizer + test harness over it
  oracle/_ref/*.so, build/oracle/libvxr_restate.so          test infrastructure (oracle/Makefile)

The built files travel to the GPU box with the repo snapshot; /root/reference does not, so everything
that needs the reference checkout (tables, drop-in, oracle) is built here and only loaded there.
"""
import os
import shutil
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("VOXELS_REFERENCE", "/root/reference")
GEN = os.path.join(REPO, "build", "gen")
LIB = os.path.join(REPO, "voxels_b200", "lib")
CSRC = os.path.join(REPO, "voxels_b200", "csrc")
TABLES = os.path.join(GEN, "vxb_tables_data.h")
LIBVXB = os.path.join(LIB, "libvxb200.so")
LIBDROPIN = os.path.join(LIB, "libvoxels_b200.so")
LIBVXH = os.path.join(REPO, "build", "libvxh_b200.so")

CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC,-fopenmp", "-ccbin", CXX]


def _run(cmd, **kw):
    print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True, **kw)


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.exists(s) and os.path.getmtime(s) <= t for s in sources)


def have_reference():
    return os.path.exists(os.path.join(REF, "src", "Transvoxel.inl"))


def build_tables(force=False):
    src = os.path.join(REPO, "tools", "gen_tables.cpp")
    inl = os.path.join(REF, "src", "Transvoxel.inl")
    if not have_reference():
        if os.path.exists(TABLES):
            return
        raise RuntimeError("build/gen/vxb_tables_data.h is missing and the reference checkout (%s) is not available to "
                           "generate it from" % REF)
    if not force and _newer(TABLES, [src, inl]):
        return
    os.makedirs(GEN, exist_ok=True)
    exe = os.path.join(GEN, "gen_tables")
    _run([CXX, "-O1", "-w", '-DVXB_TABLES_INL="%s"' % inl, src, "-o", exe])
    with open(TABLES, "w") as f:
        subprocess.run([exe], check=True, stdout=f)


def build_cuda(force=False, verbose=False):
    srcs = [os.path.join(CSRC, n) for n in os.listdir(CSRC) if n.endswith((".cu", ".cuh", ".h"))] + \
           [TABLES, os.path.join(REPO, "include", "vxb200.h")]
    if not force and _newer(LIBVXB, srcs):
        return
    os.makedirs(LIB, exist_ok=True)
    cmd = [NVCC] + NVCC_FLAGS + ["-shared", "-I", GEN, "-o", LIBVXB, os.path.join(CSRC, "vxb200.cu")]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    _run(cmd)


def build_dropin(force=False):
    """C++ drop-in (Voxels::Polygonizer & co. against the reference's own headers) + the test harness over it."""
    src = os.path.join(CSRC, "polygonizer_host.cpp")
    if not os.path.exists(src):
        return
    if not have_reference():
        if os.path.exists(LIBDROPIN):
            return
        raise RuntimeError("the drop-in is compiled against the reference's headers; %s is not available" % REF)
    _run(["make", "-C", os.path.join(REPO, "voxels_b200", "csrc"), "-f", "Makefile.dropin", "REF=" + REF])


def build_oracle():
    """Test infrastructure: the unmodified reference + the CPU restatement.  Needs the reference checkout for _ref."""
    targets = ["restate"] + (["ref"] if have_reference() else [])
    if not have_reference() and not os.path.exists(os.path.join(REPO, "oracle", "_ref", "libvxh_ref.so")):
        print("note: reference checkout absent and oracle/_ref not prebuilt - only the restatement is built")
    _run(["make", "-C", os.path.join(REPO, "oracle")] + targets + ["REF=" + REF], stdout=subprocess.DEVNULL,
         stderr=subprocess.DEVNULL if have_reference() else None)


def build_all(force=False, verbose=False):
    build_tables(force)
    build_cuda(force, verbose)
    build_dropin(force)
    build_oracle()


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
