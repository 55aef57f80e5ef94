"""Build recipe for the CPU oracle (TEST INFRASTRUCTURE ONLY).

* ``build_oracle()``  -> ``oracle/libkrige_oracle.so`` from ``oracle/krige_oracle.c`` (our restatement).
* ``build_ref_envphys()`` -> ``oracle/_ref/libenvphys_ref.so`` from the reference's ``dewpt.c`` + ``iwbt.c`` + ``topotherm.c`` in place.
* ``build_ref()``     -> ``oracle/_ref/libkrige_ref.so`` compiled from the reference's own, unmodified
  ``krige.c / lusolv.c / array.c`` where they lie under ``/root/reference`` (never copied into the repo).
  Flags are the reference's (``setup.py:39-42``: ``-fopenmp -O3``).  Only possible in the build
  container; on the GPU box the prebuilt ``.so`` travels with the snapshot.

Both are plain ``gcc`` invocations -- the reference's own build system is not run.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("SMRF_REFERENCE_ROOT", "/root/reference")
REF_DK = os.path.join(REF_ROOT, "smrf", "spatial", "dk")
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"

ORACLE_SO = os.path.join(HERE, "libkrige_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_SO = os.path.join(REF_DIR, "libkrige_ref.so")
ENV_ORACLE_SO = os.path.join(HERE, "libenvphys_oracle.so")
REF_ENV = os.path.join(REF_ROOT, "smrf", "envphys", "core")
REF_ENV_SO = os.path.join(REF_DIR, "libenvphys_ref.so")


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_oracle(force=False):
    src = os.path.join(HERE, "krige_oracle.c")
    if force or not _newer(ORACLE_SO, [src]):
        cmd = [GCC, "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", ORACLE_SO, src, "-lm"]
        subprocess.check_call(cmd)
    src = os.path.join(HERE, "envphys_oracle.c")
    if force or not _newer(ENV_ORACLE_SO, [src]):
        subprocess.check_call([GCC, "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", ENV_ORACLE_SO, src, "-lm"])
    return ORACLE_SO


def build_ref(force=False):
    """Compile the reference kriging C in place. Returns the .so path or None if unavailable."""
    srcs = [os.path.join(REF_DK, f) for f in ("krige.c", "lusolv.c", "array.c")]
    if not all(os.path.exists(s) for s in srcs):
        return REF_SO if os.path.exists(REF_SO) else None
    if not force and _newer(REF_SO, srcs):
        return REF_SO
    os.makedirs(REF_DIR, exist_ok=True)
    # -w: the reference uses K&R definitions; warnings are not ours to fix.
    cmd = [GCC, "-O3", "-fopenmp", "-w", "-shared", "-fPIC", "-I", REF_DK, "-o", REF_SO] + srcs + ["-lm"]
    subprocess.check_call(cmd)
    return REF_SO


def build_ref_envphys(force=False):
    """Compile the reference's dewpt.c + iwbt.c + topotherm.c (saturation curves) in place. Returns the .so path or None."""
    srcs = [os.path.join(REF_ENV, f) for f in ("dewpt.c", "iwbt.c", "topotherm.c")]
    if not all(os.path.exists(s) for s in srcs):
        return REF_ENV_SO if os.path.exists(REF_ENV_SO) else None
    if not force and _newer(REF_ENV_SO, srcs):
        return REF_ENV_SO
    os.makedirs(REF_DIR, exist_ok=True)
    cmd = [GCC, "-O3", "-fopenmp", "-w", "-shared", "-fPIC", "-I", REF_ENV, "-o", REF_ENV_SO] + srcs + ["-lm"]
    subprocess.check_call(cmd)
    return REF_ENV_SO


if __name__ == "__main__":
    print("oracle:", build_oracle(force=True))
    print("ref   :", build_ref(force=True))
    print("ref envphys:", build_ref_envphys(force=True))
