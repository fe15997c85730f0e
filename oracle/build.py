"""TEST INFRASTRUCTURE ONLY.  Builds oracle/_build/libnsv_oracle.so from oracle/nsv_oracle.c and,
when /root/reference is present, oracle/_ref/libnesvor_ref_cpu.so via oracle/build_ref.sh."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libnsv_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libnesvor_ref_cpu.so")
# not $CC: this image presets it to a wrapper that cannot find libgomp.spec
GCC = os.environ.get("NSV_CC", "/usr/bin/gcc")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_oracle(force=False):
    srcs = [os.path.join(HERE, f) for f in ("nsv_oracle.c", "slice_acq_oracle_impl.h", "transform_oracle_impl.h")]
    if force or _stale(ORACLE_SO, srcs):
        os.makedirs(os.path.dirname(ORACLE_SO), exist_ok=True)
        cmd = [GCC, "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-std=gnu11",
               "-o", ORACLE_SO, srcs[0], "-lm"]
        subprocess.check_call(cmd)
    return ORACLE_SO


def build_ref(force=False):
    """Returns the path of the CPU build of the reference kernels, or None if it cannot exist."""
    ref_root = os.environ.get("NSV_REFERENCE_ROOT", "/root/reference")
    srcs = [os.path.join(HERE, f) for f in ("build_ref.sh", "ref_cpu_shim.h", "ref_driver_slice_acq.inc", "ref_driver_transform.inc")]
    if os.path.isdir(ref_root) and (force or _stale(REF_SO, srcs)):
        subprocess.check_call(["bash", os.path.join(HERE, "build_ref.sh")])
    return REF_SO if os.path.exists(REF_SO) else None


REF_GPU_SO = os.path.join(HERE, "_ref", "nesvor_ref_slice_acq_cuda.so")


def build_ref_gpu(force=False):
    """The reference's own CUDA extensions (slice acquisition, pose converters) for sm_100a (oracle/build_ref_gpu.sh, ~3 min
    each with torch headers); returns the slice-acquisition module's path, or None when it cannot exist (no /root/reference and no prebuilt file).  A failed build is
    reported and tolerated: the GPU cross-check that uses it is skipped, nothing else depends on it."""
    ref_root = os.environ.get("NSV_REFERENCE_ROOT", "/root/reference")
    script = os.path.join(HERE, "build_ref_gpu.sh")
    both = [REF_GPU_SO, os.path.join(HERE, "_ref", "nesvor_ref_transform_convert_cuda.so")]
    if os.path.isdir(ref_root) and (force or any(_stale(so, [script]) for so in both)):
        try:
            subprocess.check_call(["bash", script])
        except subprocess.CalledProcessError as e:
            print(f"oracle.build: reference CUDA extension not built ({e}); its GPU cross-check will be skipped")
    return REF_GPU_SO if os.path.exists(REF_GPU_SO) else None


REF_PKG = os.path.join(os.path.dirname(HERE), "baseline", "_ref", "nesvor")
REF_TESTS = os.path.join(os.path.dirname(HERE), "baseline", "_ref", "tests")


def install_reference_package(force=False):
    """The UNMODIFIED pure-Python layer of the reference (every nesvor/**/*.py, nothing else) placed under the git-ignored
    baseline/_ref/ so that it travels to the GPU box, where tests run the reference's own INR / NeSVoR on this library
    through nesvor_b200.compat (its three native imports), and the reference's own unit tests (tests/**/*.py) next to it.  `pip install --target baseline/_ref /root/reference` cannot be
    used: setup.py builds the two CUDA extensions, which do not compile against torch 2.11 as shipped, and
    the hot path imports tinycudann.  Returns the package path or None."""
    import shutil

    ref_root = os.environ.get("NSV_REFERENCE_ROOT", "/root/reference")
    only_py = lambda d, names: [n for n in names if not (n.endswith(".py") or os.path.isdir(os.path.join(d, n)))]  # noqa: E731
    for sub, dst in (("nesvor", REF_PKG), ("tests", REF_TESTS)):  # the package and the reference's own unit tests
        src = os.path.join(ref_root, sub)
        if os.path.isdir(src) and (force or not os.path.isdir(dst)):
            if os.path.isdir(dst):
                shutil.rmtree(dst)
            shutil.copytree(src, dst, ignore=only_py)
    return REF_PKG if os.path.isdir(REF_PKG) else None


if __name__ == "__main__":
    print(build_oracle(force=True))
    print(build_ref(force=True))
    print(build_ref_gpu(force=True))
    print(install_reference_package(force=True))
