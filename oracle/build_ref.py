"""TEST INFRASTRUCTURE ONLY -- recipe that builds ``oracle/_ref``.

Compiles the reference's own ``mmdet/ops/nms/src/nms_cpu.cpp`` (from where it
lies under /root/reference, through ``ref_nms_cpu_wrap.cpp``) into
``oracle/_ref/ref_nms_cpu.so``, and its ``mmdet/ops/nms/src/soft_nms_cpu.pyx``
(Cython -> C++ -> g++, intermediate files only under ``oracle/_ref``) into
``oracle/_ref/soft_nms_cpu*.so``.  The built .so files travel to the GPU box (they
are git-ignored, not gpurun-ignored); the sources never enter this repo.
"""
import glob
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_SRC = os.path.join(os.environ.get("IOU_REFERENCE_ROOT", "/root/reference"),
                       "mmdet", "ops", "nms", "src", "nms_cpu.cpp")
NAME = "ref_nms_cpu"


def _prebuilt():
    hits = glob.glob(os.path.join(OUT, NAME + "*.so"))
    return hits[0] if hits else None


def build(verbose=False):
    """Build oracle/_ref/ref_nms_cpu.so if the reference source is present."""
    if _prebuilt():
        return _prebuilt()
    if not os.path.isfile(REF_SRC):
        return None
    from torch.utils.cpp_extension import load
    os.makedirs(OUT, exist_ok=True)
    load(name=NAME, sources=[os.path.join(HERE, "ref_nms_cpu_wrap.cpp")],
         extra_cflags=["-O2", "-w", '-DIOU_REFERENCE_NMS_CPU_SOURCE="\\"%s\\""' % REF_SRC],
         build_directory=OUT, verbose=verbose)
    return _prebuilt()


REF_SOFT_SRC = os.path.join(os.path.dirname(REF_SRC), "soft_nms_cpu.pyx")
SOFT_NAME = "soft_nms_cpu"          # the module init symbol Cython derives from the .pyx file name


def _prebuilt_soft():
    hits = glob.glob(os.path.join(OUT, SOFT_NAME + ".*.so")) + glob.glob(os.path.join(OUT, SOFT_NAME + ".so"))
    return hits[0] if hits else None


def build_soft_nms(verbose=False):
    """Build oracle/_ref/soft_nms_cpu*.so from the reference's .pyx if it is present (needs Cython + g++)."""
    if _prebuilt_soft():
        return _prebuilt_soft()
    if not os.path.isfile(REF_SOFT_SRC):
        return None
    import subprocess
    import sysconfig
    import numpy
    os.makedirs(OUT, exist_ok=True)
    cpp = os.path.join(OUT, "ref_soft_nms_cpu.cpp")
    subprocess.check_call([sys.executable, "-m", "cython", "--cplus", "-3", REF_SOFT_SRC, "-o", cpp])
    so = os.path.join(OUT, SOFT_NAME + sysconfig.get_config_var("EXT_SUFFIX"))
    cmd = ["g++", "-O2", "-w", "-shared", "-fPIC", "-I" + numpy.get_include(),
           "-I" + sysconfig.get_paths()["include"], cpp, "-o", so]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return so


def load_ref_soft_nms_cpu():
    """Return the compiled reference Cython module (exposes ``soft_nms_cpu(boxes, iou_thr, method, sigma, min_score)``)."""
    path = build_soft_nms()
    if path is None:
        raise RuntimeError("oracle/_ref/soft_nms_cpu not built and the reference .pyx is absent")
    spec = importlib.util.spec_from_file_location(SOFT_NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_ref_nms_cpu():
    """Return the compiled reference module (exposes ``nms(dets, thr)``)."""
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    path = build()
    if path is None:
        raise RuntimeError("oracle/_ref not built and reference sources absent")
    spec = importlib.util.spec_from_file_location(NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
    print(build_soft_nms(verbose="-v" in sys.argv))
