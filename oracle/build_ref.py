"""TEST INFRASTRUCTURE ONLY -- recipe that builds ``oracle/_ref``.

Compiles the reference's own ``mmdet/ops/nms/src/nms_cpu.cpp`` (from where it
lies under /root/reference, through ``ref_nms_cpu_wrap.cpp``) into
``oracle/_ref/ref_nms_cpu.so``.  The built .so travels to the GPU box (it is
git-ignored, not gpurun-ignored); the sources never enter this repo.
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
