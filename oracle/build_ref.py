"""Compiles the REFERENCE's own C++ rANS coder (third_party/CompressAI/cpp_exts/*.cpp over ryg_rans/rans64.h, the only
native code of the reference, module `mcquic.rans`) from the sources where they lie under /root/reference, directly
with g++ (its setup.py is not run).  Output only into oracle/_ref/ (git-ignored, travels to the GPU box).
TEST INFRASTRUCTURE ONLY: the checker for mcquic_b200/csrc/entropy (bit-identical streams) -- never the product.

-DNDEBUG is required: the reference's own calling convention (cdfSizes = k + 2 over a (k+1)-entry CDF,
mcquic/modules/entropyCoder.py:121) trips its debug asserts."""
import glob
import os
import subprocess
import sys
import sysconfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MCQUIC_REFERENCE_ROOT", "/root/reference")
OUT_DIR = os.path.join(ROOT, "oracle", "_ref")


def target():
    return os.path.join(OUT_DIR, "rans" + sysconfig.get_config_var("EXT_SUFFIX"))


def build(force=False):
    src_dir = os.path.join(REF, "third_party", "CompressAI", "cpp_exts")
    if not os.path.isdir(src_dir):
        return None  # reference tree absent (GPU box): use the prebuilt file if it travelled
    if os.path.exists(target()) and not force:
        return target()
    import pybind11
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["g++", "-O3", "-DNDEBUG", "-std=c++17", "-shared", "-fPIC", f"-I{sysconfig.get_paths()['include']}",
           f"-I{pybind11.get_include()}", f"-I{os.path.join(REF, 'third_party', 'CompressAI', 'ryg_rans')}",
           f"-I{src_dir}"] + sorted(glob.glob(os.path.join(src_dir, "*.cpp"))) + ["-o", target()]
    subprocess.run(cmd, check=True)
    return target()


def load():
    """Returns the reference module (RansEncoder, RansDecoder, pmfToQuantizedCDF) or None when unavailable."""
    path = target()
    if not os.path.exists(path):
        try:
            if build() is None:
                return None
        except Exception:
            return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("rans", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
