"""Build + install: compiles libfa_b200.so for sm_100a (flash-attention-v100_b200/csrc/build.sh: nvcc only, no torch
headers -- the C ABI has no torch types) and lays the Python side out the way the reference's install does
(reference setup.py:100-158): packages `flash_attn_v100` and `flash_attn`, modules `flash_attn_v100_cuda` and
`flash_attn_2_cuda`, plus `flash_attn-2.8.3` metadata so `importlib.metadata.version("flash-attn")` answers 2.8.3.

    pip install --no-build-isolation .          # or: python setup.py bdist_wheel
    FA_B200_SKIP_BUILD=1 pip install ...        # reuse an already built flash-attention-v100_b200/lib/libfa_b200.so
"""
import os
import shutil
import subprocess

from setuptools import setup
from setuptools.command.build_py import build_py

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "flash-attention-v100_b200")
IMPERSONATED = ("flash-attn", "2.8.3")  # what the reference's flash_attn/ package reports (flash_attn/__init__.py:15)


class BuildWithNativeLibrary(build_py):
    def run(self):
        lib = os.path.join(PKG, "lib", "libfa_b200.so")
        if not (os.environ.get("FA_B200_SKIP_BUILD") and os.path.exists(lib)):
            subprocess.run(["bash", os.path.join(PKG, "csrc", "build.sh")], check=True)
        super().run()
        # the shared library travels inside the flash_attn_v100 package (flash_attn_v100_cuda.load_library looks there)
        dst = os.path.join(self.build_lib, "flash_attn_v100", "lib")
        os.makedirs(dst, exist_ok=True)
        shutil.copy2(lib, os.path.join(dst, "libfa_b200.so"))
        # the extension module under its upstream name (the reference symlinks its .so to it, setup.py:149-158)
        shutil.copy2(os.path.join(PKG, "shim", "flash_attn_2_cuda.py"), os.path.join(self.build_lib, "flash_attn_2_cuda.py"))
        # upstream's metadata name, so importlib.metadata.version("flash-attn") answers 2.8.3 like after the reference's
        # InstallAttention step (setup.py:114-124, which writes a flash_attn-2.8.3.dist-info by hand). A wheel may hold
        # only one .dist-info, so the second identity ships in the legacy .egg-info form importlib.metadata also reads.
        info = os.path.join(self.build_lib, "flash_attn-%s.egg-info" % IMPERSONATED[1])
        os.makedirs(info, exist_ok=True)
        with open(os.path.join(info, "PKG-INFO"), "w") as f:
            f.write("Metadata-Version: 2.1\nName: %s\nVersion: %s\nSummary: served by flash_attn_v100 (B200 build)\n" % IMPERSONATED)
        with open(os.path.join(info, "top_level.txt"), "w") as f:
            f.write("flash_attn\nflash_attn_2_cuda\n")


setup(
    packages=["flash_attn_v100", "flash_attn"],
    package_dir={
        "flash_attn_v100": "flash-attention-v100_b200/flash_attn_v100",
        "flash_attn": "flash-attention-v100_b200/shim/flash_attn",
        "": "flash-attention-v100_b200",
    },
    py_modules=["flash_attn_v100_cuda", "sharding"],
    data_files=[],
    cmdclass={"build_py": BuildWithNativeLibrary},
    zip_safe=False,
)
