"""pip install . -- compiles the CUDA library for sm_100a with nvcc (longtermplanner_b200/_build.py:
-gencode arch=compute_100a,code=sm_100a -fmad=false) before the package is collected, and ships the
public headers next to it so that C / C++ callers of the installed package find the ABI."""
import os
import shutil
import sys

from setuptools import setup
from setuptools.command.build_py import build_py

ROOT = os.path.dirname(os.path.abspath(__file__))


class BuildWithNvcc(build_py):
    def run(self):
        sys.path.insert(0, ROOT)
        from longtermplanner_b200 import _build
        _build.build_library()
        _build.build_host_library()
        dst = os.path.join(ROOT, "longtermplanner_b200", "include")
        shutil.rmtree(dst, ignore_errors=True)
        shutil.copytree(os.path.join(ROOT, "include"), dst)
        super().run()


setup(cmdclass={"build_py": BuildWithNvcc})
