"""Packaging.  The sm_100a kernels and the C++ host extension are built IN-TREE by `realhf_b200/ops/build.py` (nvcc / g++ called
directly, `-gencode arch=compute_100a,code=sm_100a`); `build_ext` here just runs that script so that

    pip install -e . --no-build-isolation        # or: python setup.py build_ext --inplace

leaves `realhf_b200/_C/*.so` next to the sources (a source checkout on PYTHONPATH works the same way without installing).
"""
import os
import re
import sys

from setuptools import Command, find_packages, setup
from setuptools.command.build_ext import build_ext as _build_ext
from setuptools.command.develop import develop as _develop

ROOT = os.path.dirname(os.path.abspath(__file__))


def _version():
    text = open(os.path.join(ROOT, "realhf_b200", "__init__.py")).read()
    return re.search(r'__version__ = "([^"]+)"', text).group(1)


def _build_kernels(force=False):
    sys.path.insert(0, ROOT)
    from realhf_b200.ops import build
    print("built:", *build.build_all(force=force))


class build_ext(_build_ext):
    def run(self):
        _build_kernels()


class develop(_develop):
    def run(self):
        _build_kernels()
        super().run()


class build_kernels(Command):
    description = "compile the sm_100a kernels and the C++ host extension in-tree"
    user_options = [("force", "f", "rebuild everything")]

    def initialize_options(self):
        self.force = False

    def finalize_options(self):
        self.force = bool(self.force)

    def run(self):
        _build_kernels(self.force)


setup(
    name="realhf_b200",
    version=_version(),
    description="B200-native (sm_100a) RLHF training framework with the capabilities of ReaLHF",
    long_description=open(os.path.join(ROOT, "README.md")).read(),
    long_description_content_type="text/markdown",
    python_requires=">=3.10",
    packages=find_packages(include=["realhf_b200", "realhf_b200.*"]),
    package_data={"realhf_b200": ["_C/*.so", "ops/csrc/*.cu", "ops/csrc/*.cuh", "ops/csrc/*.cpp", "ops/csrc/host/*", "search/tables/*.json"]},
    install_requires=["torch>=2.5", "numpy", "networkx", "pyzmq", "psutil", "transformers", "pybind11"],
    extras_require={"attention-lib": ["flash-attn"], "logging": ["tensorboard", "wandb"]},
    entry_points={"console_scripts": ["realhf-b200=realhf_b200.apps.quickstart:main", "realhf-b200-ctl=realhf_b200.apps.main:main"]},
    cmdclass={"build_ext": build_ext, "develop": develop, "build_kernels": build_kernels},
)
