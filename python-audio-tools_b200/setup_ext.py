"""Builds the CPython 3 extension modules audiotools.pcm and audiotools.encoders in-tree
(python setup_ext.py build_ext --inplace); encoders links against ../libb200flac.so."""
import os

from setuptools import Extension, setup

HERE = os.path.dirname(os.path.abspath(__file__))
os.chdir(HERE)

setup(
    name="audiotools_b200_ext",
    version="0.1",
    packages=["audiotools"],
    ext_modules=[
        Extension("audiotools.pcm", ["audiotools/pcm.c"], extra_compile_args=["-O3", "-std=c11"]),
        Extension("audiotools.encoders", ["audiotools/encoders.c"],
                  include_dirs=[os.path.join(HERE, "..", "include")],
                  library_dirs=[HERE], libraries=["b200flac"],
                  runtime_library_dirs=["$ORIGIN/.."],
                  extra_compile_args=["-O3", "-std=c11"]),
    ],
)
