#!/usr/bin/env python
"""Build the UNMODIFIED reference (nanoporetech/tombo) hot path into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing in tombo_b200/ imports this.

Recipe:
  * the two Cython sources are compiled *where they lie* under /root/reference
    (tombo/_c_dynamic_programming.pyx, tombo/_c_helper.pyx) with
    language_level=2, cpow=True, language=c++; generated C++ goes to
    oracle/_ref/build, the extension modules to oracle/_ref/tombo/.
  * the pure-Python modules of the path are byte-compiled (sourceless .pyc)
    into oracle/_ref/tombo/ -- build outputs only, no reference source is
    copied into the repository (oracle/_ref/ is git-ignored).

The result travels to the GPU box with the gpurun snapshot and is used as
  - the generator of tests/golden/*.npz   (tests/golden/make_golden.py)
  - cpu_baseline.kind == "reference" / `bench.py --impl reference`.
/root/reference itself is only needed when (re)building.
"""
import os
import py_compile
import shutil
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get('TOMBO_REFERENCE_ROOT', '/root/reference')
OUT = os.path.join(HERE, '_ref')
PKG = os.path.join(OUT, 'tombo')
PY_MODULES = ['__init__', '_version', '_default_parameters', 'tombo_helper',
              'tombo_stats', 'resquiggle']
PYX_MODULES = ['_c_dynamic_programming', '_c_helper']


def ext_suffix():
    return sysconfig.get_config_var('EXT_SUFFIX')


def is_built():
    ok = all(os.path.exists(os.path.join(PKG, m + ext_suffix()))
             for m in PYX_MODULES)
    ok = ok and all(os.path.exists(os.path.join(PKG, m + '.pyc'))
                    for m in PY_MODULES)
    return ok


def build(force=False):
    if is_built() and not force:
        return True
    if not os.path.isdir(os.path.join(REF_ROOT, 'tombo')):
        return False
    import numpy as np
    from setuptools import Extension
    from setuptools.dist import Distribution
    from Cython.Build import cythonize

    os.makedirs(PKG, exist_ok=True)
    build_dir = os.path.join(OUT, 'build')
    os.makedirs(build_dir, exist_ok=True)
    exts = [Extension('tombo.' + m,
                      [os.path.join(REF_ROOT, 'tombo', m + '.pyx')],
                      include_dirs=[np.get_include()], language='c++',
                      define_macros=[('NPY_NO_DEPRECATED_API',
                                      'NPY_1_7_API_VERSION')])
            for m in PYX_MODULES]
    cwd = os.getcwd()
    os.chdir(REF_ROOT)
    try:
        exts = cythonize(
            exts, build_dir=build_dir, language_level=2, quiet=True,
            compiler_directives={'embedsignature': True, 'cpow': True})
    finally:
        os.chdir(cwd)
    dist = Distribution({'name': 'tombo_ref', 'ext_modules': exts})
    cmd = dist.get_command_obj('build_ext')
    cmd.build_lib = OUT
    cmd.build_temp = build_dir
    cmd.ensure_finalized()
    cmd.run()
    for m in PY_MODULES:
        py_compile.compile(os.path.join(REF_ROOT, 'tombo', m + '.py'),
                           cfile=os.path.join(PKG, m + '.pyc'), doraise=True)
    return is_built()


if __name__ == '__main__':
    ok = build(force='--force' in sys.argv)
    print('oracle/_ref built' if ok else 'oracle/_ref NOT built')
    sys.exit(0 if ok else 1)
