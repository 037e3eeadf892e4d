#!/usr/bin/env python
"""Place the UNMODIFIED reference package under ``baseline/_ref/`` so that it travels to the GPU box.

Test / measurement infrastructure (see oracle/__init__.py).  The task's recipe is
``pip install --no-index --no-build-isolation --target baseline/_ref /root/reference``; in this image that fails
because the reference's build backend (``uv_build``, pyproject.toml:1-3) is not installed and there is no network.
The package is pure Python, so a wheel install is nothing but a copy of ``src/cobel`` next to its metadata -- which
is what this script does, byte for byte, from ``/root/reference`` (read-only) into ``baseline/_ref/`` (git-ignored,
NOT gpurun-ignored).  Nothing under ``baseline/_ref`` is part of the product or of the repository's history; it is
imported only by ``bench.py --impl reference`` / the ``cpu_baseline`` leg (through oracle/ref_loader.py, behind the
GUI / gym stand-ins of oracle/_stubs) and by the CPU tests that pin the oracle.

Usage: python oracle/install_reference.py [reference root, default /root/reference]
"""
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TARGET = os.path.join(ROOT, 'baseline', '_ref')


def install(ref_root='/root/reference'):
    src = os.path.join(ref_root, 'src', 'cobel')
    if not os.path.isdir(src):
        return False
    dst = os.path.join(TARGET, 'cobel')
    if os.path.isdir(dst):
        cmp = filecmp.dircmp(src, dst, ignore=['__pycache__'])
        if not (cmp.left_only or cmp.diff_files):
            return True
        shutil.rmtree(dst)
    os.makedirs(TARGET, exist_ok=True)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    with open(os.path.join(TARGET, 'INSTALLED_FROM'), 'w') as f:
        f.write('%s (unmodified copy of src/cobel; pip install failed: build backend uv_build unavailable)\n' % ref_root)
    return True


if __name__ == '__main__':
    ok = install(sys.argv[1] if len(sys.argv) > 1 else '/root/reference')
    print('baseline/_ref: %s' % ('installed' if ok else 'reference sources not found'))
