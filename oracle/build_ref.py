"""Recipe for oracle/_ref/: the UNMODIFIED reference renderer, byte-compiled where its sources lie.

TEST INFRASTRUCTURE.  The reference is Python, so "building" it means compiling the modules on the render path
(utils/run_nerf_noscale.py, run_nerf_helpers.py and the two modules they import) from /root/reference/optimization/utils to
CPython bytecode in oracle/_ref/utils/*.pyc.  No reference source is copied or edited; oracle/_ref/ is git-ignored (it stays out
of history) but not gpurun-ignored, so it travels to the GPU box like a built .so, where oracle/ref_import.py imports it
sourceless and bench.py times it as the CPU baseline (`cpu_baseline.kind: "reference"`).  Run in the build container only:
    python oracle/build_ref.py
"""
import os
import py_compile
import sys

SRC = '/root/reference/optimization/utils'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref', 'utils')
MODULES = ['run_nerf_noscale', 'run_nerf_helpers', 'load_LINEMOD_noscale', 'gumble']


def build():
    if not os.path.isdir(SRC):
        return None
    os.makedirs(OUT, exist_ok=True)
    for m in MODULES:
        py_compile.compile(os.path.join(SRC, m + '.py'), cfile=os.path.join(OUT, m + '.pyc'), doraise=True)
    with open(os.path.join(OUT, 'BUILT_FROM'), 'w') as f:
        f.write(f'{SRC} with CPython {sys.version.split()[0]} (py_compile); bytecode only\n')
    return OUT


if __name__ == '__main__':
    print(build() or 'reference tree not present: nothing built')
