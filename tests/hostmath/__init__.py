"""ctypes loader for the host build of ``csrc/mh_math.cuh`` (CPU test helper only)."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, 'libhostmath.so')


def load():
    src = os.path.join(HERE, 'hostmath.cpp')
    hdr = os.path.join(HERE, '..', '..', 'scene-aware-3d-multi-human_b200', 'csrc', 'mh_math.cuh')
    if (not os.path.exists(SO)) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(['g++', '-O2', '-ffp-contract=off', '-shared', '-fPIC', '-x', 'c++', src, '-o', SO])
    return ctypes.CDLL(SO)
