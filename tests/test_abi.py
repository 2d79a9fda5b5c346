"""The C-ABI library loads on a machine without a GPU and exports every symbol include/mhopt.h declares; the host-side
package refuses to run without CUDA (no CPU fallback).  CPU only."""
import ctypes
import os
import re

import numpy as np
import pytest

import __graft_entry__ as ge

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'mhopt.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(mh_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_are_exported():
    lib = ctypes.CDLL(os.path.join(ge.PKG_DIR, 'libmhopt.so'))
    names = declared_symbols()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_covers_the_header():
    pkg = ge.load_package()
    import sys
    L = sys.modules[pkg.__name__ + '._lib']
    assert sorted(L.SYMBOLS) == declared_symbols()
    assert L.lib.mh_version().decode().startswith('mhopt-b200')


def test_struct_layouts_match_the_header():
    pkg = ge.load_package()
    import sys
    L = sys.modules[pkg.__name__ + '._lib']
    assert ctypes.sizeof(L.MhDims) == 12 * 4 + 8            # 12 int32 + int64 (8-byte aligned)
    assert ctypes.sizeof(L.MhCoefs) == 11 * 4
    assert ctypes.sizeof(L.MhModel) == 8 * ctypes.sizeof(ctypes.c_void_p)


def test_no_cpu_fallback(model_dir):
    import torch
    if torch.cuda.is_available():
        pytest.skip('a CUDA device is present')
    pkg = ge.load_package()
    with pytest.raises(RuntimeError):
        pkg.SMPLDepthSequenceOptimizer(image_size=(96, 64), num_frames=4, cam_K=np.eye(3, dtype=np.float32),
                                       smpl_model_parameters_path=model_dir)
    with pytest.raises(RuntimeError):
        pkg.SMPLDepthSequenceOptimizer(image_size=(96, 64), num_frames=4, cam_K=np.eye(3, dtype=np.float32), device='cpu',
                                       smpl_model_parameters_path=model_dir)


def test_product_does_not_import_the_oracle():
    # nothing under the package may import / execute oracle/ (the oracle is test infrastructure)
    for fn in os.listdir(ge.PKG_DIR):
        if fn.endswith('.py'):
            src = open(os.path.join(ge.PKG_DIR, fn)).read()
            assert not re.search(r'^\s*(from|import)\s+oracle', src, flags=re.M), fn
