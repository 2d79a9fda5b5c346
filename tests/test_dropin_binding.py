"""The unmodified reference driver binds the B200 optimiser class (runs only where /root/reference is mounted). CPU only."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'

CODE = r'''
import sys
sys.argv = ['predict']
sys.path.insert(0, %r)
import __graft_entry__ as ge
pkg = ge.load_package()
pkg.install_as_mhmocap_optimizer()
sys.path.insert(0, %r)
import mhmocap.predict as rp
assert rp.SMPLDepthSequenceOptimizer is pkg.SMPLDepthSequenceOptimizer
import inspect
ours = inspect.signature(pkg.SMPLDepthSequenceOptimizer.__init__).parameters
# every keyword Predictor passes (predict.py:290-306) is accepted
src = inspect.getsource(rp.Predictor.__init__)
import re
kws = re.findall(r'^\s+(\w+)=', src[src.index('SMPLDepthSequenceOptimizer('):], flags=re.M)
base = inspect.signature(pkg.SMPLOptimizerBase.__init__).parameters
missing = [k for k in kws if k not in ours and k not in base]
assert not missing, missing
for name in ('init_optimized_variables', 'fit', 'get_optimized_variables', 'update_scene_pointcloud', 'one_euro_filter'):
    assert hasattr(pkg.SMPLDepthSequenceOptimizer, name)
print('BOUND', len(kws))
'''


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference not mounted')
def test_reference_predictor_binds_our_optimizer():
    out = subprocess.run([sys.executable, '-c', CODE % (ROOT, REF)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert 'BOUND' in out.stdout
