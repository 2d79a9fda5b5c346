import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
MODEL_DIR = os.environ.get('MH_TEST_MODEL_DIR', '/tmp/mh_test_model')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def model_dir():
    """Synthetic SMPL-shaped model dir (seed 0 -- the seed the golden vectors were made with)."""
    from oracle import synth
    if not os.path.exists(os.path.join(MODEL_DIR, 'SMPL_NEUTRAL.pkl')):
        synth.write_model_dir(MODEL_DIR, seed=0)
    return MODEL_DIR


@pytest.fixture(scope='session')
def model(model_dir):
    from oracle import synth
    return synth.load_model_tensors(model_dir)


COEFS = dict(proj2d=1.0, depth=0.05, silhouette=0.1, reg_velocity=0.05, reg_verts_filter=0.002,
             reg_poses=0.002, reg_scales=1e-4, reg_contact=0.001, reg_foot_sliding=0.01)
