"""SMPL model files -> the float32 arrays ``libmhopt.so`` consumes.

Mirrors what the reference's ``SMPL.__init__`` registers (``mhmocap/smpl.py:179-275``): ``SMPL_NEUTRAL.pkl``
with keys ``v_template, f, shapedirs, posedirs, J_regressor, kintree_table, weights``; shapedirs cut to 10
betas (``:229``); posedirs reshaped to (207, 3V) (``:261-265``); ``parents[0] = -1`` (``:268-270``); the
AlphaPose / MuPoTS regressors are stored (V, 17) on disk and used transposed (``:249-259``).
"""
import os
import pickle

import numpy as np

REGRESSOR_ALPHAPOSE = 'SMPL_AlphaPose_Regressor_RMSprop_6.npy'
REGRESSOR_MUPOTS = 'SMPL_MuPoTs_Regressor_v1.npy'
REGRESSOR_H36M = 'J_regressor_h36m.npy'


def _dense(a):
    if hasattr(a, 'todense'):
        a = a.todense()
    if hasattr(a, 'r'):            # chumpy array
        a = a.r
    return np.array(a, dtype=np.float32)


_CACHE = {}


def load_smpl_model(model_path, alphapose_regressor=REGRESSOR_ALPHAPOSE, mupots_regressor=REGRESSOR_MUPOTS, gender='neutral',
                    h36m_regressor=REGRESSOR_H36M):
    """``model_path``: directory holding ``SMPL_<GENDER>.pkl`` (or the pickle itself) and the regressor ``.npy`` files.
    The converted arrays are cached per process (keyed by the files' paths, sizes and modification times): a process that fits many
    sequences -- one optimiser per sequence, ``predict.py:290-306`` -- reads and converts the model once.  Read-only: do not modify."""
    if os.path.isdir(model_path):
        pkl = os.path.join(model_path, 'SMPL_{}.pkl'.format(gender.upper()))
        base = model_path
    else:
        pkl = model_path
        base = os.path.dirname(model_path)
    if not os.path.exists(pkl):
        raise FileNotFoundError('Path {} does not exist!'.format(pkl))
    ap = alphapose_regressor if os.path.isabs(alphapose_regressor) else os.path.join(base, alphapose_regressor)
    mp = mupots_regressor if os.path.isabs(mupots_regressor) else os.path.join(base, mupots_regressor)
    hp = h36m_regressor if os.path.isabs(h36m_regressor) else os.path.join(base, h36m_regressor)
    key = tuple((os.path.abspath(f), os.path.getsize(f), os.path.getmtime(f)) for f in (pkl, ap, mp, hp) if os.path.exists(f))
    if key in _CACHE:
        return _CACHE[key]
    with open(pkl, 'rb') as f:
        d = pickle.load(f, encoding='latin1')
    posedirs = _dense(d['posedirs'])
    parents = np.array(d['kintree_table'][0]).astype(np.int64)
    parents[0] = -1
    model = {
        'v_template': _dense(d['v_template']),
        'faces': np.array(d['f']).astype(np.int32),
        'shapedirs': np.ascontiguousarray(_dense(d['shapedirs'])[:, :, :10]),
        'posedirs': np.ascontiguousarray(posedirs.reshape(-1, posedirs.shape[-1]).T),
        'J_regressor': _dense(d['J_regressor']),
        'lbs_weights': _dense(d['weights']),
        'parents': parents.astype(np.int32),
    }
    model['J_regressor_alphapose'] = np.ascontiguousarray(np.load(ap).T.astype(np.float32))
    if os.path.exists(mp):
        model['J_regressor_mupots'] = np.ascontiguousarray(np.load(mp).T.astype(np.float32))
    if os.path.exists(hp):
        # stored (17, V) in H36M joint order; the reference layer re-orders the rows to its 17-joint layout (smpl.py:240-242)
        h36m_rows = [6, 5, 4, 1, 2, 3, 16, 15, 14, 11, 12, 13, 8, 10, 0, 7, 9]
        model['J_regressor_h36m17'] = np.ascontiguousarray(np.load(hp)[h36m_rows].astype(np.float32))
    for v in model.values():
        v.setflags(write=False)
    _CACHE[key] = model
    return model
