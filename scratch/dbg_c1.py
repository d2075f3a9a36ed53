import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import numpy as np
import qampy_b200.equalisation as eq
g = np.load('tests/golden/g1_c1_cma.npz')
rms = lambda a: float(np.sqrt(np.mean(np.abs(a) ** 2)))
for k in ('default', 'warp', '16'):
    os.environ.pop('QB_TRAIN_KERNEL', None); os.environ.pop('QB_TRAIN_LPS', None)
    if k == 'warp': os.environ['QB_TRAIN_KERNEL'] = 'warp'
    if k == '16': os.environ['QB_TRAIN_LPS'] = '16'
    E, wxy, err = eq.equalise_signal(g["E_in"], 2, 1e-3, 4, Ntaps=11, method="cma", apply=True)
    print(k, 'E', rms(E - g["E_out"]), 'err', rms(err - g["err"]), 'w', np.max(np.abs(wxy - g["wxy"])))
    print('  err first diff idx', np.argwhere(np.abs(err - g['err']) > 1e-4)[:5].tolist(), np.isnan(err).sum())
    print('  w', wxy[0,0,:4], g['wxy'][0,0,:4])
