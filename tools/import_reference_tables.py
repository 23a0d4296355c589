"""Convert the reference's *data tables* (not sources) into the npz files the product ships.

Run once in the build container (needs /root/reference); outputs are committed because
/root/reference does not exist on the GPU box.

  rdWT.txt / idWT.txt  -> tsadar_b200/data/zprime_table.npz
      (reference: tsadar/external/files/{rdWT,idWT}.txt, loaded at form_factor.py:33-34; the
       6-digit table differs from the analytic Z' by up to 1.6e-3 so it must be used verbatim)
  angleWghtsFredfine.mat, angsFRED.mat -> tsadar_b200/data/arts_angles.npz
      (reference: calibration.py:457-458, 487-491)
  tests/test_forward/ThryE-1d.npy -> tests/golden/ThryE-1d.npy   (the only surviving golden vector)
"""
import os, shutil
import numpy as np
import scipy.io as sio

REF = "/root/reference"
HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")

rd = np.loadtxt(os.path.join(REF, "tsadar/external/files/rdWT.txt"))
im = np.loadtxt(os.path.join(REF, "tsadar/external/files/idWT.txt"))
assert rd.shape == (2001, 2) and im.shape == (2001, 2) and np.array_equal(rd[:, 0], im[:, 0])
np.savez_compressed(os.path.join(HERE, "tsadar_b200/data/zprime_table.npz"), x=rd[:, 0], re=rd[:, 1], im=im[:, 1])

w = sio.loadmat(os.path.join(REF, "tsadar/external/files/angleWghtsFredfine.mat"))["weightMatrix"]
a = sio.loadmat(os.path.join(REF, "tsadar/external/files/angsFRED.mat"))["angsFRED"][0, :]
np.savez_compressed(os.path.join(HERE, "tsadar_b200/data/arts_angles.npz"), weightMatrix=w, angsFRED=a)

shutil.copyfile(os.path.join(REF, "tests/test_forward/ThryE-1d.npy"), os.path.join(HERE, "tests/golden/ThryE-1d.npy"))
print("ok")
