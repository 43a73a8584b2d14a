"""The reference's only real-data fixture — deploy/10L.png, deploy/10R.png, a 375x1242 KITTI pair — as a
compressed array fixture (uint8 RGB, cropped to 372x1240 so that a 4x4 space-to-depth gives 93x310 feature maps):

    python tests/golden/make_golden_realpair.py        # needs /root/reference (or $DSMNET_REFERENCE)

Read exactly as the reference reads it (myDatasets_stereo/img_rw.py:28-34: cv2.imread + BGR->RGB flip == PIL RGB)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dsmnet_b200.io import imread      # noqa: E402

REF = os.environ.get("DSMNET_REFERENCE", "/root/reference")
L = imread(os.path.join(REF, "deploy", "10L.png")); R = imread(os.path.join(REF, "deploy", "10R.png"))
assert L.shape == R.shape == (375, 1242, 3) and L.dtype == np.uint8
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kitti_pair.npz")
np.savez_compressed(out, L=L[:372, :1240], R=R[:372, :1240])
print("wrote", out, os.path.getsize(out) // 1024, "KB")
