#!/usr/bin/env python
"""deploy/deploy.py of the reference on the drop-in models: predict a disparity map for one image pair.

    python tools/deploy.py --net psmnet --path_weight weight_best.pkl --path_left 10L.png --path_right 10R.png [--out dispL.pfm]

Without --path_weight the model keeps its random initialisation (useful only for timing / plumbing checks)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dsmnet_b200 import io  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--net", default="psmnet")
    ap.add_argument("--maxdisparity", default=192, type=int)
    ap.add_argument("--path_weight", default="")
    ap.add_argument("--path_left", default="10L.png")
    ap.add_argument("--path_right", default="10R.png")
    ap.add_argument("--flip", action="store_true", help="predict the right view's disparity (deploy.py:62-66)")
    ap.add_argument("--out", default="dispL.pfm")
    args = ap.parse_args()
    imgL, imgR = io.imread(args.path_left), io.imread(args.path_right)
    model = io.model_create_by_name(args.net, args.maxdisparity)
    if args.path_weight:
        io.load_weights(model, args.path_weight)
    model = model.cuda().eval()
    if args.flip:
        imgL, imgR = np.flip(imgR, axis=1), np.flip(imgL, axis=1)
    disp = io.disp_predict(model, imgL, imgR, use_cuda=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    disp = io.disp_predict(model, imgL, imgR, use_cuda=True)
    torch.cuda.synchronize()
    print("%s %dx%d: %.2f ms (second call, host arrays in -> host array out)" % (args.net, imgL.shape[0], imgL.shape[1], 1e3 * (time.perf_counter() - t0)))
    if args.flip:
        disp = np.flip(disp, axis=-1)
    io.save_pfm(args.out, np.ascontiguousarray(disp, dtype=np.float32))
    print("wrote", args.out, "min %.2f max %.2f" % (disp.min(), disp.max()))


if __name__ == "__main__":
    main()
