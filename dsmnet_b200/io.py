"""Checkpoint and image I/O of the reference, for running the drop-in models on real data.

  * ``load_weights(model, path)`` / ``save_weights`` — the reference's checkpoint format (stereo.py:61-62,81-83, deploy.py:51-53:
    ``torch.save({'state_dict': model.state_dict()}, 'weight_best.pkl')``; full checkpoints carry 'epoch', 'best_prec', 'optim' too).
    The drop-in models keep the reference's parameter names, so loading is strict.
  * ``imread`` / ``load_pfm`` / ``save_pfm`` — myDatasets_stereo/img_rw.py:28-41, img_rw_pfm.py:13-75 (PNG/JPEG as RGB uint8,
    PFM as float32, rows flipped to top-down), without the cv2 dependency.
  * ``model_create_by_name`` — models/__init__.py:6-34 for the four models that have drop-ins.
  * ``disp_predict`` — deploy/deploy.py:15-32: uint8 RGB pair -> ImageNet-normalised tensors -> model(imgL, imgR, "test") -> disparity.
"""
from __future__ import annotations

import re
import struct
import zlib

import numpy as np
import torch

IMAGENET_MEAN = (0.485, 0.456, 0.406)      # myTransforms/__init__.py:8
IMAGENET_STD = (0.229, 0.224, 0.225)


def load_pfm(fname):
    """(image float32 top-down, scale) — img_rw_pfm.py:13-48."""
    with open(fname, "rb") as f:
        header = f.readline().decode("ascii").rstrip()
        if header not in ("PF", "Pf"):
            raise ValueError("Not a PFM file.")
        m = re.match(r"^(\d+)\s(\d+)\s$", f.readline().decode("ascii"))
        if not m:
            raise ValueError("Malformed PFM header.")
        width, height = int(m.group(1)), int(m.group(2))
        scale = float(f.readline().decode("ascii").rstrip())
        endian = "<" if scale < 0 else ">"
        data = np.frombuffer(f.read(), dtype=endian + "f4")
    shape = (height, width, 3) if header == "PF" else (height, width)
    return np.flipud(data.reshape(shape)).astype(np.float32).copy(), abs(scale)


def save_pfm(fname, image, scale=1):
    """img_rw_pfm.py:50-75 (little-endian float32, bottom-up rows)."""
    image = np.asarray(image)
    if image.dtype != np.float32:
        raise ValueError("Image dtype must be float32.")
    if image.ndim == 3 and image.shape[2] == 3:
        color = True
    elif image.ndim == 2 or (image.ndim == 3 and image.shape[2] == 1):
        color = False
    else:
        raise ValueError("Image must have H x W x 3, H x W x 1 or H x W dimensions.")
    with open(fname, "wb") as f:
        f.write(b"PF\n" if color else b"Pf\n")
        f.write(("%d %d\n" % (image.shape[1], image.shape[0])).encode())
        f.write(("%f\n" % -abs(scale)).encode())
        f.write(np.flipud(image).astype("<f4").tobytes())


def _png_decode(buf: bytes) -> np.ndarray:
    """Minimal PNG decoder (non-interlaced, 8/16-bit grey / RGB / RGBA): zlib + the five row filters."""
    if buf[:8] != b"\x89PNG\r\n\x1a\n":
        raise ValueError("not a PNG file")
    pos, idat, ihdr = 8, [], None
    while pos < len(buf):
        n, typ = struct.unpack(">I4s", buf[pos:pos + 8])
        body = buf[pos + 8:pos + 8 + n]
        if typ == b"IHDR":
            ihdr = struct.unpack(">IIBBBBB", body)
        elif typ == b"IDAT":
            idat.append(body)
        elif typ == b"IEND":
            break
        pos += 12 + n
    w, h, depth, ctype, _, _, interlace = ihdr
    if interlace or depth not in (8, 16) or ctype not in (0, 2, 4, 6):
        raise ValueError("unsupported PNG variant")
    ch = {0: 1, 2: 3, 4: 2, 6: 4}[ctype]
    bpp = ch * depth // 8
    raw = np.frombuffer(zlib.decompress(b"".join(idat)), dtype=np.uint8).reshape(h, 1 + w * bpp)
    out = np.zeros((h, w * bpp), dtype=np.uint8)
    prev = np.zeros(w * bpp, dtype=np.int32)
    for y in range(h):
        ft, line = int(raw[y, 0]), raw[y, 1:].astype(np.int32)
        if ft == 0:
            cur = line
        elif ft == 2:
            cur = (line + prev) & 255
        else:                                         # Sub / Average / Paeth recur along the row: one step per pixel, vectorised over bpp
            cur = np.zeros_like(line)
            left = np.zeros(bpp, dtype=np.int32); ul = np.zeros(bpp, dtype=np.int32)
            for x in range(0, w * bpp, bpp):
                up = prev[x:x + bpp]
                if ft == 1:
                    pred = left
                elif ft == 3:
                    pred = (left + up) >> 1
                else:
                    p = left + up - ul
                    pa, pb, pc = np.abs(p - left), np.abs(p - up), np.abs(p - ul)
                    pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, up, ul))
                left = (line[x:x + bpp] + pred) & 255
                cur[x:x + bpp] = left
                ul = up
        out[y] = cur
        prev = cur
    if depth == 16:
        out = out.reshape(h, w * ch, 2)
        img = (out[..., 0].astype(np.uint16) << 8) | out[..., 1]
    else:
        img = out
    img = img.reshape(h, w, ch)
    return img[..., 0] if ch == 1 else img


def imread(fname):
    """img_rw.py:28-34: PFM -> float32 array; anything else -> RGB uint8 [H, W, 3] (uint16 for 16-bit PNGs)."""
    if str(fname).find(".pfm") > 0:
        return load_pfm(fname)[0]
    try:
        from PIL import Image
        with Image.open(fname) as im:
            if im.mode in ("I;16", "I;16B", "I"):
                return np.array(im)
            return np.array(im.convert("RGB"))
    except ImportError:
        img = _png_decode(open(fname, "rb").read())
        if img.ndim == 3 and img.shape[2] == 4:
            img = img[..., :3]
        if img.ndim == 2 and img.dtype == np.uint8:
            img = np.repeat(img[..., None], 3, axis=2)      # cv2.imread returns three channels for grey files
        return img


def load_disp(fname):
    """img_rw.py:12-23: first channel, inf / nan -> 0."""
    g = np.array(imread(fname), dtype=np.float32)
    if g.ndim > 2:
        g = g[:, :, 0]
    g[~np.isfinite(g)] = 0
    return g


def load_weights(model: torch.nn.Module, path: str, map_location="cpu", strict: bool = True):
    """stereo.py:61-62 / deploy.py:51-53.  Accepts {'state_dict': ...} files (weight_best.pkl, model_checkpoint.pkl) and bare
    state dicts; strips the 'module.' prefix nn.DataParallel adds."""
    obj = torch.load(path, map_location=map_location, weights_only=False)
    sd = obj["state_dict"] if isinstance(obj, dict) and "state_dict" in obj else obj
    sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
    return model.load_state_dict(sd, strict=strict)


def save_weights(model: torch.nn.Module, path: str):
    """stereo.py:80-83: torch.save({'state_dict': model.state_dict()}, path)."""
    torch.save({"state_dict": model.state_dict()}, path)


def model_create_by_name(name_model: str, maxdisparity: int = 192):
    """models/__init__.py:6-34 for the models with sm_100a drop-ins."""
    if name_model == "dispnetcorr":
        from .dispnetcorr import dispnetcorr
        return dispnetcorr(maxdisparity)
    if name_model == "iresnet":
        from .iresnet import iresnet
        return iresnet(maxdisparity)
    if name_model == "gcnet":
        from .gcnet import gcnet
        return gcnet(maxdisparity)
    if name_model == "psmnet":
        from .psmnet import PSMNet
        return PSMNet(maxdisparity)
    raise ValueError("no drop-in for model %r (dispnetcorr, iresnet, gcnet, psmnet)" % name_model)


def normalize_imagenet(img: torch.Tensor) -> torch.Tensor:
    """Stereo_normalize / Normalize_Imagenet (myTransforms/__init__.py:8-14): (x - mean) / std per RGB channel, x in [0, 1]."""
    mean = torch.tensor(IMAGENET_MEAN, device=img.device, dtype=img.dtype).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD, device=img.device, dtype=img.dtype).view(1, 3, 1, 1)
    return (img - mean) / std


def disp_predict(model, imgL_np, imgR_np, use_cuda=None):
    """deploy.py:15-32: uint8 RGB arrays [H, W, 3] -> disparity map [H, W] (numpy float32)."""
    use_cuda = torch.cuda.is_available() if use_cuda is None else use_cuda
    imgL = torch.from_numpy(np.ascontiguousarray(imgL_np).transpose(2, 0, 1)[None].copy()).float()
    imgR = torch.from_numpy(np.ascontiguousarray(imgR_np).transpose(2, 0, 1)[None].copy()).float()
    if use_cuda:
        imgL, imgR = imgL.cuda(), imgR.cuda()
    imgL = normalize_imagenet(imgL / 255.0); imgR = normalize_imagenet(imgR / 255.0)
    with torch.no_grad():
        _, disps = model(imgL, imgR, mode="test")
    d = disps[0]
    d = d[0, 0] if d.dim() == 4 else d[0]            # PSMNet returns (B, H, W), the others (B, 1, H, W)
    return d.detach().float().cpu().numpy()
