"""The 3-D stack's activation layout: bf16 "padded NDHWC" [B][D+2][H+2][W+2][C] with a zero rim.

The zero rim is the convolution padding (so a filter tap is a constant row offset in the
flattened voxel index and TMA needs no bounds logic); the conv kernels only ever write the
interior, so a buffer is zeroed once when it is allocated and can then be reused.
"""
from __future__ import annotations

import torch

from . import _lib


class PaddedVolume:
    __slots__ = ("data", "B", "C", "D", "H", "W")

    def __init__(self, data, B, C, D, H, W):
        self.data, self.B, self.C, self.D, self.H, self.W = data, B, C, D, H, W

    @property
    def shape5(self):
        """Logical NCDHW shape."""
        return (self.B, self.C, self.D, self.H, self.W)

    @staticmethod
    def empty(B, C, D, H, W, device, zero_rim=True):
        n = B * (D + 2) * (H + 2) * (W + 2) * C
        data = torch.zeros(n, device=device, dtype=torch.bfloat16) if zero_rim else \
            torch.empty(n, device=device, dtype=torch.bfloat16)
        return PaddedVolume(data, B, C, D, H, W)

    @staticmethod
    def empty_zero_rim(B, C, D, H, W, device):
        """Uninitialised interior, zero rim (dsm_zero_rim): for outputs a convolution kernel fills completely."""
        v = PaddedVolume.empty(B, C, D, H, W, device, zero_rim=False)
        _lib.check(_lib.lib().dsm_zero_rim(v.data.data_ptr(), B, C, D, H, W, _lib.stream_ptr(v.data.device)), "dsm_zero_rim")
        return v

    @staticmethod
    def from_ncdhw(x):
        """NCDHW fp32 (the reference's layout) -> padded NDHWC bf16 (RNE)."""
        _lib.require_cuda(x)
        x = x.contiguous().float()
        B, C, D, H, W = x.shape
        v = PaddedVolume.empty(B, C, D, H, W, x.device, zero_rim=False)
        _lib.check(_lib.lib().dsm_pack_ndhwc(x.data_ptr(), v.data.data_ptr(), B, C, D, H, W, _lib.stream_ptr(x.device)),
                   "dsm_pack_ndhwc")
        return v

    def to_ncdhw(self):
        out = torch.empty(self.B, self.C, self.D, self.H, self.W, device=self.data.device, dtype=torch.float32)
        _lib.check(_lib.lib().dsm_unpack_ndhwc(self.data.data_ptr(), out.data_ptr(), self.B, self.C, self.D, self.H, self.W,
                                               _lib.stream_ptr(self.data.device)), "dsm_unpack_ndhwc")
        return out

    def view6(self):
        """[B][D+2][H+2][W+2][C] view (for tests)."""
        return self.data.view(self.B, self.D + 2, self.H + 2, self.W + 2, self.C)
