"""Where does the bf16 disparity delta of the PSMNet 3-D stack come from?  

CPU experiment with the oracle's number-format emulation: the north-star path is evaluated in fp32 and with individual
rounding points switched to bf16, on (a) purely random BN-calibrated weights — a chaotic amplifier whose soft-argmin mixes
192 candidate disparities with broad, multi-modal weights — and (b) the `psmnet_matcher_params` stack that really matches.
Prints mean |disparity - fp32 disparity| per variant.  Runs in about a minute at the default size.

    python tests/experiments/bf16_error_budget.py [H W maxdisp]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle.ops as O  # noqa: E402

BF = torch.bfloat16


def run(params, fL, fR, maxdisp, hw, mode):
    """mode: None fp32 | 'all' | 'weights' | 'acts' | 'acts_trunk_only' (classifier inputs fp32) | 'hg_only' ..."""
    orig = O.conv3d_block
    state = {"layer": 0}

    def patched(x, weight, scale=None, shift=None, stride=1, transposed=False, residual=None, relu=False,
                operand_dtype=None, storage_dtype=None):
        cout = weight.shape[1] if transposed else weight.shape[0]
        if mode == "weights":
            return orig(x, weight.to(BF).float(), scale, shift, stride, transposed, residual, relu, None, None)
        if mode == "acts":
            return orig(x.to(BF).float(), weight, scale, shift, stride, transposed, residual, relu, None, BF if cout > 1 else None)
        if mode == "classif_fp32":       # everything bf16 except the two classifier convolutions' operands/storage
            if cout == 1 or state.get("in_classif"):
                return orig(x, weight, scale, shift, stride, transposed, residual, relu, None, None)
            return orig(x, weight, scale, shift, stride, transposed, residual, relu, BF, BF)
        if mode == "storage_fp32":        # bf16 operands, activations kept fp32 between layers (residual adds read fp32)
            return orig(x, weight, scale, shift, stride, transposed, residual, relu, BF, None)
        return orig(x, weight, scale, shift, stride, transposed, residual, relu, operand_dtype, storage_dtype)

    O.conv3d_block = patched
    try:
        od = (BF, BF) if mode == "all" else None
        with torch.no_grad():
            return O.psmnet_hotpath(params, fL, fR, maxdisp, hw, operand_dtype=od)
    finally:
        O.conv3d_block = orig


def main():
    H, W, maxdisp = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (128, 256, 96)
    torch.set_num_threads(os.cpu_count())
    fL, fR, disp = O.synthetic_stereo_features(H // 4, W // 4, d_lo=4.0, d_hi=maxdisp / 4 * 0.6, seed=4)
    cost = O.concat_volume(fL, fR, maxdisp // 4, "psm")
    sets = {"random (He init, BN calibrated)": O.psmnet_random_params(seed=21, calibrate_on=cost),
            "matcher (psmnet_matcher_params)": O.psmnet_matcher_params(seed=21)}
    for name, params in sets.items():
        ref = run(params, fL, fR, maxdisp, (H, W), None)
        # spread of the soft-argmin distribution: how many disparities carry weight
        c1, c2, c3 = O.psmnet_aggregate(params, cost)
        up = torch.nn.functional.interpolate(c3, [maxdisp, H, W], mode="trilinear", align_corners=True).squeeze(1)
        p = torch.softmax(up, 1)
        d = torch.arange(maxdisp).view(1, -1, 1, 1).float()
        mean = (p * d).sum(1, keepdim=True)
        std = ((p * (d - mean) ** 2).sum(1)).sqrt().mean()
        print("== %s: %dx%d maxdisp %d; logit std %.2f; soft-argmin distribution std %.1f px" % (name, H, W, maxdisp, float(up.std()), float(std)))
        for mode in ("all", "weights", "acts", "storage_fp32", "classif_fp32"):
            out = run(params, fL, fR, maxdisp, (H, W), mode)
            print("   %-14s mean |d - fp32| pred3 %.4f  pred2 %.4f  pred1 %.4f px" %
                  (mode, *[float((a - b).abs().mean()) for a, b in zip(out, ref)]))


if __name__ == "__main__":
    main()
