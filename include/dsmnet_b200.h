/*
 * dsmnet_b200.h — C-ABI of the B200-native (sm_100a) stereo cost-volume hot path.
 *
 * The reference (sunshinnnn/DSMnet) has no native layer and no FFI: its hot path is Python
 * that composes stock torch ops.  Each entry point below replaces one such composition; the
 * reference lines it replaces are cited next to it (paths relative to the reference root).
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer into memory owned by the caller; the library never
 *     allocates, frees or retains pointers; there is no hidden scratch;
 *   - `stream` is the caller's cudaStream_t (CUstream); all work is enqueued on it
 *     asynchronously: no host synchronisation, CUDA-graph capturable;
 *   - return 0 = ok, >0 = cudaError_t of a failed launch, <0 = DSM_E* below;
 *   - re-entrant from any host thread, no global mutable state; the caller has made the
 *     target device current;
 *   - tensors are dense, row-major in the order written (last index fastest).
 *
 * There is no CPU fallback anywhere behind this interface.
 */
#ifndef DSMNET_B200_H
#define DSMNET_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSM_ABI_VERSION 1

/* error codes (<0) */
#define DSM_EINVAL       (-1)  /* bad shape / null pointer / bad enum                   */
#define DSM_EUNSUPPORTED (-2)  /* shape or mode outside what the kernels are built for  */
#define DSM_EALIGN       (-3)  /* pointer not aligned as the kernel requires (16 B)     */
#define DSM_EDRIVER      (-4)  /* driver entry point (cuTensorMapEncodeTiled) missing   */

/* concat-volume modes */
#define DSM_VOL_PSM      0 /* models/psmnet/stackhourglass.py:124-133 : both halves 0 for x<d */
#define DSM_VOL_GC       1 /* models/gcnet.py:131-135 : left half copied for every x          */
#define DSM_VOL_GC_RIGHT 2 /* models/gcnet.py:157-164 (xR) : right-reference volume           */

/* dtypes / layouts of a cost volume */
#define DSM_F32  0
#define DSM_BF16 1
#define DSM_NCDHW        0 /* [B][2C][D][H][W]            — the reference's layout            */
#define DSM_NDHWC_PADDED 1 /* [B][D+2][H+2][W+2][2C]      — the 3-D stack's layout, zero rim  */

int         dsm_abi_version(void);
const char* dsm_strerror(int code);

/* ---- op 1: 1-D correlation. Replaces Corr1d.forward, models/util_conv.py:71-81 ---------
 * out[b,d,y,x] = sum_c fL[b,c,y,x] * fR[b,c,y,x-d*stride]  (x >= d*stride and d < W), else 0.
 * fL,fR: [B][C][H][W] fp32; out: [B][D][H][W] fp32. (The k>1 average pool of :82-85 stays a
 * stock pooling op in the caller.)                                                           */
int dsm_corr1d_fwd(const float* fL, const float* fR, float* out,
                   int B, int C, int H, int W, int D, int stride, void* stream);
/* backward of the above (autograd of util_conv.py:76-81): gL,gR: [B][C][H][W] (overwritten) */
int dsm_corr1d_bwd(const float* gout, const float* fL, const float* fR, float* gL, float* gR,
                   int B, int C, int H, int W, int D, int stride, void* stream);

/* ---- op 2: concatenation cost volume. Replaces the loops at
 * models/psmnet/stackhourglass.py:124-133, models/gcnet.py:131-135 and :156-164 -------------
 * fL,fR: [B][C][H][W] fp32. out: 2C channels, D disparities, dtype/layout as given.
 * With DSM_NDHWC_PADDED the whole buffer including the zero rim is written.                  */
int dsm_concat_volume_fwd(const float* fL, const float* fR, void* out,
                          int B, int C, int D, int H, int W,
                          int mode, int out_dtype, int out_layout, void* stream);
/* backward: gL,gR [B][C][H][W] fp32 (overwritten) from gout in the given dtype/layout        */
int dsm_concat_volume_bwd(const void* gout, float* gL, float* gR,
                          int B, int C, int D, int H, int W,
                          int mode, int dtype, int layout, void* stream);

/* ---- op 3: 3-D convolution block (k=3, pad=1), tcgen05/TMEM implicit GEMM ----------------
 * Replaces Conv3d/ConvTranspose3d + BatchNorm3d (+ReLU, +residual add) of
 * models/psmnet/submodule.py:16-19, stackhourglass.py:26-41,46-60,73-98,135-149,
 * models/util_conv.py:150-179, models/gcnet.py:38-61.
 *   x        : bf16 [B][D+2][H+2][W+2][Cin], zero rim (DSM_NDHWC_PADDED)
 *   w_packed : bf16 [27][CoutP][Cin], tap = (kd*3+kh)*3+kw, CoutP = max(16, Cout) rounded up
 *              to a multiple of 16 (extra rows zero).  For transposed convs the tap refers to
 *              the ConvTranspose3d kernel index (weight[ci][co][kd][kh][kw]).
 *   scale,shift : fp32 [CoutP] per-channel affine applied to the accumulator (folded
 *              BatchNorm + bias), either may be NULL (1 / 0)
 *   residual : same dtype/geometry as y; NULL for none
 *   relu     : 0 = none, 1 = ReLU after the residual add (PSMNet hourglass, stackhourglass.py:46-58),
 *              2 = ReLU before it (GC-Net: myAdd3d(l33(x32), x29) adds to an activated deconv, gcnet.py:78-96)
 *   y        : y_dtype DSM_BF16 -> bf16 DSM_NDHWC_PADDED [B][Do+2][Ho+2][Wo+2][Cout] (only the
 *              interior is written: the rim must have been zeroed once by the caller);
 *              y_dtype DSM_F32 (Cout==1 only) -> fp32 [B][Do][Ho][Wo]
 *   (D,H,W)  : INPUT extent; output extent is the same (stride 1), floor((n-1)/2)+1
 *              (stride 2) or 2n (transposed: k3,s2,p1,output_padding 1)
 *   ws       : unused, may be NULL (kept so the signature does not change when a split-K
 *              variant appears)                                                              */
int dsm_conv3d_fwd(const void* x, const void* w_packed, const float* scale, const float* shift,
                   const void* residual, void* y,
                   int B, int Cin, int Cout, int D, int H, int W,
                   int stride, int transposed, int relu, int y_dtype,
                   void* ws, size_t ws_bytes, void* stream);

/* ---- op 3, training: filter gradient of the k=3 convolutions (autograd of the nn.Conv3d / nn.ConvTranspose3d
 * modules cited above).  The INPUT gradient (dgrad) needs no entry point of its own: it is dsm_conv3d_fwd with
 * transformed weights (stride-1 conv: flipped taps and swapped channels; stride-2 conv: the transposed conv of
 * gy with the same filter; transposed conv: the stride-2 conv of gy) — see dsmnet_b200/conv3d.py.
 *   anchor  : the coarser of (x, gy), padded NDHWC bf16 [B][Da+2][Ha+2][Wa+2][Ca]
 *             (conv: gy, Ca = Cout;  transposed conv: x, Ca = Cin)
 *   partner : the other one, [B][Dp+2][Hp+2][Wp+2][Cb]; partner voxel = stride*anchor + tap - 1 per axis
 *   dw      : fp32 [Ca_out][Cb_out][3][3][3] = PyTorch's weight layout for both module types
 *             (Ca_out <= Ca, Cb_out <= Cb: the tensors may carry zero-padded channels)
 *   scale_a / scale_b : optional per-channel factors (the folded BatchNorm scale on the gy side), NULL = 1
 *   accumulate : 0 overwrite dw, 1 add to it;   ws: dsm_conv3d_wgrad_workspace_bytes(Ca, Cb) bytes of scratch  */
size_t dsm_conv3d_wgrad_workspace_bytes(int Ca, int Cb);
/* per-problem size: with this much scratch the stride-2 / transposed layers also run on the tcgen05 kernel (through a
 * parity gather of the finer tensor into the workspace); with only the size above they use the warp-level kernel */
size_t dsm_conv3d_wgrad_workspace_bytes_ex(int B, int Ca, int Cb, int Da, int Ha, int Wa, int stride);
int dsm_conv3d_wgrad(const void* anchor, const void* partner, float* dw,
                     int B, int Ca, int Cb, int Da, int Ha, int Wa, int Dp, int Hp, int Wp, int stride,
                     int Ca_out, int Cb_out, const float* scale_a, const float* scale_b, int accumulate,
                     void* ws, size_t ws_bytes, void* stream);

/* ---- training-mode BatchNorm3d + ReLU + skip add between the 3-D convolutions -----------------
 * Replaces nn.BatchNorm3d under model.train() with F.relu and the skip adds around it:
 * models/psmnet/submodule.py:16-19, models/psmnet/stackhourglass.py:43-62,135-149,
 * models/util_conv.py:160-178 with models/gcnet.py:65-101.
 * All volumes are padded NDHWC bf16 [B][D+2][H+2][W+2][C] with a zero rim, C in {32, 64, 128}.
 * relu: 0 none, 1 after the skip add (mask z > 0), 2 before it (mask y*scale+shift > 0).
 *   dsm_bn_stats        : sums[0..C) = sum y, sums[C..2C) = sum y*y        (double, zeroed by the call)
 *   dsm_bn_finalize_fwd : scale = gamma*rstd, shift = beta - mean*scale, mean, rstd (biased variance, eps);
 *                         running_mean/var (nullable) get torch's momentum update, conv_bias (nullable) is the
 *                         bias of the convolution in front (it only moves the running mean)
 *   dsm_bn_act_fwd      : z = act(y*scale + shift [+ residual])
 *   dsm_bn_act_bwd_reduce: g = gz*mask; sums = (sum g, sum g*y)
 *   dsm_bn_finalize_bwd : dgamma, dbeta and coef[3][C] (a, b, c) of dy = a*g + b*y + c
 *   dsm_bn_act_bwd      : dy (zero rim) and, if gres != NULL, gres = g (the gradient of the skip tensor)     */
/* zero rim of a padded bf16 volume (C % 8 == 0) whose interior a convolution is about to write          */
int dsm_zero_rim(void* data, int B, int C, int D, int H, int W, void* stream);
int dsm_bn_stats(const void* y, int B, int C, int D, int H, int W, double* sums, void* stream);
int dsm_bn_finalize_fwd(const double* sums, const float* gamma, const float* beta, const float* conv_bias,
                        int C, long long count, float eps, float momentum, float* running_mean, float* running_var,
                        float* scale, float* shift, float* mean, float* rstd, void* stream);
int dsm_bn_act_fwd(const void* y, const float* scale, const float* shift, const void* residual, int relu,
                   void* z, int B, int C, int D, int H, int W, void* stream);
int dsm_bn_act_bwd_reduce(const void* gz, const void* y, const void* z, const float* scale, const float* shift,
                          int relu, double* sums, int B, int C, int D, int H, int W, void* stream);
int dsm_bn_finalize_bwd(const double* sums, const float* gamma, const float* mean, const float* rstd,
                        int C, long long count, float* dgamma, float* dbeta, float* coef, void* stream);
int dsm_bn_act_bwd(const void* gz, const void* y, const void* z, const float* scale, const float* shift,
                   const float* coef, int relu, void* dy, void* gres, int B, int C, int D, int H, int W, void* stream);

/* backward of the single-output-channel layers (PSMNet classif*.2: Conv3d 32->1, stackhourglass.py:99-109;
 * GC-Net l37: ConvTranspose3d 32->1 stride 2, gcnet.py:60-62): input gradient and weight gradient in one pass.
 *   x, gx : padded NDHWC bf16 [B][D+2][H+2][W+2][32] (gx: interior written, rim untouched)
 *   gy    : fp32 [B][Do][Ho][Wo] (= [D][H][W] for the convolution, up to [2D][2H][2W] for the transposed one)
 *   w, dw : fp32 [32][27] (the weight with its single output channel squeezed)
 *   ws    : dsm_conv3d_c1_bwd_workspace_bytes() bytes of scratch                                        */
size_t dsm_conv3d_c1_bwd_workspace_bytes(void);
int dsm_conv3d_c1_bwd(const void* x, const float* gy, const float* w, void* gx, float* dw,
                      int B, int D, int H, int W, int Do, int Ho, int Wo, int transposed,
                      void* ws, size_t ws_bytes, void* stream);

/* The volume is never materialised: the first 3-D convolution of the stack (PSMNet dres0.0,
 * stackhourglass.py:73,135; GC-Net l19, gcnet.py:38,94 — Conv3d(64 -> 32, k3, s1) + BN + ReLU) computed straight from the two
 * 32-channel feature maps.  featL / featR: bf16 NHWC with a zero rim of one pixel, [B][H+2][W+2][32] (dsm_pack_nhwc_bf16 with
 * rim = 1, or the 2-D trunk's last layer);
 * y: padded NDHWC bf16 [B][D+2][H+2][W+2][32]; mode as dsm_concat_volume_fwd; w_packed / scale / shift / relu / variant as
 * dsm_conv3d_fwd_ex.  Same result as dsm_concat_volume_fwd (bf16, padded) followed by dsm_conv3d_fwd.               */
int dsm_conv3d_volume_fwd(const void* featL, const void* featR, const void* w_packed, const float* scale, const float* shift,
                          void* y, int B, int C, int Cout, int D, int H, int W, int mode, int relu, int variant, void* stream);
/* NCHW fp32 -> NHWC bf16 (RNE) with a zero rim of `rim` pixels: y [B][H+2*rim][W+2*rim][C], rim written too            */
int dsm_pack_nhwc_bf16(const float* x_nchw, void* y_nhwc_bf16, int B, int C, int H, int W, int rim, void* stream);
/* both feature maps of a stereo pair in one launch                                                                 */
int dsm_pack_nhwc_bf16_pair(const float* xL_nchw, const float* xR_nchw, void* yL, void* yR, int B, int C, int H, int W, int rim, void* stream);

/* cropped skip add of the training path (myadd_3d / myAdd3d crop-to-min, stackhourglass.py:10-20, util_fun.py:41-51):
 * `full` is a padded bf16 volume of extent (Dn,Hn,Wn) (the BatchNorm'ed deconv output, statistics over all of it as in
 * the reference), residual / z have the smaller extent (Do,Ho,Wo): z = act(crop(full) + residual), relu 0/1 (after the
 * add).  Backward: g = gz masked by z > 0 when relu; gfull = g inside the crop and 0 elsewhere (whole padded array
 * written); gres (optional) = g on the interior voxels (the caller zeroes its rim).                                  */
int dsm_crop_add_fwd(const void* full, const void* residual, void* z, int B, int C,
                     int Dn, int Hn, int Wn, int Do, int Ho, int Wo, int relu, void* stream);
int dsm_crop_add_bwd(const void* gz, const void* z, void* gfull, void* gres, int B, int C,
                     int Dn, int Hn, int Wn, int Do, int Ho, int Wo, int relu, void* stream);

/* diagnostics: wgrad kernel selection (0 = tcgen05 for stride-1 layers [default], 1 = warp-level everywhere;
 * returns the previous mode) and the count of bounded pipeline waits that expired (0 in a healthy run)      */
int dsm_debug_wgrad_mode(int mode);
int dsm_debug_wgrad_timeouts(void);

/* layout converters at the reference boundary (NCDHW fp32 <-> padded NDHWC bf16)             */
/* filter repacking fp32 -> bf16 [27][CoutP][Cin] (CoutP = max(16, Cout), extra rows zero), w has 27 taps innermost:
 * mode 0: w[Cout][Cin][27] (nn.Conv3d); mode 1: w[Cin][Cout][27] (nn.ConvTranspose3d);
 * mode 2: w[Cin][Cout][27] with flipped taps (the stride-1 dgrad filter: pass the Conv3d weight, Cout/Cin exchanged)
 * mode 3: w[1][Cin][27], one output channel, kw folded into the output columns (for dsm_conv3d_fwd_ex variant bit 8) */
int dsm_pack_weight(const float* w, void* w_packed_bf16, int Cout, int Cin, int mode, void* stream);
int dsm_pack_ndhwc(const float* x_ncdhw, void* y_padded_bf16, int B, int C, int D, int H, int W, void* stream);
int dsm_unpack_ndhwc(const void* x_padded_bf16, float* y_ncdhw, int B, int C, int D, int H, int W, void* stream);

/* ---- 2-D feature-extraction trunks (callers of the path) ---------------------------------
 * Replaces the convbn / BasicBlock / SPP stacks of models/psmnet/submodule.py:10-13,21-42,65-140 and the conv2d_bn /
 * BasicBlock stack of models/gcnet.py:14-29 + models/util_conv.py:119-132,180-208 (eval-mode BatchNorm folded).
 * Activations: bf16 "padded NHWC" [B][H+2r][W+2r][ld] with a zero rim of r pixels (r >= dilation of every consumer);
 * ld >= C is the channel pitch, so x / y / residual may be channel slices of a wider tensor (pass the pointer to the first
 * channel of the slice).  w_packed: bf16 [k*k][Cout][Cin], tap = kh*k + kw.  y = relu?(conv(x)*scale + shift [+ residual]);
 * relu as in dsm_conv3d_fwd.  Cin in {32, 64, 128, 192, 256, 320, ...}, Cout in {16, 32, 64, 128}; ksize 1|3; stride 1|2
 * (stride 2: rim_in = 1, dilation 1); dilation 1|2.  y_mode 0: padded NHWC bf16 with rim_out / ldy; y_mode 2: fp32 NCHW
 * [B][Cout][Ho][Wo] (the layout of the feature maps dsm_concat_volume_fwd / dsm_corr1d_fwd take; no residual).
 * tcgen05 implicit GEMM: the kernels of dsm_conv3d_fwd on a volume without rim planes.  variant: bit7 as dsm_conv3d_fwd_ex. */
int dsm_conv2d_fwd(const void* x, const void* w_packed, const float* scale, const float* shift,
                   const void* residual, void* y,
                   int B, int Cin, int Cout, int H, int W, int ksize, int stride, int dilation, int relu,
                   int rim_in, int rim_out, int ldx, int ldy, int ldr, int y_mode, int variant, void* stream);
/* The row-sharing form of dsm_conv2d_fwd (conv2d_rs.cu) for ksize 3, stride 1, Cin and Cout in {32, 64}, y_mode 0, relu 0|1,
 * ldy / ldr multiples of 16 and 32-byte aligned y / residual: a CTA keeps all nine weight tiles resident and computes a band of
 * 4 (Cout 64) or 8 (Cout 32) output rows from each input row loaded once.  dsm_conv2d_fwd routes eligible layers here unless
 * variant bit 2 (value 4) is set; DSM_EUNSUPPORTED for other channel counts.  Three warps issue the MMAs of one accumulator in
 * rotation, so the fp32 accumulation order (the last bit) can differ from run to run; variant bit 3 (value 8) uses one issuer:
 * bit-reproducible, 10-20 % slower. */
int dsm_conv2d_rs_fwd(const void* x, const void* w_packed, const float* scale, const float* shift,
                      const void* residual, void* y, int B, int Cin, int Cout, int H, int W, int dilation, int relu,
                      int rim_in, int rim_out, int ldx, int ldy, int ldr, int variant, void* stream);
/* the image-facing layer: Conv2d(3 -> 32, k 3|5, stride 2, pad k/2) + affine + ReLU, NCHW fp32 image [B][3][H][W] with the
 * PyTorch fp32 weight [32][3][k][k] -> padded NHWC bf16 [B][Ho+2r][Wo+2r][32] (submodule.py:68, gcnet.py:21); CUDA cores. */
int dsm_conv2d_first_fwd(const float* img, const float* w, const float* scale, const float* shift, void* out,
                         int B, int H, int W, int ksize, int relu, int rim_out, void* stream);
/* PSMNet's SPP branches (submodule.py:84-98,126-136): AvgPool 64/32/16/8 of the 128-channel `skip` slice, 1x1 conv
 * 128 -> 32 (+BN+ReLU; w fp32 [4][32][128], scale / shift [4][32], branch1..branch4) with the reference's one-pixel
 * padding ring, bilinear upsampling to H x W, bf16 into channels [coff, coff+128) of `cat` in the order
 * (branch4, branch3, branch2, branch1).  skip / cat: padded NHWC bf16 with rim / ld (cat may be the same buffer). */
size_t dsm_spp_workspace_bytes(int B, int H, int W);
int dsm_spp_fwd(const void* skip, const float* w, const float* scale, const float* shift, void* cat,
                int B, int H, int W, int rim, int ld, int coff, int align_corners,
                void* ws, size_t ws_bytes, void* stream);

/* ---- op 4: soft-argmin disparity regression ---------------------------------------------
 * Replaces F.softmax + disparityregression, models/psmnet/submodule.py:56-63 with
 * stackhourglass.py:155-166, and models/gcnet.py:104-111 (sign=-1).
 * cost: [B][D][H][W] fp32; disp: [B][H][W] fp32; disp = sum_d d*softmax_d(sign*cost).        */
int dsm_softargmin_fwd(const float* cost, float* disp, int B, int D, int H, int W, float sign, void* stream);
int dsm_softargmin_bwd(const float* cost, const float* disp, const float* gdisp, float* gcost,
                       int B, int D, int H, int W, float sign, void* stream);
/* plain regression of submodule.py:60-63 on probabilities: disp = sum_d d*prob[b,d,y,x]      */
int dsm_disparity_regression_fwd(const float* prob, float* disp, int B, int D, int H, int W, void* stream);
int dsm_disparity_regression_bwd(const float* gdisp, float* gprob, int B, int D, int H, int W, void* stream);
/* fused head of stackhourglass.py:152-166: trilinear upsample of cost_lr [B][Dl][Hl][Wl] to
 * (D,H,W), softmax over D, regression.  align_corners=1 is the PyTorch<=0.3 behaviour the
 * reference was written for, 0 the modern default.                                           */
int dsm_upsample_softargmin_fwd(const float* cost_lr, float* disp, int B, int Dl, int Hl, int Wl,
                                int D, int H, int W, int align_corners, void* stream);
/* training form: also writes lse2[B][H][W] = log2 of the softmax denominator (max included), which
 * dsm_upsample_softargmin_bwd needs; the backward accumulates gcost_lr = J^T gdisp (zeroed by the call,
 * float atomics) without materialising the upsampled volume.                                   */
int dsm_upsample_softargmin_fwd_lse(const float* cost_lr, float* disp, float* lse2, int B, int Dl, int Hl, int Wl,
                                    int D, int H, int W, int align_corners, void* stream);
int dsm_upsample_softargmin_bwd(const float* cost_lr, const float* disp, const float* lse2, const float* gdisp,
                                float* gcost_lr, int B, int Dl, int Hl, int Wl,
                                int D, int H, int W, int align_corners, void* stream);

/* ---- op 5: imwrap bilinear warp. Replaces imwrap_BCHW, utils/imwrap.py:59-71 -------------
 * src: [B][C][H0][W0]; disp: [B][H][W]; row: [W], col: [H] (the torch.linspace vectors of
 * imwrap.py:57-58, computed by the caller exactly as there); out: [B][C][H][W].
 * out = grid_sample(src + delt, grid, bilinear, zeros, align_corners=True),
 * grid.x = k*(row[j] - disp*2/(W0-1)), k = -1 if fliplr else 1, grid.y = col[i].             */
int dsm_warp_fwd(const float* src, const float* disp, const float* row, const float* col,
                 float delt, int fliplr, float* out,
                 int B, int C, int H0, int W0, int H, int W, void* stream);
/* backward: gsrc [B][C][H0][W0] is zeroed by the callee then scatter-added; gdisp [B][H][W]  */
int dsm_warp_bwd(const float* gout, const float* src, const float* disp, const float* row,
                 const float* col, float delt, int fliplr, float* gsrc, float* gdisp,
                 int B, int C, int H0, int W0, int H, int W, void* stream);

/* batched form: every warp of one self-supervised training step (losses/loss.py:449-452: 4 warps x 7 pyramid levels,
 * all independent) in ONE launch.  `jobs` is a HOST array of n <= DSM_WARP_MAX_JOBS descriptors (device pointers inside);
 * it travels as a kernel parameter, nothing is copied or retained.  Forward uses src/disp/row/col/out; backward uses
 * gout/src/disp/row/col/gdisp and gsrc, which may be NULL (no gradient for that source) and is NOT zeroed by the callee
 * here (the caller zeroes all gsrc buffers with one memset of their common allocation).                            */
#define DSM_WARP_MAX_JOBS 32
typedef struct DsmWarpJob {
    const float *src, *disp, *row, *col;
    float* out;
    const float* gout;
    float *gsrc, *gdisp;
    float delt;
    int fliplr;
    int B, C, H0, W0, H, W;
} DsmWarpJob;
int dsm_warp_fwd_batched(const DsmWarpJob* jobs, int n, void* stream);
int dsm_warp_bwd_batched(const DsmWarpJob* jobs, int n, void* stream);

/* ---- callers of op 5: fused SSIM map of the photometric loss (losses/SSIM.py:24-42 as used by losses/loss.py:196-236) ----
 * a, b: [B][C][H][W] fp32; s: [B][H][W] = channel-averaged 11x11 Gaussian (sigma 1.5) SSIM, zero padding.  Backward: the
 * gradient w.r.t. b (the warped image) only: gb [B][C][H][W] from gs [B][H][W]; ws: dsm_ssim_bwd_workspace_bytes.   */
int dsm_ssim_fwd(const float* a, const float* b, float* s, int B, int C, int H, int W, void* stream);
size_t dsm_ssim_bwd_workspace_bytes(int B, int H, int W);
int dsm_ssim_bwd(const float* a, const float* b, const float* gs, float* gb, int B, int C, int H, int W,
                 void* ws, size_t ws_bytes, void* stream);

/* test hook: the integer north-west source pixel (x0,y0), int32 [B][H][W], that the warp kernels
 * derive for each output pixel — lets a test assert bit-exact sampling indices against the oracle. */
int dsm_warp_indices(const float* disp, const float* row, const float* col, int fliplr,
                     int* x0, int* y0, int B, int H0, int W0, int H, int W, void* stream);

/* ---- extended / diagnostic entry points ---------------------------------------------------
 * dsm_conv3d_fwd_ex: as dsm_conv3d_fwd, with (Do,Ho,Wo) = extent of the y/residual buffers when
 * that is a crop of the natural output size (the reference's crop-to-min add, myadd_3d /
 * myAdd3d, stackhourglass.py:10-20, util_fun.py:41-51; pass 0,0,0 for the natural size) and
 * `variant` (bit0: descriptor base-offset experiment, bit1: force the per-tap kernel instead of the
 * default row-shifted-descriptor kernel for stride-1 convolutions, bit7: launch with programmatic stream
 * serialization - the kernel's prologue overlaps the previous kernel's tail, its dependent accesses wait on
 * griddepcontrol.wait; only for parameters (weights, scale, shift) that were written before the previous kernel
 * started; bit8: the weights of a single-output-channel stride-1 layer with 32 inputs are packed by dsm_pack_weight
 * mode 3 (kw folded into the output columns: a third of the MMAs); 0 = default).                                  */
int dsm_conv3d_fwd_ex(const void* x, const void* w_packed, const float* scale, const float* shift,
                      const void* residual, void* y,
                      int B, int Cin, int Cout, int D, int H, int W,
                      int stride, int transposed, int relu, int y_dtype,
                      int Do, int Ho, int Wo, int variant, void* stream);
/* A pipeline wait inside a conv / wgrad kernel that is not satisfied within 2 s is FATAL by default: the kernel traps
 * and the context reports a launch failure (nothing can be silently wrong).  dsm_debug_conv_set_trap(0) selects the
 * bring-up mode instead (the wait gives up, the kernel finishes on garbage, the counter below tells); returns the
 * previous setting.  dsm_debug_conv_timeouts: expired waits since load (0 when healthy).                          */
int dsm_debug_conv_set_trap(int on);
int dsm_debug_conv_timeouts(void);
/* bring-up aid (library built with -DDSM_CONV_TRACE): host-mapped int[4] progress slots */
int dsm_debug_conv_set_progress(int* host_mapped);

#ifdef __cplusplus
}
#endif
#endif /* DSMNET_B200_H */
