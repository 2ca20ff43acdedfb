/*
 * fusiondepth_b200 -- C ABI of the B200 (sm_100a) kernels behind the FusionDepth per-step
 * training hot path.
 *
 * The reference (AutoAILab/FusionDepth) is pure Python/PyTorch and has no FFI of its own; the
 * boundary it exposes for this path is the Python module surface `layers.*` / `networks.*`
 * (see INTEGRATION.md).  Each entry point below replaces the ATen op sequence behind one
 * reference function and cites it (paths relative to the reference repo).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is DEVICE memory unless the name ends in
 *     `_host`; the library never allocates, frees or retains device memory;
 *   - `stream` is a cudaStream_t passed as void*; launches are asynchronous, no internal syncs,
 *     capturable in CUDA graphs;
 *   - return 0 on success; non-zero => fd_last_error() describes the failure;
 *   - activations are fp32 NHWC ("channels last"), conv weights fp32 [Cout,KH,KW,Cin];
 *     images of the loss chain are fp32 NCHW as the reference's loader delivers them.
 */
#ifndef FUSIONDEPTH_B200_H
#define FUSIONDEPTH_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* fd_last_error(void);
int fd_version(void);
/* number of kernel launches issued through this library since load (for bench gpu_launches) */
long fd_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Sparse LiDAR  (bit-exact)
 * ---------------------------------------------------------------------------------------- */

/* kitti_utils.generate_depth_map (kitti_utils.py:40-102), batched over frames.
 *   points  [n_points_total,4] float32 velodyne points of all frames, concatenated
 *   offsets [n_frames+1] int32 prefix offsets into points
 *   P       [n_frames,12] float64: P_rect_0x @ R_rect_00 @ Tr_velo_to_cam (kitti_utils.py:53-56)
 *   shape_h/shape_w: the `shape` argument (0,0 = None); depth_out [n_frames,out_h,out_w] float64
 *   workspace: fd_lidar_workspace_bytes() bytes */
size_t fd_lidar_workspace_bytes(int n_frames, int W_im, int H_im);
int fd_lidar_depth_map(const float* points, const int* offsets, int n_frames, int n_points_total,
                       const double* P, int W_im, int H_im, int vel_depth, int shape_h,
                       int shape_w, double* depth_out, int out_h, int out_w, void* workspace,
                       void* stream);

/* F.max_pool2d(.,2,ceil_mode=True) -> float32 -> /100.0  (kitti_dataset.py:105-107,
 * mono_dataset.py:194-198).  depth [n,H,W] float64 -> out [n,ceil(H/2),ceil(W/2)] float32 */
int fd_lidar_pool_scale(const double* depth, int n_frames, int H, int W, float* out, void* stream);

/* get_4beam_2channel (gen2channel.py:60-117).  fourbeam [n,H,W] -> out [n,2,H,W]
 * (expanded depth, confidence); source window rows [r0,r1) cols [c0,c1)
 * (reference: 76,190,2,638 at 192x640). */
int fd_two_channel(const float* fourbeam, int n_frames, int H, int W, int r0, int r1, int c0, int c1,
                   float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused photometric loss chain
 *   trainer.generate_images_pred + compute_losses (trainer.py:425-596) and the layers.* they
 *   call (layers.py:11-20,133-162,204-226,235-281), default flags.
 * ---------------------------------------------------------------------------------------- */
typedef struct fd_photoloss_desc {
  int B, H, W;
  /* color[i][s]: frame_ids[i] (0, -1, +1) at scale s, [B,3,H>>s,W>>s] NCHW.  Only scale 0 of the
   * source frames is read (source_scale = 0, trainer.py:436,503). */
  const float* color[3][4];
  const float* disp[4];       /* [B,1,H>>s,W>>s] network output ("disp", s) */
  const float* K;             /* [B,4,4] ("K",0) */
  const float* inv_K;         /* [B,4,4] ("inv_K",0) */
  const float* T[2];          /* [B,4,4] cam_T_cam for frames -1, +1 */
  const float* noise[4];      /* [B,2,H,W] the randn of trainer.py:551 per scale */
  const float* beam;          /* [B,1,H,W] si-loss target: inputs["4beam"] (trainer.py:577-589) or
                                 inputs["inf_gdc"] (refiner.py:678-688) */
  float min_depth, max_depth; /* options.py:64-71 (0.1, 100) */
  float smoothness;           /* disparity_smoothness 1e-3 */
  float si_thresh, si_var;    /* gdc_loss_threshold 2.0, si_var 0.3 */
  /* Scale-invariant log term: D = si_pred_mul * depth, G = si_tgt_mul * target,
   * valid = G > si_lo && D < 80 && D > si_lo && |D - G| < si_thresh,
   * loss = si_weight * sqrt(mean d^2 - si_var * mean(d)^2), d = ln D - ln G over valid pixels.
   * trainer.py:577-589: (26, 100, 1, 0.1) on every scale (bit s of si_scales; only scale 0 without
   * --trainer_siloss_all_scale); refiner.py:557-563, 678-688: (1, 1, 1e-3, 10 * 0.008 * 4) on scale 0. */
  int si_scales;
  float si_pred_mul, si_tgt_mul, si_lo, si_weight;
  unsigned char* sel;         /* [4,B,H,W] out: argmin channel (0,1 identity; 2,3 warped) */
  float* out_depth[4];        /* optional [B,1,H,W] ("depth",0,s) */
  float* out_color[4][2];     /* optional [B,3,H,W] ("color",f,s) */
  float* out_to_optimise[4];  /* optional [B,H,W] */
} fd_photoloss_desc;

size_t fd_photoloss_workspace_bytes(int B, int H, int W);
/* losses[9] = loss/0..3, loss/si_loss0..3, loss.  The workspace carries the state the backward
 * needs and must be the same buffer in both calls. */
int fd_photoloss_fwd(const fd_photoloss_desc* desc_host, float* losses, void* workspace, void* stream);
/* grad_loss: device scalar d/d losses["loss"].  grad_disp[s] [B,1,H>>s,W>>s]; grad_T [B,4,4]. */
int fd_photoloss_bwd(const fd_photoloss_desc* desc_host, const float* grad_loss,
                     float* const grad_disp[4], float* grad_T0, float* grad_T1, void* workspace,
                     void* stream);

/* ------------------------------------------------------------------------------------------
 * Unfused drop-in operators: what layers.BackprojectDepth / Project3D / SSIM and the F.interpolate /
 * F.grid_sample calls of the UNCHANGED drivers resolve to (trainer.py:434-470, 479-486, 579;
 * refiner.py:325-335; evaluate_depth.py:207-218) when the fused fd_photoloss_* pair is not patched in.
 * fp32 NCHW contiguous tensors; `planes` = B*C.
 *   upsample_bilinear: F.interpolate(mode="bilinear", align_corners=False), any in/out size;
 *   backproject:  cam [B,4,HW] = [depth * (inv_K[:3,:3] @ [x,y,1]); 1]           (layers.py:133-162)
 *   project3d:    grid [B,H,W,2] = ((K@T)[:3] @ pts -> /(z+eps) -> /(W-1),(H-1) - .5) * 2  (layers.py:204-226);
 *                 bwd: dpoints [B,4,HW] (optional) and dT [B,4,4] (optional), workspace B*12 floats;
 *   grid_sample_border: bilinear, padding_mode="border", align_corners=False; bwd: dgrid and/or dimg;
 *   ssim: clamp((1 - SSIM)/2, 0, 1) over 3x3 reflect-padded windows (layers.py:251-281); bwd is the
 *         gradient wrt the first argument (swap the arguments for the second), workspace 3*planes*H*W floats.
 * ---------------------------------------------------------------------------------------- */
int fd_upsample_bilinear_fwd(const float* x, float* y, long planes, int h, int w, int H, int W, void* stream);
int fd_upsample_bilinear_bwd(const float* dy, float* dx, long planes, int h, int w, int H, int W, void* stream);
int fd_backproject_fwd(const float* depth, const float* inv_K, float* cam, int B, int H, int W, void* stream);
int fd_backproject_bwd(const float* dcam, const float* inv_K, float* ddepth, int B, int H, int W, void* stream);
int fd_project3d_fwd(const float* points, const float* K, const float* T, float* grid, int B, int H, int W,
                     float eps, void* stream);
int fd_project3d_bwd(const float* points, const float* K, const float* T, const float* dgrid, float* dpoints,
                     float* dT, int B, int H, int W, float eps, void* workspace, void* stream);
int fd_grid_sample_border_fwd(const float* img, const float* grid, float* out, int B, int C, int H, int W,
                              int Ho, int Wo, void* stream);
int fd_grid_sample_border_bwd(const float* img, const float* grid, const float* dout, float* dgrid, float* dimg,
                              int B, int C, int H, int W, int Ho, int Wo, void* stream);
/* transformation_from_parameters (layers.py:23-97): axisangle, translation [B,3] -> M [B,4,4]
 * (M = T @ R, or R^T @ T(-t) when invert); bwd: dM -> d_axisangle, d_translation [B,3]. */
int fd_pose_matrix_fwd(const float* axisangle, const float* translation, int invert, float* M, int B, void* stream);
int fd_pose_matrix_bwd(const float* axisangle, const float* translation, int invert, const float* dM,
                       float* d_axisangle, float* d_translation, int B, void* stream);
/* Cat_xy (layers.py:165-201): depth [B,1,H,W], inv_K [B,4,4] -> [B,3,H,W] = (X/30, Y/2, (Z-40)/40) */
int fd_cat_xy(const float* depth, const float* inv_K, float* out, int B, int H, int W, void* stream);
int fd_ssim_fwd(const float* x, const float* y, float* out, long planes, int H, int W, void* stream);
int fd_ssim_bwd(const float* x, const float* y, const float* dout, float* dx, long planes, int H, int W,
                void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------
 * Stage-2 (refiner.py) pseudo-3D pack -- replaces refiner.py:316-346 (default flags refine_a0='true',
 * catxy='true') and layers.Cat_xy (layers.py:165-201): per scale s, the stage-1 disparity max-pooled s
 * times -> bilinear to HxW -> depth -> rescaled by ONE ratio per batch,
 * median(beam[mask]*100) / median(depth[mask]) (torch.median = lower median; mask = beam > 0 inside
 * rows [crop_y0,crop_y1) x cols [crop_x0,crop_x1), reference 78:190 x 23:617) -> 6-channel map
 * [scaled_disp | x/30, y/2, (z-40)/40 | max-pooled 2-channel LiDAR].  No gradient (the reference
 * builds these under no_grad).
 *   disp0 [B,1,H,W]; beam [B,1,H,W]; two_cha [B,2,H,W] NCHW; inv_K[s] [B,4,4] of scale s;
 *   out[s] [B,H>>s,W>>s,6] NHWC; ratios [4] (out, one per scale).
 * fd_masked_median: out[0] = lower median of x[i]*scale over {i in window : mask_src[i] > 0} (NaN when
 * the set is empty) -- the torch.median(x[mask]) of refiner.py:332.
 * ---------------------------------------------------------------------------------------- */
size_t fd_refine_pack_workspace_bytes(int B, int H, int W);
int fd_refine_pack(const float* disp0, const float* beam, const float* two_cha, const float* const inv_K[4],
                   int B, int H, int W, int crop_y0, int crop_y1, int crop_x0, int crop_x1, float min_depth,
                   float max_depth, float* const out[4], float* ratios, void* workspace, void* stream);
size_t fd_masked_median_workspace_bytes(int B, int H, int W);
/* Depth error metrics with the masking, median scaling and clamping of Trainer.compute_depth_losses
 * (trainer.py:598-630; layers.compute_depth_errors, layers.py:284-302) and of evaluate_depth.py's per-frame
 * loop (evaluate_depth.py:42-60, 344-378, 470-478), entirely on the device:
 *   mask  = gt > mask_lo && gt < mask_hi inside rows [y0,y1) x cols [x0,x1);
 *   p     = pred_is_disp ? 1 / pred : clamp(pred, pre_min, pre_max)      (pred already has gt's size);
 *   ratio = median(gt[mask]) / median(p[mask]) when median_scaling (numpy_median: numpy.median, i.e. the mean
 *           of the middle two for even counts; else torch.median's lower median);
 *   p     = clamp(p * ratio, pred_min, pred_max);
 *   out[0..6] = abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3; out[7] = masked pixel count; out[8] = ratio. */
size_t fd_depth_errors_workspace_bytes(int B, int H, int W);
int fd_depth_errors(const float* gt, const float* pred, int B, int H, int W, int y0, int y1, int x0, int x1,
                    float mask_lo, float mask_hi, int pred_is_disp, float pre_min, float pre_max,
                    int median_scaling, int numpy_median, float pred_min, float pred_max, float* out,
                    void* workspace, void* stream);
int fd_masked_median(const float* x, const float* mask_src, int B, int H, int W, int y0, int y1, int x0,
                     int x1, float scale, float* out, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------
 * Network operators (NHWC fp32).  Replace the ATen/cuDNN ops under networks/resnet_encoder.py,
 * networks/depth_decoder.py, networks/pose_decoder.py, networks/pose_cnn.py, layers.ConvBlock /
 * Conv3x3 (layers.py:100-130) and torchvision's BasicBlock/Bottleneck.
 * ---------------------------------------------------------------------------------------- */
enum { FD_ACT_NONE = 0, FD_ACT_RELU = 1, FD_ACT_ELU = 2, FD_ACT_SIGMOID = 3, FD_ACT_TANH = 4 };

/* (x-0.45)/0.225 + NCHW->NHWC  (resnet_encoder.py:94) */
int fd_prep_input(const float* x_nchw, float* y_nhwc, int B, int C, int H, int W, float mean,
                  float std, void* stream);

/* The 7x7/2 stem as a GEMM (resnet_encoder.py:94-95): A [B*Ho*Wo, Kpad] = normalised im2col rows of
 * the NCHW input, k = (kh*KW+kw)*C + c, zero-filled up to Kpad (multiple of 32), so that the
 * tensor-core 1x1 path (fd_conv2d_fwd_tc / fd_conv2d_wgrad_tc on [B,Ho,Wo,Kpad]) runs it.
 * fd_pad_rows pads weight rows [rows,cols_src] -> [rows,cols_dst] (accumulate=0), or folds a padded
 * gradient back (cols_dst < cols_src, accumulate=1). */
int fd_stem_im2col(const float* x_nchw, float* A, int B, int C, int H, int W, int KH, int KW, int stride,
                   int pad, int Kpad, float mean, float std, void* stream);
int fd_pad_rows(const float* src, float* dst, int rows, int cols_src, int cols_dst, int accumulate,
                void* stream);

/* y = act(conv(x, w) + bias); zero padding `pad`.  x [B,H,W,Cin], w [Cout,KH,KW,Cin],
 * y [B,Ho,Wo,Cout], Ho = (H+2*pad-KH)/stride+1. */
int fd_conv2d_fwd(const float* x, const float* w, const float* bias, float* y, int B, int H, int W,
                  int Cin, int Cout, int KH, int KW, int stride, int pad, int act, void* stream);
/* dx = conv_transpose(dy, w).  wt = fd_weight_transpose(w): [Cin,KH,KW,Cout]. */
int fd_conv2d_dgrad(const float* dy, const float* wt, float* dx, int B, int H, int W, int Cin,
                    int Cout, int KH, int KW, int stride, int pad, void* stream);
/* dw [Cout,KH,KW,Cin] += sum over pixels; dw must be zero-initialised (or hold an accumulator) */
int fd_conv2d_wgrad(const float* x, const float* dy, float* dw, int B, int H, int W, int Cin,
                    int Cout, int KH, int KW, int stride, int pad, void* stream);
int fd_weight_transpose(const float* w, float* wt, int Cout, int taps, int Cin, void* stream);

/* Tensor-core path (tcgen05.mma kind::tf32, TMEM accumulators, 3xTF32 error-compensated split:
 * fp32-level accuracy).  Eligible when the gathered channel count is a multiple of 32 and the
 * produced channel count a multiple of 16 (fd_conv2d_tc_supported).  w_lo = w - tf32(w)
 * (fd_tf32_split); for the data gradient wt / wt_lo come from fd_weight_transpose_split. */
int fd_conv2d_tc_supported(int gathered_channels, int produced_channels);
int fd_tf32_split(const float* w, float* w_lo, long n, void* stream);
/* fd_conv2d_fwd_tc that also adds the per-channel sum and sum of squares of y into stats [2][Cout] (fp64,
 * zeroed by the caller): the BatchNorm batch statistics come out of the convolution epilogue (no bias,
 * no activation; stats may be NULL). */
int fd_conv2d_fwd_tc_stats(const float* x, const float* w, const float* w_lo, const float* bias, float* y,
                           int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                           int act, double* stats, void* stream);
/* ... for every conv weight of a flat parameter buffer in ONE launch: desc (device, int64 [n][5]) rows are
 * {first element of tensor j in the concatenated index space, offset of tensor j in the flat buffers,
 *  Cout, taps, Cin}; the transposed tensors land at the same offsets of flat_t / flat_tlo. */
int fd_weight_transpose_split_batched(const float* flat, float* flat_t, float* flat_tlo, const long* desc,
                                      int n, long total, void* stream);
int fd_weight_transpose_split(const float* w, float* wt, float* wt_lo, int Cout, int taps, int Cin,
                              void* stream);
int fd_conv2d_fwd_tc(const float* x, const float* w, const float* w_lo, const float* bias, float* y,
                     int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                     int act, void* stream);
int fd_conv2d_dgrad_tc(const float* dy, const float* wt, const float* wt_lo, float* dx, int B, int H,
                       int W, int Cin, int Cout, int KH, int KW, int stride, int pad, void* stream);
/* dw [Cout,KH,KW,Cin] += ... on the tensor cores (Cin and Cout multiples of 32); dw must be
 * zero-initialised or hold an accumulator. */
int fd_conv2d_wgrad_tc(const float* x, const float* dy, float* dw, int B, int H, int W, int Cin,
                       int Cout, int KH, int KW, int stride, int pad, void* stream);
/* Single-output-channel 3x3 stride-1 convolutions (the disparity heads, reference
 * networks/depth_decoder.py:52-58 + layers.py:115-130) as streaming kernels: x [B,H,W,Cin] NHWC,
 * w [1,3,3,Cin], y / dy [B,Ho,Wo] with Ho = H + 2*pad - 2.  wgrad ADDS into dw (atomics). */
int fd_conv2d_cout1_supported(int Cin, int Cout, int KH, int KW, int stride);
int fd_conv2d_cout1_fwd(const float* x, const float* w, const float* bias, float* y, int B, int H, int W,
                        int Cin, int pad, int act, void* stream);
int fd_conv2d_cout1_dgrad(const float* dy, const float* w, float* dx, int B, int H, int W, int Cin, int pad,
                          void* stream);
int fd_conv2d_cout1_wgrad(const float* x, const float* dy, float* dw, int B, int H, int W, int Cin, int pad,
                          void* stream);
/* 16-output-channel 3x3 stride-1 layers with Cin in {16, 32} (DepthDecoder upconv(0,0), upconv(0,1),
 * reference networks/depth_decoder.py:20-50) as direct convolutions on the CUDA cores; same tensor
 * conventions as fd_conv2d_*.  fwd takes Cin = 16 only (Cin = 32 forward runs on the tensor cores);
 * wgrad ADDS into dw [16,3,3,Cin]. */
int fd_conv2d_c16_supported(int Cin, int Cout, int KH, int KW, int stride);
int fd_conv2d_c16_fwd(const float* x, const float* w, const float* bias, float* y, int B, int H, int W, int Cin,
                      int pad, int act, void* stream);
int fd_conv2d_c16_dgrad(const float* dy, const float* w, float* dx, int B, int H, int W, int Cin, int pad,
                        void* stream);
int fd_conv2d_c16_wgrad(const float* x, const float* dy, float* dw, int B, int H, int W, int Cin, int pad,
                        void* stream);
/* Profiling aid: with FD_TC2_FLAGS bit 7 set, CTA (0,0) of the tensor-core conv kernels writes
 * clock64() stamps of its pipeline roles into this device buffer ((3*256*4 + 8) int64); NULL = off. */
int fd_debug_set_conv_trace(void* device_buffer);
/* dpre = dy * act'(.) computed from the activation output y; dbias[c] += sum_m dpre[m,c] */
int fd_act_bwd(const float* y, const float* dy, float* dpre, float* dbias, long M, int C, int act,
               void* stream);

/* BatchNorm2d (+ optional residual add, + optional ReLU) over M = B*H*W rows of C channels.
 * training: batch statistics, running stats updated (momentum, unbiased var).
 * ws: fd_bn_workspace_bytes(C) bytes of scratch (the first 2*C doubles end up holding the per-channel sums;
 * the rest are per-block partial rows folded by the last block to finish -- deterministic, no contended
 * atomics).  save_mean/save_rstd [C] are outputs used by the backward.
 * stat_weight < 0: running = (1-momentum)*running + momentum*stat (nn.BatchNorm2d).  stat_weight >= 0:
 * running += stat_weight*stat with atomics -- for calls of one layer that run concurrently inside an
 * optimiser step, after the caller has decayed the buffers by (1-momentum)^n_calls once.
 * stats_ready != 0: ws already holds the per-channel sum / sum of squares of x (written by
 * fd_conv2d_fwd_tc_stats into a zeroed buffer); the statistics pass over x is skipped.
 * accumulate_param_grads: dgamma/dbeta are added (atomically) into existing buffers instead of set. */
size_t fd_bn_workspace_bytes(int C);
int fd_bn_fwd(const float* x, const float* residual, const float* gamma, const float* beta,
              float* running_mean, float* running_var, int training, float momentum, float eps,
              int relu, float* y, float* save_mean, float* save_rstd, double* ws, long M, int C,
              float stat_weight, int stats_ready, void* stream);
int fd_bn_bwd(const float* x, const float* y, const float* dy, const float* gamma,
              const float* save_mean, const float* save_rstd, int relu, int training, float* dx,
              float* dresidual, float* dgamma, float* dbeta, double* ws, long M, int C,
              int accumulate_param_grads, void* stream);
/* Backward of a residual-free BatchNorm + ReLU without the saved output: the ReLU mask y > 0 is re-evaluated
 * from x, mean, rstd, gamma, beta with the forward's exact rounding (one tensor read less in each of the two
 * passes).  gamma / beta must still hold the forward's values.  fd_bn_bwd_xmask_ok(M, C) == 0: tensor too small
 * (single-kernel path of fd_bn_bwd), use fd_bn_bwd with y. */
int fd_bn_bwd_xmask_ok(long M, int C);
int fd_bn_bwd_xmask(const float* x, const float* dy, const float* gamma, const float* beta,
                    const float* save_mean, const float* save_rstd, int training, float* dx, float* dgamma,
                    float* dbeta, double* ws, long M, int C, int accumulate_param_grads, void* stream);

/* nn.MaxPool2d(3, 2, 1)  (ResNet stem) */
int fd_maxpool3x3s2_fwd(const float* x, float* y, unsigned char* idx, int B, int H, int W, int C,
                        void* stream);
int fd_maxpool3x3s2_bwd(const float* dy, const unsigned char* idx, float* dx, int B, int H, int W,
                        int C, void* stream);

/* Decoder glue: out = reflect_pad_p( concat_c( seg_0, seg_1, ... ) ), where each segment is
 * (a [+ b]) optionally nearest-upsampled x2  (layers.upsample, torch.cat, ReflectionPad2d in
 * depth_decoder.py:73-85 / layers.py:121-130).  out [B,H+2p,W+2p,sum C]. */
typedef struct fd_segment {
  const float* a;
  const float* b; /* nullable: added to a */
  int C;
  int up; /* 1: source is [B,H/2,W/2,C] */
} fd_segment;
int fd_assemble_fwd(const fd_segment* segs_host, int nseg, float* out, int B, int H, int W, int pad,
                    void* stream);
/* dseg[i] [B,Hs,Ws,C_i] = adjoint (gather form, deterministic) */
int fd_assemble_bwd(const float* dout, float* const* dsegs_host, const int* C_host,
                    const int* up_host, int nseg, int B, int H, int W, int pad, void* stream);
/* The same with an optional second destination per segment (dsegs2_host[i] nullable): the gradient of an `a + b`
 * segment as two separate tensors, one per operand.  The operands of the skip connections live on different
 * streams' autograd branches, and one tensor handed to both is open to the autograd engine accumulating into it in
 * place on one stream while the other still reads it. */
int fd_assemble_bwd2(const float* dout, float* const* dsegs_host, float* const* dsegs2_host, const int* C_host,
                     const int* up_host, int nseg, int B, int H, int W, int pad, void* stream);

int fd_add(const float* a, const float* b, float* out, long n, void* stream);
/* out = relu(a + b): the residual join of a BasicBlock / Bottleneck whose BatchNorms are folded into the
 * convolutions (inference: fusiondepth_b200.evaluation) */
int fd_add_relu(const float* a, const float* b, float* out, long n, void* stream);
/* y[b,c] = scale * mean_hw x[b,:,c]  (pose_decoder.py:44-46) */
int fd_mean_hw_fwd(const float* x, float* y, int B, int HW, int C, float scale, void* stream);
int fd_mean_hw_bwd(const float* dy, float* dx, int B, int HW, int C, float scale, void* stream);

/* torch.optim.Adam update on flat buffers (trainer.py:129): g is multiplied by grad_scale first.
 * state: 4 x int32 device words, state[0] = number of steps taken so far (incremented here, on
 * the device, so that a captured CUDA graph advances the bias correction on every replay).
 * lr < 0: the learning rate is read from state[3] (float bits) on the device instead, so that a learning
 * rate schedule (StepLR, trainer.py:131-132, 266) reaches a captured graph. */
int fd_adam_step(float* p, const float* g, float* m, float* v, long n, float lr, float beta1,
                 float beta2, float eps, int* state, float grad_scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * On-device colour side of the data producer: datasets/mono_dataset.py:85-104
 * (`preprocess`: pyramid of transforms.Resize(.., Image.ANTIALIAS), each scale from the previous one;
 * ColorJitter; ToTensor) and 156-206 (flip of the native image in get_color), on uint8 HWC images in device
 * memory, bit-identical to Pillow 12 / torchvision 0.26 `_functional_pil` (oracle/data_oracle.py).
 *
 * fd_lanczos_ksize / fd_lanczos_coeffs: HOST helpers.  Pillow's 22-bit fixed-point Lanczos-3 window for
 *   resizing `in_size` samples to `out_size`: bounds[out][2] = (first input sample, count),
 *   kk[out][ksize] int32 weights (zero padded).  The caller uploads them once per size pair.
 * fd_resize_lanczos_u8: src [B,Hin,Win,3] -> dst [B,Hout,Wout,3]; horizontal pass first into
 *   tmp [B,Hin,Wout,3] (uint8 intermediate, as PIL), then vertical.  flip [B] uint8 or NULL: image b is
 *   mirrored left-right before the resize.
 * fd_color_jitter_u8: torchvision ColorJitter in place with explicit parameters: order [B,4] = the
 *   permutation of (0 brightness, 1 contrast, 2 saturation, 3 hue) drawn for image b, factors [B,4]
 *   indexed by op: blend factors for ops 0-2, for hue the uint8 shift int32(hue * 255) & 255 as a float;
 *   NaN = op skipped.  lum_sums: B x uint64 workspace.
 * fd_image_to_tensor: [B,H,W,3] uint8 -> [B,3,H,W] float32 = u8 / 255 (transforms.ToTensor). */
int fd_lanczos_ksize(int in_size, int out_size);
int fd_lanczos_coeffs(int in_size, int out_size, int* bounds_host, int* kk_host);
int fd_resize_lanczos_u8(const unsigned char* src, unsigned char* tmp, unsigned char* dst, int B, int Hin,
                         int Win, int Hout, int Wout, const int* bounds_w, const int* kk_w, int ksize_w,
                         const int* bounds_h, const int* kk_h, int ksize_h, const unsigned char* flip,
                         void* stream);
int fd_color_jitter_u8(unsigned char* img, int B, int H, int W, const int* order, const float* factors,
                       unsigned long long* lum_sums, void* stream);
int fd_image_to_tensor(const unsigned char* img, float* out, int B, int H, int W, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Graph-based Depth Correction: gdc_old.py:74-250 as called by inf_gdc.py:81
 * (k = 10, W_tol = 3e-5, recon_tol = 5e-4, method = 'cg'); fp64 throughout, like the reference.
 *
 * fd_gdc_select: pred, gt [H,W] (gt <= 0: no return) -> cls [H*W] uint8 (0 untouched, 1 pseudo-LiDAR point to
 *   correct = gdc_old.py `pred_mask`, 2 LiDAR anchor = `gt_mask`) and points [H*W,3] (back-projection with the
 *   PREDICTED depth, kitti_util_from_pse.py:204-215; calib6 = c_u, c_v, f_u, f_v, b_x, b_y on the host;
 *   th_lo / th_hi = radians(consider_range)).  The caller orders the points class 1 first, then class 2, each in
 *   raster order (gdc_old.py:163-168).
 * fd_gdc_knn: exact k nearest neighbours of every point among the others, ascending distance (:171-172).
 * fd_gdc_weights: the (k+2)x(k+2) constrained reconstruction systems (:174-186) -> weights [n,k].
 * fd_gdc_rhs / fd_gdc_apply / fd_gdc_apply_t: b, y = A x, z = A^T y for A = [I - W_PLPL ; W_PLL] (:196-222);
 *   entries / col_ptr: the (row, slot) indices with neighbour < n_pl, sorted by neighbour (rows ascending).
 * fd_gdc_dot / fd_gdc_cg_update / fd_gdc_cg_dir: the conjugate-gradient recurrence of scipy.sparse.linalg.cg on
 *   device scalars scal = {rho, p.q, rho_new}: update: x += (rho / p.q) p, r -= (rho / p.q) q, rho_new = r.r;
 *   dir: p = r + (rho_new / rho) p, rho = rho_new.  One block each: fixed summation order. */
int fd_gdc_select(const double* pred, const double* gt, int H, int W, const double* calib6_host, double th_lo,
                  double th_hi, unsigned char* cls, double* points, void* stream);
int fd_gdc_knn(const double* points, int n, int k, int* neighbors, void* stream);
int fd_gdc_weights(const double* x_info, const int* neighbors, int n, int k, double w_tol, double* weights,
                   void* stream);
int fd_gdc_rhs(const double* weights, const int* neighbors, int n, int n_pl, int k, const double* gt_info, double* b,
               void* stream);
int fd_gdc_apply(const double* weights, const int* neighbors, int n, int n_pl, int k, const double* x, double* y,
                 void* stream);
int fd_gdc_apply_t(const double* weights, const long* entries, const long* col_ptr, int n_pl, int k,
                   const double* y, double* z, void* stream);
int fd_gdc_dot(const double* a, const double* b, int n, double* out, void* stream);
int fd_gdc_cg_update(double* x, double* r, const double* p, const double* q, int n, double* scal, void* stream);
int fd_gdc_cg_dir(double* p, const double* r, int n, double* scal, void* stream);

#ifdef __cplusplus
}
#endif
#endif
