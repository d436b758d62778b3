/* libgdrn_b200.so -- C ABI of the B200-native GDR-Net hot path.
 *
 * Drop-in boundary: the reference has no native code on this path -- every op below is
 * today an ATen / cuDNN / cuBLAS call issued by the `nn.Module`s cited per function.  The in-repo
 * precedent for a C-ABI extension is core/csrc/fps/src/ext.h:1-14 (raw pointers, caller-owned
 * buffers).  Conventions:
 *
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch allocates everything);
 *     kernels never allocate, no pointer is retained after return;
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *   - return 0 on success, <0 on error (GDRN_ERR_ARG / GDRN_ERR_CUDA) with a thread-local message
 *     from gdrn_last_error(); nothing throws across the ABI; entry points are re-entrant;
 *   - activations are NHWC bf16 "planar pairs": a hi plane and an optional lo plane with
 *     value = hi + lo.  nsplit = 1: bf16 tensor-core math on the hi plane only (lo pointers NULL);
 *     nsplit = 3: fp32-faithful mode, three tcgen05 MMAs per k-step (hi*hi + hi*lo + lo*hi).
 */
#ifndef GDRN_B200_H
#define GDRN_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define GDRN_OK 0
#define GDRN_ERR_ARG (-1)
#define GDRN_ERR_CUDA (-2)

const char* gdrn_last_error(void);
long gdrn_launch_count(void); /* kernels launched by this library since load */
int gdrn_abi_version(void);
int gdrn_storage_format(void); /* 0: bf16 (hi, lo) planes; 1: fp16 planes, lo = fp16((x - hi) * gdrn_lo_scale()) */
float gdrn_lo_scale(void);
int gdrn_set_2cta(int on); /* A/B: cta_group::2 pair tiles (256x256 1-pass / 256x128 3-pass) for eligible conv layers; default on */
long gdrn_2cta_launch_count(void); /* launches of the 2-CTA kernel since load */
int gdrn_set_pdl(int on); /* A/B: programmatic dependent launch of the GEMM kernels (prologue under the predecessor's tail); default on, GDRN_PDL=0 disables */
int gdrn_set_wgrad_2cta(int on); /* A/B: cta_group::2 pair tiles (256 co x 128/256 ci) for single-plane weight gradients with Cout % 256 == 0; default OFF (bit-equal to the 1-CTA kernel, no in-situ gain) */
long gdrn_wgrad_2cta_launch_count(void);
int gdrn_last_gemm_variant(void); /* BLOCK_N*10 + nsplit (+10000: 2-CTA kernel) of this thread's last conv/gemm forward launch */

/* ---- tcgen05 implicit-GEMM convolution, forward (and dgrad with flipped/transposed weights) ------------
 * replaces nn.Conv2d / nn.ConvTranspose2d forward: resnet_backbone.py:21-49,69-76 (torchvision BasicBlock),
 * cdpn_rot_head_region.py:82-135, conv_pnp_net.py:76-80.
 * x [N,H,W,Cin] (Cin % 64 == 0), w packed [Cout_pad][KH*KW*Cin] (gdrn_pack_weight), outputs row-major
 * [N*Ho*Wo][ldc]: bf16 planes (y_hi/y_lo) and/or fp32 (y_f32).  bias[Cout] / act (1 = LeakyReLU 0.1, 2 = ReLU) optional;
 * res_hi/res_lo: optional residual planes [N*Ho*Wo][ldc] added before the activation.  With BatchNorm folded into the
 * weights (gdrn_bn_fold_batched + row_scale of the pack jobs) and bias = the folded shift this is the eval-mode fused
 * conv + BN (+ identity) + ReLU of resnet_backbone.py:69-76 / torchvision BasicBlock / cdpn_rot_head_region.py:82-125.
 * stats (optional) [2][Cout] fp32 += per-channel sum and sum of squares of the fp32 accumulators
 * (BatchNorm batch statistics, reference get_norm("BN") layers). */
int gdrn_conv_fwd(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, void* y_hi, void* y_lo,
                  float* y_f32, const float* bias, const void* res_hi, const void* res_lo, float* stats, int N, int H, int W,
                  int Cin, int Cout, int Cout_pad, int KH, int KW, int stride, int pad, int ldc, int act, int nsplit,
                  void* stream);

/* Data gradient of a stride-2 conv (k3 p1 / k1 p0) by output-parity phases over the un-dilated dY: replaces the reference's
 * cuDNN conv-backward-data for the stride-2 layers (resnet_backbone.py layerN.0.conv1 / downsample, conv_pnp_net.py:76-80).
 * du [N][Ho][Wo][Cy], w = dgrad-packed weights (flipped taps), dx [N][2Ho][2Wo][ldc]; k1: dx must be pre-zeroed.
 * The same arithmetic is the FORWARD of nn.ConvTranspose2d(k3, s2, p1, output_padding 1) (cdpn_rot_head_region.py:82-91; its IOHW
 * weight is the OIHW weight of the transposed conv): bias / act / stats (optional) then apply to the output like gdrn_conv_fwd. */
int gdrn_conv_dgrad_s2(const void* du_hi, const void* du_lo, const void* w_hi, const void* w_lo, void* dx_hi, void* dx_lo,
                       const float* bias, float* stats, int N, int Ho, int Wo, int Cy, int Cx, int Cx_pad, int K, int pad, int ldc,
                       int act, int nsplit, void* stream);

/* ---- plain GEMM  y[M][N] = a[M][K] * w[N_pad][K]^T (+bias, act); replaces nn.Linear (conv_pnp_net.py:89-92,
 * 152-156), the 7x7 stem over its im2col matrix (resnet_backbone.py:69) and their dgrads. */
int gdrn_gemm_fwd(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, void* y_hi, void* y_lo,
                  float* y_f32, const float* bias, float* stats, int M, int N, int N_pad, int K, int ldc, int act,
                  int nsplit, void* stream);

/* ---- weight gradients (autograd wgrad of the layers above).  Split-K partials are written to the fp32
 * workspace ws [ksplit][ceil(Cout/128)*128][KH*KW*Cin]; reduce with gdrn_unpack_wgrad.
 * ws == NULL: size query only (*ws_need_out floats, *ksplit_out). ksplit <= 0: automatic. */
int gdrn_conv_wgrad(const void* dy_hi, const void* dy_lo, const void* x_hi, const void* x_lo, float* ws, long ws_floats,
                    long* ws_need_out, int* ksplit_out, int N, int H, int W, int Cin, int Cout, int KH, int KW,
                    int stride, int pad, int ksplit, int nsplit, void* stream);
int gdrn_gemm_wgrad(const void* dy_hi, const void* dy_lo, const void* x_hi, const void* x_lo, float* ws, long ws_floats,
                    long* ws_need_out, int* ksplit_out, long P, int M, int Ntot, int ksplit, int nsplit, void* stream);

/* ---- parameter layout conversion.  dst[o][tap*ipad + i] = src[o*so + i*si + r*sr + s*ss] (r,s flipped if flip).
 * Source layouts: Conv2d OIHW, ConvTranspose2d IOHW, Linear [out][in] (state_dict of the reference). */
int gdrn_pack_weight(const float* src, void* dst_hi, void* dst_lo, int O, int I, int KH, int KW, int opad, int ipad,
                     int krow, long so, long si, long sr, long ss, int flip, void* stream);
/* all per-step re-packs in one launch; jobs_dev = device array of 104-byte PackJob records (csrc/pack.cu), each with an
 * optional per-output-row scale (eval-mode BatchNorm folding) */
int gdrn_pack_weight_batched(const void* jobs_dev, int njobs, long total_blocks, void* stream);
int gdrn_unpack_wgrad(const float* ws, float* grad, int O, int I, int KH, int KW, int ipad, int krow, int ksplit, long ks_stride,
                      long so, long si, long sr, long ss, int flip, int accumulate, void* stream);
/* im2col of the 7x7/2 stem: x NCHW fp32 [B,3,H,W] -> [B*H/2*W/2][192] bf16 planes, k = (r*7+s)*3 + c */
int gdrn_stem_im2col(const float* x, void* a_hi, void* a_lo, int B, int H, int W, void* stream);

/* ---- BatchNorm2d(eps, momentum) train/eval (+ReLU, + residual add) and backward -- detectron2/torch
 * BatchNorm2d via core/utils/layer_utils.py:17-39; torchvision BasicBlock residual. */
/* eval-mode folding of ALL BatchNorm layers in one launch: jobs = device array of 64-byte records {gamma, beta,
 * running_mean, running_var, scale_out, shift_out (pointers); int C; float eps; long pad} */
int gdrn_bn_fold_batched(const void* jobs_dev, int njobs, void* stream);
int gdrn_bn_finalize(const float* stats, const float* gamma, const float* beta, float* running_mean, float* running_var,
                     float* scale, float* shift, float* mean_out, float* invstd_out, int C, float count, float eps,
                     float momentum, int train, void* stream);
/* finalize + apply in one kernel (what the engine uses): batch statistics from the conv epilogue -> y, running stats.
 * relu_mask_out (optional, with relu): uint8 [rows*C/8], bit j of byte i = [pre-activation of element 8i+j > 0]; BatchNorm
 * backward (relu_mask) reads it instead of the activation to rebuild the ReLU mask (1/8 B instead of 2 B per element). */
int gdrn_bn_fwd(const void* x_hi, const void* x_lo, const void* r_hi, const void* r_lo, void* y_hi, void* y_lo,
                const float* stats, const float* gamma, const float* beta, float* running_mean, float* running_var,
                float* mean_out, float* invstd_out, void* relu_mask_out, long rows, int C, float eps, float momentum, int train,
                int relu, void* stream);
int gdrn_bn_act(const void* x_hi, const void* x_lo, const void* r_hi, const void* r_lo, void* y_hi, void* y_lo,
                const float* scale, const float* shift, long rows, int C, int relu, void* stream);
/* flags: bit 0 = the ReLU mask is recomputed from u (plain conv-BN-ReLU: pass y_hi = NULL and beta); bit 1 = `sums` is
 * already zero (the caller cleared all layers' accumulators in one memset); bit 2 = deterministic two-stage reductions */
int gdrn_bn_bwd(const void* ga_hi, const void* ga_lo, const void* gb_hi, const void* gb_lo, const void* y_hi,
                const void* u_hi, const void* u_lo, const float* mean, const float* invstd, const float* gamma,
                const float* beta, float* sums, void* du_hi, void* du_lo, void* gout_hi, void* gout_lo, float* dgamma,
                float* dbeta, const void* relu_mask, float* det_ws, long rows, int C, int train, int flags, void* stream);
/* deterministic mode (no atomics, fixed summation order; Engine.deterministic): gdrn_bn_stats computes the BatchNorm batch
 * statistics [2][C] (sum, sum of squares) from the tensor in two ordered stages instead of the GEMM epilogue's atomics;
 * gdrn_bn_bwd with flags bit 2 does the same for its reductions.  ws / det_ws: >= 2 * C * gdrn_det_parts() floats. */
int gdrn_det_parts(void);
int gdrn_bn_stats(const void* u_hi, const void* u_lo, float* ws, float* stats, long rows, int C, void* stream);

/* ---- MaxPool2d(3,2,1) resnet_backbone.py:72; UpsamplingBilinear2d(x2) cdpn_rot_head_region.py:102;
 * zero insertion (stride-2 transposed convs); GroupNorm(32)+ReLU conv_pnp_net.py:76-80 */
/* arg_out / arg_in: uint8 [B,H/2,W/2,C] window-local arg-max codes (first maximum, like ATen's saved indices) */
int gdrn_maxpool_fwd(const void* x_hi, const void* x_lo, void* y_hi, void* y_lo, void* arg_out, int B, int H, int W, int C,
                     void* stream);
int gdrn_maxpool_bwd(const void* arg_in, const void* g_hi, const void* g_lo, void* dx_hi, void* dx_lo, int B, int H, int W,
                     int C, void* stream);
int gdrn_upsample2x_fwd(const void* x_hi, const void* x_lo, void* y_hi, void* y_lo, int B, int H, int W, int C, void* stream);
int gdrn_upsample2x_bwd(const void* g_hi, const void* g_lo, void* dx_hi, void* dx_lo, int B, int H, int W, int C, void* stream);
int gdrn_zero_insert(const void* x_hi, const void* x_lo, void* y_hi, void* y_lo, int B, int H, int W, int C, int mode,
                     void* stream);
int gdrn_gn_relu_fwd(const void* u_hi, const void* u_lo, void* y_hi, void* y_lo, const float* gamma, const float* beta,
                     float* stats, int B, int HW, int C, int G, float eps, void* stream);
int gdrn_gn_relu_bwd(const void* g_hi, const void* g_lo, const void* y_hi, const void* u_hi, const void* u_lo,
                     const float* gamma, const float* stats, void* du_hi, void* du_lo, float* dgamma, float* dbeta, int B,
                     int HW, int C, int G, void* stream);
int gdrn_add2(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, void* o_hi, void* o_lo, long n,
              void* stream);
/* bias gradients (column sums of a [rows][ld] planar tensor) and LeakyReLU(0.1) backward (conv_pnp_net.py:93,152-153) */
int gdrn_colsum(const void* x_hi, const void* x_lo, float* out, long rows, int ld, void* stream);
int gdrn_leaky_bwd(const void* g_hi, const void* g_lo, const void* y_hi, void* o_hi, void* o_lo, long n, void* stream);
int gdrn_f32_to_planes(const float* x, void* y_hi, void* y_lo, long n, void* stream);
int gdrn_planes_to_f32(const void* x_hi, const void* x_lo, float* y, long n, void* stream);

/* ---- geometry glue + per-pixel losses (GDRN.py:156-169, 341-400; conv_pnp_net.py:120-125).
 * logits fp32 [B*HW][72] (0 mask | 1..3 xyz | 4 bg | 5..68 regions), Patch-PnP input bf16 [B*HW][128]:
 * with_2d = 1: xyz | coord2d | softmax64 (69 channels, PNP_NET.WITH_2D_COORD), 0: xyz | softmax64 (67 channels). */
int gdrn_head_glue_fwd(const float* logits, const float* coord2d, const float* extents, void* out_hi, void* out_lo, int B,
                       int HW, int with_2d, void* stream);
/* standalone ConvPnPNet.forward input (conv_pnp_net.py:111-125): NCHW fp32 coor_feat [B][c_feat][HW] (+ region
 * [B][c_reg][HW]) -> NHWC planes [B*HW][128]; xyz de-normalised by extents when c_feat is 3 or 5. */
int gdrn_pnp_pack_input(const float* coor_feat, int c_feat, const float* region, int c_reg, const float* extents,
                        void* out_hi, void* out_lo, int B, int HW, void* stream);
int gdrn_pixel_loss_fwd(const float* logits, const float* gt_xyz, const float* m_visib, const float* m_trunc,
                        const long long* labels, double* sums, int B, int HW, void* stream);
int gdrn_head_bwd(const float* logits, const float* gt_xyz, const float* m_visib, const float* m_trunc,
                  const long long* labels, const double* sums, const float* gw, const void* din_hi, const void* din_lo,
                  const float* extents, void* out_hi, void* out_lo, int B, int HW, int with_2d, void* stream);

/* ---- rot6d -> R, SITE translation, allo -> ego, PM loss (+ closest symmetric GT), centroid / z losses, re/te
 * (rot_reps.py:34-49, pose_from_pred_centroid_z.py:144-227, utils.py:208-236, pm_loss.py:82-114,
 *  pose_utils.py:430-482, GDRN.py:439-471, model_utils.py:40-52).  pred [B][ld_pred]: cols 0..5 rot6d, 6..8 t.
 * syms: device-resident table of symmetry rotations [rows][3][3] (built once per object set), sym_idx [B][2] = (first row,
 * count) per crop, count 0 = asymmetric; both NULL without PM_LOSS_SYM.  The K candidates are scanned by the whole CTA. */
int gdrn_pose_loss(const float* pred, int ld_pred, const float* cams, const float* centers, const float* whs,
                   const float* ratios, const float* extents, const float* points, const float* gt_rot,
                   const float* gt_trans, const float* gt_ratio, const float* syms, const int* sym_idx, const float* gw,
                   float* out_rot, float* out_trans, double* sums, float* vis, void* dy_hi, void* dy_lo, int B, int n_pts,
                   int do_loss, float eps, void* stream);
int gdrn_loss_finalize(const double* pix_sums, const double* pose_sums, const float* vis, float* losses, float* vis_out,
                       int B, int HW, int n_pts, void* stream);

/* ---- per-instance crop / target generation of the data loader on the device (core/gdrn_modeling/data_loader.py:487-560,
 * core/utils/data_utils.py:80-137 crop_resize_by_warp_affine, :213-219 xyz_to_region), cv2.warpAffine's fixed-point sampling
 * restated.  image_u8 [B][H][W][3] (BGR uint8), centers [B][2], scales [B] (float64 like the reference's aug_bbox output) -> roi_img [B][3][R][R] = bilinear crop / pixel_std.
 * xyz [B][H][W][3] (object coordinates, 0 = background), mask_visib / mask_trunc [B][H][W] (mask_trunc may be NULL),
 * extents [B][3], fps_points [B][n_fps][3] -> roi_xyz [B][3][R][R] (normalised), roi_mask_{trunc,visib,obj} [B][R][R],
 * roi_region [B][R][R] int64 (0 = background), roi_coord_2d [B][2][R][R]. */
int gdrn_roi_crop_image(const void* image_u8, const double* centers, const double* scales, float* roi_img, int B, int H, int W,
                        int out_res, float pixel_std, void* stream);
int gdrn_roi_targets(const float* xyz, const float* mask_visib, const float* mask_trunc, const double* centers, const double* scales,
                     const float* extents, const float* fps_points, int n_fps, float* roi_xyz, float* roi_mask_trunc,
                     float* roi_mask_visib, float* roi_mask_obj, long long* roi_region, float* roi_coord_2d, int B, int H, int W,
                     int out_res, void* stream);

/* ---- evaluator metrics on the device (lib/pysixd/pose_error.py:297-337 add / adi, :400-436 re / te; the reference computes
 * them per instance in numpy with a scipy KD-tree, gdrn_evaluator.py:316-436).  R_* [B][3][3], t_* [B][3], points [B][n_pts][3]
 * (model points per instance, metres) -> out [B][4] = (ADD, ADD-S = ADI or 0 when !want_adi, re in degrees, te). */
int gdrn_pose_errors(const float* R_est, const float* t_est, const float* R_gt, const float* t_gt, const float* points, int B,
                     int n_pts, int want_adi, float* out, void* stream);

/* ---- test-time RANSAC-PnP of the evaluator on the device (core/gdrn_modeling/gdrn_evaluator.py:316-436 `process_pnp_ransac`:
 * `get_img_model_points_with_coords2d` :89-126 + lib/pysixd/misc.py:145-194 `pnp_v2` = cv2.solvePnPRansac(EPNP, reprojErr 3, 100
 * iterations) per instance on the host).  mask [B][H*W] (mask_mode 0: used as is, 1: per-ROI min-max normalised like
 * engine_utils.py:113-118 for the L1 mask head, 2: sigmoid), xyz [B][3][H*W] in [0,1], coord2d [B][2][H*W] in [0,1],
 * extents [B][3], im_wh [B][2] = (im_W, im_H), K [B][3][3].  out_pose [B][3][4] = [R | t]; out_info [B][4] = (selected points,
 * inliers, RANSAC iterations run, ok); out_inliers (optional) [B][H*W] flags indexed like cv2's inlier list (position in the
 * row-major list of selected points).  Follows OpenCV's algorithm (RNG-driven 5-point subsets, EPnP minimal + final solver,
 * adaptive iteration count): same inlier sets / poses as cv2 up to round-off. */
long gdrn_pnp_ransac_workspace_bytes(int B, int HW, int iters);
int gdrn_pnp_ransac(const float* mask, const float* xyz, const float* coord2d, const float* extents, const float* im_wh,
                    const float* K, int B, int H, int W, int mask_mode, float mask_thr, double reproj_err, int iters,
                    double confidence, void* workspace, long workspace_bytes, float* out_pose, int* out_info,
                    unsigned char* out_inliers, void* stream);

/* ---- fused Ranger step (gradient centralisation + RAdam + Lookahead), lib/torch_utils/solver/ranger.py:100-200, for all
 * tensors of a param group in ONE launch.  jobs: device array of 64-byte records {float* p; const float* g; float* m; float* v;
 * float* slow; long numel; int row_len; int pad[3]} (row_len > 0: centralise rows of that length);
 * blocks: device array of {int job; int count; long begin} (rows or elements per CTA).  The scalar RAdam rectification
 * (step_lr = step_size * lr, adaptive = N_sma > threshold) is computed by the host; wd_lr = weight_decay * lr;
 * grad_scale multiplies every gradient (1/loss-scale). */
int gdrn_ranger_step(const void* jobs_dev, const void* blocks_dev, int nblocks, float step_lr, float wd_lr, float beta1,
                     float beta2, float eps, float alpha, float grad_scale, int adaptive, int lookahead, void* stream);
/* x[i] *= s over an fp32 buffer (loss-scale removal from a slice of the flat gradient buffer) */
int gdrn_scale_f32(float* x, long n, float s, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GDRN_B200_H */
