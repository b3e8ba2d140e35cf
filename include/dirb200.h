/* dirb200.h -- C ABI of libdirb200.so: the B200-native (sm_100a) hot path of
 * YyzHarry/imbalanced-regression (ResNet-50 + FDS + LDS training step).
 *
 * The reference has no FFI: its "plugin boundary" for this path is a set of
 * Python callables.  Each entry point below names the
 * reference function it replaces (file:line under /root/reference); the
 * Python host mirrors in imbalanced-regression_b200/{fds,loss,utils,resnet,
 * datasets}.py keep the reference's names/signatures and call these through
 * ctypes.
 *
 * Conventions
 *  - plain C types only; every pointer is a DEVICE pointer unless the name
 *    ends in _host; the caller owns all memory, nothing is allocated here
 *    except inside the opaque dirb200_net object;
 *  - every function returns 0 on success, <0 on error (message via
 *    dirb200_last_error(), thread local); nothing throws, nothing
 *    synchronises the host, every launch goes to the cudaStream_t passed
 *    (as void*; NULL = legacy default stream);
 *  - there is NO CPU fallback: without a CUDA device every compute entry
 *    point fails with DIRB200_ERR_CUDA.
 */
#ifndef DIRB200_H
#define DIRB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DIRB200_OK 0
#define DIRB200_ERR_ARG (-1)
#define DIRB200_ERR_CUDA (-2)
#define DIRB200_ERR_WORKSPACE (-3)

/* bin rules (row label -> FDS table row) */
#define DIRB200_BIN_AGE 0     /* agedb-dir/fds.py:91-99 : int(label - bucket_start), edge folding */
#define DIRB200_BIN_DEPTH10 1 /* nyud2-dir/models/fds.py:51-53 : clamp(int(label*10), bucket_start, bucket_num-1) */
#define DIRB200_BIN_EDGES5 2  /* sts-b-dir/fds.py:51-57 : np.histogram edges over [0,5], bucket_num bins */
/* the rule also selects calibrate_mean_var's channel mask: v1 != 0 (age, agedb-dir/utils.py:100) or
 * v1 > 0 && v2 >= 0 (nyud2-dir/util.py:154, sts-b-dir/util.py:66) */

/* loss kinds (agedb-dir/loss.py) */
#define DIRB200_LOSS_MSE 0       /* loss.py:5-10  */
#define DIRB200_LOSS_L1 1        /* loss.py:13-18 */
#define DIRB200_LOSS_FOCAL_MSE 2 /* loss.py:21-28 */
#define DIRB200_LOSS_FOCAL_L1 3  /* loss.py:31-38 */
#define DIRB200_LOSS_HUBER 4     /* loss.py:41-48 */
#define DIRB200_ACT_SIGMOID 0
#define DIRB200_ACT_TANH 1

/* LDS re-weighting (agedb-dir/datasets.py:64-67) */
#define DIRB200_REWEIGHT_SQRT_INV 1
#define DIRB200_REWEIGHT_INVERSE 2

const char* dirb200_last_error(void);
int dirb200_version(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t dirb200_launch_count(void);

/* ---------------------------------------------------------------- FDS ---- */

/* Edge-label presence flags: flags[0] |= any(label == bucket_start),
 * flags[1] |= any(label == bucket_num-1).  The caller zeroes flags (int32[2])
 * and may OR-reduce them across ranks.  Replaces the `label == bucket_start`
 * / `label == bucket_num-1` branches of the unique-label loop,
 * agedb-dir/fds.py:94-97,124-137. */
int dirb200_fds_label_flags(const float* labels, int64_t n, int bucket_num, int bucket_start,
                            int bin_rule, int32_t* flags, void* stream);

/* Row -> table row (int32, -1 = untouched), agedb-dir/fds.py:91-99. */
int dirb200_fds_bin_rows(const float* labels, int64_t n, int bucket_num, int bucket_start,
                         int bin_rule, const int32_t* flags, int32_t* bins_out, void* stream);

/* Workspace for dirb200_fds_accumulate. */
size_t dirb200_fds_accumulate_workspace_bytes(int64_t n, int nb);

/* Segmented (per label bin) accumulation of sum / sum-of-squares in fp64 and
 * row counts over features[n,d] (fp32 row-major, each element read exactly
 * once).  ADDS into sums/sumsq [nb,d] (double) and counts [nb] (int64), so an
 * epoch can be streamed batch by batch with no host round trip; the caller
 * zeroes them at epoch start and may all-reduce(sum) them across ranks.
 * Replaces the per-label mask/gather/mean/var loop, agedb-dir/fds.py:91-102,
 * and the host round trip at agedb-dir/train.py:276-279. */
int dirb200_fds_accumulate(const float* features, const int32_t* bins, int64_t n, int d, int nb,
                           double* sums, double* sumsq, int64_t* counts,
                           void* workspace, size_t workspace_bytes, void* stream);

/* Measurement aid (bench.py's `roofline` line; no reference counterpart): when enabled, dirb200_fds_accumulate
 * records CUDA events on the caller's stream around its fds_accumulate_kernel launch alone (the counting sort
 * that precedes it is excluded); dirb200_fds_last_accumulate_kernel_ms synchronises on them and returns the
 * kernel's duration of the most recent call. */
int dirb200_fds_set_profiling(int enabled);
int dirb200_fds_last_accumulate_kernel_ms(float* ms_out);

/* mean/var (unbiased; 0 when n==1) from the accumulators, then the running
 * EMA update of every bin with count>0, agedb-dir/fds.py:100-111.
 * momentum < 0 selects the `momentum is None` rule (factor = 1 - n/tracked);
 * first_update != 0 forces factor 0 (epoch == start_update, fds.py:107). */
int dirb200_fds_finalize(const double* sums, const double* sumsq, const int64_t* counts, int nb, int d,
                         float* running_mean, float* running_var, float* num_samples_tracked,
                         double momentum, int first_update, void* stream);

/* dst[b,:] = sum_j window[j] * src[reflect(b + j - (ks-1)/2), :]; ks <= 33.
 * Replaces F.pad(reflect)+F.conv1d, agedb-dir/fds.py:58-67. */
int dirb200_fds_smooth_tables(const float* src, int nb, int d, const float* window_host, int ks,
                              float* dst, void* stream);

/* In-place whiten/re-colour of x[b,d] by label bin: FDS.smooth +
 * calibrate_mean_var (agedb-dir/fds.py:115-144, agedb-dir/utils.py:97-107).
 * Edge flags are evaluated on this batch's labels (as the reference does per
 * call).  rowbin_out[b] (int32) = table row used, or -1 when the row was left
 * untouched (label dropped, or sum(v1[row]) < 1e-10): the backward's input. */
int dirb200_fds_calibrate_fwd(float* x, const float* labels, int64_t b, int d, int bucket_num,
                              int bucket_start, int bin_rule, const float* m1, const float* v1,
                              const float* m2, const float* v2, float clip_min, float clip_max,
                              int32_t* rowbin_out, int32_t* flags_scratch /* int32[2]; may be NULL when b <= 2048 */,
                              void* stream);

/* STS-B variant (sts-b-dir/fds.py:112-125): buckets with counts[b] == 0 in this update take their neighbours'
 * running statistics (copy at the two ends, mean of both neighbours inside), in increasing bucket order. */
int dirb200_fds_fill_empty(const int64_t* counts, int nb, int d, float* running_mean, float* running_var,
                           void* stream);

/* grad_in[b,d] = grad_out[b,d] * d(calibrate)/dx (may alias). */
int dirb200_fds_calibrate_bwd(int bin_rule, const float* grad_out, const int32_t* rowbin, int64_t b, int d,
                              const float* v1, const float* v2, float clip_min, float clip_max,
                              float* grad_in, void* stream);

/* ------------------------------------------------------------- losses ---- */

size_t dirb200_loss_workspace_bytes(int64_t n);

/* Fused forward + backward of weighted_{mse,l1,focal_mse,focal_l1,huber}_loss
 * (agedb-dir/loss.py:5-48): loss_out[0] = mean(l_i * w_i); if grad_out != NULL,
 * grad_out[i] = grad_scale * d loss / d pred_i.  weight may be NULL. */
int dirb200_loss_fwd_bwd(int kind, const float* pred, const float* target, const float* weight, int64_t n,
                         float beta, float gamma, int activate, float grad_scale,
                         float* loss_out, float* grad_out, void* workspace, size_t workspace_bytes,
                         void* stream);

/* ---------------------------------------------------------------- LDS ---- */

/* hist[min(max_target-1, int(label))]++ (int64, bit exact), ADDS into hist so
 * it can be all-reduced when labels are sharded.  agedb-dir/datasets.py:60-63. */
int dirb200_lds_histogram(const float* labels, int64_t n, int max_target, int64_t* hist, void* stream);

/* hist -> per-bin value (sqrt / clip(5,1000), optional convolve1d(mode=constant)
 * with the float64 window, integer truncation on the 'inverse' path as scipy
 * does) -> per-sample float32 weight 1/value[bin] scaled to mean 1.
 * scratch: >= (2*max_target + 2) doubles.  agedb-dir/datasets.py:64-82. */
int dirb200_lds_weights(const float* labels, int64_t n, int max_target, int reweight,
                        const double* window_host, int ks, const int64_t* hist,
                        double* scratch, float* weights_out, void* stream);

/* Same, for a label column sharded over ranks: `hist` is the histogram of the WHOLE column (the per-rank
 * dirb200_lds_histogram outputs SUM-all-reduced, int64, exact) and `n_total` its length; the per-bin table and the
 * len / sum(w) normaliser (agedb-dir/datasets.py:81-82) are formed from those, the gather covers this rank's
 * `n` labels.  dirb200_lds_weights == this with n_total = n.  SURVEY.md section 8e(3). */
int dirb200_lds_weights_sharded(const float* labels, int64_t n, int64_t n_total, int max_target, int reweight,
                                const double* window_host, int ks, const int64_t* hist, double* scratch,
                                float* weights_out, void* stream);

/* Dense per-element weight lookup of nyud2-dir/loaddata.py:52-64: weights_out[i] = table[min(int(values[i] * mult),
 * max_bin)] (mult = 10, max_bin = 99 for depth maps; table = the bucket weights, device pointer). */
int dirb200_lds_table_lookup(const float* values, int64_t n, float mult, int max_bin, const float* table,
                             float* weights_out, void* stream);

/* ------------------------------------------------ BatchNorm / pooling layers ---- */
/* nn.BatchNorm2d in training mode on an NHWC bf16 tensor y [rows][c] (rows = N*H*W; c a multiple of 8, <= 2048),
 * optionally followed by ReLU (agedb-dir/resnet.py:46-51,128-130; nyud2-dir/models/modules.py:13-21): batch statistics
 * (biased variance) -> running statistics updated in place with `momentum` (unbiased variance; NULL: not tracked) ->
 * out = [relu](gamma * (y - mean) * invstd + beta).  save_mean / save_invstd [c] and scale_shift [2][c] are kept for
 * the backward.  workspace: dirb200_bn_workspace_bytes(c). */
size_t dirb200_bn_workspace_bytes(int c);
int dirb200_bn_train_fwd(const void* y, int64_t rows, int c, const float* gamma, const float* beta, float eps,
                         float momentum, float* running_mean, float* running_var, int relu, void* out, float* save_mean,
                         float* save_invstd, float* scale_shift, void* workspace, void* stream);
/* grad_out = d loss / d out -> grad_y = d loss / d y (bf16), grad_gamma / grad_beta ACCUMULATED (fp32 [c]).  With relu the
 * mask is re-derived from (y, scale_shift) -- the activation itself is not read. */
int dirb200_bn_train_bwd(const void* grad_out, const void* y, int64_t rows, int c, const float* gamma,
                         const float* save_mean, const float* save_invstd, const float* scale_shift, int relu,
                         float* grad_gamma, float* grad_beta, void* grad_y, void* workspace, void* stream);
/* nn.MaxPool2d(kernel_size=3, stride=2, padding=1) (resnet.py:82) on NHWC bf16, argmax (r*3+s of the first maximum)
 * kept as one byte per output element; backward for even H, W. */
int dirb200_maxpool3x3s2_fwd(const void* x, int n, int h, int w, int c, void* out, uint8_t* argmax, void* stream);
int dirb200_maxpool3x3s2_bwd(const void* grad_out, const uint8_t* argmax, int n, int h, int w, int c, void* grad_x,
                             void* stream);
/* nn.AvgPool2d over the whole hw-pixel map (resnet.py:87) : NHWC bf16 [n][hw][c] <-> fp32 [n][c] */
int dirb200_avgpool_fwd(const void* x, int n, int hw, int c, float* out, void* stream);
int dirb200_avgpool_bwd(const float* grad_out, int n, int hw, int c, void* grad_x, void* stream);

/* ------------------------------------------------ dense-prediction ops (NYUD2-DIR) ---- */
/* The NYUD2 decoder / feature-fusion / refinement modules (nyud2-dir/models/modules.py:6-174) are 1x1 / 3x3 / 5x5
 * convolutions (dirb200_conv_fprop / _dgrad / _wgrad take filters up to 5x5 at stride 1), bilinear up-sampling and a
 * channel concat.  NHWC bf16 like the conv stage; channel counts multiples of 8.
 * upsample: F.upsample(x, size=(ho, wo), mode='bilinear') with align_corners = False (modules.py:24), the weights
 * formed and associated as ATen does; the backward is a deterministic gather (ATen scatters with atomics). */
int dirb200_upsample_bilinear_fwd(const void* x, int n, int h, int w, int c, int ho, int wo, void* out, void* stream);
int dirb200_upsample_bilinear_bwd(const void* dy, int n, int h, int w, int c, int ho, int wo, void* dx, void* stream);
/* dst[p][dst_off + j] = src[p][src_off + j] for j < c, p < pixels (row strides in elements): torch.cat(..., 1) of
 * NHWC tensors (modules.py:120) is one call per source; its backward is the same call with the roles swapped. */
int dirb200_copy_channels(const void* src, int src_stride, int src_off, void* dst, int dst_stride, int dst_off, int c,
                          int64_t pixels, void* stream);

/* ------------------------------------------------ input pipeline ---- */
/* Batched device form of the per-sample torchvision chain agedb-dir/datasets.py:38-53 after the resize:
 * RandomCrop(size, padding=pad) -> RandomHorizontalFlip -> ToTensor -> Normalize(mean, std), bit-identical to
 * torchvision for the same draws.  images u8 [n][size][size][3] (RGB, HWC) -> out f32 [n][3][size][size].
 * crop_yx int32 [n][2] = (top, left) of the crop in the zero-padded image, 0 .. 2*pad (NULL: (pad, pad) = the
 * validation chain, :46-51); flip u8 [n] (NULL: no flips).  value = (u8 / 255 - mean) / std; padding = u8 0. */
int dirb200_augment_batch(const uint8_t* images, const int* crop_yx, const uint8_t* flip, int n, int size, int pad,
                          float mean, float stdv, float* out, void* stream);

/* ------------------------------------------------ evaluation metrics ---- */
/* hist[int(label)] += 1 for 0 <= int(label) < nbins (int64, bit-exact, ADDS; no clamping): the per-label-value
 * training counts that shot_metrics compares against, agedb-dir/train.py:339,350. */
int dirb200_int_label_histogram(const float* labels, int64_t n, int nbins, int64_t* hist, void* stream);

/* Overall and many / median / low-shot error sums of a prediction vector in one pass (replaces the host loop over
 * np.unique(labels) of shot_metrics, agedb-dir/train.py:338-391, and the MSE / L1 / G-Mean meters of validate,
 * :286-335).  A sample is "many"-shot when the training count of its label value is > many_shot_thr, "low" when
 * it is < low_shot_thr (label values absent from training or not integer valued count 0), else "median".
 * out16 (double[4][4], OVERWRITTEN): rows = overall, many, median, low; columns = count, sum (pred-label)^2,
 * sum |pred-label|, sum log|pred-label|  ->  mse = c1/c0, l1 = c2/c0, gmean = exp(c3/c0). */
int dirb200_shot_metrics(const float* preds, const float* labels, int64_t n, const int64_t* train_hist, int nbins,
                         int many_shot_thr, int low_shot_thr, double* out16, void* stream);

/* ------------------------------------------------- convolution stack ---- */
/* Activations are NHWC bf16; weights arrive in the reference's fp32
 * [Cout][Cin][KH][KW] layout (agedb-dir/resnet.py:46-51,79,112-118, i.e. the
 * nn.Conv2d parameters / state_dict tensors) and are re-laid-out to bf16 GEMM
 * operands by dirb200_conv_prep_weights.  The convolutions themselves are
 * tcgen05 implicit GEMMs (fp32 accumulate in TMEM); they replace the cuDNN
 * calls behind nn.Conv2d forward and its autograd backward.
 * Cin and Cout must be multiples of 64, except the stem (stem=1: Cin=3, 7x7,
 * stride 2, pad 3), which consumes the 16-channel space-to-depth input made
 * by dirb200_input_to_s2d. */

/* w fp32 [Cout][Cin][KH][KW] -> w_fprop bf16 [Cout][KH][KW][Cin] and (if not
 * NULL) w_dgrad bf16 [Cin][KH][KW][Cout].  stem: w_fprop bf16 [Cout][256]. */
int dirb200_conv_prep_weights(const float* w, int cout, int cin, int kh, int kw, int stem,
                              void* w_fprop, void* w_dgrad, void* stream);

/* x fp32 NCHW [n,3,h,w] (what ResNet.forward receives, resnet.py:127) ->
 * bf16 [n, h/2, w/2, 16], channel (ph*2+pw)*4+c, c==3 zero. */
int dirb200_input_to_s2d(const float* x_nchw, int n, int h, int w, void* out_bf16, void* stream);

/* y[n,ho,wo,cout] = conv2d(x[n,h,w,cin], w)  (bias-free, as every conv in resnet.py) */
int dirb200_conv_fprop(const void* x, const void* w_fprop, void* y, int n, int h, int w, int cin,
                       int cout, int kh, int kw, int stride, int pad, int stem, void* stream);

/* dx[n,h,w,cin] = d loss / d x given dy[n,ho,wo,cout] */
int dirb200_conv_dgrad(const void* dy, const void* w_dgrad, void* dx, int n, int h, int w, int cin,
                       int cout, int kh, int kw, int stride, int pad, void* stream);

/* Host-only (no CUDA call, works without a device): which GEMM form the three entry points above / below would launch for
 * a shape (op: 0 fprop, 1 dgrad, 2 wgrad).  plan7[0] tile width BN; [1] 1 = CTA pairs (cta_group::2); [2] A-operand form:
 * 0 cp.async gather, 1 tiled TMA, 2 im2col-mode TMA, 3 patch-resident (one input patch in shared memory, taps as
 * displaced descriptors); [3] image rows per tile of the patch form; [4] split-K factor (wgrad); [5] launches (a stride-2
 * dgrad runs one per non-empty output-pixel parity class); [6] 1 = this dgrad can also accumulate the BN-backward
 * moments of the previous layer in its epilogue. */
int dirb200_conv_plan(int n, int h, int w, int cin, int cout, int kh, int kw, int stride, int pad, int stem, int op,
                      int* plan7);
size_t dirb200_conv_wgrad_workspace_bytes(int n, int h, int w, int cin, int cout, int kh, int kw,
                                          int stride, int pad, int stem);

/* dw fp32 [Cout][Cin][KH][KW] (=, or += when accumulate) d loss / d w from x and dy
 * (split-K partials in `workspace`, then a deterministic reduce). */
int dirb200_conv_wgrad(const void* x, const void* dy, float* dw, void* workspace, size_t workspace_bytes,
                       int n, int h, int w, int cin, int cout, int kh, int kw, int stride, int pad,
                       int stem, int accumulate, void* stream);

/* ------------------------------------------------ ResNet backbone runner ---- */
/* Opaque native runner of the bottleneck ResNet of agedb-dir/resnet.py:41-70,
 * 73-138 (conv1/bn1/relu/maxpool, layer1-4, avgpool, view) for one fixed
 * input shape.  It owns every activation / gradient / operand buffer (NHWC
 * bf16); parameters, their gradients and the BN running statistics stay in
 * caller-owned flat fp32 buffers laid out in the reference's
 * named_parameters() order: conv1.weight, bn1.weight, bn1.bias, then per
 * block conv1.weight, bn1.{weight,bias}, conv2.weight, bn2.*, conv3.weight,
 * bn3.*, [downsample.0.weight, downsample.1.{weight,bias}]; running stats per
 * BN in the same order as (running_mean, running_var). */
typedef struct dirb200_net dirb200_net;

int dirb200_resnet_create(int n, int h, int w, const int* blocks_per_stage, int num_stages, dirb200_net** out);
void dirb200_resnet_destroy(dirb200_net* net);
int64_t dirb200_resnet_param_count(const dirb200_net* net);   /* floats in the flat parameter buffer */
int64_t dirb200_resnet_running_count(const dirb200_net* net); /* floats in the flat BN running buffer */
int64_t dirb200_resnet_feature_dim(const dirb200_net* net);   /* 2048 for ResNet-50 */
int64_t dirb200_resnet_device_bytes(const dirb200_net* net);

/* ResNet.forward up to `encoding` (resnet.py:128-138): x fp32 NCHW -> enc fp32
 * [n, feature_dim].  training: batch statistics + running-stat update. */
int dirb200_resnet_forward(dirb200_net* net, const float* x_nchw, const float* params, float* bn_running,
                           int training, float* enc_out, void* stream);

/* Backward of the above: d_enc fp32 [n, feature_dim]; ACCUMULATES into grads. */
int dirb200_resnet_backward(dirb200_net* net, const float* d_enc, const float* params, float* grads, void* stream);

/* The same backward, one stage per call, so that the caller can overlap the gradient all-reduce of a finished stage
 * (agedb-dir/train.py:143 DataParallel's reduction) with the stages still running:
 * stage = dirb200_resnet_num_stages() (avg-pool backward + last layer group; reads d_enc), then stage-1 ... 1, then 0
 * (max-pool, stem).  In exactly that order after a training-mode forward.  dirb200_resnet_stage_param_range: the
 * [lo, hi) float range of the flat parameter / gradient buffers that stage owns (0 = conv1 + bn1). */
int dirb200_resnet_num_stages(const dirb200_net* net);
int dirb200_resnet_backward_stage(dirb200_net* net, int stage, const float* d_enc, const float* params, float* grads,
                                  void* stream);
int dirb200_resnet_stage_param_range(const dirb200_net* net, int stage, int64_t* lo, int64_t* hi);

/* Per-kernel-class device timing of forward/backward (CUDA events around every launch group).  Classes:
 * 0 prep (weight re-layout, s2d), 1 conv fprop, 2 conv dgrad, 3 conv wgrad GEMM, 4 wgrad split-K reduce,
 * 5 BN statistics, 6 BN apply, 7 BN backward reduce, 8 BN backward apply, 9 pooling.
 * read_profile synchronises, fills ms_by_kind[10] / groups_by_kind[10] and clears the log. */
int dirb200_resnet_set_profiling(dirb200_net* net, int enabled);
int dirb200_resnet_read_profile(dirb200_net* net, double* ms_by_kind, int64_t* groups_by_kind);

/* Test / debugging aid: device pointer + shape ([rows][channels] bf16, NHWC) of an internal activation.
 * block = -1: stem (0 = conv1 raw, 1 = relu(bn1), 6 = max-pool output); block >= 0: 0/1 = conv1 raw / act,
 * 2/3 = conv2 raw / act, 4 = conv3 raw, 5 = downsample raw, 6 = block output. */
int dirb200_resnet_peek(dirb200_net* net, int block, int which, void** ptr, int64_t* rows, int* channels);

/* nn.Linear(feature_dim, 1) (resnet.py:88,148): pred[n] = x[n,d] . w[d] + bias */
int dirb200_linear1_fwd(const float* x, const float* w, const float* bias, int64_t n, int d, float* pred,
                        void* stream);
/* dx[n,d] (may be NULL), dw[d], dbias[1] (all overwritten) */
int dirb200_linear1_bwd(const float* grad_pred, const float* x, const float* w, int64_t n, int d, float* dx,
                        float* dw, float* dbias, void* stream);

/* Fused optimizer steps over flat fp32 buffers (torch.optim.Adam / SGD semantics,
 * agedb-dir/train.py:163-164,262); grads are multiplied by grad_scale first
 * (1/world_size after a sum all-reduce) and, when clip_coef != NULL, by the device scalar *clip_coef
 * (dirb200_grad_clip_coef below: gradient-norm clipping without a host round trip). */
int dirb200_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                      float beta1, float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                      const float* clip_coef, void* stream);
int dirb200_sgd_step(float* params, const float* grads, float* momentum_buf, int64_t n, float lr, float momentum,
                     float weight_decay, int first_step, float grad_scale, const float* clip_coef, void* stream);

/* torch.nn.utils.clip_grad_norm_ over one flat gradient buffer (sts-b-dir/trainer.py:147-149, --max_grad_norm):
 * out[0] = min(1, max_norm / (||grad_scale * grads||_2 + 1e-6)), out[1] = that norm.  The gradients themselves are
 * not rewritten -- the optimizer step applies out[0] while it reads them.  workspace: >= 8 KiB + 16 B, fp64 partials. */
size_t dirb200_grad_clip_workspace_bytes(void);
int dirb200_grad_clip_coef(const float* grads, int64_t n, float grad_scale, float max_norm, void* workspace,
                           size_t workspace_bytes, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIRB200_H */
