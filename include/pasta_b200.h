/*
 * pasta_b200.h — C ABI of the B200-native PASTA-GAN operator hot path.
 *
 * The reference has no C ABI: its native code is reached through two pybind11 modules
 * JIT-built by torch.utils.cpp_extension (torch_utils/custom_ops.py:46-124):
 *
 *     _plugin.upfirdn2d(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain)
 *                                                    torch_utils/ops/upfirdn2d.cpp:16-94, :98-101
 *     _plugin.bias_act(x, b, xref, yref, dy, grad, dim, act, alpha, gain, clamp)
 *                                                    torch_utils/ops/bias_act.cpp:32-90, :94-97
 *
 * and every convolution is torch.nn.functional.conv2d / conv_transpose2d (cuDNN) called from
 * torch_utils/ops/conv2d_gradfix.py:35-43 via conv2d_resample.py:29-54.
 *
 * This header is what a binding for that path binds instead: plain pointers, sizes and
 * scalars, no torch types.  Conventions shared by every entry point:
 *
 *   - All data pointers are DEVICE pointers on the current CUDA device.  Inputs are borrowed and
 *     never written; outputs are caller-allocated (ownership stays with the caller, exactly like
 *     the torch::empty tensors of upfirdn2d.cpp:35 / bias_act.cpp:55).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Launches are
 *     asynchronous; no entry point synchronises the host (upfirdn2d.cpp:92, bias_act.cpp:88).
 *   - Return value: 0 on success, a PG_ERR_* code otherwise; pg_last_error() returns a
 *     thread-local human-readable message for the last failing call on the calling thread
 *     (the reference raises RuntimeError through TORCH_CHECK / AT_CUDA_CHECK).
 *   - Entry points are re-entrant and keep no mutable global state (bar a statistics counter): they are called under the GIL
 *     from the forward pass and from autograd worker threads in backward.
 *   - There is no CPU implementation behind any of these symbols.
 */
#ifndef PASTA_B200_H_
#define PASTA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PG_ABI_VERSION 2

enum {
    PG_OK = 0,
    PG_ERR_INVALID_ARGUMENT = 1,   /* TORCH_CHECK failures of the reference launchers            */
    PG_ERR_UNSUPPORTED = 2,        /* valid request this build has no kernel for                 */
    PG_ERR_CUDA = 3,               /* cudaLaunchKernel / runtime error (AT_CUDA_CHECK)           */
    PG_ERR_NO_DEVICE = 4           /* no sm_100 device / kernel image not loadable on this device */
};

/* element types of activations (the reference dispatches AT_DISPATCH_FLOATING_TYPES_AND_HALF,
 * upfirdn2d.cpp:59, bias_act.cpp:77); all arithmetic is fp32 (fp64 for PG_F64), cf. upfirdn2d.cu:15-18 */
enum { PG_F32 = 0, PG_F16 = 1, PG_F64 = 2 };

/* activation indices == the reference's `cuda_idx` (bias_act.py:23-33) */
enum {
    PG_ACT_LINEAR = 1, PG_ACT_RELU = 2, PG_ACT_LRELU = 3, PG_ACT_TANH = 4, PG_ACT_SIGMOID = 5,
    PG_ACT_ELU = 6, PG_ACT_SELU = 7, PG_ACT_SOFTPLUS = 8, PG_ACT_SWISH = 9
};

int         pg_abi_version(void);
const char* pg_last_error(void);
/* 0 if the current device can run the sm_100a kernel images in this library. */
int         pg_check_device(void);
/* Number of kernels this library has launched in this process so far (statistics only; relaxed atomic). */
int64_t     pg_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * bias_act  — replaces _plugin.bias_act (bias_act.cpp:32-90; kernel bias_act.cu:23-147).
 *
 *   grad == 0 :  y = clamp( act(x + b) * gain )
 *   grad == 1 :  y = x * gain * act'(.)          with `x` the incoming gradient, act' evaluated from
 *                yref (or xref + b for swish); zero where |yref| >= clamp
 *   grad == 2 :  y = x * dy * gain * act''(.)    (tanh, sigmoid, elu, selu, softplus, swish only)
 *
 * NULL for b / xref / yref / dy means "absent" (the reference passes empty tensors).
 * clamp < 0 means "no clamp".  The bias element used for flat index i is b[(i / step_b) % size_b]
 * (bias_act.cu:44) — step_b is x.stride(dim) in elements; x, xref, yref, dy, y share one dense
 * layout of size_x elements.  size_x must be <= INT32_MAX ("x is too large", bias_act.cpp:41).
 */
int pg_bias_act(const void* x, const void* b, const void* xref, const void* yref, const void* dy, void* y,
                int64_t size_x, int32_t size_b, int64_t step_b,
                int32_t grad, int32_t act, float alpha, float gain, float clamp,
                int32_t dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * upfirdn2d — replaces _plugin.upfirdn2d (upfirdn2d.cpp:16-94; kernels upfirdn2d.cu:29-200).
 *
 *   y = decimate_down( FIR_f( pad( zero_insert_up(x) ) ) ) * gain          per (n, c) plane
 *
 * sizes are {N, C, H, W}; strides are in ELEMENTS in the same order (NCHW-contiguous and
 * channels_last are both expressed this way, upfirdn2d.cpp:48-55).  f is float32 [fh, fw] with
 * element strides {f_stride_h, f_stride_w}; flip == 0 means true convolution (the filter is
 * flipped before correlating), flip != 0 means correlation (upfirdn2d.py:195-196).  Padding may be
 * negative (crop).  out_size must equal
 *   outW = (W*upx + padx0 + padx1 - fw + downx) / downx,   outH likewise     (upfirdn2d.cpp:32-33)
 * and every tensor must hold <= INT32_MAX elements.
 */
int pg_upfirdn2d(const void* x, const float* f, void* y,
                 const int32_t in_size[4], const int64_t in_stride[4],
                 const int32_t out_size[4], const int64_t out_stride[4],
                 int32_t fh, int32_t fw, int64_t f_stride_h, int64_t f_stride_w,
                 int32_t upx, int32_t upy, int32_t downx, int32_t downy,
                 int32_t padx0, int32_t padx1, int32_t pady0, int32_t pady1,
                 int32_t flip, float gain, int32_t dtype, void* stream);

/* Same resampling with the bias_act epilogue of the calling layer fused in (north_star kernel 1):
 *   y = clamp( act( upfirdn2d(x) + b[c] ) * act_gain )
 * i.e. upfirdn2d(...) followed by bias_act(dim=1) in one pass over HBM (Conv2dLayer.forward,
 * training/networks.py:170-179; SynthesisLayer.forward :296-315).  b may be NULL.
 * act must be linear / relu / lrelu (the activations whose backward needs only y). */
int pg_upfirdn2d_bias_act(const void* x, const float* f, const void* b, void* y,
                          const int32_t in_size[4], const int64_t in_stride[4],
                          const int32_t out_size[4], const int64_t out_stride[4],
                          int32_t fh, int32_t fw, int64_t f_stride_h, int64_t f_stride_w,
                          int32_t upx, int32_t upy, int32_t downx, int32_t downy,
                          int32_t padx0, int32_t padx1, int32_t pady0, int32_t pady1,
                          int32_t flip, float gain,
                          int32_t act, float alpha, float act_gain, float clamp,
                          int32_t dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * conv2d_igemm — tcgen05 / TMEM implicit-GEMM convolution, fp32 NCHW in and out, fp16 or bf16 tensor-core
 * operands with fp32 accumulation.  Replaces, in one launch, what the reference spreads over
 *   conv2d_gradfix.conv2d / conv_transpose2d (cuDNN)           torch_utils/ops/conv2d_gradfix.py:35-43
 *   conv2d_resample's up-2 lowering (convT + FIR, gain 4)      torch_utils/ops/conv2d_resample.py:125-142
 *   weight modulation / demodulation / noise                   training/networks.py:64-93 (modulated_conv2d)
 *   bias_act of the calling layer                              training/networks.py:176-179, :311-315, :4345-4349
 *
 *   y[n,o] = clamp( act( dcoefs[n,o] * conv( styles[n,c] * in_gain * in_act(x[n,c]) , w[o,c] ) + noise + bias[o] ) * gain )
 *
 * x [N,Cin,H,W] and y [N,Cout,H*up,W*up] are dense NCHW fp32.  w [Cout,Cin,k,k] fp32, k in {1,3}, stride 1, padding k/2
 * ("same").  flip_weight != 0: cross-correlation (what conv2d computes); 0: true convolution (conv2d_resample.py:35-36).
 * up == 2 (k == 3 only): the zero-insert + 4x4 FIR (gain 4) + 3x3 convolution of SynthesisLayer.conv0 evaluated in
 * polyphase form on the low-resolution input; `fir` is the [4,4] float32 filter (ignored when up == 1).
 * styles [N,Cin], dcoefs [N,Cout], bias [Cout], noise ([H*up,W*up] with noise_batch_stride 0, or per-sample with stride in
 * elements) may each be NULL.  in_act / act: PG_ACT_LINEAR / RELU / LRELU.  clamp < 0: none.
 * operand_format: 0 = fp16 (10-bit mantissa, saturating), 1 = bf16, 2 = tf32 (10-bit mantissa, fp32 exponent range; half the tensor rate, dense fp32 / fp16
 *                 NCHW inputs only: no channel-blocked input, packed weights take twice the bytes -> pg_conv2d_igemm_workspace_bytes_fmt).
 * workspace: device scratch of at least pg_conv2d_igemm_workspace_bytes(...) bytes (packed weights), caller-owned.
 * Forward only: gradients of convolutions stay on conv2d_gradfix in this release.
 */
/* `up` selects the resampling fused into the convolution: 1 = none, 2 = up-2 (above), PG_CONV_DOWN2 = the down-2 form of
 * conv2d_resample.py:119-122 (4x4 FIR with padding 2, then the 3x3 convolution at stride 2; y is [N,Cout,H/2,W/2]) evaluated as a
 * 'same' 3x3 convolution over the four space-to-depth planes of x with the 6x6 composite kernel w (*) f — no filtered intermediate. */
#define PG_CONV_DOWN2 (-2)
/* The same down-2 form for a channel-blocked float16 input (PG_LAYOUT_C8, loaded by a strided TMA box): the packed weights order the 4 * Cin
 * space-to-depth channels parity-major, so prepack and launch must both be given this value. */
#define PG_CONV_DOWN2_C8 (-3)
int64_t pg_conv2d_igemm_workspace_bytes(int32_t Cin, int32_t Cout, int32_t ksize, int32_t up);                 /* operand_format 0 / 1 */
int64_t pg_conv2d_igemm_workspace_bytes_fmt(int32_t Cin, int32_t Cout, int32_t ksize, int32_t up, int32_t operand_format);
/* The same operation in two steps, so that inference can pack the weights once per parameter version:
 *   prepack: w * w_scale (the layer's weight_gain, training/networks.py:171) -> fp16/bf16 GEMM tiles in `workspace`
 *   run:     the convolution proper on packed weights.  pg_conv2d_igemm_fwd == prepack(w_scale = 1) + run. */
int pg_conv2d_igemm_prepack(const float* w, const float* fir, float w_scale, int32_t Cin, int32_t Cout, int32_t ksize, int32_t up,
                            int32_t flip_weight, int32_t operand_format, void* workspace, int64_t workspace_bytes, void* stream);
int pg_conv2d_igemm_run(const float* x, const void* wpack, const float* styles, const float* dcoefs,
                        const float* noise, int64_t noise_batch_stride, const float* bias, float* y,
                        int32_t N, int32_t Cin, int32_t Cout, int32_t H, int32_t W, int32_t ksize, int32_t up,
                        int32_t in_act, float in_alpha, float in_gain,
                        int32_t act, float alpha, float gain, float clamp, int32_t operand_format, void* stream);
/* pg_conv2d_igemm_run with two more fusions around the same GEMM:
 *   x2 != NULL: the input is the channel concatenation [x ; x2] without materialising it (x [N,Cin1,H,W], x2 [N,Cin-Cin1,H,W], Cin1 % 8 == 0;
 *               replaces torch.cat + conv of SynthesisBlockFull's merge_conv, training/networks.py:5705-5706); not with PG_CONV_DOWN2.
 *   residual != NULL: y += residual after activation, gain and clamp (the `y.add_(x)` that closes every residual block,
 *               training/networks.py:986-990, :5268-5272).
 *   x_dtype / y_dtype = PG_F16: x and / or y are dense NCHW float16 (pass the pointers cast to float*).  For tensors that only travel between two
 *               of these convolutions (the SPADE blocks' `actv` and normalised maps, training/networks.py:4371-4379): the consumer would round
 *               the fp32 value to an fp16 operand anyway, so results are bit-identical while the tensor costs half the bytes.  A float16 input
 *               needs a plain layer (styles == NULL, in_act linear, in_gain 1, no x2 / down-2), operand_format 0 and even W; a float16 output
 *               excludes `residual`. */
int pg_conv2d_igemm_run2(const float* x, const float* x2, int32_t Cin1, const void* wpack, const float* styles, const float* dcoefs,
                         const float* noise, int64_t noise_batch_stride, const float* bias, const float* residual, float* y,
                         int32_t N, int32_t Cin, int32_t Cout, int32_t H, int32_t W, int32_t ksize, int32_t up,
                         int32_t in_act, float in_alpha, float in_gain,
                         int32_t act, float alpha, float gain, float clamp, int32_t operand_format, int32_t x_dtype, int32_t y_dtype, void* stream);
/* Batched weight packing: `batch` weight sets in one launch, set i written at workspace + i * pg_conv2d_igemm_workspace_bytes(...).
 *   w_batch_stride > 0: set i is read from w + i * w_batch_stride elements — the per-sample weights [N*O, I, k, k] that the reference's FUSED
 *                       modulated convolution hands to conv2d_resample with groups = N (training/networks.py:84-94);
 *   w_batch_stride == 0 with styles [batch, Cin]: set i = w * styles[i, c] — the modulation of networks.py:64-66 applied while packing, for layers
 *                       whose activations are read by TMA and therefore cannot be scaled on the way in.
 *   n_tile: 0 = the library's default N tile (min(Cout rounded to 16, 256)); otherwise the GEMM's N-tile width (multiple of 16 dividing Cout).  The
 *           same value must be passed in pg_conv_args.n_tile.  Used by the SPADE layer: with rows ordered [gamma_t ; beta_t] per 64-channel tile t
 *           and n_tile = 128, two CTAs share an SM and one's epilogue overlaps the other's main loop. */
int pg_conv2d_igemm_prepack_batched(const float* w, int64_t w_batch_stride, const float* styles, int32_t batch, const float* fir, float w_scale,
                                    int32_t Cin, int32_t Cout, int32_t ksize, int32_t up, int32_t flip_weight, int32_t operand_format, int32_t n_tile,
                                    void* workspace, int64_t workspace_bytes, void* stream);

/* The general entry point (everything pg_conv2d_igemm_run2 / _spade_run do, plus):
 *   wpack_sample_stride > 0: sample n multiplies with the packed weight set at wpack + n * wpack_sample_stride bytes.  This is the
 *               `groups = N` convolution of the reference's fused modulated conv (x [1, N*I, H, W] x w [N*O, I, k, k], conv2d_resample.py:59 with
 *               groups = N called from networks.py:88-90) with x viewed as [N, I, H, W] and y as [N, O, H', W'], including its up-2 form.
 *   x_layout / y_layout = PG_LAYOUT_C8: the tensor is float16 in the channel-blocked layout [N][C/8][H][W][8] (C % 16 == 0).  Such an input is
 *               loaded by the Tensor Memory Accelerator straight into the tensor-core operand layout (a 3-D box of rows x strip positions x 2
 *               channel blocks per 16-channel chunk; halo and padding are the tensor map's out-of-bounds zero fill) — no conversion pass at all.
 *               It must be a plain layer (no styles, input activation or down-2; fold a modulation into the weights with
 *               pg_conv2d_igemm_prepack_batched).  x2 (the fused concat) must then be channel-blocked too, with Cin1 % 16 == 0.  Such an output is written 16 bytes per (pixel, 8 channels) by the epilogue.
 *   spade_x != NULL: the SPADE epilogue of pg_conv2d_igemm_spade_run (Cout = 2C).  wpack = [gamma ; beta], or with n_tile = 2 Ct < 2C the rows
 *               ordered tile by tile: [gamma[0:Ct] ; beta[0:Ct] ; gamma[Ct:2Ct] ; beta[Ct:2Ct] ; ...].
 * struct_bytes must be sizeof(pg_conv_args) (guards against a binding built for another header). */
enum { PG_LAYOUT_NCHW = 0, PG_LAYOUT_C8 = 1 };
typedef struct pg_conv_args {
    uint32_t struct_bytes;
    int32_t  N, Cin, Cout, H, W, ksize, up;
    const void* x;  int32_t x_dtype, x_layout;
    const void* x2; int32_t cin1, residual_layout;   /* residual_layout: PG_LAYOUT_NCHW (float32) or PG_LAYOUT_C8 (float16, Cout % 16 == 0) */
    const void* wpack; int64_t wpack_sample_stride;
    const float* styles; const float* dcoefs; const float* noise; int64_t noise_batch_stride; const float* bias; const float* residual;
    void* y; int32_t y_dtype, y_layout;
    int32_t in_act; float in_alpha, in_gain;
    int32_t act; float alpha, gain, clamp;
    int32_t operand_format, n_tile;      /* n_tile: N-tile width the weights were packed with (0 = the library's default) */
    const float* spade_x; const float* spade_mean; const float* spade_rstd;
    void* stream;
} pg_conv_args;
int pg_conv2d_igemm_launch(const pg_conv_args* args);

/* Plan / loader choices that were measured against each other (column bands, strip length, loader variants, TMA on / off ...).  Defaults are the
 * measured best; the PASTA_B200_CONV_* environment variables are read once at first use and this call overrides a value for the rest of the
 * process.  Keys: conv_bands, conv_band_tw, conv_band_minw, conv_band_ratio10, conv_persist, conv_nacc, conv_pair, conv_vec2, conv_lean,
 * conv_cgroups, conv_tma.  Every setting computes the same result (bit-identical for a fixed operand format). */
int pg_set_tuning(const char* key, int32_t value);

int pg_conv2d_igemm_fwd(const float* x, const float* w, const float* fir, const float* styles, const float* dcoefs,
                        const float* noise, int64_t noise_batch_stride, const float* bias, float* y,
                        int32_t N, int32_t Cin, int32_t Cout, int32_t H, int32_t W, int32_t ksize, int32_t up,
                        int32_t flip_weight, int32_t in_act, float in_alpha, float in_gain,
                        int32_t act, float alpha, float gain, float clamp, int32_t operand_format,
                        void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * conv2d_wgrad — weight gradient of the stride-1 'same' convolution on tcgen05 / TMEM (k in {1, 3}):
 *
 *   dw[o, c, kh, kw] (+)= scale * sum_{n,h,w} dy[n, o, h, w] * x[n, c, h + kh - k/2, w + kw - k/2]
 *
 * Replaces the cuDNN call behind the reference's conv2d_gradfix weight gradient (torch_utils/ops/conv2d_gradfix.py:140-148,
 * aten::cudnn_convolution_backward_weight).  x [N,Cin,H,W], dy [N,Cout,H,W] dense NCHW float32; dw [Cout,Cin,k,k] float32 in the layout of
 * F.conv2d's weight (cross-correlation taps).  bf16 tensor-core operands (gradients need the exponent range), fp32 accumulation; the pixel
 * dimension is split over CTAs and summed by a second kernel (deterministic, no atomics).  accumulate != 0 adds to dw instead of overwriting.
 * workspace: at least pg_conv2d_wgrad_workspace_bytes(...) bytes, 16-byte aligned, caller-owned.
 * The INPUT gradient of the same convolution needs no entry point of its own: it is pg_conv2d_igemm_* on dy with the weights transposed
 * (Cin <-> Cout) and flip_weight = 0. */
int64_t pg_conv2d_wgrad_workspace_bytes(int32_t N, int32_t Cin, int32_t Cout, int32_t H, int32_t W, int32_t ksize);
int pg_conv2d_wgrad(const float* x, const float* dy, float* dw, int32_t N, int32_t Cin, int32_t Cout, int32_t H, int32_t W, int32_t ksize,
                    float scale, int32_t accumulate, void* workspace, int64_t workspace_bytes, void* stream);

/* SPADE normalisation fused into the epilogue of the convolution that produces its modulation maps (reference Spade_Norm_Block.forward,
 * training/networks.py:4371-4379, plus the pre-activation of the Spade_Conv2dLayer that consumes the result, :4345-4349):
 *
 *   [gamma | beta] = conv(feat, [w_gamma ; w_beta])          (one GEMM, 2C output columns, k in {1,3}, 'same')
 *   y[n,c]         = act( (x[n,c] - mean[n,c]) * rstd[n,c] * (1 + gamma[n,c]) + beta[n,c] ) * gain
 *
 * feat [N,Cin,H,W]; wpack_gamma_beta = pg_conv2d_igemm_prepack of the [2C,Cin,k,k] concatenation; x, y [N,C,H,W]; mean, rstd [N,C]
 * (instance-norm statistics of x); 2C <= 256, C % 16 == 0.  gamma and beta never reach HBM.  feat_dtype / y_dtype: PG_F32 or PG_F16 as in
 * pg_conv2d_igemm_run2 (x, mean, rstd are always float32). */
int pg_conv2d_igemm_spade_run(const float* feat, const void* wpack_gamma_beta, const float* x, const float* mean, const float* rstd,
                              float* y, int32_t N, int32_t Cin, int32_t C, int32_t H, int32_t W, int32_t ksize,
                              int32_t act, float alpha, float gain, int32_t operand_format, int32_t feat_dtype, int32_t y_dtype, void* stream);

/* Instance-norm statistics of x [planes = N*C, hw] (dense, fp32): mean and rstd = rsqrt(biased variance + eps), one streaming pass.
 * Feeds pg_conv2d_igemm_spade_run; replaces nn.InstanceNorm2d(affine=False) inside Spade_Norm_Block (training/networks.py:4363, :4377). */
int pg_instance_norm_stats(const float* x, float* mean, float* rstd, int64_t planes, int64_t hw, float eps, void* stream);
/* nn.InstanceNorm2d(affine=False) + activation in one pass for planes of <= 24576 elements (staged in shared memory):
 *   y[p, :] = act((x[p, :] - mean_p) * rsqrt(var_p + eps)) * gain,  act in linear / relu / lrelu(alpha).
 * The InstanceNorm -> LeakyReLU tail of the style encoder's Dense blocks (training/networks.py:594-611). */
int pg_instance_norm_act(const float* x, float* y, int64_t planes, int64_t hw, float eps, int32_t act, float alpha, float gain, void* stream);

/* Garment-feature completion of SynthesisNetworkFull.get_spade_feat (training/networks.py:5777-5800), the two passes over the feature map:
 *   pg_masked_plane_sum: out[n,c] = sum_hw feat[n,c,hw] * mask[n,hw]                      feat [N,C,hw], mask [N,hw], out [N,C]
 *   pg_masked_fill:      out[n,c,hw] = feat[n,c,hw] * (1 - rest[n,hw]) + fill[n,c] * rest[n,hw], out addressed with a batch stride in elements
 *                        (>= C*hw) so that it can be a channel slice of the concatenated upper|lower tensor (replaces torch.cat, :5831). */
int pg_masked_plane_sum(const float* feat, const float* mask, float* out, int64_t N, int64_t C, int64_t hw, void* stream);
int pg_masked_fill(const float* feat, const float* rest, const float* fill, float* out, int64_t N, int64_t C, int64_t hw,
                   int64_t out_batch_stride, int32_t out_dtype, void* stream);   /* out_dtype: PG_F32 or PG_F16 (out cast to float*) */

/* pg_masked_fill writing channel-blocked fp16 (PG_LAYOUT_C8): channels land in blocks [cb_offset, cb_offset + C/8) of out [N][cb_total][hw][8]
 * (C % 8 == 0) — the concatenated upper|lower garment feature map in the layout the SPADE convolutions load by TMA. */
int pg_masked_fill_c8(const float* feat, const float* rest, const float* fill, void* out, int64_t N, int64_t C, int64_t hw,
                      int64_t cb_total, int64_t cb_offset, void* stream);

/* Boundary of a chain of channel-blocked layers: dense NCHW (float32 or float16) [N, C, hw] <-> float16 C8 [N][ceil(C/8)][hw][8] (channels past C are
 * zero).  Tensors from other code enter through pg_nchw_to_c8; results leave through pg_c8_to_nchw.  Both HBM-bound, one pass. */
int pg_nchw_to_c8(const void* x, void* y, int64_t N, int64_t C, int64_t hw, int32_t x_dtype, void* stream);
int pg_c8_to_nchw(const void* x, void* y, int64_t N, int64_t C, int64_t hw, int32_t y_dtype, void* stream);

/* Device-side input / output pipeline around the generator (reference test.py:105-115, :131-135).
 *   pg_u8_normalize: njobs (<= 8) tensors in one launch; job i converts rows[i] rows of row_len[i] uint8 elements (row r at src[i] + r*src_stride[i])
 *                    to float32 rows at dst[i] + r*dst_stride[i] (strides in elements): x / 127.5 - 1 when normalize[i] != 0, else x.  A destination
 *                    stride larger than the row writes a channel slice of a wider tensor (`pose || retain`, test.py:115).  The job arrays are HOST arrays.
 *   pg_image_to_u8_bgr: img [N,3,H,W] float32 -> out [N,H,x1-x0,3] uint8 = uint8(clip((img[:, ::-1, :, x0:x1] + 1) * 127.5, 0, 255)), channel-last. */
int pg_u8_normalize(const void* const* src, void* const* dst, const int64_t* rows, const int64_t* row_len, const int64_t* src_stride,
                    const int64_t* dst_stride, const int32_t* normalize, int32_t njobs, void* stream);
int pg_image_to_u8_bgr(const float* img, void* out, int32_t N, int32_t H, int32_t W, int32_t x0, int32_t x1, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Patch routing — replaces the cv2.warpPerspective calls of the reference's data loader
 * (training/dataset.py:838-927, UvitonDatasetFull*.normalize: 28 rectifying warps + 28 back-warps per sample on the CPU).
 * uint8 images, fixed-point INTER_LINEAR exactly as OpenCV evaluates it (1/32-pixel coordinates, 2^15-scaled int16 weights).
 *
 *   pg_warp_perspective_u8: njobs independent warps in one launch.  Job i is one
 *       cv2.warpPerspective(src, M, (dst_w, dst_h), flags=INTER_LINEAR, borderMode=border, borderValue=0)
 *     with m = the DESTINATION -> SOURCE matrix (cv2 inverts M first: pass inv(M), row-major, as doubles).  Strides are in bytes, so a job can
 *     read a channel slice of a wider image and write one (the reference concatenates the ten patches on the channel axis, dataset.py:920-923).
 *     channels <= 4; border: 0 = BORDER_CONSTANT (value 0), 1 = BORDER_REPLICATE.  The job table lives in DEVICE memory;
 *     max_dst_pixels = max over jobs of dst_h * dst_w (sizes the grid).
 *   pg_patch_denorm_u8: the composite of dataset.py:882-886 / :892-897 for P parts (<= 16) of B samples:
 *       for p in 0..P-1 (valid[b][p] != 0):   keep = warp(masks[b][..., 3p], m[b][p], BORDER_CONSTANT) == 255
 *                                             denorm[b] = keep ? warp(patches[b][..., 3p:3p+3], m[b][p], BORDER_CONSTANT) : denorm[b]
 *     patches / masks [B][h][w][3P] uint8 (what pg_warp_perspective_u8 wrote), m [B][P][9] doubles (image pixel -> patch coordinates, i.e. inv(M_inv)),
 *     valid [B][P] uint8, denorm [B][H][W][3] (written in full; 0 where no part claims the pixel), part_masks [B][P][H][W] 0 / 1 or NULL. */
typedef struct pg_warp_job {
    double m[9];
    const uint8_t* src; uint8_t* dst;
    int32_t src_h, src_w, src_row_stride, src_pix_stride;
    int32_t dst_h, dst_w, dst_row_stride, dst_pix_stride;
    int32_t channels, border;
} pg_warp_job;                                                  /* 128 bytes */
int pg_warp_perspective_u8(const pg_warp_job* jobs_device, int32_t njobs, int32_t max_dst_pixels, void* stream);
int pg_patch_denorm_u8(const void* patches, const void* masks, const double* m, const void* valid, void* denorm, void* part_masks,
                       int32_t B, int32_t P, int32_t h, int32_t w, int32_t H, int32_t W, void* stream);
/*   pg_patch_crop_transforms: HOST function (no kernel, all pointers are host memory) — the crop geometry of the whole batch: for every sample b and
 *     body part p (the ten parts of dataset.py:847-857) what get_crop (dataset.py:751-836) returns,
 *       M[b][p]     = cv2.getPerspectiveTransform(part quadrilateral, patch corners)     M_inv[b][p] = the transform back,
 *     and the two matrices cv2.warpPerspective derives from them, to_patch = inv(M) and to_image = inv(M_inv) (either may be NULL), all row-major
 *     doubles [B][10][9]; valid[b][p] = 0 (and zero matrices) where the part's joints are not confident (>= 0.1) even after the reference's fall-backs.
 *     keypoints [B][18][3] doubles (x, y, confidence) in the un-padded 192-wide frame, OpenPose-18 order of dataset.py:859-861; patch h x w;
 *     o_h = image height; ar = the box aspect (0.5 in the reference).  Bit-equal to OpenCV's results (same operations, same order, no FMA). */
int pg_patch_crop_transforms(const double* keypoints, int32_t B, int32_t h, int32_t w, int32_t o_h, double ar,
                             double* M, double* M_inv, double* to_patch, double* to_image, uint8_t* valid);

/* ---------------------------------------------------------------------------------------------
 * torgb_skip — the ToRGB skip path of a synthesis block in one streaming kernel (north_star kernel 3):
 *
 *   out[n,o] = upsample2d(img_in[n,o], fir)  +  clamp( sum_c w[o,c] * styles[n,c] * x[n,c] + bias[o] )
 *
 * Replaces upfirdn2d.upsample2d + modulated_conv2d(demodulate=False, 1x1) + bias_act(linear, clamp) + img.add_(y)
 * (training/networks.py:5709-5715, ToRGBLayerFull.forward :5601-5611; upfirdn2d.py:308-343).
 * x [N,C,H,W] fp32 dense NCHW (16-byte aligned, W % 4 == 0); w [O,C] (the 1x1 kernel), O <= 8; styles [N,C] or NULL (already
 * multiplied by the layer's weight_gain); bias [O] or NULL; img_in [N,O,H/2,W/2] or NULL (first block: no skip image);
 * fir: the [4,4] float32 resample filter (needed with img_in); clamp < 0: none; out [N,O,H,W].  Forward only. */
int pg_torgb_skip(const float* x, const float* w, const float* styles, const float* bias, const float* img_in, const float* fir,
                  float* out, int32_t N, int32_t C, int32_t O, int32_t H, int32_t W, float clamp, void* stream);
/* The same with x as channel-blocked float16 (PG_LAYOUT_C8, [N][C/8][H*W][8], C % 8 == 0): half the bytes of the feature-map read. */
int pg_torgb_skip_c8(const void* x_c8, const float* w, const float* styles, const float* bias, const float* img_in, const float* fir,
                     float* out, int32_t N, int32_t C, int32_t O, int32_t H, int32_t W, float clamp, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PASTA_B200_H_ */
