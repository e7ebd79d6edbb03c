/* rovr_b200.h — C ABI of the B200-native ROVR hot path (librovr_b200.so).
 *
 * The reference (arjvik/Reinformcement-Optimized-Video-Reconstruction) has NO native code and no
 * FFI: its boundary is the Python nn.Module surface. This header is therefore a
 * new design: one plain-C entry point per fused operator and direction, bound from Python with
 * ctypes by the drop-in modules. Each declaration cites the reference lines whose arithmetic it
 * replaces.
 *
 * Conventions
 *   - every pointer is a raw CUDA device pointer (tensor.data_ptr()); the caller allocates all
 *     outputs and workspaces; nothing here allocates or synchronises;
 *   - activations are NHWC bf16; `*_ld` is the channel stride of one pixel in elements, so a
 *     pointer may address a channel slice of a wider (concat) buffer; slices start on multiples
 *     of 8 channels and ld is a multiple of 8;
 *   - parameters and parameter gradients are fp32 in the PyTorch layouts of the reference
 *     state_dict (Conv2d [Cout][Cin][kh][kw], ConvTranspose2d [Cin][Cout][kh][kw]);
 *   - `stream` is a cudaStream_t passed as void*;
 *   - return 0 on success, negative on error; rovr_last_error() describes the last failure on the
 *     calling thread. There is no CPU fallback: without an sm_100 device every compute call fails.
 */
#ifndef ROVR_B200_H
#define ROVR_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

int rovr_abi_version(void);
const char* rovr_last_error(void);
/* 0 if the current device is sm_100 (B200); negative otherwise. */
int rovr_device_check(void);
/* last mbarrier-timeout code recorded by a kernel (0 = none); debugging aid. */
int rovr_hang_code(unsigned int* code);
/* number of kernels this library has launched in this process (bench.py reports it). */
unsigned long long rovr_launch_count(void);

/* ---- layout packing ------------------------------------------------------------------------
 * cat of up to three NCHW fp32 tensors along C, converted to NHWC bf16 padded to cpad channels.
 * Replaces torch.cat + rearrange 'b n c h w -> b (n c) h w' (rovr/local_net.py:48-49) and
 * torch.cat([x, context], 1) (rovr/policy_net_1.py:88). */
int rovr_pack_nchw_to_nhwc(const float* s0, int c0, const float* s1, int c1, const float* s2, int c2,
                           void* dst, int B, int H, int W, int cpad, void* stream);
int rovr_unpack_nhwc_to_nchw(const void* src, int ld, float* dst, int B, int H, int W, int C,
                             void* stream);
/* torchvision ToTensor on the device: dst[i] = src[i] / denom (uint8 -> fp32, denom = 255; IEEE division,
 * bit-identical to the host's .div(255)). The reference
 * decodes frames to uint8 and converts on the host (rovr/video_ds.py:107-121); feeding the uint8 frames and
 * converting here cuts the host -> device bytes of a step 4x. */
int rovr_u8_to_f32(const void* src, float* dst, long long n, float denom, void* stream);
/* the dataset's mask corruption on the device (rovr/video_ds.py:62-87, the deterministic box of every difficulty):
 * image i (NCHW fp32, clip-frame index frame_index[i], device int64) gets a zeroed box_w x box_h box at
 * x0 = (n % 8) * W / 8, y0 = (n / 8) * H / 3, clipped; out = clean * mask; mask (optional) as in the dataset. */
int rovr_corrupt_frames(const float* clean, const long long* frame_index, float* out, float* mask, int N, int C,
                        int H, int W, int box_w, int box_h, void* stream);

/* ---- weight repacking (fp32 parameter -> bf16 K-major GEMM operand) ------------------------- */
/* Conv2d 3x3 weight [Cout][Cin][3][3] -> [Cout][9][cin_pad] */
int rovr_repack_conv3x3_fprop(const float* w, void* wk, int Cout, int Cin, int cin_pad, void* stream);
/* Conv2d 3x3 weight -> [cin_pad][9][Cout] with the taps flipped (t -> 8 - t; rows >= Cin are zero):
 * the data gradient is then the forward kernel applied to dy with these weights */
int rovr_repack_conv3x3_dgrad(const float* w, void* wk, int Cout, int Cin, int cin_pad, void* stream);
/* ConvTranspose2d 2x2 weight [Cin][Cout][2][2] -> [4*Cout][Cin] (row = q*Cout+co, q = 2*ky+kx) */
int rovr_repack_convT2x2_fprop(const float* w, void* wk, int Cin, int Cout, void* stream);
/* ConvTranspose2d 2x2 weight -> [Cin][4*Cout] */
int rovr_repack_convT2x2_dgrad(const float* w, void* wk, int Cin, int Cout, void* stream);

/* Every weight tensor of a network re-packed by ONE launch (after an optimizer step; same layouts as the
 * single calls above). kind: 0 = Conv2d 3x3 forward operand (a = Cout, b = Cin, c = cin_pad), 1 = Conv2d 3x3
 * data-gradient operand (same), 2 = ConvTranspose2d 2x2 forward operand (a = Cin, b = Cout), 3 = its
 * data-gradient operand (same). n <= 40. */
typedef struct rovr_repack_item {
  const float* w;
  void* wk;
  int kind, a, b, c;
} rovr_repack_item;
int rovr_repack_batch(const rovr_repack_item* items, int n, void* stream);
/* nn.Linear / Conv2d 1x1 weight [N][K] -> bf16 [n_pad][k_pad], or transposed [k_pad][n_pad] */
int rovr_repack_linear(const float* w, void* wk, int N, int K, int n_pad, int k_pad, int transpose,
                       void* stream);

/* ---- Conv2d 3x3, padding 1 (+bias, +ReLU) ---------------------------------------------------
 * Replaces nn.Conv2d(k=3,p=1) + F.relu: rovr/local_net.py:12-18,26,31,36,52-68;
 * rovr/policy_net_1.py:19-47,61-81; rovr/policy_net_2.py:42-54. Cin, Cout multiples of 16. */
int rovr_conv3x3_fprop(const void* x, int x_ld, const void* wk, const float* bias, void* y, int y_ld,
                       int B, int H, int W, int Cin, int Cout, int relu, void* stream);
/* same, and additionally pooled = nn.MaxPool2d(2, 2)(y) from the same epilogue (the encoder's
 * conv -> ReLU -> pool of rovr/local_net.py:52-55 in one kernel); needs W >= 8, H >= 16, even H, W. */
int rovr_conv3x3_fprop_pool2(const void* x, int x_ld, const void* wk, const float* bias, void* y, int y_ld,
                             void* pooled, int pooled_ld, int B, int H, int W, int Cin, int Cout, int relu,
                             void* stream);
/* conv7 + ReLU with the LocalNet tail fused into its epilogue: out = sigmoid(conv8_1x1(y)) (NCHW fp32
 * [B][3][H][W]) and, if target != NULL, *loss = mean((out - target)^2) — rovr/local_net.py:68,71 and
 * nn.MSELoss of rovr/train_local_net_unet.py:90,107 without a second pass over y. Cout must be 64.
 * ws >= rovr_conv3x3_fprop_tail_workspace(B, H, W) when a loss is requested. */
size_t rovr_conv3x3_fprop_tail_workspace(int B, int H, int W);
int rovr_conv3x3_fprop_tail(const void* x, int x_ld, const void* wk, const float* bias, void* y, int y_ld,
                            const float* w8, const float* b8, float* out, const float* target, float* loss,
                            void* ws, size_t ws_bytes, int B, int H, int W, int Cin, int Cout, void* stream);
/* dx = conv3x3^T(dy); if mask != NULL, dx[..., c] *= (mask[..., c] > 0) for c < mask_cols (ReLU of the
 * producer of x; mask_cols <= 0 means all Cin channels — the skip half of a U-Net concat gradient is
 * masked by the pool backward that consumes it, so only the up-conv half needs it here).
 * If colsum != NULL it receives sum over pixels of dx[pixel][c] for c < colsum_cols <= Cin (fp32) —
 * the bias gradient of the layer that produced those channels of x — computed in the epilogue;
 * ws >= rovr_dgrad_colsum_workspace(B,H,W,Cin). */
size_t rovr_dgrad_colsum_workspace(int B, int H, int W, int C);
int rovr_conv3x3_dgrad(const void* dy, int dy_ld, const void* wk_d, void* dx, int dx_ld,
                       const void* mask, int mask_ld, int mask_cols, int B, int H, int W, int Cin, int Cout,
                       float* colsum, int colsum_cols, void* ws, size_t ws_bytes, void* stream);
size_t rovr_conv3x3_wgrad_workspace(int B, int H, int W, int Cin, int Cout);
/* dw[Cout][cin_keep][3][3] (fp32) = sum_pixels dy (x) x_shifted; Cin is the padded channel count
 * of x, cin_keep <= Cin the true one. */
int rovr_conv3x3_wgrad(const void* dy, int dy_ld, const void* x, int x_ld, float* dw, int B, int H,
                       int W, int Cin, int cin_keep, int Cout, void* ws, size_t ws_bytes,
                       void* stream);

/* ---- ConvTranspose2d k=2 s=2 (+bias, +ReLU), H and W are the INPUT resolution ---------------
 * Replaces nn.ConvTranspose2d(k=2,s=2) + F.relu: rovr/local_net.py:24,29,34,58,62,66;
 * rovr/policy_net_1.py:32,37,42,67-77. y is written pixel-shuffled straight into its (concat)
 * destination: the torch.cat of rovr/local_net.py:59,63,67 never materialises. */
int rovr_convT2x2_fprop(const void* x, int x_ld, const void* wk, const float* bias, void* y, int y_ld,
                        int B, int H, int W, int Cin, int Cout, int relu, void* stream);
int rovr_convT2x2_dgrad(const void* dy, int dy_ld, const void* wk_d, void* dx, int dx_ld,
                        const void* mask, int mask_ld, int B, int H, int W, int Cin, int Cout,
                        float* colsum, int colsum_cols, void* ws, size_t ws_bytes, void* stream);
size_t rovr_convT2x2_wgrad_workspace(int B, int H, int W, int Cin, int Cout);
int rovr_convT2x2_wgrad(const void* dy, int dy_ld, const void* x, int x_ld, float* dw, int B, int H,
                        int W, int Cin, int Cout, void* ws, size_t ws_bytes, void* stream);

/* ---- plain GEMM on the same engine: y[M][N] = x[M][K] . wk[N][K]^T + bias (bf16 in, bf16 or fp32 out)
 * Replaces nn.Linear / 1x1 convolutions on bf16 activations. K multiple of 16, N multiple of 16. */
int rovr_gemm_bf16(const void* x, int x_ld, const void* wk, const float* bias, void* y_bf16,
                   float* y_f32, int y_ld, int M, int N, int K, int relu, void* stream);

/* weight gradient of the same: dw[n_keep][k_keep] (fp32, dense) = sum_m dy[m][n] x[m][k] over bf16
 * row-major operands (N, K multiples of 16 are the padded widths). Used for 1x1 convolutions
 * (rovr/policy_net_1.py:46,48) and the attention / feed-forward projections
 * (rovr/common_layers.py:58,70,84-85). */
size_t rovr_gemm_wgrad_workspace(long long M, int N, int K);
int rovr_gemm_wgrad(const void* dy, int dy_ld, const void* x, int x_ld, float* dw, long long M, int N,
                    int n_keep, int K, int k_keep, void* ws, size_t ws_bytes, void* stream);

/* ---- max pooling -----------------------------------------------------------------------------
 * nn.MaxPool2d: rovr/local_net.py:21,53-55; rovr/policy_net_1.py:29; rovr/policy_net_2.py:45-58. */
int rovr_maxpool_fwd(const void* x, int x_ld, void* y, int y_ld, int B, int H, int W, int C, int kh,
                     int kw, int sh, int sw, void* stream);
/* gx = [relu_mask: (x > 0) *] (gskip + maxpool_backward(gp)); gskip may be NULL. If colsum != NULL
 * (non-overlapping windows, C <= 256) it receives sum over pixels of gx[pixel][c] (fp32 [C]) — the bias
 * gradient of the convolution that produced x — from the same pass;
 * ws >= rovr_maxpool_bwd_colsum_workspace(...). */
size_t rovr_maxpool_bwd_colsum_workspace(int B, int H, int W, int C, int kh, int kw);
int rovr_maxpool_bwd(const void* x, int x_ld, const void* gp, int gp_ld, const void* gskip, int gs_ld,
                     void* gx, int gx_ld, int B, int H, int W, int C, int kh, int kw, int sh, int sw,
                     int relu_mask, float* colsum, void* ws, size_t ws_bytes, void* stream);

/* ---- LocalNet tail: conv8 1x1 (64->3) + sigmoid (+ fused L2 loss) ------------------------------
 * rovr/local_net.py:39,71; nn.MSELoss of rovr/train_local_net_unet.py:90,107.
 * out: NCHW fp32 [B][3][H][W]. If target != NULL, *loss = mean((out-target)^2) (ws >= rovr_tail_workspace). */
size_t rovr_tail_workspace(int B, int H, int W);
int rovr_tail_fwd(const void* y7, const float* w8, const float* b8, float* out, const float* target,
                  float* loss, void* ws, size_t ws_bytes, int B, int H, int W, void* stream);
/* g7 (NHWC bf16, ld 64), dw8[3][64], db8[3], db7[64] (= column sums of g7, conv7's bias
 * gradient; may be NULL). dL/dout = gout (NCHW fp32, may be NULL) + mse_scale * (*gloss) *
 * (out - target) when target != NULL (mse_scale = 2 / numel; gloss is the device scalar
 * dL/dloss, NULL = 1). */
int rovr_tail_bwd(const void* y7, const float* w8, const float* out, const float* gout,
                  const float* target, float mse_scale, const float* gloss, void* g7, float* dw8,
                  float* db8, float* db7, void* ws, size_t ws_bytes, int B, int H, int W,
                  void* stream);

/* ---- train-mode BatchNorm2d (+ReLU) on NHWC bf16 ----------------------------------------------
 * nn.BatchNorm2d in training mode followed by F.relu: rovr/policy_net_1.py:20-49,61-81;
 * rovr/policy_net_2.py:43-55. x is the convolution output (bias included), y the activation.
 * Batch statistics over npix = B*H*W pixels; running_mean / running_var (momentum, unbiased
 * variance) and num_batches_tracked (int64) are updated in place when non-NULL. Channels
 * >= c_valid are zero padding: y is written as 0 there and they carry no statistics.
 * mean / rstd (fp32 [C]) are saved for the backward pass. ws >= rovr_bn_workspace(C). */
size_t rovr_bn_workspace(int C);
int rovr_bn_train_fwd(const void* x, int x_ld, void* y, int y_ld, long long npix, int C, int c_valid,
                      const float* gamma, const float* beta, float eps, float momentum,
                      float* running_mean, float* running_var, long long* num_batches_tracked,
                      float* mean, float* rstd, int relu, void* ws, size_t ws_bytes, void* stream);
/* dx = dBN(dy * (y > 0)); dgamma / dbeta fp32 [c_valid]. */
int rovr_bn_train_bwd(const void* dy, int dy_ld, const void* y, int y_ld, const void* x, int x_ld,
                      void* dx, int dx_ld, long long npix, int C, int c_valid, const float* gamma,
                      const float* mean, const float* rstd, float* dgamma, float* dbeta, int relu,
                      void* ws, size_t ws_bytes, void* stream);

/* Train-mode BatchNorm with PER-FRAME statistics over `frames` groups of `pix` pixels (frame-major NHWC; x is the
 * convolution output, bf16 or — x_f32 != 0, x_ld in floats — fp32, so that the statistics and the centring see the
 * unrounded values; y is always bf16):
 * the reference encodes one frame per call (rovr/resnet_extractor.py:42-47, `.unsqueeze(0)`), so a trunk left in
 * training mode (the default constructor, rovr/resnet_extractor.py:6-8) normalises each frame with its own
 * statistics and updates running_mean / running_var once per frame, in frame order; num_batches_tracked += frames.
 * mean / rstd: fp32 [frames][C]. x == y is allowed for a bf16 x. ws >= rovr_bn_frames_workspace(C, frames, pix). */
size_t rovr_bn_frames_workspace(int C, int frames, long long pix);
int rovr_bn_train_fwd_frames(const void* x, int x_f32, int x_ld, void* y, int y_ld, int frames, long long pix, int C,
                             int c_valid, const float* gamma, const float* beta, float eps, float momentum,
                             float* running_mean, float* running_var, long long* num_batches_tracked,
                             float* mean, float* rstd, int relu, void* ws, size_t ws_bytes, void* stream);

/* eval mode (module.eval()): running statistics, buffers untouched. rstd (fp32 [C]) is produced for the
 * backward pass, whose statistics are constants: dx = gamma * rstd * dy * (y > 0). */
int rovr_bn_eval_fwd(const void* x, int x_ld, void* y, int y_ld, long long npix, int C, int c_valid,
                     const float* gamma, const float* beta, float eps, const float* running_mean,
                     const float* running_var, float* rstd, int relu, void* stream);
int rovr_bn_eval_bwd(const void* dy, int dy_ld, const void* y, int y_ld, const void* x, int x_ld, void* dx,
                     int dx_ld, long long npix, int C, int c_valid, const float* gamma,
                     const float* running_mean, const float* rstd, float* dgamma, float* dbeta, int relu,
                     void* ws, size_t ws_bytes, void* stream);

/* ---- LayerNorm over the last dim of fp32 rows [rows][E] ---------------------------------------
 * nn.LayerNorm: rovr/common_layers.py:59,63,71-72,75-76,86,90. y_f32 and / or y_bf16 may be NULL. */
int rovr_layernorm_fwd(const float* x, long long rows, int E, float eps, const float* gamma,
                       const float* beta, float* y_f32, void* y_bf16, float* mean, float* rstd,
                       void* stream);
size_t rovr_layernorm_workspace(int E);
int rovr_layernorm_bwd(const float* g, const float* x, long long rows, int E, const float* gamma,
                       const float* mean, const float* rstd, float* dx, int accumulate, float* dgamma,
                       float* dbeta, void* ws, size_t ws_bytes, void* stream);

/* ---- fp32 nn.Linear for small batches (weight-read-bound GEMV family) ---------------------------
 * rovr/policy_net_1.py:54-57,94; rovr/policy_net_2.py:63-69,79; rovr/resnet_extractor.py:9,46;
 * rovr/action_lstm.py:13-14,33,35. w is the PyTorch [N][K] weight. */
int rovr_linear_f32_fwd(const float* x, int x_ld, const float* w, const float* bias, float* y, int y_ld,
                        int M, int N, int K, int accumulate, void* stream);
size_t rovr_linear_f32_dgrad_workspace(int M, int N, int K);
/* dx[M][K] (dense) = dy[M][N] . w[N][K] */
int rovr_linear_f32_dgrad(const float* dy, int dy_ld, const float* w, float* dx, int M, int N, int K,
                          int accumulate, void* ws, size_t ws_bytes, void* stream);
/* dw[N][K] = dy^T x, db[N] = column sums of dy (db may be NULL) */
int rovr_linear_f32_wgrad(const float* dy, int dy_ld, const float* x, int x_ld, float* dw, float* db,
                          int M, int N, int K, void* stream);

/* ---- standardisation (x - mean) / (std_unbiased + eps_add) along one strided dimension -----------
 * per-sample over 400 features (eps 0): rovr/policy_net_1.py:91-93; over the batch dim (eps .001):
 * rovr/policy_net_2.py:104-106. Element (o, i) lives at o*outer_stride + i*inner_stride. */
int rovr_standardize_fwd(const float* x, float* y, float* sig, int outer, int len, long long outer_stride,
                         long long inner_stride, float eps_add, void* stream);
int rovr_standardize_bwd(const float* g, const float* y, const float* sig, float* dx, int outer, int len,
                         long long outer_stride, long long inner_stride, float eps_add, void* stream);

/* ---- policy heads on logits [b][n], n <= 32, b <= 1024 --------------------------------------------
 * mask_std: in-place scatter of 0 at target[b][tk] (int64; tk may be 0), then if standardize
 * out = (l - mean(dim=1)[no keepdim]) / (std(dim=1) + 0.1) exactly as rovr/policy_net_2.py:121-122 /
 * rovr/policy_net_1.py:100 broadcast it (b == 1 or b == n only). */
int rovr_head_mask_std_fwd(float* logits, const long long* target, int tk, float* out, float* sig, int b,
                           int n, int standardize, void* stream);
int rovr_head_mask_std_bwd(const float* g, const float* logits, const float* out, const float* sig,
                           const long long* target, int tk, float* dl, int b, int n, int standardize,
                           void* stream);
/* probs = softmax((logits - log(expo)) / tau) — F.gumbel_softmax(hard=False) with the Exp(1) draw
 * `expo` supplied by the caller. mode 0: probs only; 1: idx[b] = argmax, val = log p_max
 * (rovr/policy_net_1.py:102-103); 2: idx[b][2] = top-2, val = (log p1 + log p2)/2 + 0.69314
 * (rovr/policy_net_2.py:100-102); 3: val = log p[action[b]] (policy_net_1.py:114);
 * 4: val = log(p[a0] p[a1])/2 + 0.69314 for action[b][2] (policy_net_2.py:139-141). */
int rovr_head_gumbel_fwd(const float* logits, const float* expo, float tau, float* probs, int b, int n,
                         int mode, const long long* action, long long* idx, float* val, void* stream);
int rovr_head_gumbel_bwd(const float* probs, const float* gval, float tau, int b, int n, int mode,
                         const long long* action, float* dl, void* stream);

/* ---- nn.LSTMCell pointwise part (rovr/action_lstm.py:33); gates = W_ih x + b_ih + W_hh h + b_hh,
 * order i, f, g, o; act [B][4*Hd] receives the activated gates for the backward pass. */
int rovr_lstm_pointwise_fwd(const float* gates, const float* c_prev, float* h, float* c, float* act,
                            int B, int Hd, void* stream);
int rovr_lstm_pointwise_bwd(const float* act, const float* c_prev, const float* c, const float* dh,
                            const float* dc, float* dgates, float* dc_prev, int B, int Hd, void* stream);

/* ---- small data movers ----------------------------------------------------------------------
 * NHWC bf16 -> rows [B][C*H*W] fp32 in NCHW order (nn.Flatten / rearrange 'b c h w -> b (c h w)':
 * rovr/policy_net_2.py:59; rovr/policy_net_1.py:90) and back (zero-padding channels to cpad). */
int rovr_flatten_nhwc(const void* src, int ld, float* dst, int dst_ld, int B, int HW, int C, void* stream);
int rovr_unflatten_nhwc(const float* src, int src_ld, void* dst, int ld, int B, int HW, int C, int cpad,
                        void* stream);
/* dst[r][c] (+)= scale * src[r][c] on fp32 row-strided matrices (torch.cat / slicing of feature rows,
 * rovr/policy_net_2.py:92; rovr/action_lstm.py:31). */
int rovr_copy2d_f32(const float* src, int src_ld, float* dst, int dst_ld, int rows, int cols, float scale,
                    int accumulate, void* stream);

/* ---- ResNet-50 frame-feature extractor helpers ---------------------------------------------------
 * torchvision resnet50 children[:-1] + Linear(2048, 768) as used by rovr/resnet_extractor.py:5-67.
 * Convolutions run on rovr_conv3x3_fprop / rovr_gemm_bf16 with eval-mode BatchNorm folded in. */
/* wf[Cout][K] = w * gamma/sqrt(var+eps), bf[Cout] = beta - mean*gamma/sqrt(var+eps) (frozen + eval
 * trunk, rovr/resnet_extractor.py:11-14) */
int rovr_fold_bn(const float* w, const float* gamma, const float* beta, const float* mean, const float* var,
                 float eps, float* wf, float* bf, int Cout, int K, void* stream);
/* 7x7 stride-2 pad-3 stem: NCHW fp32 [B][3][H][W] -> [B*Ho*Wo][kpad] bf16, k = c*49 + r*7 + s;
 * quantise != 0 applies the ToPILImage -> ToTensor uint8 round trip of rovr/resnet_extractor.py:18-23 */
int rovr_stem_im2col(const float* src, void* dst, int B, int H, int W, int kpad, int quantise, void* stream);
/* nn.MaxPool2d(k, s, pad) with -inf padding (resnet50.maxpool = 3, 2, 1) */
int rovr_maxpool_pad_fwd(const void* x, int x_ld, void* y, int y_ld, int B, int H, int W, int C, int k, int s,
                         int pad, void* stream);
/* y[b][oy][ox] = x[b][oy*s][ox*s]: the sampling pattern of a stride-s convolution */
int rovr_subsample(const void* x, int x_ld, void* y, int y_ld, int B, int H, int W, int C, int s, void* stream);
/* out = relu(a + b), dense bf16 (bottleneck residual join) */
int rovr_add_relu(const void* a, const void* b, void* out, long long n, void* stream);
/* AdaptiveAvgPool2d(1): NHWC bf16 [B][HW][C] -> fp32 [B][C] */
int rovr_avgpool(const void* x, int ld, float* out, int B, int HW, int C, void* stream);
/* feature rows [n][ch*tile*tile] fp32 <-> mosaics [nb][ch][side][side] at (slot/per_row*tile,
 * slot%per_row*tile) (rovr/resnet_extractor.py:30-40,49-55); batch / slot (int64 [n]) may be NULL
 * (row r -> mosaic r / slots_per_mosaic, slot r % slots_per_mosaic). gather != 0 reads tiles back. */
int rovr_mosaic_paste(const float* feat, float* mosaic, const long long* batch, const long long* slot, int n,
                      int slots_per_mosaic, int ch, int tile, int per_row, int side, int gather, void* stream);
/* ToPILImage -> transforms.Resize((Ho, Wo)) -> ToTensor of rovr/resnet_extractor.py:18-23: PIL's
 * two-pass 8-bit bilinear (antialiased) resampler; tmp holds BC*H*Wo floats. */
int rovr_resize_antialias(const float* src, float* dst, float* tmp, int BC, int H, int W, int Ho, int Wo,
                          void* stream);

/* ---- attention blocks (rovr/common_layers.py:54-118) -----------------------------------------------
 * batched GEMM on the tcgen05 engine: y[i2][i1][m][n] = sum_k a[i2][i1][m][k] * b[i2][i1][n][k]
 * (QK^T, PV and their gradients; i1 = head, i2 = batch element). Strides in elements, multiples
 * of 8; exactly one of y_bf16 / y_f32 is non-NULL. K, N multiples of 16; b holds n_rows_b <= N rows per
 * matrix (output columns beyond them are zero: key-token padding). */
int rovr_gemm_batched_bf16(const void* a, long long a_ld, long long a_s1, long long a_s2, const void* b,
                           long long b_ld, long long b_s1, long long b_s2, void* y_bf16, float* y_f32,
                           long long y_ld, long long y_s1, long long y_s2, int M, int N, int n_rows_b, int K,
                           int n1, int n2, void* stream);
/* out[i2][i1][c][r] = in[i2][i1][r][c], columns r in [R, r_pad) zero-filled (per-head operand transposes) */
int rovr_transpose_bf16(const void* in, long long in_ld, long long in_s1, long long in_s2, void* out,
                        long long out_ld, long long out_s1, long long out_s2, int R, int C, int r_pad, int n1,
                        int n2, void* stream);
/* P = softmax_t(scale * s) over t < T of fp32 rows [rows][t_pad] -> bf16 (zero in the padding) */
int rovr_softmax_fwd(const float* s, void* p, long long rows, int T, int t_pad, float scale, void* stream);
/* dS = scale * P * (dP - sum_t dP P) */
int rovr_softmax_bwd(const float* dp, const void* p, void* ds, long long rows, int T, int t_pad, float scale,
                     void* stream);
/* exact GELU on bf16 (F.gelu, rovr/common_layers.py:91) and its gradient */
int rovr_gelu_fwd(const void* h, void* a, long long n, void* stream);
int rovr_gelu_bwd(const void* da, const void* h, void* dh, long long n, void* stream);
/* Host-only self-test of the index arithmetic shared by host and kernels (multiply-shift division); no GPU
 * needed. 0 = ok. */
int rovr_host_selftest(void);
/* Tuning knob (no reference counterpart): CTA-pair (cta_group::2, 256-row MMA) launches of the igemm engine.
 * 0 = never, 1 = where measured to pay (default; env ROVR_PAIR overrides at load), 2 = whenever the shape
 * allows (even number of 128-pixel tiles, N tile a multiple of 32). Returns the previous mode. */
int rovr_set_pair_mode(int mode);
int rovr_cast_f32_bf16(const float* src, void* dst, long long n, void* stream);
/* out = a + b, dense fp32 (residual joins, rovr/common_layers.py:101-102,113-115); n multiple of 4 */
int rovr_add_f32(const float* a, const float* b, float* out, long long n, void* stream);
/* positional encodings (rovr/common_layers.py:7-52): out[b][p][j] = x + w1[j]*(p % n1) + b1[j]
 * [+ w2[j]*(p / n1) + b2[j]] with w, b the weight / bias of Linear(1, D); grad: which = 0 -> (w1, b1),
 * 1 -> (w2, b2) */
int rovr_posenc_add(const float* x, float* out, const float* w1, const float* b1, const float* w2,
                    const float* b2, int B, int P, int D, int n1, void* stream);
int rovr_posenc_grad(const float* g, int B, int P, int D, int n1, int which, float* gw, float* gb, void* stream);

/* ---- bias gradient: out[c] = sum over pixels of g[pixel][c]; C even, C <= 512 ------------------ */
size_t rovr_colsum_workspace(int C);
int rovr_colsum(const void* g, int ld, long long npix, int C, float* out, void* ws, size_t ws_bytes,
                void* stream);

/* the same for a wide bf16 row-major matrix [M][C] with row stride ld, any even C (bias gradients of
 * nn.Linear / in-projection layers, rovr/common_layers.py:58,70,84-85) */
size_t rovr_colsum_rows_workspace(int C);
int rovr_colsum_rows(const void* g, long long ld, long long M, int C, float* out, void* ws, size_t ws_bytes,
                     void* stream);

/* ---- emulated-fp32 policy trunks (csrc/fp32x.cuh) ------------------------------------------------
 * PolicyNetwork1UNet.unet (rovr/policy_net_1.py:60-84) and PolicyNetwork2UNet.video_conv
 * (rovr/policy_net_2.py:41-60) keep fp32 NHWC activations (`*_ld` in floats, multiples of 4) and run
 * their convolutions as split-bf16 products on the tensor cores: an fp32 value is v = hi + mid + lo
 * (three bf16 pieces) and the pieces are stacked along the contraction dimension, so one pass of the
 * implicit-GEMM kernel accumulates hi*hi + mid*hi + lo*hi + hi*mid + mid*mid + hi*lo in fp32. */
/* dst[b][oy][ox][t*cb + c_off + c] = piece(t) of max over the kh x kw window (stride sh, sw) of
 * src[b][.][.][c] for c < C, zero for C <= c < cw; t < nterms. nterms 6: pieces [h m l h m h] (forward
 * operand), 3: [h m h] (gradient operand), 2: [h m]. src is addressed through element strides (batch, row,
 * column, channel), so NHWC views and NCHW tensors both work; with src2 != NULL channels c_split .. C are
 * read from src2 (same strides): torch.cat([image, context], 1) of rovr/policy_net_1.py:88 without the
 * copy. relu_mask != NULL (same strides as src, no pooling / second source): values are taken as 0 where the
 * mask is <= 0 — the ReLU backward of the producing layer folded into the gradient split.
 * kh = kw = sh = sw = 1: no pooling.
 * Replaces F.max_pool2d + the operand conversion in front of every convolution of the trunks. */
int rovr_split_stack(const float* src, const float* src2, int c_split, const float* relu_mask, long long s_b, long long s_y,
                     long long s_x, long long s_c, int B, int H, int W, int C, int kh, int kw, int sh, int sw,
                     void* dst, int dst_ld, int cb, int c_off, int cw, int nterms, void* stream);
/* weight side of the same products: w fp32 [d0][d1][inner] -> out fp32 with dim stack_dim (0 / 1) replaced
 * by nterms blocks of cb entries holding pieces [h h h m m l] (nterms 6) or [h h m] (3); zero padding for
 * entries >= the original extent. The rovr_repack_* calls then make the bf16 operands (exactly). */
int rovr_split_weights(const float* w, float* out, int d0, int d1, int inner, int stack_dim, int nterms,
                       int cb, void* stream);
/* weight gradient of a stacked product: dwp fp32 [2*cb0][2*cb1][inner] -> dw[i][j][t] = sum of its four
 * blocks (i < d0, j < d1) */
int rovr_blocksum4(const float* dwp, float* dw, int d0, int d1, int inner, int cb0, int cb1, void* stream);
/* Conv2d 3x3, padding 1, stride 2 (torchvision ResNet-50 layer{2,3,4}.0.conv2, rovr/resnet_extractor.py:8,16): x is
 * [B][H][W][Cin], y is [B][(H-1)/2+1][(W-1)/2+1][Cout]; the operand map samples every other input pixel (TMA element
 * strides), so nothing is computed at the discarded positions. */
int rovr_conv3x3_fprop_s2(const void* x, int x_ld, const void* wk, const float* bias, void* y, int y_ld, int B, int H,
                          int W, int Cin, int Cout, int relu, void* stream);
/* the same with an fp32 NHWC output (y_ld in floats): the train-mode trunk's BatchNorm reads the unrounded values */
int rovr_conv3x3_fprop_s2_f32out(const void* x, int x_ld, const void* wk, const float* bias, float* y, int y_ld, int B,
                                 int H, int W, int Cin, int Cout, int relu, void* stream);
/* Conv2d 3x3 pad 1 with fp32 NHWC output (forward; or, with the rovr_repack_conv3x3_dgrad operand and
 * Cin / Cout swapped by the caller, the data gradient) */
int rovr_conv3x3_f32out(const void* x, int x_ld, const void* wk, const float* bias, float* y, int y_ld,
                        int B, int H, int W, int Cin, int Cout, int relu, void* stream);
/* ConvTranspose2d k2 s2 with fp32 NHWC output / input gradient */
int rovr_convT2x2_fprop_f32out(const void* x, int x_ld, const void* wk, const float* bias, float* y, int y_ld,
                               int B, int H, int W, int Cin, int Cout, int relu, void* stream);
int rovr_convT2x2_dgrad_f32out(const void* dy, int dy_ld, const void* wk_d, float* dx, int dx_ld, int B, int H,
                               int W, int Cin, int Cout, void* stream);
/* nn.BatchNorm2d (+ReLU) on fp32 NHWC views, semantics of rovr_bn_train_fwd / rovr_bn_eval_fwd /
 * rovr_bn_train_bwd (two-pass centred variance, fp64 combine). ws >= rovr_bn_workspace(C). */
int rovr_bn_f32_train_fwd(const float* x, int x_ld, float* y, int y_ld, long long npix, int C, int c_valid,
                          const float* gamma, const float* beta, float eps, float momentum,
                          float* running_mean, float* running_var, long long* num_batches_tracked,
                          float* mean, float* rstd, int relu, void* ws, size_t ws_bytes, void* stream);
int rovr_bn_f32_eval_fwd(const float* x, int x_ld, float* y, int y_ld, long long npix, int C, int c_valid,
                         const float* gamma, const float* beta, float eps, const float* running_mean,
                         const float* running_var, float* rstd, int relu, void* stream);
int rovr_bn_f32_bwd(const float* dy, int dy_ld, const float* y, int y_ld, const float* x, int x_ld, float* dx,
                    int dx_ld, long long npix, int C, int c_valid, const float* gamma, const float* mean,
                    const float* rstd, float* dgamma, float* dbeta, int relu, int eval_mode, void* ws,
                    size_t ws_bytes, void* stream);
/* out[c] = sum over pixels of an fp32 NHWC view; ws >= rovr_bn_workspace(C) */
int rovr_colsum_f32(const float* g, int ld, long long npix, int C, float* out, void* ws, size_t ws_bytes,
                    void* stream);
/* nn.MaxPool2d on fp32 NHWC views (arg-max = first maximum in row-major window order, like ATen);
 * bwd: gx = [gskip +] scatter of gp to the arg-max of each window */
int rovr_maxpool_f32_fwd(const float* x, int x_ld, float* y, int y_ld, int B, int H, int W, int C, int kh,
                         int kw, int sh, int sw, void* stream);
int rovr_maxpool_f32_bwd(const float* x, int x_ld, const float* gp, int gp_ld, const float* gskip, int gs_ld,
                         float* gx, int gx_ld, int B, int H, int W, int C, int kh, int kw, int sh, int sw,
                         void* stream);
/* nn.Flatten on the NCHW view of an fp32 NHWC tensor and its inverse: rows[b][c*HW + p] <-> x[b][p][c] */
int rovr_flatten_f32(const float* x, int ld, float* rows, long long rows_ld, int B, int HW, int C, void* stream);
int rovr_unflatten_f32(const float* rows, long long rows_ld, float* x, int ld, int B, int HW, int C, int cpad,
                       void* stream);

/* ---- LPIPS(net='vgg') perceptual loss (csrc/lpips.cuh) ------------------------------
 * lpips.LPIPS(net='vgg')(in0, in1[, normalize=True]) of rovr/train_local_net_unet.py:91,109 and
 * rovr/rovr.py:54,255. The VGG16 convolutions run through rovr_conv3x3_fprop[_pool2] / rovr_conv3x3_dgrad /
 * rovr_maxpool_bwd on a 2N batch ([in0 ; in1]); these are the remaining pieces. */
/* ScalingLayer + pack: dst [2N][H][W][16] bf16, channel c < 3 = ((normalize ? 2v-1 : v) - shift[c]) / scale[c];
 * shift3 / scale3 are HOST pointers to three floats. */
int rovr_lpips_pack(const float* in0, const float* in1, void* dst, int N, int H, int W, const float* shift3,
                    const float* scale3, int normalize, void* stream);
/* one tap: feats [2N][hw][C] bf16, lin_w fp32 [C]; partial [N][nblocks] receives per-block sums of
 * sum_c w_c (f0_c/(|f0|+1e-10) - f1_c/(|f1|+1e-10))^2; grad (optional, [N][hw][C] bf16) = (1/hw) * d/d f0 of it,
 * masked by f0 > 0. nblocks = rovr_lpips_head_blocks(N, hw). */
int rovr_lpips_head_blocks(int N, long long hw);
int rovr_lpips_head(const void* feats, int N, long long hw, int C, const float* lin_w, void* grad, float* partial,
                    int nblocks, void* stream);
/* fp32 variants for LPIPS(precision="fp32x") (VGG16 on the emulated-fp32 path of csrc/fp32x.cuh): packed input
 * [2N][H][W][4] fp32, fp32 features / feature gradients, gradient w.r.t. the packed input [N][H][W][ld] fp32 */
int rovr_lpips_pack_f32(const float* in0, const float* in1, float* dst, int N, int H, int W, const float* shift3,
                        const float* scale3, int normalize, void* stream);
int rovr_lpips_head_f32(const float* feats, int N, long long hw, int C, const float* lin_w, float* grad,
                        float* partial, int nblocks, void* stream);
int rovr_lpips_unpack_grad_f32(const float* gx, int ld, const float* gval, float* gout, int N, int H, int W,
                               const float* shift3, const float* scale3, int normalize, void* stream);
/* val[n] = sum over taps of (sum of partial blocks) / hw; partials / nblocks / hws are HOST arrays of ntaps entries */
int rovr_lpips_finalize(const float* const* partials, const int* nblocks, const long long* hws, int ntaps, int N,
                        float* val, void* stream);
/* gradient w.r.t. the packed input [N][H][W][16] bf16 -> d/d in0 NCHW fp32, times the per-image upstream
 * gradient gval[n] (device, fp32) */
int rovr_lpips_unpack_grad(const void* gx16, const float* gval, float* gout, int N, int H, int W,
                           const float* shift3, const float* scale3, int normalize, void* stream);

/* ---- dropout (rovr/common_layers.py:58,70: nn.MultiheadAttention(dropout=p) on the attention probabilities;
 * :87,91: nn.Dropout after the GELU). out = x * keep / (1 - p); the keep mask is a pure function of
 * state = {seed, counter} (device int64 [2]), the call site and the element index, so backward recomputes it
 * instead of storing it. rovr_dropout_advance increments the counter (stream-ordered, graph-capturable). */
int rovr_dropout(const void* x, void* out, long long n, int is_f32, float p, const long long* state, int site,
                 void* stream);
int rovr_dropout_advance(long long* state, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ROVR_B200_H */
