/*
 * yad_b200.h - C ABI of the B200-native audio-activity-detection hot path.
 *
 * The reference (ches-001/YOLO-inspired-audio-activity-detection) is pure
 * Python and has no FFI layer; every entry point below replaces a *library
 * call site* of the reference (file:line relative to the reference root,
 * "[ta]" = torchaudio 2.11, "[tv]" = torchvision 0.26).  The Python host
 * package (yad_b200) binds these with ctypes; INTEGRATION.md shows the stub a
 * reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; the message is in
 *     yad_last_error() (thread local).  No C++ exception crosses the ABI.
 *   - every pointer is a DEVICE pointer owned by the caller (inputs, outputs
 *     and workspaces); the library never allocates or frees device memory.
 *   - every call is asynchronous on `stream` (a cudaStream_t / CUstream).
 *   - no mutable global state after yad_init(): calls are re-entrant (the
 *     reference shares one model across 10 threads, inference.py:212-236).
 *   - no CPU fallback: unsupported shapes / architectures return an error.
 *   - activations are NHWC ("pixel-major"): element (b,h,w,c) of a tensor with
 *     channel pitch `ld` lives at ((b*H + h)*W + w)*ld + c.
 */
#ifndef YAD_B200_H_
#define YAD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* yad_stream_t;              /* cudaStream_t */

enum { YAD_F32 = 0, YAD_BF16 = 1 };       /* element types */
enum { YAD_ACT_NONE = 0, YAD_ACT_RELU = 1, YAD_ACT_LRELU02 = 2 };

enum {
  YAD_OK = 0,
  YAD_ERR_ARG = -1,        /* bad argument / unsupported shape */
  YAD_ERR_CUDA = -2,       /* a CUDA runtime / driver call failed */
  YAD_ERR_ARCH = -3,       /* device is not sm_100 */
  YAD_ERR_WORKSPACE = -4   /* caller workspace too small */
};

int yad_version(void);
const char* yad_last_error(void);
/* Checks the device is compute capability 10.x, resolves cuTensorMapEncodeTiled,
 * raises the dynamic shared-memory limits of the kernels.  Idempotent. */
int yad_init(int device);

/* ------------------------------------------------------------------ frontend
 * Replaces, fused: torchaudio Resample (modules/_architecture.py:25-28,84 ->
 * [ta] functional.py:1405-1431), MelSpectrogram (:30-33,98 -> torch.stft,
 * abs().pow(2), matmul fb), MFCC (:34-37,99), AmplitudeToDB x2 (:29,100-101 ->
 * [ta] functional.py:390-403) and scale_input (:103-105,182-189).
 *
 * Stage A: PCM -> mel power.  Polyphase resample (new_rate/orig_rate reduced to
 * P/O, P % 4 == 0; only the non-zero taps of each phase), Hann window, 1000-point real FFT,
 * |X|^2, sparse mel filterbank.  n_fft = hop = 1000, 32 mels, center=False.
 *   pcm        [B, L]                 f32
 *   taps       [P/4][4][YAD_FE_QW]    f32  resampler.kernel rows of the phase quad (4u .. 4u+3), all four
 *                                          cut to the quad's common window [tap_base[u], +YAD_FE_QW)
 *                                          of the 2*width+O wide tap axis, zero padded
 *   tap_base   [P/4]                  i32
 *   lane_map   [256] or NULL          i32  resample-role thread -> work item: quad | (hop slice << 16), -1 = idle lane.
 *                                          Built by the host so that the 32 lanes of a warp read 32 distinct
 *                                          shared-memory banks ((tap_base[quad] + slice * O) mod 32 distinct per
 *                                          warp); every (quad, slice), slice < min(256 / (P/4), hops per group),
 *                                          must occur exactly once.  NULL: built-in (conflicting) order.
 *   window_len                        max(tap_base) + YAD_FE_QW (span of the padded signal one hop reads)
 *   window     [1000]                 f32  analysis window
 *   twiddle    [1000][2]              f32  exp(-2*pi*i*k/1000), built in fp64 by the host
 *   fb_val [nnz] f32, fb_bin [nnz] i32, fb_start [33] i32   mel filterbank in CSR form over mel bands; the
 *                                          bins of one band must be contiguous (triangular filters are)
 *   mel        [B, 32, T]             f32  (T frames; T*1000 <= ceil(P*L/O))
 */
#define YAD_FE_QW 21
int yad_frontend_mel_power(const float* pcm, int64_t B, int64_t L, int32_t P, int32_t O, int32_t width,
                           const float* taps, const int32_t* tap_base, const int32_t* lane_map,
                           int32_t window_len, const float* window, const float* twiddle, const float* fb_val,
                           const int32_t* fb_bin, const int32_t* fb_start, int32_t fb_nnz, float* mel,
                           int64_t T, yad_stream_t stream);

/* The same from 16-bit PCM [B, L] (the sample format of audio files): x = pcm / 32768 exactly as torchaudio.load(normalize=True)
 * hands it to the reference (inference.py:137-149), so the result is bit-identical to the fp32 entry point on those values;
 * half the bytes over PCIe and out of HBM. */
int yad_frontend_mel_power_i16(const int16_t* pcm, int64_t B, int64_t L, int32_t P, int32_t O, int32_t width,
                               const float* taps, const int32_t* tap_base, const int32_t* lane_map,
                               int32_t window_len, const float* window, const float* twiddle, const float* fb_val,
                               const int32_t* fb_bin, const int32_t* fb_start, int32_t fb_nnz, float* mel,
                               int64_t T, yad_stream_t stream);

/* Stage A with `taper_input: true` (modules/_architecture.py:87-94): the resampled 16 kHz signal is multiplied by a clip-long
 * window (taper [taper_len] f32, taper_len >= T*1000; the reference's registered buffer `taper_window`, a non-periodic
 * hann / hamming / ... window of the resampled length) before the Hann analysis window, in the reference's order
 * (x * taper) * hann.  pcm is fp32 (pcm_is_i16 = 0) or 16-bit PCM. */
int yad_frontend_mel_power_taper(const void* pcm, int32_t pcm_is_i16, const float* taper, int64_t taper_len, int64_t B, int64_t L,
                                 int32_t P, int32_t O, int32_t width, const float* taps, const int32_t* tap_base,
                                 const int32_t* lane_map, int32_t window_len, const float* window, const float* twiddle,
                                 const float* fb_val, const int32_t* fb_bin, const int32_t* fb_start, int32_t fb_nnz, float* mel,
                                 int64_t T, yad_stream_t stream);

/* Stage B: mel power [B,32,T] -> x_spectral [B,2,32,T] f32 (channel 0 = standardised
 * dB-mel, channel 1 = standardised dB-of-MFCC).  One CTA per clip; T <= 1024.
 * Optional taps (may be NULL): meldb, mfcc, mfdb, each [B,32,T] f32 (pre-standardise). */
int yad_frontend_finish(const float* mel, int64_t B, int64_t T, const float* dct /*[32,32]*/,
                        float top_db, int32_t standardise, float* x_spectral, float* tap_meldb,
                        float* tap_mfcc, float* tap_mfdb, yad_stream_t stream);

/* Stage B that additionally writes x_spectral as channel-interleaved bf16 words (ch0 in the low half) into
 * xs_bf16 [B][32][bf_pitch] at word offset bf_margin + t: the input form of yad_conv_stem_fused (bf_margin = 9 = its left
 * padding; the caller zero-fills the tensor once, the margins are never written).  T <= 1024. */
int yad_frontend_finish_bf16(const float* mel, int64_t B, int64_t T, const float* dct, float top_db, int32_t standardise,
                             float* x_spectral, float* tap_meldb, float* tap_mfcc, float* tap_mfdb, void* xs_bf16,
                             int64_t bf_pitch, int32_t bf_margin, yad_stream_t stream);

/* ------------------------------------------------------------------ convolutions
 * Replace F.conv2d + batch_norm(eval) + activation (+ residual add) call sites:
 * modules/_backbone.py:143-151, [tv] models/resnet.py:89-105, modules/_common.py:43-48,86-95.
 * BatchNorm is folded into (weight, bias) by the host at pack time.
 */
typedef struct {
  int32_t B, H, W;            /* input spatial */
  int32_t Cin, ld_in;         /* input channels used / channel pitch (elements) */
  int32_t Cout, ld_out;       /* output channels written / channel pitch */
  int32_t co_off;             /* first output channel inside the pitch (concat-free writes) */
  int32_t kh, kw, sh, sw, ph, pw;
  int32_t act;                /* YAD_ACT_* */
  int32_t ld_res;             /* residual channel pitch (0 = no residual) */
  /* Optional pixel strides (in pixels; all three 0 = dense NHWC): element (b,h,w,c) of the input lives at
   * (b*in_sb + h*in_sh + w*in_sw)*ld_in + c, of the output / residual at (b*out_sb + h*out_sh + w*out_sw)*ld + c.
   * Used to read / write the halo-padded "flat" layout of yad_conv_flat.  yad_conv_tc needs in_sw <= in_sh <= in_sb. */
  int32_t in_sw, in_sh, in_sb;
  int32_t out_sw, out_sh, out_sb;
} yad_conv_desc;

/* Stem conv1 (2 -> 64, 7x7, stride 2, pad 3): reads x_spectral NCHW f32 directly,
 * writes NHWC (out_dtype) [B, (H-1)/2+1, (W-1)/2+1, 64].  weight [7][7][2][64] f32 (tap-major). */
int yad_conv_stem(const float* x_nchw, int64_t B, int32_t H, int32_t W, const float* weight,
                  void* out, int32_t out_dtype, yad_stream_t stream);

/* The same stem on the tcgen05 tensor cores (bf16 operands, fp32 accumulate, bf16 output).  The CTA keeps the raw
 * channel-interleaved input patch in shared memory and the UMMA descriptor reads overlapping 32-byte windows of it
 * (no-swizzle K-major layout with LBO = 16 B): no im2col copy.
 * weight_packed: bf16 [64][112] (K index = kh*16 + kw*2 + c, zero for kw = 7) stored as 8x8 core matrices:
 * element (n, k) at byte (n/8)*1792 + (k/8)*128 + (n%8)*16 + (k%8)*2.
 * Output: s2d_hp = s2d_wp = 0 -> dense NHWC [B, Ho, Wo, 64].  Otherwise the space-to-depth FLAT layout that lets the next
 * 7x7 stride-2 conv run as a stride-1 conv (yad_conv_flat_taps): pixel (b, ho, wo) goes to flat index
 * (b*s2d_wp + wo/2)*s2d_hp + ho/2 of a 256-channel tensor, channel block ((ho%2)*2 + wo%2)*64; the caller zero-fills
 * the buffer once (halo cells are never written). */
int yad_conv_stem_tc(const float* x_nchw, int64_t B, int32_t H, int32_t W, const void* weight_packed,
                     void* out_bf16, int32_t s2d_hp, int32_t s2d_wp, yad_stream_t stream);

/* conv1 o conv2 of the stem as ONE 19x19 / stride-4 / pad-9 convolution 2 -> 64 (+ bias, ReLU) on the tensor cores: the two
 * reference convolutions (modules/_backbone.py:143-146) have nothing between them, so W12 = conv_transpose(W2 * bn1, W1) and
 * the intermediate [B,64,16,480] tensor never exists.  Reads x_spectral in the padded channel-interleaved bf16 form of
 * yad_frontend_finish_bf16 ([B][32][x_pitch] words, 9-word zero margin on the left, zeros up to x_pitch on the right,
 * x_pitch % 4 == 0 and >= 512*(n_seg-1) + 536: every patch row is one 16-byte aligned bulk copy), H = 32; writes the FLAT
 * halo-padded bf16 layout of yad_conv_flat ([B, Wp, Hp, 64], pixel (b,h,w) at (b*Wp + w)*Hp + h).
 * w_classes: bf16 [4 row classes][64][912] (K = dh*48 + dw*2 + c, zero for dw >= 19) in the 8x8 core-matrix layout of
 * yad_conv_stem_tc (element (n,k) of a class at byte (n/8)*14592 + (k/8)*128 + (n%8)*16 + (k%8)*2); the classes hold the
 * composite built from the conv2 row taps that stay inside conv1's output: 0 = rows 2..6 (all), 1 = row 0 (kh2 >= 3),
 * 2 = row 1 (kh2 >= 1), 3 = row 7 (kh2 <= 4).  n_cta_interior: CTAs given to class 0 (0 = built-in split).
 * Border COLUMNS (0, 1 and the last one or two) are computed with the interior-column weights here; yad_conv_stem_fused_fixup
 * must follow on the same stream: it recomputes them exactly in fp32 from x_spectral NCHW f32.  cols / col_var are HOST arrays (n_cols <= 8);
 * w_var f32 [n_var][4 row classes][19][19][2][64]. */
int yad_conv_stem_fused(const void* x_bf16_padded, int64_t x_pitch, int64_t B, int32_t H, int32_t W, const void* w_classes,
                        const float* bias, void* out_flat_bf16, int32_t Hp, int32_t Wp, int32_t n_cta_interior,
                        yad_stream_t stream);
int yad_conv_stem_fused_fixup(const float* x_nchw, int64_t B, int32_t H, int32_t W, const float* w_var, const float* bias,
                              const int32_t* cols, const int32_t* col_var, int32_t n_cols, void* out_flat_bf16,
                              int32_t Hp, int32_t Wp, yad_stream_t stream);
/* The overlapped pair: yad_conv_stem_fused_skip leaves output columns [0, skip_lo) and [Wo - skip_hi, Wo) unwritten and
 * yad_conv_stem_fused_fixup_bf16 computes exactly those columns (cols must be that set) from the SAME padded bf16 input the
 * tensor-core kernel reads (20 KB of shared memory: its CTAs fit next to the resident stem CTA).  The two write disjoint cells
 * and depend only on the frontend's output, so they may run on different streams at the same time. */
int yad_conv_stem_fused_skip(const void* x_bf16_padded, int64_t x_pitch, int64_t B, int32_t H, int32_t W, const void* w_classes,
                             const float* bias, void* out_flat_bf16, int32_t Hp, int32_t Wp, int32_t n_cta_interior,
                             int32_t skip_lo, int32_t skip_hi, yad_stream_t stream);
int yad_conv_stem_fused_fixup_bf16(const void* x_bf16_padded, int64_t x_pitch, int64_t B, int32_t H, int32_t W, const float* w_var,
                                   const float* bias, const int32_t* cols, const int32_t* col_var, int32_t n_cols,
                                   void* out_flat_bf16, int32_t Hp, int32_t Wp, yad_stream_t stream);

/* CUDA-core implicit GEMM (fp32 accumulate; in/out dtype f32 or bf16).
 * weight [kh][kw][Cin][Cout_pad] in `dtype`; bias [Cout] f32; residual/out NHWC. */
int yad_conv_simt(const yad_conv_desc* d, int32_t dtype, const void* in, const void* weight,
                  int32_t ld_w, const float* bias, const void* residual, void* out, yad_stream_t stream);

/* tcgen05 / TMEM implicit GEMM, bf16 in, fp32 accumulate in tensor memory, TMA-fed.
 * Requirements: Cin % 64 == 0 (pad with zero channels), Cout_pad % 16 == 0, Cout_pad <= 256 per
 * call tile (larger Cout is tiled over N by the grid), stride 1 or 2.
 * weight [Cout_pad][kh*kw*Cin] bf16 (K-major, K ordered (kh, kw, cin)).
 * out_dtype: YAD_BF16 or YAD_F32.  */
int yad_conv_tc(const yad_conv_desc* d, const void* in, const void* weight, int32_t cout_pad,
                const float* bias, const void* residual, void* out, int32_t out_dtype,
                float* out2_f32 /* optional fp32 copy of the output, pitch ld_out2, may be NULL */,
                int32_t ld_out2, yad_stream_t stream);

/* yad_conv_tc plus the BasicBlock's downsample branch in the same launch (torchvision resnet.py:92-100 via
 * modules/_backbone.py:148: `identity = self.downsample(x)` next to `out = relu(bn1(conv1(x)))`, both strided convolutions of
 * the same x): the 1x1 stride-(sh, sw) pad-0 convolution reads exactly the input pixels of the main convolution's centre tap,
 * so it runs as Cin / 64 extra K-blocks into a second TMEM accumulator and is written by the same epilogue.
 *   weight_ds [cout_pad][Cin] bf16, bias_ds [cout_pad], act_ds its activation (YAD_ACT_NONE for the reference's downsample),
 *   out_ds: bf16, same pixel strides / channel pitch (d->ld_out) as `out`.  Requires Cout == cout_pad, a multiple of 32, co_off 0.
 * Both outputs are bitwise what two yad_conv_tc calls produce. */
int yad_conv_tc_dual(const yad_conv_desc* d, const void* in, const void* weight, int32_t cout_pad, const float* bias, void* out,
                     const void* weight_ds, const float* bias_ds, int32_t act_ds, void* out_ds, yad_stream_t stream);

/* Patch-resident tcgen05 implicit GEMM for stride-1 'same' convolutions on the FLAT halo-padded layout:
 * pixel (b,h,w) of an activation lives at flat index f = (b*Wp + w)*Hp + h (h fastest), channel pitch ld; cells with
 * h >= H or w >= W are the (shared) zero halo: they are read as padding and only ever written with zeros, so the caller
 * zero-fills the buffer once.  Hp - H / Wp - W must cover the filter reach (taps that can only see padding, e.g. the top and
 * bottom rows of a 3x3 at H = 1, are skipped and need no halo).  Every tap is then a constant shift in f, so a CTA
 * loads each input pixel once per tile instead of once per tap.  in / out / residual share the geometry.
 * Requirements: Cin % 64 == 0, Cout % 32 == 0, cout_pad % 64 == 0 (weight rows / bias zero padded), bf16 in and out.
 * weight [cout_pad][kh*kw*Cin] bf16, K ordered (kh, kw, cin) as for yad_conv_tc.  flags: reserved, 0. */
typedef struct {
  int32_t B, H, W;            /* image size */
  int32_t Hp, Wp;             /* padded pitches */
  int32_t Cin, ld_in;
  int32_t Cout, ld_out, co_off;
  int32_t kh, kw, ph, pw;     /* 'same' convolution: kh = 2*ph + 1, kw = 2*pw + 1 */
  int32_t act;                /* YAD_ACT_* */
  int32_t ld_res;             /* residual channel pitch (0 = no residual) */
  int32_t Hp_out, Wp_out;     /* pitches of out / residual (0, 0 = same as the input) */
} yad_flat_desc;
int yad_conv_flat(const yad_flat_desc* d, const void* in, const void* weight, int32_t cout_pad, const float* bias,
                  const void* residual, void* out, int32_t flags, yad_stream_t stream);
/* yad_conv_flat_taps over TWO input tensors of the same flat geometry: step i reads 64-channel chunk chunk[i] of input src[i]
 * (0: `in`, Cin / ld_in of the descriptor; 1: `in2`, cin2 / ld_in2).  Used to fold the BasicBlock's downsample branch into the
 * block's second convolution (torchvision resnet.py:96-103 via modules/_backbone.py:148: out = relu(bn2(conv2(t)) + downsample(x))):
 * both are linear in their inputs and meet before the activation, so downsample(x) is Cin(x) / 64 more K steps of the same GEMM
 * (weights concatenated along K, biases added by the caller) - no separate launch, no identity tensor written and re-read. */
int yad_conv_flat_taps2(const yad_flat_desc* d, int32_t n_steps, const int32_t* src, const int32_t* chunk, const int32_t* dh,
                        const int32_t* dw, const int32_t* wk, int64_t k_total, const void* in, const void* in2, int32_t cin2,
                        int32_t ld_in2, const void* weight, int32_t cout_pad, const float* bias, const void* residual, void* out,
                        yad_stream_t stream);

/* Debug aid of the yad_conv_flat* family: dev_buf (16 int64 per CTA, one CTA per SM; device) receives on the following
 * launches the CTA's coarse timeline - [8] entry, [9] set-up done, [10] / [11] MMA loop begin / end, [12] epilogue done, [13] exit
 * (clock64), [14] / [15] globaltimer at entry / exit - and, with YAD_FLAT_DBG bit 4, the MMA-issuing warp's cycle budget - [0] total,
 * [1] waiting for an accumulator stage, [2] for a patch, [3] for a weight block, [4] inside the issue blocks
 * (tools/bench_conv.py --timeline); NULL switches it off (the default). */
int yad_conv_flat_set_timeline(void* dev_buf);

/* yad_conv_flat that also writes a SPACE-TO-DEPTH copy of its output for the stride-2 block that follows (torchvision
 * resnet.py:92-100: the first BasicBlock of layer2..4 reads its input with stride 2 twice, conv1 and downsample):
 *   out_s2d [B, Wp2, Hp2, 4 * Cout] bf16, pixel (b, h, w) -> cell (b, w / 2, h / 2), channel plane (h & 1) * 2 + (w & 1).
 * On that copy a stride-2 convolution is a stride-1 flat convolution over (plane, shift) steps (yad_conv_flat_taps), i.e. it
 * runs in the patch-resident kernel and reads every input element once.  Zero-filled by the caller once (halo cells). */
int yad_conv_flat_s2d(const yad_flat_desc* d, const void* in, const void* weight, int32_t cout_pad, const float* bias,
                      const void* residual, void* out, void* out_s2d, int32_t Hp2, int32_t Wp2, yad_stream_t stream);

/* The same kernel with an explicit step list (kh/kw/ph/pw of d are ignored): step i reads the input channels
 * [64*chunk[i], 64*chunk[i]+64) at the flat shift dw[i]*Hp + dh[i] and multiplies them with the [cout_pad x 64] block
 * at K offset wk[i] of weight [cout_pad][k_total].  Steps must be grouped by ascending chunk; at most 64 steps.
 * This is how the stem's 7x7 stride-2 conv2 (modules/_backbone.py:144) runs as a stride-1 conv: its input is the
 * space-to-depth output of yad_conv_stem_tc (4 parity planes x 64 channels), tap (kh, kw) = plane ((kh-3)&1, (kw-3)&1)
 * shifted by (floor((kh-3)/2), floor((kw-3)/2)). */
int yad_conv_flat_taps(const yad_flat_desc* d, int32_t n_steps, const int32_t* chunk, const int32_t* dh, const int32_t* dw,
                       const int32_t* wk, int64_t k_total, const void* in, const void* weight, int32_t cout_pad,
                       const float* bias, const void* residual, void* out, yad_stream_t stream);

/* ------------------------------------------------------------------ the whole neck as one kernel (bf16, H = 1, deploy form)
 * Replaces MultiScaleFmapModule.forward, modules/_common.py:241-265, for the default ResNet backbone: H-means (:248-252),
 * CSPSPPF (:204-215), both BiC blocks (:179-185), the four RepBlocks (:148-158, re-parameterised), the two stride-(1,2)
 * downsample convs (:238-239) and the [B, G, 15] permute (:259-264).  One persistent CTA per SM runs a host-compiled program
 * (yad_b200/neck_fused.py) clip by clip with every intermediate activation in shared memory (csrc/neck_fused.cu).
 *   fmaps[4]            the backbone maps layer1..layer4 in the flat layout of yad_conv_flat ([B, Wp, Hp, C] bf16)
 *   fmap_k[i]           Hp * C of map i (channels of one (b, w) column incl. the zero halo row: the H-mean is folded into K)
 *   fmap_rows_per_clip  Wp of map i
 *   fmap_box_rows       rows of one TMA box of map i (multiple of 8, <= 128; chosen by the program: see neck_fused.py)
 *   clips_per_unit      clips one CTA pass processes together (the program's G: 1 or 2)
 *   wblob [wrows, 64]   bf16 weight blocks ([N x 64] per K block, stacked), bias [n_bias] f32
 *   ops / kbs           the program (n_ops x 24 int32, n_kb x 2 int32; layouts in csrc/neck_fused.cu), device memory
 *   pool_bytes, n_slots shared-memory plan: activation pool size and ring depth (16 KB slots, 3..8)
 *   heads[3]            fp32 outputs [B, head_W[i], head_ld] (sm, md, lg), first 3 * (3 + nc) channels valid
 *   dbg                 optional bf16 buffer for the program's DUMP ops (NULL in production) */
int yad_neck_fused(const void* const* fmaps, const int32_t* fmap_k, const int32_t* fmap_rows_per_clip,
                   const int32_t* fmap_box_rows, int64_t B, int32_t clips_per_unit,
                   const void* wblob, int64_t wrows, const float* bias, int32_t n_bias, const void* ops, int32_t n_ops,
                   const void* kbs, int32_t n_kb, int32_t pool_bytes, int32_t n_slots, float* const* heads,
                   const int32_t* head_W, int32_t head_ld, void* dbg, yad_stream_t stream);

/* ------------------------------------------------------------------ file-rate resampling in front of the network
 * Replaces the torchaudio.transforms.Resample(orig_freq = file rate, new_freq = model sample_rate) of inference.py:152-159
 * (default Hann-windowed sinc bank, [ta] functional.py:1305-1431).  x [B, L] f32 or int16 PCM (x / 32768), kernel [P][2 * width + O]
 * f32 (P / O = new / orig rate over their gcd; built by yad_b200.frontend_consts.sinc_resample_kernel), out [B, Lout] f32 with
 * Lout <= ceil(P * L / O). */
int yad_resample_sinc(const void* x, int32_t x_is_i16, int64_t B, int64_t L, int32_t O, int32_t P, int32_t width,
                      const float* kernel, float* out, int64_t Lout, yad_stream_t stream);

/* Debug aid of yad_neck_fused: dev_buf (8 * n_ops int64, device, zeroed by the caller) receives CTA 0's per-op clock64 stamps of
 * one of its clips on the following launches (tools/neck_timeline.py); NULL switches it off (the default).
 * yad_neck_fused_set_timeline_iter: which clip of CTA 0 (0 = its first: cold; 1 = its second: steady state, maps prefetched). */
int yad_neck_fused_set_timeline(void* dev_buf);
int yad_neck_fused_set_timeline_iter(int32_t iter);

/* ------------------------------------------------------------------ neck glue (NHWC, dtype f32|bf16)
 * adaptive_avg_pool2d(H->1) modules/_common.py:248-252; F.interpolate bilinear x2 / x0.5
 * :173-174,181-182; cascaded max_pool2d k5 s1 p2 :207-209.  All write into a channel slice
 * (co_off, ld_out) of the destination so torch.cat never materialises. */
/* yad_hmean input pixel strides (in pixels): element (b,h,w,c) at (b*in_sb + h*in_sh + w*in_sw)*ld_in + c; pass
 * (W, 1, H*W) ... i.e. in_sh = W, in_sw = 1, in_sb = H*W for dense NHWC.  Output is dense [B, W, ld_out]. */
int yad_hmean(const void* in, int32_t dtype, int64_t B, int32_t H, int32_t W, int32_t C, int32_t ld_in,
              int32_t in_sw, int32_t in_sh, int32_t in_sb, void* out, int32_t ld_out, int32_t co_off, yad_stream_t stream);
int yad_resize_w(const void* in, int32_t dtype, int64_t B, int32_t W, int32_t C, int32_t ld_in, int32_t ci_off,
                 int32_t up /*1: x2, 0: x0.5*/, void* out, int32_t ld_out, int32_t co_off, yad_stream_t stream);
/* in: slice [ci_off, ci_off+C) of a [B,W,ld] tensor; writes pool5, pool5^2, pool5^3 of it to channel
 * slices co_off, co_off+C, co_off+2C of out. */
int yad_sppf_pools(const void* in, int32_t dtype, int64_t B, int32_t W, int32_t C, int32_t ld_in, int32_t ci_off,
                   void* out, int32_t ld_out, int32_t co_off, yad_stream_t stream);

/* Height half of the 2-D SPPF pools (modules/_common.py:197,207-209 when the neck keeps its height, i.e. backbone: custom):
 * out[b,h,w,c] = max over rows h-radius..h+radius (clipped) of the channel slice [ci_off, ci_off+C) of in [B,H,W,ld_in].
 * k cascaded 5x5 stride-1 max pools = a (4k+1)^2 box max = yad_maxpool_h(radius 2k) of the k-fold W cascade that
 * yad_sppf_pools (called with B*H rows) writes.  Out of place. */
int yad_maxpool_h(const void* in, int32_t dtype, int64_t B, int32_t H, int32_t W, int32_t C, int32_t ld_in, int32_t ci_off,
                  int32_t radius, void* out, int32_t ld_out, int32_t co_off, yad_stream_t stream);
/* x_spectral [B,C,H,W] f32 -> NHWC [B,H,W,ld_out] (out_dtype): input layout of CustomBackBone.first_conv
 * (modules/_backbone.py:97-101,109). */
int yad_nchw_to_nhwc(const float* in, int64_t B, int32_t C, int32_t H, int32_t W, void* out, int32_t out_dtype, int32_t ld_out,
                     yad_stream_t stream);

/* RepVGG train-form merge (modules/_common.py:90-95): out = act(a + b [+ scale*x + shift]); a, b are the
 * activated 3x3 / 1x1 branches [npix, ld_ab]; (scale, shift) the eval-mode identity BatchNorm (x may be NULL). */
int yad_repvgg_merge(const void* a, const void* b, const void* x, const float* scale, const float* shift,
                     int32_t dtype, int64_t npix, int32_t C, int32_t ld_ab, int32_t ld_x, void* out, int32_t ld_out,
                     int32_t co_off, int32_t act, yad_stream_t stream);

/* ------------------------------------------------------------------ anchor decode
 * Replaces get_scale_pred + combine (modules/_architecture.py:113-156).
 * head_s [B, G_s, ld_s] (dtype) with A*(3+nc) valid channels; anchors_s [A] f32 in seconds.
 * preds [B, sum_s G_s*A, 3+nc] f32, scale-major, row = g*A + a.
 * center = ((sigmoid(tc)*2 - 0.5) + g) * stride_s / center_scaler ; width = (sigmoid(tw)*2)^2 * anchor
 * both clipped to [0, duration]. */
int yad_decode(const void* const* heads, const int32_t* G, const int32_t* ld, const int32_t* stride,
               int32_t n_scales, int32_t dtype, const float* anchors /*[n_scales*A]*/, int32_t A, int32_t nc,
               float center_scaler, float duration, int64_t B, float* preds, yad_stream_t stream);

/* yad_decode with the anchors read from DEVICE memory ([n_scales*A] f32, seconds): train mode, where the anchors are
 * parameters that change every step (modules/_architecture.py:39-41) and the launch is replayed from a CUDA graph. */
int yad_decode_dev(const void* const* heads, const int32_t* G, const int32_t* ld, const int32_t* stride,
                   int32_t n_scales, int32_t dtype, const float* anchors_dev, int32_t A, int32_t nc,
                   float center_scaler, float duration, int64_t B, float* preds, yad_stream_t stream);

/* ------------------------------------------------------------------ segment NMS + post-processing
 * Replaces inference.py:42-110 (process_model_outputs) incl. torchvision.ops.batched_nms
 * ([tv] ops/boxes.py:48-120 -> torchvision::nms): class-agnostic per-clip greedy NMS on 1-D
 * segments dressed as boxes of height _h; fp32 IoU; suppress iff iou > (double)thr; stable
 * descending score order; NaN never suppresses.
 *   preds      [B, P, 3+nc] f32 (P <= 1024)
 *   keep       [B, P] i32   kept local indices in descending score order (first n_keep[b] valid); keep and n_keep may both
 *   n_keep     [B]    i32   be NULL when only seg_rows is wanted: the greedy scan then stops at the first surviving box
 *                           with conf <= conf_thr (lower-scored boxes cannot change the segments)
 *   conf/boxes [B, P] / [B, P, 2] f32  optional taps (may be NULL): confidence, clipped (x1,x2)
 *   seg_rows   [B, P, 5] f32  per clip: rows passing conf > conf_thr, sorted by centre ascending,
 *                             = [conf, obj_logit, label, start, end] (or centre,width if !start_end)
 *   n_seg      [B] i32
 */
int yad_nms(const float* preds, int64_t B, int32_t P, int32_t nc, double iou_thr, float conf_thr,
            float duration, float box_h, int32_t return_start_end, int32_t* keep, int32_t* n_keep,
            float* conf, float* boxes, float* seg_rows, int32_t* n_seg, yad_stream_t stream);

/* Compacts seg_rows/n_seg into the reference's return layout:
 * segments [K,5] f32, batch_idxs [K] i64, total[0] = K (i64, device).  */
int yad_compact_segments(const float* seg_rows, const int32_t* n_seg, int64_t B, int32_t P,
                         float* segments, int64_t* batch_idxs, int64_t* total, yad_stream_t stream);

/* ------------------------------------------------------------------ training-side (config 5)
 * Anchor matching: dataset.py:286-365 (build_target_by_scale).  targets [T,4] f32 =
 * (batch_idx, cls, centre_s, dur_s).  Outputs have capacity 3*A*T rows; rows are in the
 * reference order (base matches anchor-major, then left-neighbour copies, then right).
 * n_out[0] = M. */
int yad_build_targets(const float* targets, int32_t T, const float* anchors, int32_t A, int32_t G,
                      float anchor_t, float duration, float edge_t, int64_t* batch_idx, int64_t* grid_idx,
                      int64_t* anchor_idx, int64_t* classes, float* cw, int32_t* n_out, yad_stream_t stream);

/* Batch builder: the device half of AudioDataset.__getitem__ / collate_fn (dataset.py:132-155,276-283).
 * packed: B ragged clips back to back, clip b = n_channels[b] x n_samples[b] elements (channel-major) starting at element
 * offset[b]; fp32 or 16-bit PCM (dtype_i16 = 1: x / 32768).  out [B, 1, L] f32 = channel mean, zero padded to L samples. */
int yad_collate_clips(const void* packed, int32_t dtype_i16, const int64_t* offset, const int32_t* n_samples,
                      const int32_t* n_channels, int64_t B, int64_t L, float* out, yad_stream_t stream);

/* Detection loss of ONE scale: modules/_loss.py:115-228 (AudioDetectionLoss.loss_fn + compute_ciou) for the default
 * train_config (multi_label BCE class loss with label smoothing, BCEWithLogits objectness, no focal loss).
 *   pred [B, G, A, 3+nc] f32 (decoded predictions: obj, cls.., centre_s, width_s); matches (bi, gi, ai, cl i64, cw [M,2] f32)
 *   as produced by yad_build_targets, M known to the host.
 * Outputs: grad [B, G, A, 3+nc] f32 (overwritten) = d(box_w * box + conf_scale * conf + class_w * cls) / d pred, where
 * conf_scale = conf_w * the scale's weight (4 / 2 / 1); acc [16] f64 (9 used) sums: {sum(1-ciou), sum ciou, sum BCE objectness over all
 * cells, sum BCE class, sum sigmoid(obj) over matches, sum sigmoid(obj) over cells with t_conf == 0, count of those cells,
 * count of matches whose class != ignore_index, sum of class_weights[target] over those matches}; confusion [nc][nc] i32 (target class x argmax class) for the
 * accuracy / precision / recall / f1 metrics.  Duplicate (b,g,a) matches: the last one owns t_conf (index_put_ order, Q12),
 * all of them receive box / class gradients.  Workspaces: owner_ws [B*G*A] i32, ciou_ws [M] f32. */
int yad_loss_scale(const float* pred, int64_t B, int32_t G, int32_t A, int32_t nc, const int64_t* bi, const int64_t* gi,
                   const int64_t* ai, const int64_t* cl, const float* cw, int32_t M, float box_w, float conf_scale,
                   float class_w, float label_smoothing, int64_t ignore_index, int32_t* owner_ws, float* ciou_ws,
                   int32_t* confusion, double* acc, float* grad, yad_stream_t stream);

/* The other two branches of the reference constructor (modules/_loss.py:74-81):
 *   cls_mode 1: class loss = nn.CrossEntropyLoss(weight=class_weights, ignore_index) over the class-valid matches (multi_label:
 *               false, :157-158); class_weights [nc] f32 on the device or NULL; acc[3] then holds sum_m w[c_m] (logsumexp - x_c)
 *               and acc[8] the divisor sum_m w[c_m];
 *   focal_gamma > 0: objectness = FocalLoss(alpha, gamma, with_logits=True) = mean(alpha (1 - exp(-bce))^gamma bce) (:9-37). */
int yad_loss_scale_ex(const float* pred, int64_t B, int32_t G, int32_t A, int32_t nc, const int64_t* bi, const int64_t* gi,
                      const int64_t* ai, const int64_t* cl, const float* cw, int32_t M, float box_w, float conf_scale,
                      float class_w, float label_smoothing, int64_t ignore_index, int32_t cls_mode, const float* class_weights,
                      float focal_alpha, float focal_gamma, int32_t* owner_ws, float* ciou_ws, int32_t* confusion, double* acc,
                      float* grad, yad_stream_t stream);

/* ------------------------------------------------------------------ train-mode network (fp32, NHWC rows x channels)
 * What autograd does for TrainerPipeline.__feed (pipeline/_trainer.py:94-108): the backward of F.conv2d, BatchNorm2d in
 * training mode, the activations, the neck glue and the anchor decode.  Correctness-first CUDA-core kernels. */

/* Data gradient of the convolution described by d (the FORWARD descriptor): dX [B,H,W,ld_in] = dx_in (may be NULL) +
 * sum_taps dY[B,Ho,Wo,ld_out slice co_off] * W.  weight_t: [kh][kw][Cout][Cin] f32. */
int yad_conv_dgrad(const yad_conv_desc* d, const float* dy, const float* weight_t, const float* dx_in, float* dx,
                   yad_stream_t stream);
/* Weight gradient: dw [kh][kw][Cin][Cout] f32 += X^T dY (fp32 atomics); dbias_ws [Cout] f64 += column sums of dY (may be NULL). */
int yad_conv_wgrad(const yad_conv_desc* d, const float* x, const float* dy, float* dw, double* dbias_ws, yad_stream_t stream);
/* BatchNorm2d, training mode, over N rows x C channels: batch mean / biased variance (fp64 sums), y = act((x - mean) * invstd
 * * gamma + beta); running statistics updated in place with `momentum` and the unbiased variance (may be NULL).
 * ws: [2*C] f64 scratch. */
int yad_bn_train_fwd(const float* x, int32_t ld_x, int64_t N, int32_t C, const float* gamma, const float* beta, float eps,
                     float momentum, float* running_mean, float* running_var, int32_t act, float* y, int32_t ld_y,
                     float* save_mean, float* save_invstd, double* ws, yad_stream_t stream);
/* The same from pre-computed moments: sums [2*C] f64 = (sum x, sum x^2) over the N rows (yad_corr_tf32's `stats`): one launch
 * that finalises mean / invstd, updates the running statistics and applies. */
int yad_bn_train_apply(const float* x, int32_t ld_x, int64_t N, int32_t C, const float* gamma, const float* beta, float eps,
                       float momentum, float* running_mean, float* running_var, int32_t act, float* y, int32_t ld_y,
                       float* save_mean, float* save_invstd, const double* sums, yad_stream_t stream);
/* Its backward, the activation's included: g = dy * act'(y); dx (overwritten, or += when accumulate) = gamma * invstd * (g - mean(g) - xhat * mean(g xhat));
 * dgamma += sum g xhat; dbeta += sum g. */
int yad_bn_train_bwd(const float* x, int32_t ld_x, const float* y, int32_t ld_y, const float* dy, int32_t ld_dy, int64_t N, int32_t C,
                     const float* gamma, const float* save_mean, const float* save_invstd, int32_t act, float* dx, int32_t ld_dx,
                     int32_t accumulate, float* dgamma, float* dbeta, double* ws, yad_stream_t stream);
/* y = act(a + b [+ c]) and its backward (g = dy * act'(y) ACCUMULATED into da, db, dc; any of them may be NULL). */
int yad_add_act(const float* a, int32_t ld_a, const float* b, int32_t ld_b, const float* c, int32_t ld_c, int64_t N, int32_t C,
                int32_t act, float* y, int32_t ld_y, yad_stream_t stream);
int yad_add_act_bwd(const float* y, int32_t ld_y, const float* dy, int32_t ld_dy, int64_t N, int32_t C, int32_t act, float* da,
                    int32_t ld_a, float* db, int32_t ld_b, float* dc, int32_t ld_c, yad_stream_t stream);
/* Dropout with a counter-based mask: y = (accumulate ? y : 0) + x * keep(seed, i) / (1 - p); calling it on dy gives the backward. */
int yad_dropout(const float* x, int64_t n, float p, uint64_t seed, int32_t accumulate, float* y, yad_stream_t stream);
/* The same with the effective seed = seed + *seed_dev (a device-resident step counter, so that a captured CUDA graph draws a new
 * mask on every replay). */
int yad_dropout_dev(const float* x, int64_t n, float p, uint64_t seed, const uint64_t* seed_dev, int32_t accumulate, float* y,
                    yad_stream_t stream);
/* Backward of yad_hmean / yad_resize_w (gradients ACCUMULATED into din) and MaxPool(5,1,2) along W forward / backward
 * (first maximum of the window receives the gradient, accumulated with atomics). */
int yad_hmean_bwd(const float* dout, int32_t ld_o, int64_t B, int32_t H, int32_t W, int32_t C, float* din, int32_t ld_i,
                  yad_stream_t stream);
int yad_resize_w_bwd(const float* dout, int32_t ld_o, int64_t B, int32_t W, int32_t C, int32_t up, float* din, int32_t ld_i,
                     yad_stream_t stream);
int yad_maxpool5_w(const float* x, int32_t ld_x, int64_t B, int32_t W, int32_t C, float* y, int32_t ld_y, yad_stream_t stream);
int yad_maxpool5_w_bwd(const float* x, int32_t ld_x, const float* dy, int32_t ld_dy, int64_t B, int32_t W, int32_t C, float* dx,
                       int32_t ld_dx, yad_stream_t stream);
/* Backward of yad_decode for one scale: head [B,G,ld_h] f32, dpred [B,G,A,3+nc]; dhead (overwritten, first A*(3+nc) channels),
 * danchor_s [A] += d loss / d anchor (in seconds). */
int yad_decode_bwd(const float* head, int32_t ld_h, const float* dpred, int64_t B, int32_t G, int32_t A, int32_t nc,
                   const float* anchors_s, float stride_over_scaler, float duration, float* dhead, int32_t ld_dh, float* danchor_s,
                   yad_stream_t stream);

/* out (dense, row-major over sizes[4]) (+)= in[sum_k i_k * in_strides[k]] (element strides): weight re-packing between the
 * OIHW master layout and the kernels' layouts, NCHW -> NHWC, gradient un-packing.  yad_add_f64_to_f32: out[i] += (float)a[i]. */
int yad_permute4(const float* in, const int64_t* in_strides, float* out, const int32_t* sizes, int32_t accumulate,
                 yad_stream_t stream);
int yad_add_f64_to_f32(const double* a, int32_t n, float* out, yad_stream_t stream);

/* ------------------------------------------------------------------ TF32 tensor-core convolutions of the train step
 * fp32 NHWC tensors, tf32 operands, fp32 accumulation in tensor memory (tcgen05.mma kind::tf32): what cuDNN runs for the
 * reference's F.conv2d forward / backward on a GPU by default (torch.backends.cudnn.allow_tf32).
 *
 * yad_corr_tf32: generic tap-list correlation
 *     out[b,i,j,n] (+)= bias[n] + sum_t sum_c in[b, i*sh + tap_dh[t], j*sw + tap_dw[t], c] * weight[n][tap_k[t] + c]
 * in: dense NHWC [B,H,W,ld_in] (Cin % 32 == 0, zero-padded channels); weight [cout_pad][k_total] f32 (K-major); out pixel
 * (b,i,j) at (b*out_sb + i*out_sh + j*out_sw)*ld_out (all three 0 = dense [B,Ho,Wo,ld_out]); bias may be NULL.  The host
 * builds from it the forward conv, the data gradient of a stride-1 conv (flipped taps, transposed weight) and of a
 * stride-2 conv (one call per output parity class).
 * yad_wgrad_tf32: dw[tap_dst[t]][ci][co] += sum_{b,i,j} x[b, i*sh + tap_dh[t], j*sw + tap_dw[t], ci] * dy[b,i,j,co]
 * x [B,H,W,ld_in], dy [B,Ho,Wo,ld_out] dense with ld % 32 == 0 (pad channels must hold finite values); dw [taps][Cin][Cout]. */
typedef struct {
  int32_t B, H, W;            /* input spatial */
  int32_t Cin, ld_in;
  int32_t Ho, Wo;             /* output spatial (explicit) */
  int32_t Cout, ld_out;
  int32_t sh, sw;             /* input step per output pixel: 1 or 2 */
  int32_t out_sw, out_sh, out_sb;
  int32_t n_taps;
  int32_t act;                /* YAD_ACT_* */
  int32_t accumulate;         /* out += instead of out = */
  int32_t whole_rows;         /* out is a dense buffer of which this call owns every channel of the pitch: lets a small
                                 problem zero-fill it and split the K loop over CTAs (fp32 red.add of partial sums) */
} yad_corr_desc;
/* stats (may be NULL): [2*Cout] f64, += per-channel sum and sum of squares of the values written (the batch moments of the
 * BatchNorm that follows the conv, fused into the epilogue; feed them to yad_bn_train_apply). */
int yad_corr_tf32(const yad_corr_desc* d, const int32_t* tap_dh, const int32_t* tap_dw, const int32_t* tap_k, const float* in,
                  const float* weight, int32_t cout_pad, int64_t k_total, const float* bias, float* out, double* stats,
                  yad_stream_t stream);
int yad_wgrad_tf32(const yad_corr_desc* d, const int32_t* tap_dh, const int32_t* tap_dw, const int32_t* tap_dst, const float* x,
                   const float* dy, float* dw, yad_stream_t stream);
/* im2col of the stem conv1 (C -> 64, 7x7, stride 2, pad 3): x NCHW f32 [B,C,H,W] -> patches [B,(H-1)/2+1,(W-1)/2+1,K] with
 * k = (kh*7 + kw)*C + c, zero for k >= 49*C (K a multiple of 32): conv1 and its weight gradient then run as a 1x1 conv on the
 * TF32 kernels. */
int yad_stem_im2col(const float* x_nchw, int64_t B, int32_t C, int32_t H, int32_t W, int32_t K, float* patches, yad_stream_t stream);
/* ws[c] += sum over N rows of x[r*ld + c] (fp64): bias gradients. */
int yad_colsum_f64(const float* x, int32_t ld, int64_t N, int32_t C, double* ws, yad_stream_t stream);

/* Fused Adam (L2 weight decay) + EMA over a flat fp32 parameter arena:
 * torch.optim.Adam (train.py:83-90, config.yaml:75-80) and smoothener/_ema.py:20-26.
 * ema may be NULL.  step >= 1. */
int yad_adam_ema_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* ema,
                      int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                      int32_t step, float ema_momentum, yad_stream_t stream);

/* Train step: every OIHW master weight -> the two K-major fp32 operands of the TF32 convolutions (forward / weight gradient
 * [ceil16(O)][kh*kw][cin_pad] and data gradient [ceil16(I)][kh*kw][coutk]) in ONE launch (the parameters change every step:
 * pipeline/_trainer.py:105).  table: DEVICE array of n_entries 56-byte records
 *   { const float* src; float* wf; float* wt; int32 O, I, KK, cin_pad, coutk, pad; int64 start }   (start = prefix sum of O*I*KK),
 * total_elems = sum of O*I*KK.  Pad rows / columns of wf, wt are not written (the caller zero-fills them once). */
int yad_pack_weights_tf32(const void* table, int32_t n_entries, int64_t total_elems, yad_stream_t stream);

/* ------------------------------------------------------------------ anchor clustering
 * Replaces the iterations of sklearn.cluster.KMeans(algorithm="lloyd") in compute_anchors.py:72-86 for 1-D data (the segment
 * durations), fp64, one CTA iterating to convergence on the device.
 *   x        [n]  f64   centred data (device)
 *   centers  [k]  f64   in: initial centres (k-means++ / random seeding is done by the host), out: final centres; k <= 16
 *   tol_abs             tol * var(x): stop when sum_j |c_j' - c_j|^2 <= tol_abs, or when no label changes
 *   labels   [n]  i32   out: label of every point for the final centres
 *   n_iter   [1]  i32,  inertia [1] f64   out (device)
 * E-step: argmin_j (c_j^2 - 2 x c_j), first minimum on ties; empty clusters are re-seeded with the farthest points. */
int yad_kmeans1d_lloyd(const double* x, int64_t n, double* centers, int32_t k, int32_t max_iter, double tol_abs,
                       int32_t* labels, int32_t* n_iter, double* inertia, yad_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif  /* YAD_B200_H_ */
