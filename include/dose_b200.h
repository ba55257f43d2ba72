/* dose_b200.h — C ABI of libdose_b200.so: the sm_100a kernels behind the OAR-TRANSEG -> DOSE-PYFER
 * hot path of GhTara/Dose_Prediction.
 *
 * The reference has no FFI / operator API of its own: every device op on the path is an ATen call made
 * from nn.Module.forward.  Each entry point below therefore names the ATen call
 * site(s) it replaces (paths relative to the reference root; "monai:" = monai==0.7.0).
 *
 * Conventions
 *   - plain C types only: device pointers as void* / float* / double*, sizes as int / long long,
 *     cudaStream_t for ordering.  No torch types.
 *   - every function returns 0 on success, non-zero on failure; dp_last_error() gives the message.
 *   - asynchronous on the given stream, never synchronises, never allocates device memory.
 *   - activation tensors use the "c8" layout [N][C/8][D][H][W][8] (fp16, optionally as a hi/lo pair of
 *     tensors whose sum carries ~22 mantissa bits; fp32 for pre-normalisation "raw" tensors);
 *     (cb_total, cb_off) address a channel-block slice of a larger buffer, which is how torch.cat
 *     (base_blocks.py:139, blocks_MDUNet.py:154, c3d.py:103-112, dose_pyfer.py:357) disappears.
 *   - per-(n,c) InstanceNorm statistics are double[N][C][2] = {sum, sum of squares}, accumulated by the
 *     producing kernel and finalised (biased variance, eps 1e-5) by the consumer.
 *   - err_flag: optional device int set to 1 if an in-kernel mbarrier wait times out (protocol bug).
 */
#ifndef DOSE_B200_H_
#define DOSE_B200_H_

#include <cuda_runtime.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* activation ids used by the *_act arguments */
#define DP_ACT_NONE 0
#define DP_ACT_RELU 1
#define DP_ACT_LRELU 2 /* LeakyReLU(0.01): monai UnetResBlock.lrelu */
#define DP_ACT_MISH 3  /* blocks_MDUNet.py:139,143,148 (act='mish') */
#define DP_ACT_GELU 4  /* erf GELU: monai MLPBlock.fn */

const char* dp_last_error(void);
int dp_abi_version(void);
int dp_device_sm_count(void);

/* Per-device library state (SURVEY 8b "opaque dp_handle*").  The entry points below act on the caller's CURRENT device and
 * stream (as the CUDA runtime does) and keep no global mutable state beyond what a handle stands for: per-device kernel
 * attributes, configured on first use on each device, and the process-wide tensor-map (CUtensorMap) cache, dropped when
 * the last handle is destroyed.  One handle per device / rank; NULL on an invalid device (see dp_last_error).        */
typedef struct dp_handle dp_handle;
dp_handle* dp_handle_create(int device);
int dp_handle_device(const dp_handle* h);
void dp_handle_destroy(dp_handle* h);

/* nn.Conv3d stride 1, odd k<=7, dilation dil, "same" padding, on tcgen05 tensor cores.
 * Replaces: blocks_MDUNet.py:68,71 (3^3), :102,105 (7^3), :166-187 (dilated), c3d.py:16,30 (stride-1
 * SingleConv/UpConv convs), monai UnetResBlock.conv1/conv2.  Fused epilogue: y = acc*scale[c]+shift[c]
 * (conv bias and/or eval-mode BatchNorm3d fold, blocks_MDUNet.py:103,106), optional ReLU, statistics.
 *   in_c8      input c8 fp16 buffer (cb_total_in channel blocks per image)
 *   chunk_cb   host array [n_chunks]: first channel block of every 16-channel K chunk (hi/lo operand
 *              splitting = listing hi blocks, lo blocks, hi blocks again against [Whi;Whi;Wlo])
 *   wpack      fp16 [k(kd)][n_chunks][k(kh)][k(kw)][2][cout][8]
 *   out_f32    c8 fp32 output or NULL;  out_hi/out_lo  c8 fp16 output (lo optional) or NULL
 *   tap_mask   optional host array [n_chunks] (k <= 3): bit (kd*k+kh)*k+kw set = tap present; lets the
 *              stride-2 convs of c3d.py:49-61 run as sparse 3^3 convs over a space-to-depth input
 *   fold       1: wpack holds 2*cout columns per tap, [W_hi | W_lo] for hi chunks and [W_hi | 0] for lo chunks; the
 *              epilogue adds column c and c + cout (the 3-term operand split with each hi chunk read once)
 *              2 (cout <= 64, dilation 1): depth-pair mode — a tile covers output planes (d, d + 1); wpack fp16
 *              [k + 1 virtual depth taps v][n_chunks][kh][kw][2][2*cout][8], rows [W[kd = v] | W[kd = v - 1]] (zero where
 *              the tap does not exist): every input plane is read once for both output planes with N = 2*cout columns */
int dp_conv3d_tc(const void* in_c8, int cb_total_in, const uint8_t* chunk_cb, int n_chunks, const void* wpack,
                 int N, int D, int H, int W, int cout, int k, int dil, const float* scale, const float* shift,
                 int relu, float* out_f32, void* out_hi, void* out_lo, int cb_total_out, int cb_out_off,
                 double* stats, int* err_flag, int max_ctas, const uint32_t* tap_mask, int fold, cudaStream_t stream);

/* Same convolution for C_out in {16,32}, k in {3,7}, dilation 1, with the k depth taps stacked into the MMA
 * N dimension (N = k*C_out) and a sliding ring of output-plane accumulators in TMEM (conv_stack.cu): the
 * shared-memory A operand is read once per k taps instead of once per tap.
 *   wpack_stack  fp16 [n_chunks][k(kh)][k(kw)][2][k*cout][8], row j*cout+co holding W[co, :, kd=k-1-j, kh, kw]
 *   seg_len      output planes per work item along D (0 = choose for load balance)
 *   tiles_per_cta  adjacent W tiles sharing one weight stream (0 = default)
 *   fold         1 (cout == 16): weights carry 2*cout rows per slot ([W_hi | W_lo] for hi chunks, [W_hi | 0] for
 *                lo chunks); the epilogue adds the two halves — the 3-term operand split in 2 MMAs per chunk
 *                2 (cout == 16, k == 3; chunks = x_hi blocks then x_lo blocks): split-half layout, wpack_stack fp16
 *                [4 rotations][n_chunks][kh][kw][2][W_hi rows: 4 slots x 16 | W_lo rows: 4 slots x 16][8] followed by
 *                [4 rotations][2][128][8] = tap (0,0) of chunk 0 with the newest plane's rows zeroed too; x_lo chunks
 *                issue N = 64 (the W_hi rows) instead of N = 128 against a zero half                               */
int dp_conv3d_stack(const void* in_c8, int cb_total_in, const uint8_t* chunk_cb, int n_chunks, const void* wpack_stack,
                    int N, int D, int H, int W, int cout, int k, const float* scale, const float* shift, int relu,
                    float* out_f32, void* out_hi, void* out_lo, int cb_total_out, int cb_out_off, double* stats,
                    int* err_flag, int seg_len, int tiles_per_cta, int fold, cudaStream_t stream);

/* Generic direct convolution (any stride): c3d.py:49,53,57,61 (stride-2 SingleConv convs).
 *   w_packed   fp32 [k^3 taps][cin][cout]                                                              */
int dp_conv3d_direct(const void* in_hi, const void* in_lo, int in_cb_total, int in_cb_off, int cin, int N, int D,
                     int H, int W, int k, int stride, int dil, const float* w_packed, const float* scale,
                     const float* shift, int relu, int cout, float* out_raw, void* out_hi, void* out_lo,
                     int out_cb_total, int out_cb_off, double* stats_out, cudaStream_t stream);

/* nn.Linear / attention contractions: C[M,N] = alpha * A[M,K] . B[N,K]^T (+bias, +rowvec, act, +resid).
 * Replaces: monai PatchEmbeddingBlock Linear (+position_embeddings), SABlock.qkv / out_proj and both
 * einsums, MLPBlock.linear1 (+GELU) / linear2 (+residual).  mode_qkv scatters the [M, 3*heads*hd] result
 * into q [B,heads,T,hd] (scaled by q_scale), k [B,heads,T,hd], v^T [B,heads,hd,vt_ld] (vt_ld >= T, 8 | vt_ld)
 * (einops "b h (qkv l d) -> qkv b l h d").  Batched: A/B rows advance by a/b_batch_rows per batch entry z,
 * outputs by z*c_batch_stride, or (z/period)*c_batch_stride + (z%period)*c_batch_stride2 if period>0.
 * split_k > 1: plain fp32 partial sums go to out_f32[split][M][ldc]; finish with dp_splitk_reduce.      */
int dp_gemm_tc(const void* A, const void* B, int M, int N, int K, int batch, int a_batch_rows, int b_batch_rows,
               long long c_batch_stride, int c_batch_period, long long c_batch_stride2, int ldc, int split_k, const float* bias, const float* rowvec,
               int row_period, const float* resid, float alpha, int act, float* out_f32, int atomic,
               void* out_f16, int mode_qkv, int heads, int hd, int T, int vt_ld, void* q, void* k, void* vt, float q_scale,
               int* err_flag, cudaStream_t stream);

/* Fused attention of monai SABlock.forward for one ViT layer: out[b*T + t][h*hd + d] = softmax_y(q . k^T)[t][y] v[y][d]
 * (einsum "blxd,blyd->blxy" * scale, softmax(-1), einsum "bhxy,bhyd->bhxd", rearrange "b h l d -> b l (h d)"), with
 * q (pre-scaled), k [batch*heads][T][hd] and v^T [batch*heads][hd][vt_ld] as written by dp_gemm_tc(mode_qkv).
 * Scores and probabilities stay in TMEM / shared memory.  hd in {64, 128}; other head sizes use
 * dp_gemm_tc + dp_softmax + dp_gemm_tc.                                                                     */
int dp_attention(const void* q, const void* k, const void* vt, int batch, int heads, int T, int vt_ld, int hd,
                 void* out, int ld_out, int* err_flag, cudaStream_t stream);

/* Deterministic split-K finish: out[m][n] = sum_s ws[s][m][n] + bias[n] + rowvec[m % row_period][n]
 * (dp_gemm_tc with split_k > 1 writes the fp32 partials ws[split_k][M][N]; no atomics anywhere). */
int dp_splitk_reduce(const float* ws, int splits, int M, int N, const float* bias, const float* rowvec, int row_period,
                     float* out, cudaStream_t stream);

/* NCDHW fp32 <-> c8 fp16 (module-boundary conversion; batch['Input'].float(), train_light_pyfer.py:124) */
int dp_pack_ncdhw(const float* src, int N, int C, long long vox, void* hi, void* lo, int cb_total, int cb_off,
                  cudaStream_t stream);
int dp_unpack_c8(const void* hi, const void* lo, int cb_total, int cb_off, int N, int C, long long vox, float* dst,
                 cudaStream_t stream);

/* nn.InstanceNorm3d (+affine: c3d.py:17,31) -> activation -> optional residual add (+activation), with
 * optional statistics of the result for a chained InstanceNorm (blocks_MDUNet.py:136-139).
 * Residual = c8 fp16 tensor, or raw fp32 tensor normalised with res_stats (monai UnetResBlock.norm3).
 * s2d_*: optional second copy of the result in space-to-depth layout (channel block parity*C/8 + cb of a
 * half-resolution tensor) that lets the stride-2 convs (c3d.py:49-61) run on the tensor-core conv kernel. */
int dp_norm_act(const float* raw_f32, const void* raw_hi, const void* raw_lo, int in_cb_total, int in_cb_off,
                const double* stats, const float* gamma, const float* beta, int act, const void* res_hi,
                const void* res_lo, const float* res_raw, const double* res_stats, int res_cb_total, int res_cb_off,
                int act_after_res, void* out_hi, void* out_lo, int out_cb_total, int out_cb_off, double* stats_out,
                int N, int C, long long vox, void* s2d_hi, void* s2d_lo, int s2d_cb_total, int s2d_cb_off, int D, int H,
                int W, cudaStream_t stream);

/* 3x3x3 convolution (pad 1, stride 1) of a ONE-channel planar fp32 volume to 16 channels, exact fp32 on the CUDA cores
 * (csrc/conv_small.cu): monai UnetResBlock.conv1 of the seg net's encoder1 (oar_transeg.py:92-100, in_channels = 1).
 * w_host / bias_host are HOST arrays ([16][1][3][3][3], [16] or NULL) passed as kernel parameters.  Output: raw c8 fp32
 * (+ statistics [N][16][2]); xstats [N][2] receives {sum x, sum x^2} of the input per image (zero it first).     */
int dp_conv3d_c1(const float* x_planar, const float* w_host, const float* bias_host, int N, int D, int H, int W,
                 float* out_raw, int out_cb_total, double* stats, double* xstats, cudaStream_t stream);

/* dp_norm_act for the tail of that res block: out = act(IN(raw) + norm3(conv3(x))) where conv3 is the 1x1x1 conv
 * 1 -> C of the same one-channel input x (res_w [C]) — its InstanceNorm is the affine map
 * w_c (x - mean_x) / sqrt(w_c^2 var_x + eps) of x, evaluated from res_x (planar fp32) and res_xstats [N][2].   */
int dp_norm_act_resx(const float* raw_f32, int in_cb_total, const double* stats, const float* res_x, const float* res_w,
                     const double* res_xstats, int act_after_res, void* out_hi, void* out_lo, int out_cb_total,
                     int out_cb_off, int N, int C, long long vox, cudaStream_t stream);

/* dp_norm_act (raw fp32 in, no residual) fused with the 1x1x1 head that consumes its result: the dose heads
 * (dose_pyfer.py:290-300,316-317), conv_out_A (:353,359) and the seg logits (base_blocks.py:151-165).
 * head_w fp32 [head_co][C], head_b [head_co] or NULL, head_out NCDHW fp32 [N][head_co][vox]; head_co <= 8, C <= 128.
 * The head contracts the values exactly as stored (fp16, or the fp16 hi+lo pair when out_lo is given).    */
int dp_norm_act_head(const float* raw_f32, int in_cb_total, const double* stats, const float* gamma, const float* beta,
                     int act, void* out_hi, void* out_lo, int out_cb_total, int out_cb_off, const float* head_w,
                     const float* head_b, int head_co, float* head_out, int N, int C, long long vox, cudaStream_t stream);

/* nn.Conv3d k=1 over the channel-concatenation of up to three sources, each normalised/activated on
 * load: blocks_MDUNet.py:145-157 (cat(x3,x7) -> 1^3), monai UnetResBlock.conv3, heads dose_pyfer.py:290-300,353,
 * base_blocks.py:151 (seg logits).  Output: raw c8 fp32 (+stats), c8 fp16, or NCDHW fp32 (out_planar).
 *   w          fp32 [cout][sum of source C]                                                            */
int dp_pointwise_conv(int nsrc, const void* const* src_hi, const void* const* src_lo, const float* const* src_raw,
                      const int* src_cb_total, const int* src_cb_off, const int* src_C,
                      const double* const* src_stats, const int* src_act, const float* w, const float* bias, int cout,
                      int N, long long vox, float* out_raw, void* out_hi, void* out_lo, int out_cb_total,
                      int out_cb_off, float* out_planar, double* stats_out, int out_act, cudaStream_t stream);

/* The same 1^3 convolution for weights that are known on the HOST at launch time (inference plans): w_host
 * [cout][sum of source C] and bias_host [cout] (or NULL) are host pointers; 16 output channels' weights per launch
 * travel as kernel parameters and are consumed from the constant bank as FFMA operands (no weight loads: the
 * shared-memory return path was what bounded dp_pointwise_conv).  Padded input channel blocks (8 channels each, per
 * source) must number 1, 2, 3, 4, 6 or 8.                                                                     */
/* The same 1x1x1 convolution on tcgen05 tensor cores (csrc/pointwise_tc.cu): the staging threads apply each source's
 * pending InstanceNorm + activation, split the fp32 result into an fp16 hi+lo pair and write it straight into the UMMA
 * operand layout; 3-term operand split (A_hi.[W_hi|W_lo] + A_lo.W_hi), fp32 accumulation in TMEM.
 *   wpack  fp16 [K/8][2*cout][8]: rows 0..cout-1 = fp16(W), rows cout..2cout-1 = fp16(W - fp16(W)); K = every source's
 *          channels padded to whole 8-channel blocks, then to a multiple of 16; cout in {16, 32, 64}, K <= 128.
 *   src_stats0 / src_act0 (arrays or NULL; entries may be NULL): an optional FIRST normalisation stage per source,
 *          x -> act0(IN(x; stats0)) before IN(.; stats) + act — conv_block_3's closing "IN, ReLU" (blocks_MDUNet.py:64-78)
 *          followed by conv_3_1's IN + act (:150-157), both applied straight from the raw conv output.
 * Output: raw c8 fp32 (+ statistics) or c8 fp16 hi[/lo].                                                      */
int dp_pointwise_tc(int nsrc, const void* const* src_hi, const void* const* src_lo, const float* const* src_raw,
                    const int* src_cb_total, const int* src_cb_off, const int* src_C, const double* const* src_stats,
                    const int* src_act, const double* const* src_stats0, const int* src_act0, const void* wpack,
                    const float* bias, int cout, int N, long long vox,
                    float* out_raw, void* out_hi, void* out_lo, int out_cb_total, int out_cb_off, double* stats_out,
                    int* err_flag, cudaStream_t stream);

int dp_pointwise_conv_cw(int nsrc, const void* const* src_hi, const void* const* src_lo, const float* const* src_raw,
                         const int* src_cb_total, const int* src_cb_off, const int* src_C,
                         const double* const* src_stats, const int* src_act, const float* w_host, const float* bias_host,
                         int cout, int N, long long vox, float* out_raw, void* out_hi, void* out_lo, int out_cb_total,
                         int out_cb_off, float* out_planar, double* stats_out, int out_act, cudaStream_t stream);

/* nn.ConvTranspose3d k=2 s=2 no bias (monai get_conv_layer(is_transposed=True): base_blocks.py:118-127,
 * UnetrPrUpBlock).  Input addressed by element strides so ViT tokens [B,T,C] are read in place
 * (proj_feat, dose_pyfer.py:118-122, becomes a no-op).  w_packed fp32 [8 parity][cin][cout].          */
int dp_deconv2x(const void* in_hi, const void* in_lo, long long in_nstride, long long in_vstride,
                long long in_cbstride, int cin, int cout, int N, int D, int H, int W, const float* w_packed,
                void* out_hi, void* out_lo, int out_cb_total, int out_cb_off, cudaStream_t stream);

/* dp_deconv2x for a c8 input and weights known on the HOST (w_host_packed fp32 [8 parity][cin][cout], host pointer):
 * weights travel as kernel parameters (constant bank), the voxel's input vector stays in registers for all
 * parities.  C_in in {32, 64}.                                                                                  */
int dp_deconv2x_cw(const void* in_hi, const void* in_lo, long long in_nstride, long long in_cbstride, int cin, int cout,
                   int N, int D, int H, int W, const float* w_host_packed, void* out_hi, void* out_lo, int out_cb_total,
                   int out_cb_off, cudaStream_t stream);

/* The same transposed convolution (monai get_conv_layer(is_transposed=True, kernel 2, stride 2), base_blocks.py:118-127;
 * UnetrPrUpBlock / UnetrUpBlock transp_conv) for a c8 input on the tensor cores: a 1^3 implicit GEMM of dp_conv3d_tc
 * with nq * cout columns (parities q0 .. q0+nq-1; column (q - q0) * cout + co = W[:, co, q>>2, (q>>1)&1, q&1]) whose
 * epilogue scatters every 8-channel group to its output voxel.  wpack / chunk_cb as for dp_conv3d_tc with k = 1
 * (operand-split chunks included); nq * cout <= 256, 16 | cout, q0 and nq even (the W-neighbour parities q, q+1 are
 * written together as full 32-byte sectors).                                                                     */
int dp_deconv2x_tc(const void* in_c8, int cb_total_in, const uint8_t* chunk_cb, int n_chunks, const void* wpack, int N,
                   int D, int H, int W, int cout, int q0, int nq, const float* scale, const float* shift, void* out_hi,
                   void* out_lo, int cb_total_out, int cb_out_off, int* err_flag, cudaStream_t stream);

/* Same transposed convolution as a tcgen05 GEMM [B*Dg*Hg*Wg, cin] x [cin, 8*cout] with a pixel-shuffle
 * scatter epilogue; used when the input is a ViT token matrix [B, Dg*Hg*Wg, cin] fp16 (cin = 768).
 *   w_nk  fp16 [8*cout][cin], row (i*4+j*2+l)*cout + co = W[:, co, i, j, l]                           */
int dp_deconv2x_gemm(const void* tokens, const void* w_nk, int B, int Dg, int Hg, int Wg, int cin, int cout,
                     void* out_hi, void* out_lo, int out_cb_total, int out_cb_off, int* err_flag,
                     cudaStream_t stream);

/* F.interpolate(scale_factor=2, mode='trilinear', align_corners=True): c3d.py:36 */
int dp_upsample2x(const void* in_hi, const void* in_lo, int in_cb_total, int in_cb_off, int ncb, int N, int D, int H,
                  int W, void* out_hi, void* out_lo, int out_cb_total, int out_cb_off, cudaStream_t stream);

/* nn.LayerNorm(hidden), eps 1e-5 (monai TransformerBlock.norm1/norm2, ViT.norm) */
int dp_layernorm(const float* x, const float* gamma, const float* beta, int rows, int cols, void* out_f16,
                 float* out_f32, cudaStream_t stream);

/* softmax(dim=-1) of attention scores (monai SABlock.forward) */
int dp_softmax(const float* s, int rows, int cols, int ld_in, void* p, int ld_out, cudaStream_t stream);

/* einops Rearrange "b c (h p1)(w p2)(d p3) -> b (h w d)(p1 p2 p3 c)", p=16 (monai PatchEmbeddingBlock),
 * restated for c8 input: K order (c/8, p1, p2, p3, c%8)                                                */
int dp_patchify(const void* in_c8, int cb_total, int cb_off, int ncb, int N, int S0, int S1, int S2, void* out,
                cudaStream_t stream);

/* Patch-embedding Linear with the A operand GATHERED by TMA from the c8 activation (no dp_patchify copy):
 * out[token][hidden] = A[token][(c/8, p1, p2, p3, c%8)] . w_nk[hidden][same K order]^T (+bias, +rowvec[token % row_period] =
 * position embeddings) for split_k == 1; for split_k > 1 plain fp32 partials [split_k][N*tokens][hidden] (finish with
 * dp_splitk_reduce).  Volumes with H = W = 128, D a multiple of 32 (one TMA box = 2 x 8 x 8 patches x 64 K elements, the
 * 16-voxel patch pitch in the tensor map's byte strides); other shapes use dp_patchify + dp_gemm_tc.              */
int dp_gemm_patch_embed(const void* in_c8, int cb_total, int cb_off, int ncb, int N, int D, int H, int W, const void* w_nk,
                        int hidden, int split_k, const float* bias, const float* rowvec, int row_period, float* out_f32,
                        int* err_flag, cudaStream_t stream);

/* The same Rearrange for a ONE-channel planar fp32 volume [N][S0][S1][S2] (the seg net's CT input): out fp16
 * [N * tokens][4096], K order (p1, p2, p3) = the reference Linear weight's own column order.                */
int dp_patchify_planar(const float* in_planar, int N, int S0, int S1, int S2, void* out, cudaStream_t stream);

/* Cascade hand-off: argmax(8) -> one-hot -> drop background -> permute(0,3,2,1) -> cat(ptv, oars, ct^T)
 * (train_light_linked_model.py:156-167, OARSegmentation/config.py:70).  Writes the dose net's c8 input
 * (2 channel blocks: [PTV,7 OARs] [CT,0...]) and optionally NCDHW fp32 structures [N,9,S,S,S].        */
int dp_handoff(const float* logits, int ncls, const float* ptv, const float* ct, int N, int S, void* out_hi,
               void* out_lo, int out_cb_total, int out_cb_off, float* structures, cudaStream_t stream);

/* monai.inferers.sliding_window_inference (constant blending; train_light_linked_model.py:152-154):
 * dp_crop_pack     ROI windows (b, x0, y0, z0) of an NCDHW fp32 volume -> c8 predictor input, one batch entry per window
 * dp_window_add    out[b, :, window] += win[slot]   (one launch per window, reference accumulation order)
 * dp_div_count     data[r][v] /= count[v]           (count_map of the reference)                        */
int dp_crop_pack(const float* src, int C, int S0, int S1, int S2, int R, int n_win, const int* win_b, const int* win_x0,
                 const int* win_y0, const int* win_z0, void* hi, void* lo, int cb_total, int cb_off, cudaStream_t stream);
int dp_window_add(const float* win, int ncls, int R, int n_win, const int* win_b, const int* win_x0, const int* win_y0,
                  const int* win_z0, const int* win_slot, float* out, int S0, int S1, int S2, cudaStream_t stream);
int dp_div_count(float* data, const float* count, long long vol, int rows, cudaStream_t stream);

/* ======================================================================== training step
 * Pyfer.training_step (train_light_pyfer.py:122-143): train-mode forward, GenLoss (loss.py:69-119), backward
 * through net_B (net_A frozen, train_light_pyfer.py:85-88) and the optimizer update (:194-197).  The forward
 * convolutions, dgrad convolutions (same kernels, transposed/flipped weights) and the token GEMMs run through
 * the entry points above; the entry points below are the autograd / loss / optimizer pieces ATen provides
 * to the reference.
 *
 * "gradient sum": the upstream gradient of a tensor with several consumers arrives as up to three fp32 c8
 * tensors (n_g, g_f32[], g_cb_total[], g_cb_off[]) and/or one fp16 c8 tensor (g_f16, g16_cb_total,
 * g16_cb_off); the kernels add them on load (what autograd's AccumulateGrad / add_ nodes do).            */

/* Backward of dp_norm_act (native_batch_norm_backward / instance_norm backward + threshold / leaky_relu /
 * mish backward + the residual add of monai UnetResBlock.forward).  Same forward operands as dp_norm_act; the
 * forward values are recomputed from the saved pre-normalisation tensor.  phase 0 accumulates
 * bsum[N][C][6] = {sum g, sum g*xh, sum g_res, sum g_res*eh, dgamma, dbeta} (zero it first); phase 1 writes
 * dx (fp16 c8 for the tensor-core dgrad/wgrad kernels, or fp32 c8) and the residual-branch gradient.
 * For train-mode BatchNorm3d (blocks_MDUNet.py:103,106) run dp_batch_combine on stats / bsum between the
 * phases.                                                                                              */
int dp_norm_act_bwd(const float* raw_f32, const void* raw_hi, const void* raw_lo, int in_cb_total, int in_cb_off,
                    const double* stats, const float* gamma, const float* beta, int act, const void* res_hi,
                    const void* res_lo, const float* res_raw, const double* res_stats, int res_cb_total,
                    int res_cb_off, int act_after_res, int n_dy, const float* const* dy_f32, const int* dy_cb_total,
                    const int* dy_cb_off, const void* dy_f16, int dy16_cb_total, int dy16_cb_off, double* bsum,
                    int phase, void* dx_hi, float* dx_f32, int dx_cb_total, int dx_cb_off, void* dres_hi,
                    float* dres_f32, int dres_cb_total, int dres_cb_off, int N, int C, long long vox,
                    cudaStream_t stream);

/* nn.BatchNorm3d in train mode: replace the per-(n,c) slots [N][C][k] by their mean over n (so the
 * per-instance consumers see batch statistics) and, for k == 2 with running_mean != NULL, update the running
 * statistics (momentum, unbiased variance) exactly like torch.batch_norm(training=True).                */
int dp_batch_combine(double* slots, int N, int C, int k, long long vox, float* running_mean, float* running_var,
                     float momentum, cudaStream_t stream);

/* dgamma / dbeta of an affine norm from bsum slots 4, 5 (summed over n), times scale */
int dp_affine_grad(const double* bsum, int N, int C, float* dgamma, float* dbeta, float scale, cudaStream_t stream);

/* grad[i] = acc[i] * scale: converts the fp64 accumulation arena of the small parameter gradients */
int dp_grad_finalize(const double* acc, float* grad, long long n, float scale, cudaStream_t stream);

/* Weight gradient of nn.Conv3d (stride 1, k in {1,3,7}, dilation dil): slow_conv3d / cudnn wgrad of
 * blocks_MDUNet.py:68-71,102-105, monai UnetResBlock.conv1/conv2.  x: c8 fp16 forward input (16-channel chunks
 * chunk_cb[] holding logical channels chunk_ci0[] .. +chunk_nci[]), g: c8 fp16 gradient of the conv output.
 * Writes `splits` fp32 partials ws[splits][cout][cin][k][k][k] (sum them with dp_splitk_reduce).        */
int dp_conv3d_wgrad(const void* x_c8, int x_cb_total, const uint8_t* chunk_cb, const int* chunk_ci0,
                    const int* chunk_nci, int n_chunks, const void* g_c8, int g_cb_total, int g_cb_off, int N, int D,
                    int H, int W, int cin, int cout, int k, int dil, float* ws, int splits, cudaStream_t stream);

/* The same weight gradient on tcgen05 tensor cores (k in {3,7}, dilation 1): both operands are read MN-major
 * straight from the c8 layout (K = 16 voxels along W), the 16 M blocks are the x row at 16 overlapping one-voxel
 * shifts (the kw taps), the N dimension stacks the k rows of the g plane (the kh taps); accumulators stay in
 * TMEM over the CTA's whole share of the volume (wgrad_tc.cu).  Same partial-sum output as dp_conv3d_wgrad. */
int dp_conv3d_wgrad_tc(const void* x_c8, int x_cb_total, const uint8_t* chunk_cb, const int* chunk_ci0,
                       const int* chunk_nci, int n_chunks, const void* g_c8, int g_cb_total, int g_cb_off, int N, int D,
                       int H, int W, int cin, int cout, int k, float* ws, int splits, int* err_flag, cudaStream_t stream);

/* Weight (+bias) gradient of the 1x1x1 convolutions (blocks_MDUNet.py:145-157, monai UnetResBlock.conv3) and,
 * with deconv = 1, of nn.ConvTranspose3d k2 s2 (base_blocks.py:118-127, monai UnetrPrUpBlock):
 * dw[co*dw_co_stride + ci*dw_ci_stride + o*dw_o_stride] += sum g[co][child_o(v)] * x[ci][v]  (fp64 atomics).
 * x is a c8 fp16 tensor, or a token matrix [N][D*H*W][x_C] fp16 (x_tok).                                 */
int dp_small_wgrad(int n_g, const float* const* g_f32, const int* g_cb_total, const int* g_cb_off, const void* g_f16,
                   int g16_cb_total, int g16_cb_off, int g_C, const void* x_hi, const void* x_lo, int x_cb_total,
                   int x_cb_off, int x_C, const void* x_tok, int N, int D, int H, int W, int deconv, double* dw,
                   long long dw_co_stride, long long dw_ci_stride, long long dw_o_stride, double* dbias,
                   cudaStream_t stream);

/* Data gradient of nn.ConvTranspose3d k2 s2: dx[ci][v] = sum_{o,co} g[co][child_o(v)] * W[ci][co][o].
 * w fp32 [8][Co][Ci]; output fp32 c8 (dx_c8) or fp32 tokens [N][D*H*W][Ci] (dx_tok).                     */
int dp_deconv2x_bwd_data(int n_g, const float* const* g_f32, const int* g_cb_total, const int* g_cb_off,
                         const void* g_f16, int g16_cb_total, int g16_cb_off, int Co, const float* w, int Ci, int N,
                         int D, int H, int W, float* dx_c8, int dx_cb_total, int dx_cb_off, float* dx_tok,
                         cudaStream_t stream);

/* Backward of the dose heads (1x1x1 conv C -> 1, dose_pyfer.py:290-300,316-317): g planar fp32 [N][vox];
 * dx (fp32 c8) = g*w[c]; dw[c] += sum g*x[c]; db += sum g  (fp64 accumulators).                          */
int dp_head_bwd(const float* g, const void* x_hi, const void* x_lo, int x_cb_total, int x_cb_off, int C,
                const float* w, int N, long long vox, float* dx, int dx_cb_total, int dx_cb_off, double* dw,
                double* db, cudaStream_t stream);

/* GenLoss (loss.py:57-119): masked L1 of one prediction scale s^3 against the GT [N][2][S^3] (dose, possible-dose
 * mask), the GT resampled like loss.py:63-64 (trilinear align_corners=True / nearest-exact).  phase 0:
 * acc[0] += sum |p-t|, acc[1] += #mask;  phase 1: dpred = coef * sign(p-t) / acc[1].                     */
int dp_masked_l1(const float* pred, const float* gt, int N, int S, int s, double* acc, int phase, float coef,
                 float* dpred, cudaStream_t stream);
/* loss = delta1 * acc[0]/acc[1] + delta2 * mean_{i>=1} acc[2i]/acc[2i+1]   (loss.py:96-112)
 *        [+ weight_a * acc_a[0]/acc_a[1]: the net_A term of `casecade and not freez`, loss.py:114-115; acc_a may be NULL] */
int dp_genloss_finalize(const double* acc, int n_scales, float delta1, float delta2, const double* acc_a, float weight_a,
                        float* loss, cudaStream_t stream);
/* Adjoint of the 2x linear interpolation (align_corners=True) along one axis: g fp32 [outer][2*len_in][inner][8] ->
 * out [outer][len_in][inner][8].  Three passes (D, H, W) = backward of F.interpolate(scale_factor=2, 'trilinear',
 * align_corners=True) in UpConv.forward (c3d.py:35-38) on c8 fp32 tensors; deterministic gather. */
int dp_lerp2x_bwd(const float* g, long long outer, int len_in, long long inner, float* out, cudaStream_t stream);

/* Seg training loss: monai 0.7.0 DiceCELoss(to_onehot_y=True, softmax=True) as used at
 * OARSegmentation/train_light_transeg.py:148,186.  logits: c8 fp32 with C <= 8 classes in channel block 0; label:
 * fp32 class index per voxel.  acc = double[N*24 + 1].  phase 0 accumulates {sum p*t, sum t, sum p} per (n,c) and
 * the cross-entropy sum; phase 1 writes coef * dLoss/dlogit as fp16 c8.                                         */
int dp_dice_ce(const float* logits_c8, int cb_total, const float* label, int N, int C, long long vox, double* acc,
               int phase, float coef, void* g_f16, int g_cb_total, cudaStream_t stream);
int dp_dice_ce_finalize(const double* acc, int N, int C, long long vox, float* loss, cudaStream_t stream);

/* Fused AdamW over a flat fp32 parameter buffer (configure_optimizers, train_light_pyfer.py:194-197; fp32
 * optimizer state, decoupled weight decay).  inv_scale undoes the static loss scaling; when *found_inf != 0
 * (dp_grad_check) the update is skipped.                                                                */
int dp_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
             float weight_decay, int step, float inv_scale, const int* found_inf, cudaStream_t stream);
/* dp_adamw with the step counter on the device: state = int[3] {completed steps, consecutive skipped steps, total
 * skipped steps}; a step skipped for non-finite gradients does not advance the bias correction (torch.optim.AdamW
 * under GradScaler behaves the same way).                                                                       */
int dp_adamw_dev(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                 float weight_decay, float inv_scale, const int* found_inf, int* state, cudaStream_t stream);
int dp_grad_check(const float* g, long long n, int* found_inf, cudaStream_t stream);

/* bitsandbytes 8-bit optimizer-state layout (bnb.optim.Adam8bit, train_light_pyfer.py:194-197): block-wise (2048 elements)
 * quantisation of a state tensor to uint8 codes into a sorted 256-entry code book times the block's absmax, and back.
 * codes_u8 [n], absmax [ceil(n / 2048)], qmap256 fp32 [256] (bnb's "dynamic" map; optim8bit.create_dynamic_map).      */
int dp_quantize_blockwise(const float* x, long long n, const float* qmap256, void* codes_u8, float* absmax, cudaStream_t stream);
int dp_dequantize_blockwise(const void* codes_u8, const float* absmax, const float* qmap256, long long n, float* x,
                            cudaStream_t stream);

/* Per-step re-packing of the live fp32 parameters into the kernels' fp16 operand layouts (what the inference
 * plans do once on the host side): conv weights [cout][cin][k][k][k] -> dp_conv3d_tc / dp_conv3d_stack layout for
 * 16-channel chunks holding logical input channels chunk_ci0[] .. +chunk_nci[]; transpose_flip = 1 reads the
 * stored tensor as the dgrad weights W'[ci][co][k-1-kd][k-1-kh][k-1-kw] (cout/cin are the packed conv's dims). */
int dp_pack_conv_weight(const float* w, int cout, int cin, int k, int transpose_flip, const int* chunk_ci0,
                        const int* chunk_nci, int n_chunks, int stacked, void* out, cudaStream_t stream);
int dp_cast_f16(const float* src, long long n, void* dst, cudaStream_t stream);

/* Token-side backward pieces of monai ViT (nn.LayerNorm, softmax, GELU backward; layout shuffles that feed
 * the dgrad / wgrad GEMMs of dp_gemm_tc, which wants both operands K-major).                             */
int dp_layernorm_bwd(const float* x, const float* gamma, const float* dy, const float* add, int rows, int cols,
                     float* dx, double* dgamma, double* dbeta, cudaStream_t stream);
int dp_softmax_bwd(const void* probs, int ld_p, const float* dprobs, int ld_dp, int rows, int cols, void* ds, int ld_ds,
                   cudaStream_t stream);
int dp_act_fwd(const float* u, long long n, int act, void* y_f16, cudaStream_t stream);
int dp_act_bwd(const float* u, const float* dh, long long n, int act, float* du_f32, void* du_f16, cudaStream_t stream);
/* dst[b][c][r] = scale * src[b][r][c] (fp32 or fp16 source, fp16 destination with row pitch ld_dst) */
int dp_transpose(const void* src, int src_f32, long long src_batch_stride, int ld_src, int R, int C, void* dst_f16,
                 long long dst_batch_stride, int ld_dst, int batch, float scale, cudaStream_t stream);
/* split (merge = 0) rows [B][T][ld] (columns col0 + head*hd + d) into [B*heads][T][hd], or merge back */
int dp_heads(const void* src, int src_f32, void* dst, int dst_f32, int B, int T, int heads, int hd, int ld, int col0,
             int merge, float scale, cudaStream_t stream);
/* out[c] += sum_r a[r][c] (bias / position-embedding gradients, fp64 accumulators) */
int dp_colsum(const float* a, int rows, long long cols, double* out, cudaStream_t stream);
/* y = a + b (b optional), fp32 and/or fp16 outputs */
int dp_add(const float* a, const float* b, long long n, float* y_f32, void* y_f16, cudaStream_t stream);

/* ======================================================================== on-device evaluation
 * What Pyfer.test_step / LinkedNet.test_step do with numpy after a D2H copy (train_light_pyfer.py:210-216,
 * train_light_linked_model.py:171-176, DosePrediction/Evaluate/evaluate_openKBP.py).                       */

/* out = (mask < 1 || pred < 0) ? 0 : pred * scale      (train_light_pyfer.py:210-213) */
int dp_dose_postprocess(const float* pred, const float* mask, long long n, float scale, float* out, cudaStream_t stream);

/* get_3D_Dose_dif (evaluate_openKBP.py:42-48) and IVS (:17-39) at the n_levels ascending isodose levels (device
 * double[n_levels], np.linspace(0, 70, 101) in the reference).  acc: double[2], hist: uint64[3*(n_levels+1)]
 * scratch; ivs: float[n_levels], dose_dif: float[1].                                                          */
int dp_dose_stats(const float* pred, const float* gt, const float* mask, long long n, const double* levels, int n_levels,
                  double* acc, unsigned long long* hist, float* ivs, float* dose_dif, cudaStream_t stream);

/* get_DVH_metrics (evaluate_openKBP.py:51-81) for n_struct ROI masks [n_struct][vox] (fp32, > 0 = inside) of the
 * predicted and the ground-truth dose at once: exact np.percentile(linear) order statistics by a three-pass
 * radix select.  is_target: device int[n_struct] (1 = PTV: D1, D95, D99, mean; 0 = OAR: D_0.1cc, mean).
 * out: float[n_struct][2 (pred, gt)][5] = {q0, q1, q2, unused, mean}; dvh_dif: float[1] = mean |gt - pred| over the
 * metrics of all non-empty structures (:206-222).  workspace: dp_dvh_workspace_bytes() bytes.                 */
long long dp_dvh_workspace_bytes(void);
int dp_dvh_metrics(const float* pred, const float* gt, const float* masks, int n_struct, const int* is_target, long long vox,
                   float voxels_in_tenth_of_cc, void* workspace, float* out, float* dvh_dif, cudaStream_t stream);

/* Seg validation / test metric: monai 0.7.0 HausdorffDistanceMetric(include_background=False, percentile=95) on
 * one-hot(argmax(logits)) vs the label map for ONE volume (OARSegmentation/train_light_transeg.py:158-166,199-216):
 * surface voxels (mask ^ binary_erosion(mask)), exact Euclidean distance to the other surface, np.percentile in both
 * directions, the larger one.  logits [C][D][H][W] fp32, label [D][H][W] fp32 class indices, hd float[C] (class 0 = NaN;
 * NaN when a class has no surface on either side as source, inf when only the target surface is missing).          */
long long dp_hd95_workspace_bytes(int C, int D, int H, int W);
int dp_hd95(const float* logits, const float* label, int C, int D, int H, int W, float percentile, void* workspace, float* hd,
            cudaStream_t stream);

/* ======================================================================== input pipeline
 * The numpy / monai transforms of DosePrediction/DataLoader/dataloader_OpenKBP_monai.py:160-243 after the files
 * are read.  dp_prepare_input: raw arrays [A][B][C] (masks uint8, NULL = structure absent; CT int16 or fp32 HU;
 * dose fp32 Gy) -> Input [9][C][B][A] = [PTV (70/70, 63/70, 56/70 merge, :113-125), 7 OARs, clip(CT)/1000 + ct_shift
 * (:137-146, RandShiftIntensityd :189-193)] and GT [2][C][B][A] = [dose/70, dose_mask] (:128-134,:195-201), including
 * Transposed(indices=[2,1,0]) (:173).  ptv_u8 / oar_u8: HOST arrays of 3 / 7 device pointers.
 * dp_flip_rot90: out = np.rot90(np.flip(in, flipped spatial axes), k, axes=(0,1)) per channel (RandFlipd x3 :214-228,
 * RandRotate90d :229-233) for in [C][S0][S1][S2]; out is [C][S1][S0][S2] when k is odd.                       */
int dp_prepare_input(const void* const* ptv_u8, const void* const* oar_u8, const void* ct_i16, const float* ct_f32,
                     const float* dose, const void* dose_mask_u8, int A, int B, int C, float a_min, float a_max,
                     float ct_shift, float* input, float* gt, cudaStream_t stream);
int dp_flip_rot90(const float* in, float* out, int C, int S0, int S1, int S2, int flip0, int flip1, int flip2, int k,
                  cudaStream_t stream);

/* Orientationd(axcodes="RAS") (dataloader_OpenKBP_monai.py:182): out axis a = in axis perm[a], reversed when flip[a]
 * (the transposes + flips nibabel's ornt_transform prescribes; the host derives them from the NIfTI affine). */
int dp_permute_flip(const float* in, float* out, int C, int S0, int S1, int S2, int perm0, int perm1, int perm2, int flip0,
                    int flip1, int flip2, cudaStream_t stream);

/* RandCropByPosNegLabeld (monai 0.7.0; dataloader_OpenKBP_monai.py:206-215, OARSegmentation/DataLoader/provided_dataset.py:
 * 158-167).  Foreground = any label channel > 0, background = any image channel > image_threshold and not foreground
 * (image NULL: every non-foreground voxel).  dp_posneg_count: per 4096-voxel block {foreground, background} counts
 * (block_counts int[blocks][2]); the host draws the samples with numpy RandomState semantics from the totals and names,
 * per sample, {block, rank inside the block, want_foreground} (pick_dev int[n][3], device memory).  dp_posneg_crop: finds
 * those voxels, clamps the centres like correct_crop_centers, writes the crop origins (roi_start_dev int[n][3]) and crops
 * every source [C_i][S0][S1][S2] to dst_i [n][C_i][R][R][R].  src / src_channels / dst: HOST arrays.                */
int dp_posneg_count(const float* label, int label_channels, const float* image, int image_channels, float image_threshold,
                    long long vox, int* block_counts, cudaStream_t stream);
int dp_posneg_crop(const float* label, int label_channels, const float* image, int image_channels, float image_threshold,
                   int S0, int S1, int S2, int R, int n_samples, const int* pick_dev, int* roi_start_dev, int n_src,
                   const float* const* src, const int* src_channels, float* const* dst, cudaStream_t stream);

/* Seg validation metric: monai 0.7.0 DiceMetric(include_background=False, reduction="mean") on one-hot(argmax(logits))
 * vs the label map (OARSegmentation/train_light_transeg.py:199-216).  logits: NCDHW fp32 [N][C][vox]; label: fp32 class
 * index [N][vox]; counts: uint64[N*48] scratch; dice: float[N][C] (NaN where the class is absent from the label);
 * mean_dice: float[1], mean over the non-NaN foreground entries.                                               */
int dp_dice_metric(const float* logits, const float* label, int N, int C, long long vox, unsigned long long* counts,
                   float* dice, float* mean_dice, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DOSE_B200_H_ */
