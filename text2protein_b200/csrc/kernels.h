// Internal C++ launch API of the t2p CUDA kernels (one function per kernel family).
// Activations are NHWC ("pixels x channels") so that every convolution / projection is a
// row-major [M = B*H*W, K] x [N, K]^T contraction; weights are packed [N][taps][Cin].
#pragma once
#include "common.h"

namespace t2p {

// ----------------------------------------------------------------------------- implicit GEMM
// out[m, n] = alpha * ( sum_k A[m, k] * Wt[n, k] + bias[n] + rowbias[m / rows_per_sample, n]
//                       + residual[m, n] )
// A is the im2col view (ksize 1 or 3, stride 1, zero pad ksize/2) of up to two NHWC sources
// concatenated along channels (the UNet skip concat is never materialised).
struct ConvGemmArgs {
  const void* a0 = nullptr;
  int c0 = 0;
  const void* a1 = nullptr;
  int c1 = 0;
  int B = 1, H = 1, W = 1;  // geometry of the A sources; M = B*H*W
  int ksize = 1;
  const void* w = nullptr;  // [N][ksize*ksize*(c0+c1)], same dtype as A
  int N = 0;
  const float* bias = nullptr;     // [N]
  const float* rowbias = nullptr;  // [M / rows_per_sample][N]
  int rows_per_sample = 0;
  const void* residual = nullptr;  // [M][N], out dtype
  int res_up = 0;                  // 1: residual is [B][H/2][W/2][N] and is nearest-upsampled x2
  float alpha = 1.f;
  void* out = nullptr;
  int out_dtype = kBF16;
  // optional fused GroupNorm statistics of the stored output: per pixel tile and channel {sum, sum of squares},
  // [M / T][N][2] == part[B][rows_per_sample / T][N][2], T = conv_gemm_tc_stat_tile(args) (tcgen05 kernels only)
  float* stat_part = nullptr;
  // extra sources entering through the centre tap only -- a 1x1 convolution over x0|x1 summed with the
  // ksize x ksize one over a0|a1 (the ResBlock skip path folded into Conv_1): K grows by xc0 + xc1 and a weight
  // row is [taps * (c0 + c1) | xc0 | xc1].  Channel-major tcgen05 kernel only (conv_gemm_tc_channel_major()).
  const void* x0 = nullptr;
  int xc0 = 0;
  const void* x1 = nullptr;
  int xc1 = 0;
  int rowbias_ld = 0;  // row pitch of rowbias (0 -> N)
  // 1: process the pixel tiles from the last to the first.  Consecutive streaming kernels alternate direction so
  // that each starts on the part of its input the previous kernel wrote last (still resident in the 126 MB L2).
  int reverse = 0;
  int out_nchw = 0;    // 1: store out as [B][N][H*W] (fp32 only; used by the final conv)
  // GroupNorm-apply + SiLU fused into the operand path: a0 | a1 are RAW activations and the kernel feeds
  // silu(x * gn_scale[b][c] + gn_shift[b][c]) (affine over the channel concat, [B][c0 + c1] fp32) to the tensor
  // core.  Only where conv_gemm_tc_fuses_gn(args) says so (3x3 on 128-pixel-wide images: the halo kernel).
  const float* gn_scale = nullptr;
  const float* gn_shift = nullptr;
  // GroupNorm + SiLU of the OUTPUT inside the epilogue: out = silu(GroupNorm(acc * alpha + bias + rowbias)) with the
  // CONSUMER's parameters (ResBlock: GroupNorm_1 after Conv_0).  Only where conv_gemm_tc_gn_out_ok(args); scratch the
  // CALLER keeps in a fixed state between launches: gno_part >= conv_gemm_tc_gn_out_part_floats(args) floats with every
  // byte 0xff (the "not written yet" sentinel), gno_flags conv_gemm_tc_gn_out_flag_ints(args) ints that are zero; the
  // kernel restores both before it ends.
  // Split-K scratch (channel-major tcgen05 kernel, launches of few tiles; see conv_gemm_tc_splits): fp32 partial
  // accumulators, kSplitKPartFloats floats, and kSplitKTicketInts arrival counters that are ZERO between launches.
  // Without them the launch does not split.
  float* sk_part = nullptr;
  int* sk_ticket = nullptr;
  // Optional: a device range the launch asks the L2 to fetch while it runs (tcgen05 kernels) -- the engine passes the
  // NEXT GEMM's weights: at small batches a layer is a chain of k-blocks at the latency of its weight loads, and the
  // 200 MB of weights do not survive in the 126 MB L2 from one forward to the next.
  const void* l2_prefetch = nullptr;
  long long l2_prefetch_bytes = 0;
  const float* gno_gamma = nullptr;
  const float* gno_beta = nullptr;
  int gno_groups = 0;
  float gno_eps = 1e-6f;
  void* gno_part = nullptr;
  int* gno_flags = nullptr;
};

// bf16 tcgen05 / TMEM / TMA path (sm_100a).  A, W are bf16.
void conv_gemm_tc(const ConvGemmArgs& a, cudaStream_t st);
// pixel-tile size T the tcgen05 launch for these arguments uses for fused GroupNorm statistics
// (rows_per_sample % T == 0), or 0 when it cannot produce them
int conv_gemm_tc_stat_tile(const ConvGemmArgs& a);
// true when these arguments run on the channel-major kernel (the one that accepts x0 / x1)
bool conv_gemm_tc_channel_major(const ConvGemmArgs& a);
// true when this launch can apply GroupNorm + SiLU to its 3x3 sources itself (gn_scale / gn_shift)
bool conv_gemm_tc_fuses_gn(const ConvGemmArgs& a);
// true when this launch can normalise + activate its own output (gno_*): channel-major kernel, whole pixel tiles per
// sample, N % 128 == 0, `groups` groups of 4 / 8 / 16 / 32 channels, and few enough tiles per sample for the inter-CTA wait
bool conv_gemm_tc_gn_out_ok(const ConvGemmArgs& a, int groups);
long long conv_gemm_tc_gn_out_part_floats(const ConvGemmArgs& a, int groups);
long long conv_gemm_tc_gn_out_flag_ints(const ConvGemmArgs& a);  // size of gno_flags
// K splits conv_gemm_tc uses for these arguments when given the split-K scratch (1 = none).  Host-only.
int conv_gemm_tc_splits(const ConvGemmArgs& a);
constexpr long long kSplitKMaxItems = 160;                                // (tile, split) work items of a split launch
constexpr long long kSplitKPartFloats = kSplitKMaxItems * 128 * 256;      // [item][128 channels][<= 256 pixels]
constexpr long long kSplitKTicketInts = kSplitKMaxItems * 8;              // [tile][8 epilogue warps]
// fp32 or bf16 SIMT path (verification mode and tiny-channel edge layers). dtype of A and W.
void conv_gemm_simt(const ConvGemmArgs& a, int in_dtype, cudaStream_t st);

// ----------------------------------------------------------------------------- normalisation
// GroupNorm statistics: per-(sample, pixel block, channel) {sum, sum of squares}, part[B][nblk][C][2] floats,
// nblk = gn_stats_blocks(B, HW); reduced without atomics so results are run-to-run deterministic.
int gn_stats_blocks(int B, int HW);
void gn_stats(const void* a0, int c0, const void* a1, int c1, int B, int HW, int dtype, float* part,
              cudaStream_t st);
// -> per-(sample, channel) affine  y = x * scale + shift  (fp32 [B][C] each) over the channel concat of two
// statistics sources (part0: channels [0, c0), part1: [c0, c0 + c1)); blocks are added in order, in double
void gn_finalize(const float* part0, int nblk0, int c0, const float* part1, int nblk1, int c1, const float* gamma,
                 const float* beta, int B, int G, int HW, float eps, float* scale, float* shift, cudaStream_t st);
// y = act(x * scale + shift) over the (virtual) channel concat of a0|a1; mode 0 same size, 1 = 2x2 mean
// after the activation (raw_out, optional, receives the 2x2 mean of the raw input), 2 = nearest x2 upsample
// (raw_out, optional, receives the upsampled raw input).
void gn_apply(const void* a0, int c0, const void* a1, int c1, int B, int H, int W, int dtype, const float* scale,
              const float* shift, int act, int mode, void* out, void* raw_out, cudaStream_t st, int reverse = 0);
// statistics + finalize + apply in ONE launch for small tensors (HW <= 1024, groups of 4 / 8 / 16 / 32 channels)
bool gn_small_supported(int c0, int c1, int HW, int G);
void gn_small(const void* a0, int c0, const void* a1, int c1, int B, int HW, int dtype, int G, float eps,
              const float* gamma, const float* beta, int act, void* out, cudaStream_t st);
void layernorm(const void* x, const float* gamma, const float* beta, long long M, int C, float eps, int dtype,
               void* y, cudaStream_t st);
void geglu(const void* z, long long M, int D, int dtype, void* out, cudaStream_t st);

// ----------------------------------------------------------------------------- attention
struct AttnArgs {
  const void* q = nullptr;
  const void* k = nullptr;
  const void* v = nullptr;
  void* out = nullptr;
  int B = 0, heads = 1, Tq = 0, Tk = 0, d = 0;
  long long ldq = 0, ldk = 0, ldv = 0, ldo = 0;  // row pitches in elements
  float scale = 1.f;
};
void attention_simt(const AttnArgs& a, int dtype, cudaStream_t st);
void attention_mma(const AttnArgs& a, cudaStream_t st);  // bf16 mma.sync kernel (head dims the tcgen05 kernel lacks)
bool attention_mma_supported(const AttnArgs& a);
// bf16 tcgen05 kernel (attention_tc.cu): scores and output accumulator in TMEM, operands by TMA; d in {32, 64, 128}
// or a multiple of 256
void attention_tc(const AttnArgs& a, cudaStream_t st);
bool attention_tc_supported(const AttnArgs& a);

// ----------------------------------------------------------------------------- helpers
// timesteps: optional fp32 [B] time conditioning to embed instead of float(labels)
void temb_mlp(const long long* labels, const float* timesteps, int B, int nf, const float* w0, const float* b0, const float* w1,
              const float* b1, float* out, cudaStream_t st);
// `ld` = row pitch of `out` in elements (0: dense); lets several K segments share one packed row
void pack_conv_weight(const float* w, int cout, int cin, int k, int cin_pad, int out_dtype, void* out,
                      cudaStream_t st, long long ld = 0);
void pack_matrix(const float* w, int rows, int cols, int transpose, int out_dtype, void* out, cudaStream_t st,
                 long long ld = 0);
void pack_identity(int n, long long ld, int out_dtype, void* out, cudaStream_t st);
void add_vectors_f32(const float* a, const float* b, int n, float* out, cudaStream_t st);
void broadcast_row_f32(float* buf, int n, int B, cudaStream_t st);  // rows 1..B-1 of [B][n] = row 0
void convert_f32(const float* in, long long n, int out_dtype, void* out, cudaStream_t st);
void nhwc_to_nchw_f32(const void* in, int dtype, int B, int HW, int C, float* out, cudaStream_t st);
void nchw_f32_to_nhwc(const float* in, int B, int HW, int C, int cpad, int out_dtype, void* out, cudaStream_t st);
// first conv (Cin = 5 / 8) as a K = kpad GEMM on the tensor cores: im2col of the fp32 NCHW state + matching weights
void im2col3x3_nchw(const float* x, int B, int C, int H, int W, int kpad, void* out, cudaStream_t st);
void pack_first_conv(const float* w, int cout, int C, int kpad, void* out, cudaStream_t st);
void scale_by_sigma(const float* h_nchw, const long long* labels, const double* sigmas, int B, int HW, int C,
                    int do_scale, int out_dtype, void* out, cudaStream_t st);

// ----------------------------------------------------------------------------- PC sampler steps
struct PeerGroup;
struct PcStepArgs {
  float* x = nullptr;           // [B][C][HW] fp32 state, updated in place
  const void* score = nullptr;  // fp32 or fp64; NCHW or NHWC
  int score_dtype = kF32;
  int score_nhwc = 0;
  const double* sigmas = nullptr;   // optional: score = raw / sigmas[labels[b]]
  const long long* labels = nullptr;
  const float* G = nullptr;         // [B]
  const float* sqrt_alpha = nullptr;
  const float* alpha = nullptr;
  int probability_flow = 0;
  float snr = 0.f;
  const unsigned char* mask = nullptr;
  const float* x_init = nullptr;
  float* x_mean_out = nullptr;
  unsigned long long seed = 0;
  long long stream_base = 0, stream_mul = 0;
  const long long* iter_ptr = nullptr;
  long long sample_offset = 0;
  int B = 0, C = 0, HW = 0;
  double* partial = nullptr; // corrector only: pc_corrector_workspace_doubles(B, C*HW) doubles
  int conditioned_in_place = 0;  // caller guarantees x_out (and x_mean_out) already hold x_init where mask == 0
  float* x_out = nullptr;        // new state (nullptr: x, in place); must differ from x with `symmetrize`
  // 1: channels 0 and 1 (Cb-Cb distance, omega: symmetric maps) of the new state and of x_mean take their
  // symmetric part 0.5 (u[i][j] + u[j][i]) wherever both positions are free -- needs a square image (W * W == HW)
  int symmetrize = 0;
  int W = 0;
  // x_mean_out is written only when *iter_ptr == *last_iter_ptr (the sampler returns the x_mean of its LAST
  // predictor step, sampling.py:289); nullptr: always
  const long long* last_iter_ptr = nullptr;
  // Langevin step size over the GLOBAL batch of a sharded run: after the norm phase every rank writes
  // its (sum_b ||grad_b||, sum_b ||noise_b||) into the mailbox of every peer over NVLink and waits for theirs
  const PeerGroup* peers = nullptr;
  const long long* tag_base_ptr = nullptr;  // device: tag of this run's first corrector step
};

// mailboxes of the ranks of one node, mapped into this process (cudaIpcOpenMemHandle); see pc_step.cu
struct PeerGroup {
  static constexpr int kMaxWorld = 16;
  int world = 1, rank = 0;
  long long global_batch = 0;
  unsigned long long* box[kMaxWorld] = {};  // box[r]: mailbox of rank r, [2 parities][world] slots of 4 x u64
};
void pc_predictor_step(const PcStepArgs& a, cudaStream_t st);
void pc_corrector_step(const PcStepArgs& a, cudaStream_t st);
long long pc_corrector_workspace_doubles(int B, long long E);
void philox_normal_fill(unsigned long long seed, unsigned long long stream, long long first_element, long long count,
                        float scale, float* out, cudaStream_t st);
void philox_bits_fill(unsigned long long seed, unsigned long long stream, long long first_quad, long long quads,
                      unsigned int* out, cudaStream_t st);
void run_prep(long long* state, const long long* label_table, const float* g_table, int B, long long* labels, float* G,
              cudaStream_t st);
void apply_mask(float* x, const unsigned char* mask, const float* fixed, long long n, cudaStream_t st);

// ----------------------------------------------------------------------------- last layer (final_conv.cu)
// out[B][nout][H][W] fp32 = conv3x3(silu(x * scale + shift)) + bias on the RAW bf16 NHWC activation x; w is the
// packed [nout][9][cin] bf16 weight.  GroupNorm-apply, SiLU and the convolution in one pass over the tensor.
bool final_conv_fused_supported(int cin, int nout, int H, int W);
void final_conv_fused(const void* x, const float* scale, const float* shift, const void* w, const float* bias, float* out,
                      int B, int H, int W, int cin, int nout, cudaStream_t st);

// ----------------------------------------------------------------------------- callers of the loop (conditions.cu)
void length_mask(const int* lengths, int B, int N, unsigned char* out, cudaStream_t st);
void inpaint_mask(const int* ranges, int R, int per_sample, int B, int N, unsigned char* out, cudaStream_t st);
void condition_mask(const int* lengths, const int* ranges, int R, int per_sample, int has_ss, int B, int C, int N,
                    unsigned char* out, cudaStream_t st);
void embed_gather(const void* table, int table_dtype, long long V, int D, const long long* tokens, long long n,
                  float* out_f32, void* out_bf16, cudaStream_t st);
void postprocess_6d(const float* x, int B, int C, int N, float* out, int* L_out, cudaStream_t st);

}  // namespace t2p
