/* t2p.h -- C ABI of the B200-native Text2Protein sampling path (libt2p.so).
 *
 * The reference (szhan227/text2protein) has no FFI layer: its boundary is the Python API of
 * score_sde_pytorch/{sampling,sde_lib,utils}.py and score_sde_pytorch/models/{ncsnpp,utils}.py.  The Python
 * mirror in text2protein_b200/ keeps those names and signatures and forwards to the entry points below through
 * ctypes.  Every entry point cites the reference code it replaces (paths relative to the reference root).
 *
 * Conventions: plain C types only; all tensor pointers are DEVICE pointers owned by the caller; `stream` is a
 * cudaStream_t passed as void*; no entry point synchronises the host except where stated; return value 0 = ok,
 * non-zero = failure with the message available from t2p_last_error() (thread-local).
 */
#ifndef T2P_H_
#define T2P_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define T2P_ABI_VERSION 3

enum t2p_dtype { T2P_F32 = 0, T2P_BF16 = 1, T2P_F64 = 2, T2P_I64 = 3, T2P_U8 = 4 };

const char* t2p_last_error(void);
int t2p_abi_version(void);

/* Layout of the argument structs as THIS library was compiled: `which` = 0 t2p_unet_cfg, 1 t2p_step_args,
 * 2 t2p_run_args, 3 t2p_conv_args, 4 t2p_gemm_record.  t2p_sizeof returns sizeof (or -1); t2p_struct_layout fills
 * the byte offset of every field in declaration order (up to `cap`) and returns the field count.  A binding checks
 * both against its own declaration before the first call (tests/test_abi.py does, for the ctypes binding). */
int t2p_sizeof(int which);
int t2p_struct_layout(int which, int32_t* offsets, int cap);

/* ---------------------------------------------------------------------------------------------------------
 * Score network.  Replaces score_sde_pytorch/utils.py:4-9 get_model + models/ncsnpp.py:74-263 UNetModel. */
typedef struct t2p_unet t2p_unet;

typedef struct t2p_unet_cfg {
  int32_t num_channels;        /* data.num_channels                 ncsnpp.py:137 */
  int32_t max_res_num;         /* data.max_res_num                  ncsnpp.py:87  */
  int32_t nf;                  /* model.nf                          ncsnpp.py:80  */
  int32_t n_ch_mult;
  int32_t ch_mult[16];         /* model.ch_mult                     ncsnpp.py:81  */
  int32_t num_res_blocks;      /* model.num_res_blocks              ncsnpp.py:82  */
  int32_t n_attn_resolutions;
  int32_t attn_resolutions[16];/* model.attn_resolutions            ncsnpp.py:83  */
  int32_t n_heads;             /* model.n_heads                     ncsnpp.py:94  */
  int32_t context_dim;         /* model.context_dim                 ncsnpp.py:95  */
  int32_t num_scales;          /* model.num_scales (sigmas buffer)  ncsnpp.py:78  */
  int32_t scale_by_sigma;      /* model.scale_by_sigma              ncsnpp.py:259 */
  int32_t compute_dtype;       /* T2P_BF16: tcgen05 path; T2P_F32: CUDA-core verification path */
} t2p_unet_cfg;

int t2p_unet_create(const t2p_unet_cfg* cfg, t2p_unet** out);
void t2p_unet_destroy(t2p_unet* u);

/* Parameter tree introspection, in the reference's state_dict order (705 entries for cond_length.yml). */
int t2p_unet_num_params(const t2p_unet* u);
int t2p_unet_param_info(const t2p_unet* u, int index, char* name_buf, int name_cap, int64_t* shape4, int* ndim,
                        int* dtype);

/* Copies one state_dict tensor (fp32, or fp64 for "sigmas"; a leading "module." is ignored, as produced by
 * DataParallel, score_sde_pytorch/utils.py:8) into the engine.  Replaces nn.Module.load_state_dict /
 * ExponentialMovingAverage.copy_to (models/ema.py:51-61) on the device side. */
int t2p_unet_load(t2p_unet* u, const char* name, const void* dev_ptr, const int64_t* shape, int ndim, int dtype,
                  void* stream);
/* Repacks all weights into kernel layout ([Cout][kh][kw][Cin], fused QKV, stacked Dense_0).  Call after loads. */
int t2p_unet_finalize(t2p_unet* u, void* stream);

/* Text context, fp32 [B][L][context_dim].  Projects to_k / to_v of every cross-attention ONCE per run
 * (model/attention.py:174-175 recomputes them in each of the 2*num_scales forwards).  Synchronises `stream`. */
int t2p_unet_set_context(t2p_unet* u, const float* ctx, int B, int L, void* stream);

/* UNetModel.forward(x, time_cond, text_emb), ncsnpp.py:220-263.  x fp32 [B][C][N][N]; labels int64 [B];
 * out [B][C][N][N] in `out_dtype` (T2P_F64 reproduces the reference's promoted dtype). */
int t2p_unet_forward(t2p_unet* u, const float* x, const int64_t* labels, void* out, int out_dtype, int B,
                     void* stream);
/* Same with a FLOATING time conditioning (models/utils.py:151-152: the VP branch of get_score_fn passes
 * labels = t * (N - 1)): ncsnpp.py:221-223 embeds the float value (`timesteps`, fp32 [B]) and uses its truncation
 * (`labels`, = time_cond.long()) only for the sigma lookup. */
int t2p_unet_forward_t(t2p_unet* u, const float* x, const int64_t* labels, const float* timesteps, void* out,
                       int out_dtype, int B, void* stream);

/* Option, default off: apply GroupNorm + SiLU inside the operand path of the 3x3 convolutions on 128-pixel-wide
 * images (the layers holding most of the FLOPs) instead of a separate pass over the activation -- h = act(GroupNorm(x)),
 * layers.py:305,318, never reaches HBM.  Same results to bf16 rounding; measured break-even in time on
 * cond_length.yml at B = 64 (profiles/r02_fused_gn_ab.txt), 1.1 GB less activation arena. */
int t2p_unet_set_fused_groupnorm(t2p_unet* u, int enable);

/* Option, default ON: a ResBlock's Conv_0 applies GroupNorm_1 + SiLU to its own output in its epilogue (per-sample
 * statistics exchanged between the CTAs of the launch), h = act(GroupNorm_1(Conv_0(.) + temb)), layers.py:314-318 -- the
 * raw Conv_0 output and the separate statistics / finalize / apply passes over it disappear wherever
 * t2p_conv2d_normalises_output holds for the launch.  Off = the three-kernel sequence of rounds 1-2 (A/B, debugging). */
int t2p_unet_set_epilogue_groupnorm(t2p_unet* u, int enable);

/* Debug taps: with debug on, every top-level block's output is kept as fp32 NCHW ("pre_conv",
 * "input_blocks.<i>", "mid_blocks", "out_blocks.<i>", "out"). */
int t2p_unet_set_debug(t2p_unet* u, int enable);
int t2p_unet_tap(t2p_unet* u, const char* name, float* dst, int64_t capacity, int64_t* shape4, void* stream);
/* Profile mode: CUDA events around every implicit-GEMM launch of the following forward passes (eager only).
 * t2p_unet_profile_read synchronises, fills up to `cap` records and returns the number recorded (or -1). */
typedef struct t2p_gemm_record {
  int64_t M; int32_t N; int32_t K; int32_t ksize;
  int32_t tensor_core;  /* 0 = CUDA-core kernel, 1 = tcgen05 kernel, 2 = tcgen05 kernel that also applied GroupNorm + SiLU to its output */
  int32_t H; int32_t W; float ms;
} t2p_gemm_record;
int t2p_unet_set_profile(t2p_unet* u, int enable);
int t2p_unet_profile_read(t2p_unet* u, t2p_gemm_record* out, int cap);
int64_t t2p_unet_workspace_bytes(const t2p_unet* u);
int64_t t2p_unet_launches_per_forward(const t2p_unet* u);

/* ---------------------------------------------------------------------------------------------------------
 * Fused predictor / corrector half-steps.  Replace ReverseDiffusionPredictor.update_fn (sampling.py:162-167,
 * with sde_lib.py:96-101 RSDE.discretize and :237-245 VESDE.discretize / :149-157 VPSDE.discretize),
 * LangevinCorrector.update_fn (sampling.py:179-199) and the mask + .float() lines (sampling.py:283-287). */
typedef struct t2p_step_args {
  float* x;                 /* [B][C][N*N] fp32 state, updated in place */
  const void* score;        /* model output: fp32 or fp64, NCHW or NHWC */
  int32_t score_dtype;      /* T2P_F32 | T2P_F64 */
  int32_t score_nhwc;
  const double* sigmas;     /* optional: score = raw / sigmas[labels[b]] (ncsnpp.py:259-261) */
  const int64_t* labels;
  const float* G;           /* [B] predictor: discretised diffusion coefficient */
  const float* sqrt_alpha;  /* [B] predictor, VP only: f = sqrt_alpha * x - x; NULL for VE (f = 0) */
  const float* alpha;       /* [B] corrector, VP only; NULL = 1 */
  int32_t probability_flow;
  float snr;
  const uint8_t* mask;      /* [B][C][N*N] conditional_mask (1 = free), or NULL */
  const float* x_init;      /* x_initial, required with mask */
  float* x_mean_out;        /* optional: masked x_mean as float */
  uint64_t seed;
  int64_t stream_id;        /* Philox stream of this half-step */
  int64_t sample_offset;    /* global index of local sample 0 (batch sharding) */
  int32_t B, C, HW;
  double* workspace;        /* corrector: >= t2p_corrector_workspace_bytes(B, C*HW) bytes */
  int32_t conditioned_in_place; /* caller guarantees that x_out (and x_mean_out) already hold x_init wherever mask == 0
                                 * (true inside a sampling run after the first mask application, sampling.py:283-287
                                 * being idempotent there): fully conditioned quads are then not read or written */
  int32_t symmetrize;       /* 0 (default) = the reference's update, bit for bit.  1: channels 0 and 1 (Cb-Cb distance
                             * and omega, symmetric maps) of the new state and of x_mean are replaced by their symmetric
                             * part 0.5 (u[i][j] + u[j][i]) -- in float64, before the single float rounding, so the
                             * stored maps are exactly symmetric -- wherever (i, j) and (j, i) are both free.  The
                             * reference has no such option (downstream takes np.triu, rosetta_min/utils.py:140,157).
                             * Needs x_out != x and a square map (W * W == HW). */
  float* x_out;             /* where the new state goes; NULL = x (in place) */
  int32_t W;                /* map width (symmetrize only) */
  int32_t reserved;
} t2p_step_args;

int64_t t2p_corrector_workspace_bytes(int B, int64_t elems_per_sample);
int t2p_predictor_step(const t2p_step_args* a, void* stream);
int t2p_corrector_step(const t2p_step_args* a, void* stream);

/* Philox4x32-10 + Box-Muller normals of stream `stream_id`, global elements [first, first+count), times scale.
 * Replaces torch.randn_like / VESDE.prior_sampling (sde_lib.py:229-230).  first, count multiples of 4. */
int t2p_philox_normal(uint64_t seed, int64_t stream_id, int64_t first, int64_t count, float scale, float* out,
                      void* stream);
int t2p_philox_bits(uint64_t seed, int64_t stream_id, int64_t first_quad, int64_t quads, uint32_t* out,
                    void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Whole sampling loop.  Replaces the body of pc_sampler (sampling.py:279-289) for the native score network and
 * the VE SDE: K iterations of corrector (n_steps inner steps) then predictor, graph-captured once and replayed. */
typedef struct t2p_run_args {
  float* x;                  /* [B][C][N][N] fp32: in = masked prior sample, out = final x */
  float* x_mean;             /* [B][C][N][N] fp32 out: masked x_mean of the last predictor step */
  const uint8_t* mask;       /* or NULL */
  const float* x_init;
  const int64_t* label_table;/* [num_iters] host: labels per iteration (models/utils.py:166-169) */
  const float* g_table;      /* [num_iters] host: G per iteration (sde_lib.py:237-245) */
  int32_t num_iters;
  int32_t n_steps;           /* corrector steps per iteration (sampling.n_steps_each) */
  float snr;
  int32_t probability_flow;
  uint64_t seed;
  int64_t sample_offset;
  int32_t B;
  int32_t use_graph;         /* 1: capture one iteration into a CUDA graph and replay it */
  int32_t symmetrize;        /* see t2p_step_args.symmetrize; 0 = the reference's behaviour */
  int32_t reserved;
  struct t2p_peer_group* peers; /* NULL (default): the Langevin step size is this call's batch mean (every shard is an
                             * independent reference run, sampling.py:193-195).  Else: the mean over the GLOBAL batch of
                             * all ranks of the group -- a sharded run then equals one reference run of the whole batch */
} t2p_run_args;

int t2p_pc_run(t2p_unet* u, const t2p_run_args* a, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Opt-in global-batch step size for batch-sharded runs on one node: between the norm phase and
 * the update phase of the fused corrector kernel every rank writes its two norm sums into the mailbox of every
 * peer through NVLink peer memory and reads theirs -- no host round trip, no extra launch, graph-replayable.
 * Set-up: every rank creates a mailbox and exports its CUDA IPC handle (T2P_IPC_HANDLE_BYTES bytes); the handles
 * are exchanged by the host (torch.distributed.all_gather_object in text2protein_b200/distributed.py) and opened.
 * All ranks must then make the same sequence of t2p_pc_run calls with their group. */
#define T2P_IPC_HANDLE_BYTES 64
typedef struct t2p_peer_group t2p_peer_group;
int t2p_peer_mailbox_create(int world, void** mailbox, void* ipc_handle_out);
int t2p_peer_group_open(void* own_mailbox, const void* ipc_handles, int world, int rank, int64_t global_batch,
                        t2p_peer_group** out);
void t2p_peer_group_close(t2p_peer_group* g);

/* ---------------------------------------------------------------------------------------------------------
 * Per-kernel entry points (unit-testable against their torch counterparts).  NHWC / token-major tensors. */
typedef struct t2p_conv_args {
  const void* a0; int32_t c0;      /* NHWC source 0 */
  const void* a1; int32_t c1;      /* optional NHWC source 1 (channel concat), c1 = 0 if unused */
  int32_t B, H, W, ksize;          /* ksize 1 or 3, stride 1, zero padding ksize/2 */
  const void* w; int32_t N;        /* [N][ksize*ksize*(c0+c1)] (taps outer, channels inner) */
  const float* bias;               /* [N] or NULL */
  const float* rowbias;            /* [B][rowbias_ld] per-sample bias or NULL */
  int32_t rowbias_ld;
  const void* residual;            /* [M][N] in out dtype or NULL */
  int32_t res_up;
  float alpha;
  void* out; int32_t out_dtype;
  int32_t in_dtype;                /* T2P_BF16 -> tcgen05 kernel (needs c % 64 == 0); T2P_F32 -> CUDA-core kernel */
  float* stat_part;                /* optional fused GroupNorm statistics, [B*H*W/T][N][2] {sum, sumsq} per
                                      T-pixel tile, T = t2p_conv2d_stat_tile(args) (tcgen05 kernel, bf16 out) */
  const void* x0; int32_t xc0;     /* optional NHWC sources entering through the centre tap only: a 1x1 convolution */
  const void* x1; int32_t xc1;     /* over x0|x1 summed with the one over a0|a1 (ResnetBlockBigGANpp Conv_2 folded into
                                      Conv_1, layers.py:318-327); w rows are [k*k*(c0+c1) | xc0 | xc1].  bf16, N >= 128 */
  const float* gn_scale;           /* optional: a0|a1 are RAW and the kernel feeds silu(x * gn_scale + gn_shift) to the tensor */
  const float* gn_shift;           /* core (h = act(GroupNorm(x)), layers.py:305,318): per-(sample, channel) affine over the
                                      concat, fp32 [B][c0+c1].  Only where t2p_conv2d_fuses_groupnorm(args) != 0 */
  const float* gno_gamma;          /* optional (ABI 3): the launch normalises its OWN output, out = silu(GroupNorm(conv + bias + */
  const float* gno_beta;           /* rowbias)) with the CONSUMER's GroupNorm(gno_groups, N, gno_eps) weight / bias [N] -- h =   */
  int32_t gno_groups;              /* act(GroupNorm_1(Conv_0(.) + temb)), layers.py:314-318, without the raw tensor ever reaching */
  float gno_eps;                   /* HBM.  Only where t2p_conv2d_normalises_output(args) != 0; excludes stat_part / gn_scale    */
} t2p_conv_args;
int t2p_conv2d(const t2p_conv_args* a, void* stream);        /* nn.Conv2d / NIN / nn.Linear: layers.py:82-95,128-137 */
/* Pixel-tile size T of the fused GroupNorm statistics t2p_conv2d would write for these arguments (H*W % T == 0),
 * or 0 when this launch cannot produce them (then stat_part must be NULL).  Host-only, no GPU work. */
int t2p_conv2d_stat_tile(const t2p_conv_args* a);
/* 1 when t2p_conv2d can apply GroupNorm + SiLU to its 3x3 sources itself for these arguments (3x3, 128-pixel-wide
 * images, bf16, N >= 128: the halo kernel), else 0.  Host-only. */
int t2p_conv2d_fuses_groupnorm(const t2p_conv_args* a);
/* 1 when t2p_conv2d can apply GroupNorm(gno_groups) + SiLU to its own output for these arguments (channel-major tcgen05
 * kernel: bf16, N % 128 == 0, groups of 4 / 8 / 16 / 32 channels, whole pixel tiles per sample, no residual, and few enough
 * tiles per sample for the inter-CTA statistics exchange), else 0.  Host-only; reads gno_groups. */
int t2p_conv2d_normalises_output(const t2p_conv_args* a);

/* Last layer, ncsnpp.py:212-216,257: out fp32 NCHW [B][nout][H][W] = Conv3x3(SiLU(x * scale + shift)) + bias, with x the
 * raw bf16 NHWC activation [B][H][W][cin], scale / shift the per-(sample, channel) GroupNorm affine [B][cin] and w the
 * packed bf16 weight [nout][9][cin] (cin in {64, 128, 256}, nout <= 8, W % 16 == 0). */
int t2p_final_conv(const void* x, const float* scale, const float* shift, const void* w, const float* bias, float* out,
                   int B, int H, int W, int cin, int nout, void* stream);

/* nn.GroupNorm(G, C, eps) [+ SiLU] [+ 2x2 mean | nearest x2] over the concat of a0|a1: layers.py:282-311 */
int t2p_groupnorm(const void* a0, int c0, const void* a1, int c1, int B, int H, int W, int dtype, int groups,
                  float eps, const float* gamma, const float* beta, int silu, int resample_mode, void* out,
                  void* raw_out, void* stream);
/* The apply pass alone, given the per-(sample, channel) affine [B][c0+c1]: y = act(x * scale + shift) [+ resampling].
 * This is the HBM-bound kernel the engine launches after the GEMM epilogue has produced the statistics. */
int t2p_groupnorm_apply(const void* a0, int c0, const void* a1, int c1, int B, int H, int W, int dtype, const float* scale,
                        const float* shift, int silu, int resample_mode, void* out, void* raw_out, void* stream);
/* The same GroupNorm [+ SiLU] (no resampling) as ONE launch for small tensors -- H*W <= 1024, (c0+c1) % 32 == 0,
 * c0 % 32 == 0, groups of 4 / 8 / 16 / 32 channels: what the engine runs at 32 x 32 and below (statistics, finalize and
 * apply of the other entry points are three launches there).  Fails for other shapes. */
int t2p_groupnorm_small(const void* a0, int c0, const void* a1, int c1, int B, int H, int W, int dtype, int groups,
                        float eps, const float* gamma, const float* beta, int silu, void* out, void* stream);
int t2p_layernorm(const void* x, const float* gamma, const float* beta, int64_t M, int C, float eps, int dtype,
                  void* y, void* stream);                     /* attention.py:203-205 */
int t2p_geglu(const void* z, int64_t M, int D, int dtype, void* out, void* stream); /* attention.py:42-44 */
/* softmax(scale * q k^T) v on strided token-major views: layers.py:160-176, attention.py:170-193.
 * use_tensor_cores: 0 = CUDA-core kernel (fp32 or bf16), 1 = bf16 mma.sync kernel, 2 = bf16 tcgen05 / TMEM / TMA kernel
 * (what the engine runs for head dims 32, 64, 128 and multiples of 256). */
int t2p_attention(const void* q, const void* k, const void* v, void* out, int B, int heads, int Tq, int Tk, int d,
                  int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, float scale, int dtype, int use_tensor_cores,
                  void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Callers either side of the sampling loop, generated on the device from small integer inputs. */

/* Text context from token ids: rows of llm.model.embed_tokens (sampling_6d.py:134-137) gathered straight into the
 * engine's context buffer, then the hoisted K|V projections (attention.py:174-175).  table: [vocab][context_dim],
 * T2P_F32 or T2P_BF16; tokens: int64 [B][L]. */
int t2p_unet_set_context_tokens(t2p_unet* u, const void* table, int table_dtype, int64_t vocab, const int64_t* tokens,
                                int B, int L, void* stream);
/* Stand-alone gather: out fp32 [n][D] = table[tokens[i]] (what `llm.model.embed_tokens(tokens)` returns). */
int t2p_embed_tokens(const void* table, int table_dtype, int64_t vocab, int D, const int64_t* tokens, int64_t n,
                     float* out, void* stream);
/* out[b][i][j] = (i < lengths[b] && j < lengths[b]) as bytes: the "length" condition, utils.py:89-93,139-148. */
int t2p_length_mask(const int32_t* lengths, int B, int N, uint8_t* out, void* stream);
/* out[b][i][j] = sel(i) | sel(j), sel = union of the inclusive residue ranges [R][2] (shared) or [B][R][2]
 * (per_sample): selected_mask_batch / the outer-OR of random_mask_batch, utils.py:56-58,62-81. */
int t2p_inpaint_mask(const int32_t* ranges, int R, int per_sample, int B, int N, uint8_t* out, void* stream);
/* The sampler's conditional mask [B][C][N][N] (1 = free to evolve), sampling.py:258-281, for the condition keys
 * length (lengths != NULL), ss (has_ss) and inpainting (ranges != NULL). */
int t2p_condition_mask(const int32_t* lengths, const int32_t* ranges, int R, int per_sample, int has_ss, int B, int C,
                       int N, uint8_t* out, void* stream);
/* sampling_rosetta.py:69-96: mask = round(sample[:, -1]) == 1; L_out[b] = sqrt(count) (or -1 when not an integer);
 * out fp32 [B][8][N*N]: per channel the masked entries in raster order (first L*L values, rest 0) of
 * dist, omega, theta, phi clipped to [-1, 1], then dist_abs, omega_abs, theta_abs, phi_abs. */
int t2p_postprocess_6d(const float* sample, int B, int C, int N, float* out, int32_t* L_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* T2P_H_ */
