// extern "C" surface of libt2p.so (include/t2p.h): argument marshalling, error plumbing and the
// graph-captured sampling loop.
#include <cstddef>
#include <cstring>
#include <memory>

#include "../../include/t2p.h"
#include "unet.h"

namespace t2p {
namespace {
thread_local std::string g_last_error;
}
void set_last_error(const std::string& m) { g_last_error = m; }
}  // namespace t2p

struct RunKey {
  const void *x, *x_mean, *mask, *x_init;
  int B, n_steps;
  float snr;
  int pf;
  unsigned long long seed;
  long long sample_offset;
  long long generation;
  // every device buffer the captured iteration points into that the caller does not own: activation arena and
  // time-embedding buffer of the network (resource_epoch), loop buffers of this handle (buf_epoch)
  long long resource_epoch, buf_epoch;
  int symmetrize;
  const void* peers;
  bool operator==(const RunKey& o) const {
    return x == o.x && x_mean == o.x_mean && mask == o.mask && x_init == o.x_init && B == o.B &&
           n_steps == o.n_steps && snr == o.snr && pf == o.pf && seed == o.seed && sample_offset == o.sample_offset &&
           generation == o.generation && resource_epoch == o.resource_epoch && buf_epoch == o.buf_epoch &&
           symmetrize == o.symmetrize && peers == o.peers;
  }
};

struct t2p_peer_group {
  t2p::PeerGroup g;
  void* opened[t2p::PeerGroup::kMaxWorld] = {};  // peer mappings to close (own mailbox excluded)
  void* own = nullptr;
  unsigned long long runs = 0;  // t2p_pc_run calls made with this group (all ranks make the same sequence)
};

struct t2p_unet {
  std::unique_ptr<t2p::UNet> net;
  cudaGraphExec_t graph_exec = nullptr;
  RunKey graph_key{};
  // sampling-loop state (device), sized lazily
  int run_B = 0, run_K = 0;
  long long buf_epoch = 0;
  float* x_alt = nullptr;  // second state buffer of an out-of-place (symmetrised) run
  long long* labels = nullptr;
  float* G = nullptr;
  long long* state = nullptr;
  long long* label_table = nullptr;
  float* g_table = nullptr;
  float* h = nullptr;
  double* partial = nullptr;
  cudaStream_t run_stream = nullptr;  // capture is illegal on the legacy default stream: the loop runs here
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  ~t2p_unet() {
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
    if (run_stream) cudaStreamDestroy(run_stream);
    if (ev_in) cudaEventDestroy(ev_in);
    if (ev_out) cudaEventDestroy(ev_out);
    for (void* p : {static_cast<void*>(labels), static_cast<void*>(G), static_cast<void*>(state),
                    static_cast<void*>(label_table), static_cast<void*>(g_table), static_cast<void*>(h),
                    static_cast<void*>(partial), static_cast<void*>(x_alt)})
      if (p) cudaFree(p);
  }
};

#define T2P_API_BEGIN try {
#define T2P_API_END                         \
  return 0;                                 \
  }                                         \
  catch (const std::exception& e) {         \
    t2p::set_last_error(e.what());          \
    return -1;                              \
  }                                         \
  catch (...) {                             \
    t2p::set_last_error("unknown error");   \
    return -1;                              \
  }

using namespace t2p;

static cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

extern "C" {

const char* t2p_last_error(void) { return g_last_error.c_str(); }
int t2p_abi_version(void) { return T2P_ABI_VERSION; }

#define T2P_OFF(T, f) static_cast<int32_t>(offsetof(T, f))
static std::vector<int32_t> struct_offsets(int which) {
  switch (which) {
    case 0: return {T2P_OFF(t2p_unet_cfg, num_channels), T2P_OFF(t2p_unet_cfg, max_res_num), T2P_OFF(t2p_unet_cfg, nf),
                    T2P_OFF(t2p_unet_cfg, n_ch_mult), T2P_OFF(t2p_unet_cfg, ch_mult), T2P_OFF(t2p_unet_cfg, num_res_blocks),
                    T2P_OFF(t2p_unet_cfg, n_attn_resolutions), T2P_OFF(t2p_unet_cfg, attn_resolutions),
                    T2P_OFF(t2p_unet_cfg, n_heads), T2P_OFF(t2p_unet_cfg, context_dim), T2P_OFF(t2p_unet_cfg, num_scales),
                    T2P_OFF(t2p_unet_cfg, scale_by_sigma), T2P_OFF(t2p_unet_cfg, compute_dtype)};
    case 1: return {T2P_OFF(t2p_step_args, x), T2P_OFF(t2p_step_args, score), T2P_OFF(t2p_step_args, score_dtype),
                    T2P_OFF(t2p_step_args, score_nhwc), T2P_OFF(t2p_step_args, sigmas), T2P_OFF(t2p_step_args, labels),
                    T2P_OFF(t2p_step_args, G), T2P_OFF(t2p_step_args, sqrt_alpha), T2P_OFF(t2p_step_args, alpha),
                    T2P_OFF(t2p_step_args, probability_flow), T2P_OFF(t2p_step_args, snr), T2P_OFF(t2p_step_args, mask),
                    T2P_OFF(t2p_step_args, x_init), T2P_OFF(t2p_step_args, x_mean_out), T2P_OFF(t2p_step_args, seed),
                    T2P_OFF(t2p_step_args, stream_id), T2P_OFF(t2p_step_args, sample_offset), T2P_OFF(t2p_step_args, B),
                    T2P_OFF(t2p_step_args, C), T2P_OFF(t2p_step_args, HW), T2P_OFF(t2p_step_args, workspace),
                    T2P_OFF(t2p_step_args, conditioned_in_place), T2P_OFF(t2p_step_args, symmetrize),
                    T2P_OFF(t2p_step_args, x_out), T2P_OFF(t2p_step_args, W), T2P_OFF(t2p_step_args, reserved)};
    case 2: return {T2P_OFF(t2p_run_args, x), T2P_OFF(t2p_run_args, x_mean), T2P_OFF(t2p_run_args, mask),
                    T2P_OFF(t2p_run_args, x_init), T2P_OFF(t2p_run_args, label_table), T2P_OFF(t2p_run_args, g_table),
                    T2P_OFF(t2p_run_args, num_iters), T2P_OFF(t2p_run_args, n_steps), T2P_OFF(t2p_run_args, snr),
                    T2P_OFF(t2p_run_args, probability_flow), T2P_OFF(t2p_run_args, seed),
                    T2P_OFF(t2p_run_args, sample_offset), T2P_OFF(t2p_run_args, B), T2P_OFF(t2p_run_args, use_graph),
                    T2P_OFF(t2p_run_args, symmetrize), T2P_OFF(t2p_run_args, reserved), T2P_OFF(t2p_run_args, peers)};
    case 3: return {T2P_OFF(t2p_conv_args, a0), T2P_OFF(t2p_conv_args, c0), T2P_OFF(t2p_conv_args, a1),
                    T2P_OFF(t2p_conv_args, c1), T2P_OFF(t2p_conv_args, B), T2P_OFF(t2p_conv_args, H), T2P_OFF(t2p_conv_args, W),
                    T2P_OFF(t2p_conv_args, ksize), T2P_OFF(t2p_conv_args, w), T2P_OFF(t2p_conv_args, N),
                    T2P_OFF(t2p_conv_args, bias), T2P_OFF(t2p_conv_args, rowbias), T2P_OFF(t2p_conv_args, rowbias_ld),
                    T2P_OFF(t2p_conv_args, residual), T2P_OFF(t2p_conv_args, res_up), T2P_OFF(t2p_conv_args, alpha),
                    T2P_OFF(t2p_conv_args, out), T2P_OFF(t2p_conv_args, out_dtype), T2P_OFF(t2p_conv_args, in_dtype),
                    T2P_OFF(t2p_conv_args, stat_part), T2P_OFF(t2p_conv_args, x0), T2P_OFF(t2p_conv_args, xc0),
                    T2P_OFF(t2p_conv_args, x1), T2P_OFF(t2p_conv_args, xc1), T2P_OFF(t2p_conv_args, gn_scale),
                    T2P_OFF(t2p_conv_args, gn_shift), T2P_OFF(t2p_conv_args, gno_gamma), T2P_OFF(t2p_conv_args, gno_beta),
                    T2P_OFF(t2p_conv_args, gno_groups), T2P_OFF(t2p_conv_args, gno_eps)};
    case 4: return {T2P_OFF(t2p_gemm_record, M), T2P_OFF(t2p_gemm_record, N), T2P_OFF(t2p_gemm_record, K),
                    T2P_OFF(t2p_gemm_record, ksize), T2P_OFF(t2p_gemm_record, tensor_core), T2P_OFF(t2p_gemm_record, H),
                    T2P_OFF(t2p_gemm_record, W), T2P_OFF(t2p_gemm_record, ms)};
  }
  return {};
}

int t2p_sizeof(int which) {
  switch (which) {
    case 0: return static_cast<int>(sizeof(t2p_unet_cfg));
    case 1: return static_cast<int>(sizeof(t2p_step_args));
    case 2: return static_cast<int>(sizeof(t2p_run_args));
    case 3: return static_cast<int>(sizeof(t2p_conv_args));
    case 4: return static_cast<int>(sizeof(t2p_gemm_record));
  }
  return -1;
}

int t2p_struct_layout(int which, int32_t* offsets, int cap) {
  const std::vector<int32_t> o = struct_offsets(which);
  if (o.empty()) return -1;
  for (size_t i = 0; i < o.size() && static_cast<int>(i) < cap; ++i) offsets[i] = o[i];
  return static_cast<int>(o.size());
}

int t2p_unet_create(const t2p_unet_cfg* c, t2p_unet** out) {
  T2P_API_BEGIN
  T2P_CHECK(c && out, "null argument");
  UNetConfig cfg;
  cfg.num_channels = c->num_channels;
  cfg.max_res_num = c->max_res_num;
  cfg.nf = c->nf;
  T2P_CHECK(c->n_ch_mult > 0 && c->n_ch_mult <= 16 && c->n_attn_resolutions >= 0 && c->n_attn_resolutions <= 16,
            "bad ch_mult / attn_resolutions length");
  cfg.ch_mult.assign(c->ch_mult, c->ch_mult + c->n_ch_mult);
  cfg.attn_resolutions.assign(c->attn_resolutions, c->attn_resolutions + c->n_attn_resolutions);
  cfg.num_res_blocks = c->num_res_blocks;
  cfg.n_heads = c->n_heads;
  cfg.context_dim = c->context_dim;
  cfg.num_scales = c->num_scales;
  cfg.scale_by_sigma = c->scale_by_sigma;
  cfg.compute_dtype = c->compute_dtype;
  auto u = std::make_unique<t2p_unet>();
  u->net = std::make_unique<UNet>(cfg);
  *out = u.release();
  T2P_API_END
}

void t2p_unet_destroy(t2p_unet* u) { delete u; }

int t2p_unet_num_params(const t2p_unet* u) { return u ? static_cast<int>(u->net->params().size()) : -1; }

int t2p_unet_param_info(const t2p_unet* u, int index, char* name_buf, int name_cap, int64_t* shape4, int* ndim,
                        int* dtype) {
  T2P_API_BEGIN
  T2P_CHECK(u && index >= 0 && index < static_cast<int>(u->net->params().size()), "bad parameter index");
  const Param& p = *u->net->params()[index];
  T2P_CHECK(static_cast<int>(p.name.size()) < name_cap, "name buffer too small");
  std::strcpy(name_buf, p.name.c_str());
  *ndim = static_cast<int>(p.shape.size());
  for (size_t i = 0; i < p.shape.size() && i < 4; ++i) shape4[i] = p.shape[i];
  *dtype = p.dtype;
  T2P_API_END
}

int t2p_unet_load(t2p_unet* u, const char* name, const void* dev_ptr, const int64_t* shape, int ndim, int dtype,
                  void* stream) {
  T2P_API_BEGIN
  T2P_CHECK(u && name && dev_ptr, "null argument");
  u->net->load(name, dev_ptr, std::vector<int64_t>(shape, shape + ndim), dtype, S(stream));
  T2P_API_END
}

int t2p_unet_finalize(t2p_unet* u, void* stream) {
  T2P_API_BEGIN
  u->net->finalize(S(stream));
  T2P_API_END
}

int t2p_unet_set_context(t2p_unet* u, const float* ctx, int B, int L, void* stream) {
  T2P_API_BEGIN
  T2P_CHECK(u && ctx && B > 0 && L > 0, "bad context");
  u->net->set_context(ctx, B, L, S(stream));
  T2P_API_END
}

int t2p_unet_set_context_tokens(t2p_unet* u, const void* table, int table_dtype, int64_t vocab, const int64_t* tokens,
                                int B, int L, void* stream) {
  T2P_API_BEGIN
  T2P_CHECK(u && table && tokens && B > 0 && L > 0 && vocab > 0, "bad token context");
  u->net->set_context_tokens(table, table_dtype, vocab, reinterpret_cast<const long long*>(tokens), B, L, S(stream));
  T2P_API_END
}

int t2p_embed_tokens(const void* table, int table_dtype, int64_t vocab, int D, const int64_t* tokens, int64_t n,
                     float* out, void* stream) {
  T2P_API_BEGIN
  T2P_CHECK(table && tokens && out && n > 0, "null argument");
  embed_gather(table, table_dtype, vocab, D, reinterpret_cast<const long long*>(tokens), n, out, nullptr, S(stream));
  T2P_API_END
}

int t2p_length_mask(const int32_t* lengths, int B, int N, uint8_t* out, void* stream) {
  T2P_API_BEGIN
  T2P_CHECK(lengths && out && B > 0 && N > 0, "null argument");
  length_mask(lengths, B, N, out, S(stream));
  T2P_API_END
}

int t2p_inpaint_mask(const int32_t* ranges, int R, int per_sample, int B, int N, uint8_t* out, void* stream) {
  T2P_API_BEGIN
  T2P_CHECK(out && B > 0 && N > 0 && R >= 0 && (R == 0 || ranges), "null argument");
  inpaint_mask(ranges, R, per_sample, B, N, out, S(stream));
  T2P_API_END
}

int t2p_condition_mask(const int32_t* lengths, const int32_t* ranges, int R, int per_sample, int has_ss, int B, int C,
                       int N, uint8_t* out, void* stream) {
  T2P_API_BEGIN
  T2P_CHECK(out && B > 0 && C > 0 && N > 0, "null argument");
  condition_mask(lengths, ranges, R, per_sample, has_ss, B, C, N, out, S(stream));
  T2P_API_END
}

int t2p_postprocess_6d(const float* sample, int B, int C, int N, float* out, int32_t* L_out, void* stream) {
  T2P_API_BEGIN
  T2P_CHECK(sample && out && L_out && B > 0 && N > 0, "null argument");
  postprocess_6d(sample, B, C, N, out, L_out, S(stream));
  T2P_API_END
}

int t2p_unet_forward(t2p_unet* u, const float* x, const int64_t* labels, void* out, int out_dtype, int B,
                     void* stream) {
  T2P_API_BEGIN
  T2P_CHECK(u && x && labels && out && B > 0, "bad forward arguments");
  u->net->forward(x, reinterpret_cast<const long long*>(labels), out, out_dtype, B, S(stream));
  T2P_API_END
}

int t2p_unet_forward_t(t2p_unet* u, const float* x, const int64_t* labels, const float* timesteps, void* out,
                       int out_dtype, int B, void* stream) {
  T2P_API_BEGIN
  T2P_CHECK(u && x && labels && out && B > 0, "bad forward arguments");
  u->net->forward(x, reinterpret_cast<const long long*>(labels), out, out_dtype, B, S(stream), timesteps);
  T2P_API_END
}

int t2p_unet_set_epilogue_groupnorm(t2p_unet* u, int enable) {
  T2P_API_BEGIN
  T2P_CHECK(u != nullptr, "null handle");
  u->net->set_epilogue_groupnorm(enable != 0);
  T2P_API_END
}

int t2p_unet_set_fused_groupnorm(t2p_unet* u, int enable) {
  T2P_API_BEGIN
  T2P_CHECK(u != nullptr, "null handle");
  u->net->set_fused_groupnorm(enable != 0);
  T2P_API_END
}

int t2p_unet_set_debug(t2p_unet* u, int enable) {
  T2P_API_BEGIN
  u->net->set_debug(enable != 0);
  T2P_API_END
}

int t2p_unet_tap(t2p_unet* u, const char* name, float* dst, int64_t capacity, int64_t* shape4, void* stream) {
  T2P_API_BEGIN
  T2P_CHECK(u->net->tap(name, dst, capacity, shape4, S(stream)), std::string("no tap named '") + name + "'");
  T2P_API_END
}

int t2p_unet_set_profile(t2p_unet* u, int enable) {
  T2P_API_BEGIN
  u->net->set_profile(enable != 0);
  T2P_API_END
}

int t2p_unet_profile_read(t2p_unet* u, t2p_gemm_record* out, int cap) {
  try {
    std::vector<GemmRecord> tmp(cap > 0 ? cap : 0);
    const int n = u->net->profile_records(tmp.data(), cap);
    for (int i = 0; i < n && i < cap; ++i) {
      out[i].M = tmp[i].M; out[i].N = tmp[i].N; out[i].K = tmp[i].K; out[i].ksize = tmp[i].ksize;
      out[i].tensor_core = tmp[i].tc; out[i].H = tmp[i].H; out[i].W = tmp[i].W; out[i].ms = tmp[i].ms;
    }
    return n;
  } catch (const std::exception& e) {
    t2p::set_last_error(e.what());
    return -1;
  }
}

int64_t t2p_unet_workspace_bytes(const t2p_unet* u) { return static_cast<int64_t>(u->net->workspace_bytes()); }
int64_t t2p_unet_launches_per_forward(const t2p_unet* u) { return u->net->launches_per_forward(); }

// ---------------------------------------------------------------------------------------------- steps
static PcStepArgs step_args(const t2p_step_args* a) {
  PcStepArgs p;
  p.x = a->x; p.score = a->score; p.score_dtype = a->score_dtype; p.score_nhwc = a->score_nhwc;
  p.sigmas = a->sigmas; p.labels = reinterpret_cast<const long long*>(a->labels);
  p.G = a->G; p.sqrt_alpha = a->sqrt_alpha; p.alpha = a->alpha;
  p.probability_flow = a->probability_flow; p.snr = a->snr;
  p.mask = a->mask; p.x_init = a->x_init; p.x_mean_out = a->x_mean_out;
  p.seed = a->seed; p.stream_base = a->stream_id; p.stream_mul = 0; p.iter_ptr = nullptr;
  p.sample_offset = a->sample_offset;
  p.B = a->B; p.C = a->C; p.HW = a->HW;
  p.conditioned_in_place = a->conditioned_in_place;
  p.x_out = a->x_out;
  p.symmetrize = a->symmetrize;
  p.W = a->W;
  return p;
}

int64_t t2p_corrector_workspace_bytes(int B, int64_t elems_per_sample) {
  return static_cast<int64_t>(sizeof(double)) * pc_corrector_workspace_doubles(B, elems_per_sample);
}

int t2p_predictor_step(const t2p_step_args* a, void* stream) {
  T2P_API_BEGIN
  T2P_CHECK(a && a->x && a->score, "null argument");
  pc_predictor_step(step_args(a), S(stream));
  T2P_API_END
}

int t2p_corrector_step(const t2p_step_args* a, void* stream) {
  T2P_API_BEGIN
  T2P_CHECK(a && a->x && a->score && a->workspace, "null argument");
  PcStepArgs p = step_args(a);
  p.partial = a->workspace;
  pc_corrector_step(p, S(stream));
  T2P_API_END
}

int t2p_philox_normal(uint64_t seed, int64_t stream_id, int64_t first, int64_t count, float scale, float* out,
                      void* stream) {
  T2P_API_BEGIN
  philox_normal_fill(seed, static_cast<unsigned long long>(stream_id), first, count, scale, out, S(stream));
  T2P_API_END
}

int t2p_philox_bits(uint64_t seed, int64_t stream_id, int64_t first_quad, int64_t quads, uint32_t* out,
                    void* stream) {
  T2P_API_BEGIN
  philox_bits_fill(seed, static_cast<unsigned long long>(stream_id), first_quad, quads, out, S(stream));
  T2P_API_END
}

// ---------------------------------------------------------------------------------------------- run
// Device buffers of the loop.  Whenever one of them moves, buf_epoch changes and with it the graph key: a captured
// iteration holds raw pointers into them (the cache used to be keyed on the caller's buffers only, so a run with a
// larger K could replay a graph pointing at freed label / G tables).
static void ensure_run_buffers(t2p_unet* u, int B, int K, bool need_alt) {
  const UNetConfig& c = u->net->cfg();
  const long long E = static_cast<long long>(c.num_channels) * c.max_res_num * c.max_res_num;
  if (u->run_B != B) {
    ++u->buf_epoch;
    if (u->x_alt) { T2P_CUDA(cudaFree(u->x_alt)); u->x_alt = nullptr; }
    for (void** p : {reinterpret_cast<void**>(&u->labels), reinterpret_cast<void**>(&u->G),
                     reinterpret_cast<void**>(&u->state), reinterpret_cast<void**>(&u->h),
                     reinterpret_cast<void**>(&u->partial)}) {
      if (*p) T2P_CUDA(cudaFree(*p));
      *p = nullptr;
    }
    T2P_CUDA(cudaMalloc(reinterpret_cast<void**>(&u->labels), sizeof(long long) * B));
    T2P_CUDA(cudaMalloc(reinterpret_cast<void**>(&u->G), sizeof(float) * B));
    T2P_CUDA(cudaMalloc(reinterpret_cast<void**>(&u->state), sizeof(long long) * 4));
    T2P_CUDA(cudaMalloc(reinterpret_cast<void**>(&u->h), sizeof(float) * B * E));
    T2P_CUDA(cudaMalloc(reinterpret_cast<void**>(&u->partial), sizeof(double) * pc_corrector_workspace_doubles(B, E)));
    u->run_B = B;
  }
  if (need_alt && !u->x_alt) {
    ++u->buf_epoch;
    T2P_CUDA(cudaMalloc(reinterpret_cast<void**>(&u->x_alt), sizeof(float) * B * E));
  }
  if (u->run_K < K) {
    ++u->buf_epoch;
    const int cap = std::max(K, c.num_scales);  // one allocation covers every run of the model's own schedule
    if (u->label_table) T2P_CUDA(cudaFree(u->label_table));
    if (u->g_table) T2P_CUDA(cudaFree(u->g_table));
    u->label_table = nullptr;
    u->g_table = nullptr;
    u->run_K = 0;
    T2P_CUDA(cudaMalloc(reinterpret_cast<void**>(&u->label_table), sizeof(long long) * cap));
    T2P_CUDA(cudaMalloc(reinterpret_cast<void**>(&u->g_table), sizeof(float) * cap));
    u->run_K = cap;
  }
}

int t2p_pc_run(t2p_unet* u, const t2p_run_args* a, void* stream) {
  T2P_API_BEGIN
  T2P_CHECK(u && a && a->x && a->x_mean && a->label_table && a->g_table, "null argument");
  T2P_CHECK(a->num_iters > 0 && a->n_steps >= 0 && a->B > 0, "bad run arguments");
  cudaStream_t user = S(stream);
  if (!u->run_stream) {
    T2P_CUDA(cudaStreamCreateWithFlags(&u->run_stream, cudaStreamNonBlocking));
    T2P_CUDA(cudaEventCreateWithFlags(&u->ev_in, cudaEventDisableTiming));
    T2P_CUDA(cudaEventCreateWithFlags(&u->ev_out, cudaEventDisableTiming));
  }
  // order the loop after everything already queued on the caller's stream, and the caller after the loop
  cudaStream_t st = u->run_stream;
  T2P_CUDA(cudaEventRecord(u->ev_in, user));
  T2P_CUDA(cudaStreamWaitEvent(st, u->ev_in, 0));
  struct Rejoin {
    t2p_unet* u; cudaStream_t user;
    ~Rejoin() {
      u->net->set_reuse_temb(false);  // also on the error path
      u->net->set_uniform_labels(false);
      if (cudaEventRecord(u->ev_out, u->run_stream) == cudaSuccess) cudaStreamWaitEvent(user, u->ev_out, 0);
    }
  } rejoin{u, user};
  UNet& net = *u->net;
  net.set_uniform_labels(true);  // run_prep gives every sample of an iteration the same label
  const UNetConfig& c = net.cfg();
  const int B = a->B, K = a->num_iters;
  const int HW = c.max_res_num * c.max_res_num;
  const bool sym = a->symmetrize != 0;
  ensure_run_buffers(u, B, K, sym);
  T2P_CUDA(cudaMemcpyAsync(u->label_table, a->label_table, sizeof(long long) * K, cudaMemcpyHostToDevice, st));
  T2P_CUDA(cudaMemcpyAsync(u->g_table, a->g_table, sizeof(float) * K, cudaMemcpyHostToDevice, st));
  // state: [0] iteration, [1] next iteration, [2] last iteration of THIS run (x_mean is stored there only),
  // [3] first mailbox tag of this run.  Uploaded from a pageable host array: the copy is staged before the call returns.
  t2p_peer_group* pg = a->peers;
  if (pg) T2P_CHECK(pg->g.world > 1 && pg->g.global_batch >= B, "bad peer group");
  const long long state0[4] = {0, 0, K - 1, pg ? static_cast<long long>(++pg->runs << 32) : 0};
  T2P_CUDA(cudaMemcpyAsync(u->state, state0, sizeof(state0), cudaMemcpyHostToDevice, st));

  PcStepArgs base;
  base.x = a->x; base.score = u->h; base.score_dtype = kF32; base.score_nhwc = 0;
  base.sigmas = c.scale_by_sigma ? net.sigmas() : nullptr;
  base.labels = u->labels; base.G = u->G;
  base.probability_flow = a->probability_flow; base.snr = a->snr;
  base.mask = a->mask; base.x_init = a->x_init;
  base.seed = a->seed; base.stream_mul = a->n_steps + 1; base.iter_ptr = u->state;
  base.sample_offset = a->sample_offset;
  base.B = B; base.C = c.num_channels; base.HW = HW;
  base.partial = u->partial;
  base.last_iter_ptr = u->state + 2;
  base.symmetrize = sym ? 1 : 0;
  base.W = c.max_res_num;
  if (pg) { base.peers = &pg->g; base.tag_base_ptr = u->state + 3; }
  if (a->mask) {
    // sampling.py:283-287 re-applies the condition after every half-step; x and x_mean take x_initial at the
    // conditioned positions ONCE here and the step kernels leave those positions alone for the rest of the run
    const long long n = static_cast<long long>(B) * c.num_channels * HW;
    apply_mask(a->x, a->mask, a->x_init, n, st);
    apply_mask(a->x_mean, a->mask, a->x_init, n, st);
    base.conditioned_in_place = 1;
  }
  // Out-of-place (symmetrised) runs ping-pong between the caller's x and x_alt: the update of (i, j) reads the OLD
  // state at (j, i).  Both buffers start conditioned; an iteration with an odd number of half-steps copies back.
  float* cur = a->x;
  float* alt = sym ? u->x_alt : nullptr;
  if (sym)
    T2P_CUDA(cudaMemcpyAsync(alt, a->x, sizeof(float) * static_cast<size_t>(B) * c.num_channels * HW,
                             cudaMemcpyDeviceToDevice, st));

  auto iteration = [&]() {
    run_prep(u->state, u->label_table, u->g_table, B, u->labels, u->G, st);
    // every score evaluation of one iteration is at the same noise level: the time-embedding path (pre_blocks
    // MLP + 42 Dense_0 projections) is computed by the first one and reused by the rest
    net.set_reuse_temb(false);
    float* xin = cur;
    float* xout = sym ? alt : cur;
    for (int j = 0; j < a->n_steps; ++j) {  // Langevin corrector, sampling.py:188-197
      net.forward_raw(xin, u->labels, u->h, B, st);
      net.set_reuse_temb(true);
      PcStepArgs s = base;
      s.x = xin; s.x_out = xout;
      s.stream_base = 1 + j;
      if (a->n_steps > 1 && a->mask) {
        // the reference re-applies the condition after the WHOLE corrector update (sampling.py:282-283), not
        // between its inner steps: the inner steps run unmasked (conditioned positions drift and are seen by the
        // next score evaluation), the last one restores x_initial everywhere it is conditioned
        if (j + 1 < a->n_steps) { s.mask = nullptr; s.x_init = nullptr; }
        s.conditioned_in_place = 0;
      }
      pc_corrector_step(s, st);
      std::swap(xin, xout);
      if (!sym) xout = xin;
    }
    net.forward_raw(xin, u->labels, u->h, B, st);  // reverse-diffusion predictor, sampling.py:162-167
    net.set_reuse_temb(false);
    PcStepArgs s = base;
    s.x = xin; s.x_out = xout;
    s.stream_base = 1 + a->n_steps;
    s.x_mean_out = a->x_mean;
    pc_predictor_step(s, st);
    if (sym && xout != cur)  // odd number of half-steps: bring the state back to the caller's buffer
      T2P_CUDA(cudaMemcpyAsync(cur, xout, sizeof(float) * static_cast<size_t>(B) * c.num_channels * HW,
                               cudaMemcpyDeviceToDevice, st));
  };

  if (!a->use_graph) {
    for (int i = 0; i < K; ++i) iteration();
  } else {
    // The captured iteration is cached: a later run with the same buffers, batch, seed and weights replays
    // it for all K iterations.  On a miss, iteration 0 runs eagerly (sizes the workspace, builds TMA
    // descriptors, sets kernel attributes outside the capture), then iterations 1..K-1 replay the capture.
    RunKey key{a->x, a->x_mean, a->mask, a->x_init, a->B, a->n_steps, a->snr, a->probability_flow, a->seed,
               a->sample_offset, net.generation(), 0, u->buf_epoch, a->symmetrize, pg};
    int first = 0;
    // (the network's resource epoch is compared AFTER the eager iteration of a miss would have grown its buffers;
    // on a hit nothing can have moved since the capture, or the epochs would differ)
    key.resource_epoch = net.resource_epoch();
    if (!u->graph_exec || !(key == u->graph_key)) {
      if (u->graph_exec) { cudaGraphExecDestroy(u->graph_exec); u->graph_exec = nullptr; }
      iteration();
      first = 1;
      key.resource_epoch = net.resource_epoch();
      if (K > 1) {
        cudaGraph_t graph = nullptr;
        T2P_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        try {
          iteration();
        } catch (...) {
          cudaStreamEndCapture(st, &graph);
          if (graph) cudaGraphDestroy(graph);
          throw;
        }
        T2P_CUDA(cudaStreamEndCapture(st, &graph));
        cudaError_t e = cudaGraphInstantiate(&u->graph_exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) { u->graph_exec = nullptr; T2P_CUDA(e); }
        u->graph_key = key;
      }
    }
    for (int i = first; i < K; ++i) T2P_CUDA(cudaGraphLaunch(u->graph_exec, st));
  }
  T2P_API_END
}

// ---------------------------------------------------------------------------------------------- peers
int t2p_peer_mailbox_create(int world, void** mailbox, void* ipc_handle_out) {
  T2P_API_BEGIN
  T2P_CHECK(mailbox && ipc_handle_out && world > 1 && world <= PeerGroup::kMaxWorld, "bad mailbox arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == T2P_IPC_HANDLE_BYTES, "IPC handle size");
  void* p = nullptr;
  const size_t bytes = sizeof(unsigned long long) * 4 * 2 * PeerGroup::kMaxWorld;
  T2P_CUDA(cudaMalloc(&p, bytes));
  T2P_CUDA(cudaMemset(p, 0, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); T2P_CUDA(e); }
  std::memcpy(ipc_handle_out, &h, sizeof(h));
  *mailbox = p;
  T2P_API_END
}

int t2p_peer_group_open(void* own_mailbox, const void* ipc_handles, int world, int rank, int64_t global_batch,
                        t2p_peer_group** out) {
  T2P_API_BEGIN
  T2P_CHECK(own_mailbox && ipc_handles && out && world > 1 && world <= PeerGroup::kMaxWorld && rank >= 0 && rank < world &&
            global_batch > 0, "bad peer group arguments");
  auto g = std::make_unique<t2p_peer_group>();
  g->g.world = world;
  g->g.rank = rank;
  g->g.global_batch = global_batch;
  g->own = own_mailbox;
  for (int r = 0; r < world; ++r) {
    if (r == rank) { g->g.box[r] = static_cast<unsigned long long*>(own_mailbox); continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, static_cast<const char*>(ipc_handles) + static_cast<size_t>(r) * sizeof(h), sizeof(h));
    void* p = nullptr;
    T2P_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    g->opened[r] = p;
    g->g.box[r] = static_cast<unsigned long long*>(p);
  }
  *out = g.release();
  T2P_API_END
}

void t2p_peer_group_close(t2p_peer_group* g) {
  if (!g) return;
  for (void* p : g->opened)
    if (p) cudaIpcCloseMemHandle(p);
  if (g->own) cudaFree(g->own);
  delete g;
}

// ---------------------------------------------------------------------------------------------- per-op
static ConvGemmArgs conv_args_from_abi(const t2p_conv_args* a) {
  ConvGemmArgs g;
  g.a0 = a->a0; g.c0 = a->c0; g.a1 = a->a1; g.c1 = a->c1;
  g.B = a->B; g.H = a->H; g.W = a->W; g.ksize = a->ksize; g.w = a->w; g.N = a->N;
  g.bias = a->bias; g.rowbias = a->rowbias; g.rowbias_ld = a->rowbias_ld; g.rows_per_sample = a->H * a->W;
  g.residual = a->residual; g.res_up = a->res_up; g.alpha = a->alpha;
  g.out = a->out; g.out_dtype = a->out_dtype;
  g.stat_part = a->stat_part;
  g.x0 = a->x0; g.xc0 = a->xc0; g.x1 = a->x1; g.xc1 = a->xc1;
  g.gn_scale = a->gn_scale; g.gn_shift = a->gn_shift;
  return g;
}

int t2p_conv2d_normalises_output(const t2p_conv_args* a) {
  try {
    if (!a || a->in_dtype != T2P_BF16 || a->c0 % 64 != 0 || a->c1 % 64 != 0 || a->stat_part || a->gn_scale) return 0;
    return conv_gemm_tc_gn_out_ok(conv_args_from_abi(a), a->gno_groups) ? 1 : 0;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return 0;
  }
}

int t2p_conv2d_fuses_groupnorm(const t2p_conv_args* a) {
  try {
    if (!a || a->in_dtype != T2P_BF16 || a->c0 % 64 != 0 || a->c1 % 64 != 0 || a->ksize != 3) return 0;
    return conv_gemm_tc_fuses_gn(conv_args_from_abi(a)) ? 1 : 0;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return 0;
  }
}

int t2p_conv2d_stat_tile(const t2p_conv_args* a) {
  try {
    if (!a || a->in_dtype != T2P_BF16 || a->c0 % 64 != 0 || a->c1 % 64 != 0) return 0;
    return conv_gemm_tc_stat_tile(conv_args_from_abi(a));
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return 0;
  }
}

int t2p_conv2d(const t2p_conv_args* a, void* stream) {
  T2P_API_BEGIN
  T2P_CHECK(a && a->a0 && a->w && a->out, "null argument");
  ConvGemmArgs g = conv_args_from_abi(a);
  if (a->xc0 > 0 || a->xc1 > 0)
    T2P_CHECK(a->in_dtype == T2P_BF16 && a->c0 % 64 == 0 && a->c1 % 64 == 0, "centre-tap sources are tcgen05-only");
  if (a->gn_scale || a->gn_shift)
    T2P_CHECK(a->gn_scale && a->gn_shift && t2p_conv2d_fuses_groupnorm(a), "this launch cannot fuse GroupNorm");
  if (a->gno_gamma || a->gno_beta) {
    // the launch normalises its own output: scratch for the statistics exchange lives for this call only (the engine
    // keeps its own, unet.cu)
    T2P_CHECK(a->gno_gamma && a->gno_beta && t2p_conv2d_normalises_output(a), "this launch cannot normalise its output");
    g.gno_gamma = a->gno_gamma; g.gno_beta = a->gno_beta; g.gno_groups = a->gno_groups; g.gno_eps = a->gno_eps;
    const size_t part_bytes = sizeof(float) * static_cast<size_t>(conv_gemm_tc_gn_out_part_floats(g, a->gno_groups));
    const size_t flag_bytes = sizeof(int) * static_cast<size_t>(conv_gemm_tc_gn_out_flag_ints(g));
    char* scratch = nullptr;
    T2P_CUDA(cudaMalloc(reinterpret_cast<void**>(&scratch), part_bytes + flag_bytes));
    g.gno_part = scratch;
    g.gno_flags = reinterpret_cast<int*>(scratch + part_bytes);
    cudaError_t e = cudaMemsetAsync(g.gno_part, 0xff, part_bytes, S(stream));
    if (e == cudaSuccess) e = cudaMemsetAsync(g.gno_flags, 0, flag_bytes, S(stream));
    try {
      T2P_CUDA(e);
      conv_gemm_tc(g, S(stream));
      T2P_CUDA(cudaStreamSynchronize(S(stream)));
    } catch (...) {
      cudaFree(scratch);
      throw;
    }
    T2P_CUDA(cudaFree(scratch));
    return 0;
  }
  if (a->in_dtype == T2P_BF16 && a->c0 % 64 == 0 && a->c1 % 64 == 0) {
    if (conv_gemm_tc_splits(g) > 1) {
      // a launch of few tiles splits K over the idle SMs: partial accumulators + arrival counters for this call only
      // (the engine keeps its own, unet.cu)
      const size_t part_bytes = sizeof(float) * static_cast<size_t>(kSplitKPartFloats);
      const size_t ticket_bytes = sizeof(int) * static_cast<size_t>(kSplitKTicketInts);
      char* scratch = nullptr;
      T2P_CUDA(cudaMalloc(reinterpret_cast<void**>(&scratch), part_bytes + ticket_bytes));
      g.sk_part = reinterpret_cast<float*>(scratch);
      g.sk_ticket = reinterpret_cast<int*>(scratch + part_bytes);
      cudaError_t e = cudaMemsetAsync(g.sk_ticket, 0, ticket_bytes, S(stream));
      try {
        T2P_CUDA(e);
        conv_gemm_tc(g, S(stream));
        T2P_CUDA(cudaStreamSynchronize(S(stream)));
      } catch (...) {
        cudaFree(scratch);
        throw;
      }
      T2P_CUDA(cudaFree(scratch));
      return 0;
    }
    conv_gemm_tc(g, S(stream));
  } else {
    conv_gemm_simt(g, a->in_dtype, S(stream));
  }
  T2P_API_END
}

int t2p_final_conv(const void* x, const float* scale, const float* shift, const void* w, const float* bias, float* out,
                   int B, int H, int W, int cin, int nout, void* stream) {
  T2P_API_BEGIN
  T2P_CHECK(x && scale && shift && w && bias && out, "null argument");
  final_conv_fused(x, scale, shift, w, bias, out, B, H, W, cin, nout, S(stream));
  T2P_API_END
}

int t2p_groupnorm_apply(const void* a0, int c0, const void* a1, int c1, int B, int H, int W, int dtype, const float* scale,
                        const float* shift, int silu, int resample_mode, void* out, void* raw_out, void* stream) {
  T2P_API_BEGIN
  T2P_CHECK(a0 && scale && shift && out, "null argument");
  gn_apply(a0, c0, a1, c1, B, H, W, dtype, scale, shift, silu, resample_mode, out, raw_out, S(stream));
  T2P_API_END
}

int t2p_groupnorm(const void* a0, int c0, const void* a1, int c1, int B, int H, int W, int dtype, int groups,
                  float eps, const float* gamma, const float* beta, int silu, int resample_mode, void* out,
                  void* raw_out, void* stream) {
  T2P_API_BEGIN
  const int C = c0 + c1;
  float* sums = nullptr;
  float* affine = nullptr;
  const int nblk = gn_stats_blocks(B, H * W);
  T2P_CUDA(cudaMalloc(reinterpret_cast<void**>(&sums), sizeof(float) * 2 * B * C * nblk));
  T2P_CUDA(cudaMalloc(reinterpret_cast<void**>(&affine), sizeof(float) * 2 * B * C));
  try {
    gn_stats(a0, c0, a1, c1, B, H * W, dtype, sums, S(stream));
    gn_finalize(sums, nblk, C, nullptr, 0, 0, gamma, beta, B, groups, H * W, eps, affine, affine + static_cast<size_t>(B) * C, S(stream));
    gn_apply(a0, c0, a1, c1, B, H, W, dtype, affine, affine + static_cast<size_t>(B) * C, silu, resample_mode, out,
             raw_out, S(stream));
    T2P_CUDA(cudaStreamSynchronize(S(stream)));
  } catch (...) {
    cudaFree(sums);
    cudaFree(affine);
    throw;
  }
  cudaFree(sums);
  cudaFree(affine);
  T2P_API_END
}

int t2p_groupnorm_small(const void* a0, int c0, const void* a1, int c1, int B, int H, int W, int dtype, int groups,
                        float eps, const float* gamma, const float* beta, int silu, void* out, void* stream) {
  T2P_API_BEGIN
  T2P_CHECK(a0 && gamma && beta && out && B > 0, "null argument");
  gn_small(a0, c0, a1, c1, B, H * W, dtype, groups, eps, gamma, beta, silu, out, S(stream));
  T2P_API_END
}

int t2p_layernorm(const void* x, const float* gamma, const float* beta, int64_t M, int C, float eps, int dtype,
                  void* y, void* stream) {
  T2P_API_BEGIN
  layernorm(x, gamma, beta, M, C, eps, dtype, y, S(stream));
  T2P_API_END
}

int t2p_geglu(const void* z, int64_t M, int D, int dtype, void* out, void* stream) {
  T2P_API_BEGIN
  geglu(z, M, D, dtype, out, S(stream));
  T2P_API_END
}

int t2p_attention(const void* q, const void* k, const void* v, void* out, int B, int heads, int Tq, int Tk, int d,
                  int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, float scale, int dtype, int use_tensor_cores,
                  void* stream) {
  T2P_API_BEGIN
  AttnArgs a;
  a.q = q; a.k = k; a.v = v; a.out = out;
  a.B = B; a.heads = heads; a.Tq = Tq; a.Tk = Tk; a.d = d;
  a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo; a.scale = scale;
  if (use_tensor_cores == 2) {
    T2P_CHECK(dtype == T2P_BF16 && attention_tc_supported(a), "tcgen05 attention needs bf16 and a supported head dim");
    attention_tc(a, S(stream));
  } else if (use_tensor_cores) {
    T2P_CHECK(dtype == T2P_BF16 && attention_mma_supported(a), "tensor-core attention needs bf16 and a supported head dim");
    attention_mma(a, S(stream));
  } else {
    attention_simt(a, dtype, S(stream));
  }
  T2P_API_END
}

}  // extern "C"
