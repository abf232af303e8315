// Native score-network engine: builds the module tree of the reference UNetModel from the config
// (score_sde_pytorch/models/ncsnpp.py:74-217), owns kernel-layout copies of the weights and runs the
// forward pass (ncsnpp.py:220-263) as a fixed sequence of t2p kernels on one CUDA stream.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "kernels.h"

namespace t2p {

struct UNetConfig {
  int num_channels = 5;
  int max_res_num = 128;
  int nf = 128;
  std::vector<int> ch_mult;
  int num_res_blocks = 2;
  std::vector<int> attn_resolutions;
  int n_heads = 8;
  int context_dim = 4096;
  int num_scales = 2000;
  int scale_by_sigma = 1;
  int compute_dtype = kBF16;  // kBF16 (tcgen05 path) or kF32 (verification path)
};

struct Param {
  std::string name;
  std::vector<int64_t> shape;
  int dtype = kF32;       // kF32 for parameters, kF64 for the sigmas buffer
  void* data = nullptr;   // device master copy in `dtype`
  bool loaded = false;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

// Deterministic first-fit pool over one device arena: the alloc / free sequence of a forward pass is a
// pure function of the batch size, so replays (and CUDA-graph captures) see identical addresses.
class Workspace {
 public:
  ~Workspace();
  void begin(bool dry);
  void* alloc(size_t bytes);
  void free(void* p);
  size_t peak() const { return peak_; }
  void reserve(size_t bytes);
  size_t capacity() const { return cap_; }
  long long epoch() const { return epoch_; }  // bumped whenever the arena is re-allocated

 private:
  struct Block { size_t off, size; bool used; };
  std::vector<Block> blocks_;
  char* base_ = nullptr;
  size_t cap_ = 0, top_ = 0, peak_ = 0;
  long long epoch_ = 0;
  bool dry_ = false;
};

struct Act {  // NHWC activation in the compute dtype (+ optional producer-side GroupNorm statistics)
  void* p = nullptr;
  int B = 0, H = 0, W = 0, C = 0;
  float* spart = nullptr;  // [B][snblk][C][2] per-tile {sum, sum of squares} left by the producer GEMM
  int snblk = 0;
  bool want_stats = false;  // gemm() may attach spart when the producing launch can emit statistics
  long long rows() const { return static_cast<long long>(B) * H * W; }
};

struct Linear {  // packed [N][K] weight in compute dtype (or fp32 when force_f32), fp32 bias
  Param* w = nullptr;
  Param* b = nullptr;
  std::vector<Param*> w_cat;  // several source matrices concatenated along N (fused projections)
  std::vector<Param*> b_cat;
  bool nin = false;           // sources are NIN.W ([in, out]) and need a transpose
  int ksize = 1, cin = 0, N = 0;
  bool force_f32 = false;
  // skip path folded into this GEMM (channel-major tcgen05 kernel): xk extra K columns holding the 1x1 skip
  // convolution's weight (w_x, bias b_x added to the bias) or, for an identity skip, the identity matrix
  int xk = 0;
  Param* w_x = nullptr;
  Param* b_x = nullptr;
  float* bsum = nullptr;  // b + b_x
  void* wp = nullptr;
  float* bp = nullptr;
  int K() const { return ksize * ksize * cin; }
  int Ktot() const { return K() + xk; }
  int Kalg() const { return K() + (w_x ? xk : 0); }  // algorithmic K: identity columns are not counted as FLOPs
};

struct GroupNormP { Param* w = nullptr; Param* b = nullptr; int C = 0, G = 0; };
struct LayerNormP { Param* w = nullptr; Param* b = nullptr; int C = 0; };

struct ResBlockM {
  int in_ch = 0, out_ch = 0;
  bool up = false, down = false, has_skip_conv = false;
  bool folded = false;  // the skip path runs inside Conv_1's GEMM (bf16 engine, out_ch >= 128)
  GroupNormP gn0, gn1;
  Linear conv0, conv1, conv2;
  Param* dense_w = nullptr;
  Param* dense_b = nullptr;
  int temb_off = 0;
};
struct AttnBlockM {
  int C = 0;
  GroupNormP gn;
  Linear qkv, proj;
};
struct TransformerM {
  int C = 0, heads = 0;
  GroupNormP norm;
  Linear proj_in, proj_out, qkv1, out1, q2, kv2, out2, ff_in, ff_out;
  LayerNormP ln1, ln2, ln3;
  void* kv = nullptr;  // hoisted K|V projection of the text context, [B*L][2C] (grow-only buffer)
  size_t kv_bytes = 0;
};
struct ModuleM {
  int kind = 0;  // 0 ResBlock, 1 AttnBlock, 2 SpatialTransformer
  std::unique_ptr<ResBlockM> res;
  std::unique_ptr<AttnBlockM> attn;
  std::unique_ptr<TransformerM> st;
};
using BlockM = std::vector<ModuleM>;

struct GemmRecord {  // one implicit-GEMM launch of a profiled forward pass
  long long M = 0;
  int N = 0, K = 0, ksize = 1, tc = 0, H = 0, W = 0;
  float ms = 0.f;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
};

class UNet {
 public:
  explicit UNet(const UNetConfig& cfg);
  ~UNet();
  const UNetConfig& cfg() const { return cfg_; }
  const std::vector<std::unique_ptr<Param>>& params() const { return params_; }
  void load(const std::string& name, const void* dev_ptr, const std::vector<int64_t>& shape, int dtype,
            cudaStream_t st);
  void finalize(cudaStream_t st);
  // ctx: fp32 [B][L][context_dim] on the device.  Projects K|V of every cross-attention once (the
  // reference recomputes them in all 2*num_scales forwards, attention.py:174-175).
  void set_context(const float* ctx, int B, int L, cudaStream_t st);
  // same, from token ids [B][L] and the embedding table [V][context_dim] (fp32 or bf16) -- SURVEY 8f rank 3
  void set_context_tokens(const void* table, int table_dtype, long long V, const long long* tokens, int B, int L,
                          cudaStream_t st);
  // x: fp32 NCHW [B][C][N][N]; labels: int64 [B]; h_out: fp32 NCHW [B][C][N][N] un-scaled final conv
  // (what the fused PC-step kernels consume).  Returns through `h_out` only.
  // timesteps: optional fp32 [B] time conditioning embedded instead of float(labels) (VP models, ncsnpp.py:221-223)
  void forward_raw(const float* x, const long long* labels, float* h_out, int B, cudaStream_t st,
                   const float* timesteps = nullptr);
  // true: the next forward passes may reuse the time-embedding biases of the previous one (same labels, same B);
  // the PC loop sets it for the predictor evaluation that follows a corrector evaluation at the same noise level
  void set_reuse_temb(bool on) { reuse_temb_ = on; }
  // every sample carries the same noise label (the PC loop): the time-embedding path is evaluated for one sample
  // and its rows replicated -- bit-identical to evaluating it per sample
  void set_uniform_labels(bool on) { uniform_labels_ = on; }
  // Reference-shaped output: NCHW, divided by sigmas[labels] in double (ncsnpp.py:259-261), fp64 or fp32.
  void forward(const float* x, const long long* labels, void* out, int out_dtype, int B, cudaStream_t st,
               const float* timesteps = nullptr);
  const double* sigmas() const { return static_cast<const double*>(sigmas_->data); }
  void set_debug(bool on) { debug_ = on; }
  // GroupNorm-apply + SiLU inside the operand path of the 3x3 convolutions on 128-pixel-wide images (gemm_tc.cu,
  // conv_gemm_tcHF_kernel) instead of a separate pass; changes the launch sequence, hence the plan and any captured graph
  void set_fused_groupnorm(bool on) {
    if (on != fuse_gn_) { fuse_gn_ = on; planned_B_ = -1; ++generation_; }
  }
  // GroupNorm_1 + SiLU inside Conv_0's EPILOGUE (gemm_tc.cu, epilogue_role_gn): the raw Conv_0 output, its statistics
  // pass, gn_finalize and gn_apply of every ResBlock whose Conv_0 qualifies disappear.  Changes the launch sequence.
  void set_epilogue_groupnorm(bool on) {
    if (on != gn_out_) { gn_out_ = on; planned_B_ = -1; ++generation_; }
  }
  // copies a recorded block output (fp32 NCHW) to dst; returns its shape
  bool tap(const std::string& name, float* dst, int64_t capacity, int64_t shape[4], cudaStream_t st);
  // profile mode: CUDA events around every implicit-GEMM launch of the following forward passes
  void set_profile(bool on);
  int profile_records(GemmRecord* out, int cap);
  long long generation() const { return generation_; }  // bumped by finalize() / set_context()
  // bumped whenever a device buffer a captured forward may point into (activation arena, time-embedding buffer)
  // is re-allocated: a CUDA graph captured under another epoch must not be replayed
  long long resource_epoch() const { return resource_epoch_ + lane_.ws.epoch(); }
  size_t workspace_bytes() const { return lane_.ws.capacity(); }
  long long launches_per_forward() const { return launches_; }

 private:
  Param* add_param(const std::string& name, std::vector<int64_t> shape, int dtype = kF32);
  GroupNormP make_gn(const std::string& key, int C, int G = 0);
  LayerNormP make_ln(const std::string& key, int C);
  Linear make_conv(const std::string& key, int cin, int cout, int k);
  Linear make_linear(const std::string& key, int cin, int cout, bool bias);
  ModuleM make_res(const std::string& key, int in_ch, int out_ch, bool up, bool down);
  ModuleM make_attn(const std::string& key, int C);
  ModuleM make_st(const std::string& key, int C);
  void pack(Linear& l, cudaStream_t st);
  void set_context_impl(const float* ctx, const void* table, int table_dtype, long long V, const long long* tokens, int B,
                        int L, cudaStream_t st);

  // forward helpers (all honour dry_)
  Act new_act(int B, int H, int W, int C, bool with_stats);
  void free_act(Act& a);
  void gemm(const Linear& l, const Act& a0, const Act* a1, Act& out, const float* rowbias, int rowbias_ld,
            const void* residual, int res_up, float alpha, int out_dtype = -1, int out_nchw = 0,
            const Act* x0 = nullptr, const Act* x1 = nullptr, const float* gn_affine = nullptr,
            const GroupNormP* gn_out = nullptr);
  bool fuses_gn(const Linear& l, const Act& a0, const Act* a1) const;
  bool normalises_output(const Linear& l, const Act& a0, const GroupNormP& gn) const;
  void group_norm(const GroupNormP& g, const Act& a0, const Act* a1, int act, int mode, Act& out, Act* raw_out,
                  float** affine_out = nullptr);
  void attention(const void* q, const void* k, const void* v, void* out, int B, int heads, int Tq, int Tk, int d,
                 long long ldq, long long ldk, long long ldv, long long ldo, float scale);
  Act run_res(ResBlockM& m, const Act& a0, const Act* a1);
  Act finish_res(ResBlockM& m, const Act& a0, const Act* a1, Act& h2, Act& xr);
  Act run_attn(AttnBlockM& m, const Act& x);
  Act run_st(TransformerM& m, const Act& x);
  Act run_block(BlockM& blk, const Act& a0, const Act* a1, const std::string& tapname);
  void record_tap(const std::string& name, const Act& a, int dtype = -1);
  void forward_impl(const float* x, const long long* labels, float* h_out, int B);

  UNetConfig cfg_;
  std::vector<std::unique_ptr<Param>> params_;
  std::unordered_map<std::string, Param*> by_name_;
  Param* sigmas_ = nullptr;
  Param *pre0_w_, *pre0_b_, *pre1_w_, *pre1_b_;
  Linear pre_conv_, out_conv_;
  GroupNormP out_gn_;
  std::vector<BlockM> input_blocks_, out_blocks_;
  BlockM mid_block_;
  std::vector<ResBlockM*> all_res_;
  std::vector<TransformerM*> all_st_;
  std::vector<Linear*> all_linear_;
  Linear dense_all_;  // every ResBlock's Dense_0 stacked along N
  void* first_wp_ = nullptr;  // pre_conv weights as [nf][first_kpad_] bf16 (k = tap * C + c)
  int first_kpad_ = 0;
  int temb_total_ = 0;
  bool finalized_ = false;
  std::vector<void*> owned_;  // device allocations freed in the destructor

  // per-forward state: the execution context (stream + activation arena) of the forward being recorded
  struct Lane {
    Workspace ws;
    cudaStream_t st = nullptr;
    float* temb_all = nullptr;
    unsigned seq = 0;  // serpentine direction counter
  };
  Lane lane_;
  Lane* ln_ = &lane_;
  bool dry_ = false;
  bool debug_ = false;
  bool fuse_gn_ = false;
  bool gn_out_ = true;
  // statistics exchange of the epilogue GroupNorm (gemm_tc.cu, epilogue_role_gn): per-slot counters (idle: zero) and the
  // per-part statistics (idle: all ones); every launch leaves them idle
  int* gno_flags_ = nullptr;
  size_t gno_flags_bytes_ = 0;
  void* gno_part_ = nullptr;
  size_t gno_part_bytes_ = 0;
  // split-K scratch of the channel-major GEMM (launches of few tiles): partial accumulators, arrival counters (idle: zero)
  float* sk_part_ = nullptr;
  int* sk_ticket_ = nullptr;
  // weights of the forward's GEMM launches in launch order (recorded by the dry pass): launch i asks the L2 for the
  // weights of launch i + 1
  std::vector<std::pair<const void*, long long>> gemm_seq_;
  size_t gemm_idx_ = 0;
  bool profile_ = false;
  long long generation_ = 0;
  std::vector<GemmRecord> profile_log_;
  int planned_B_ = -1;
  long long launches_ = 0;
  // Serpentine sweep: consecutive streaming kernels (GEMM, GroupNorm apply) walk their tiles in alternating
  // direction, so each begins on the data its predecessor wrote last -- still resident in the 126 MB L2.
  bool serpentine_ = true;
  int ctx_B_ = 0, ctx_L_ = 0;
  std::map<std::string, std::pair<float*, std::vector<int64_t>>> taps_;
  void* ctx_buf_ = nullptr;  // text context in the compute dtype (grow-only staging buffer of set_context)
  size_t ctx_buf_bytes_ = 0;
  float* temb_persist_ = nullptr;  // [B][temb_total_] Dense_0(act(temb)) of the last evaluated labels
  size_t temb_persist_bytes_ = 0;
  int temb_valid_B_ = 0;
  bool reuse_temb_ = false;
  bool uniform_labels_ = false;
  float* h_scratch_ = nullptr;
  size_t h_scratch_bytes_ = 0;
  long long resource_epoch_ = 0;
  const float* timesteps_ = nullptr;  // time conditioning of the forward being recorded (null: float(labels))
};

}  // namespace t2p
