// UNet engine: turns the reference UNetModel topology (backbones/unet_openai.py:553-744)
// into a static launch plan over the kernels of this library.
//
//   create   : walk the constructor loops of the reference (:607-737) to enumerate layers and
//              the state-dict entries they own (same key names as the reference).
//   finalize : repack weights into kernel layouts, plan the activation workspace (arena with
//              plan-time liveness reuse -- no allocator runs inside forward) and build the op
//              list for one arithmetic mode:
//                EO_MODE_FP32  fp32 NHWC activations, SIMT kernels, GroupNorm folded into the
//                              consumer's operand load;
//                EO_MODE_BF16  bf16 NHWC activations, tcgen05 implicit-GEMM convs + tcgen05
//                              attention, GroupNorm statistics + apply as bandwidth kernels.
//   forward  : enqueue the op list on the caller's stream.
//
// Data layout in HBM: activations NHWC ([B,H,W,C], C contiguous) so that a conv tap is a
// dense [pixels, channels] K-major GEMM operand that TMA can fetch as one box; weights
// [Cout][K] with K ordered (source segment, tap, channel); network input/output NCHW fp32 as
// in the reference, converted inside the stem / output convolutions.
#include "kernels.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <vector>

namespace eo {

// ---------------------------------------------------------------------------------------
// plan-time arena: best-fit free list over one device allocation
// ---------------------------------------------------------------------------------------
struct Arena {
  struct Blk { size_t off, size; };
  std::vector<Blk> free_list;
  size_t top = 0, peak = 0;
  bool keep = false;   // debug: never reuse (activations stay readable after forward)
  static size_t align(size_t v) { return (v + 1023) & ~(size_t)1023; }
  size_t alloc(size_t bytes) {
    bytes = align(bytes ? bytes : 1);
    int best = -1;
    for (int i = 0; i < (int)free_list.size(); ++i)
      if (free_list[i].size >= bytes && (best < 0 || free_list[i].size < free_list[best].size)) best = i;
    if (best >= 0) {
      size_t off = free_list[best].off;
      if (free_list[best].size == bytes) free_list.erase(free_list.begin() + best);
      else { free_list[best].off += bytes; free_list[best].size -= bytes; }
      return off;
    }
    size_t off = top;
    top += bytes;
    peak = std::max(peak, top);
    return off;
  }
  void release(size_t off, size_t bytes) {
    if (keep) return;
    bytes = align(bytes ? bytes : 1);
    free_list.push_back({off, bytes});
    std::sort(free_list.begin(), free_list.end(), [](const Blk& a, const Blk& b) { return a.off < b.off; });
    for (size_t i = 0; i + 1 < free_list.size();) {
      if (free_list[i].off + free_list[i].size == free_list[i + 1].off) {
        free_list[i].size += free_list[i + 1].size;
        free_list.erase(free_list.begin() + i + 1);
      } else ++i;
    }
    if (!free_list.empty() && free_list.back().off + free_list.back().size == top) {
      top = free_list.back().off;
      free_list.pop_back();
    }
  }
};

struct Act {   // NHWC activation living in the arena
  size_t off = 0;
  int C = 0, H = 0, W = 0;
  size_t bytes = 0;
  bool valid = false;
  long long stats = -1;   // offset (floats) of this tensor's per-channel [Bmax][C][2] sums in ch_stats, or -1
};

struct WSpec {
  std::string name;
  std::vector<int64_t> shape;
  const float* src = nullptr;   // caller-owned fp32 device pointer (read only during finalize)
  const float* priv = nullptr;  // engine-owned copy of vectors / matrices that kernels read at run time
};

enum LayerKind { L_CONV_IN, L_RES, L_ATTN, L_DOWN, L_UP };
struct Layer {
  LayerKind kind;
  std::string prefix;   // e.g. "input_blocks.1.0."
  int cin = 0, cout = 0, heads = 0;
  int tb_off = -1;      // ResBlock: column offset into the per-step embedding table
  int updown = 0;       // ResBlock of resblock_updown: 1 = down (2x2 average), 2 = up (nearest x2)
};
struct Block { std::vector<Layer> layers; std::string name; };

struct Op {
  std::string name;
  std::function<int()> prepare;                        // after the arena is allocated
  std::function<int(int, cudaStream_t)> run;           // (B, stream)
  const char* kernel = "";                             // kernel family (for bench.py's roofline)
  double flops = 0;                                    // algorithmic FLOPs per image
  double exec_flops = 0;                               // FLOPs the launch executes per image (padded rows, folded taps)
  double bytes = 0;                                    // algorithmic HBM bytes per image
};

}  // namespace eo

using namespace eo;

struct eo_unet {
  eo_unet_cfg cfg{};
  std::vector<Block> in_blocks, out_blocks;
  Block mid;
  int final_ch = 0;
  int ted = 0;           // time_embed_dim
  int tb_total = 0;      // sum of ResBlock out channels
  std::vector<WSpec> wspecs;
  std::map<std::string, int> windex;

  // ---- finalized state
  bool finalized = false;
  int mode = EO_MODE_FP32, Bmax = 0, H = 0, W = 0;
  int act_dt = DT_F32;
  Arena arena;
  uint8_t* arena_base = nullptr;
  std::vector<void*> owned;            // packed weights and persistent buffers
  std::vector<TcConvPlan*> tc_plans;
  std::vector<TcAttnPlan*> attn_plans;
  std::vector<Op> ops;
  std::map<std::string, Act> named;
  int64_t dev_bytes = 0;
  int n_launches = 0;
  // persistent small buffers
  float *e0 = nullptr, *l1 = nullptr, *emb = nullptr, *tb = nullptr;
  float *w_emb_cat = nullptr, *b_emb_cat = nullptr;
  double* gn_sums = nullptr;
  int n_gn = 0;
  // timestep tables (eo_unet_build_time_tables): row t = the per-step embedding table `tb` of timestep value t,
  // computed once per sampling loop instead of once per step (unet_openai.py:763, :374-376)
  float* tt_table = nullptr;
  int tt_n = 0;
  bool tt_on = false;               // forwards gather from the table (set by build, cleared by clear / finalize; the
                                    // whole-loop entry points switch it on for their own loop only)
  void release_time_tables() {
    if (tt_table) { cudaFree(tt_table); dev_bytes -= (int64_t)tt_n * tb_total * sizeof(float); }
    tt_table = nullptr; tt_n = 0; tt_on = false;
  }
  int build_time_tables(int n, cudaStream_t st);
  double* ch_stats = nullptr;       // per-channel GroupNorm sums emitted by conv epilogues (zeroed per forward)
  size_t ch_stats_floats = 0;       // element count
  // per-forward io (read by ops at launch time)
  const float* io_x = nullptr; int io_cx = 0;
  const float* io_cond = nullptr; int io_cc = 0;
  const int64_t* io_t = nullptr; const int64_t* io_y = nullptr;
  float* io_out = nullptr;
  // CUDA-graph replay of the forward: decisive for launch-bound sizes (64x64 batch 1:
  // 2.69 -> 2.35 ms per step), and still 0.9 ms of a 69.6 ms step at 256x256 batch 64 (171 launch gaps).  The graph
  // reads its inputs from / writes its output to engine-owned staging buffers, so one instantiation serves every
  // step of a sampling loop; key = (B, Cx, Cc, has y).
  struct GraphSlot { int seen = 0; bool failed = false; cudaGraphExec_t exec = nullptr; };
  std::map<long long, GraphSlot> graphs;
  cudaStream_t graph_stream = nullptr;          // capture stream (the caller's may be the legacy default stream)
  float *gs_x = nullptr, *gs_cond = nullptr, *gs_out = nullptr;
  int64_t *gs_t = nullptr, *gs_y = nullptr;
  void release_graphs_only() {
    for (auto& kv : graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    graphs.clear();
  }
  void release_graphs() {
    release_graphs_only();
    if (graph_stream) { cudaStreamDestroy(graph_stream); graph_stream = nullptr; }
    gs_x = gs_cond = gs_out = nullptr; gs_t = gs_y = nullptr;   // freed with `owned`
  }
  int run_ops(int B, cudaStream_t st);
  int forward_graph(const float* x, int Cx, const float* cond, int Cc, const int64_t* t, const int64_t* y, float* out,
                    int B, cudaStream_t st, bool* handled);

  ~eo_unet() { release_plan(); }

  void release_plan() {
    release_time_tables();
    release_graphs();
    for (auto* p : tc_plans) tc_conv_plan_destroy(p);
    for (auto* p : attn_plans) tc_attn_plan_destroy(p);
    tc_plans.clear(); attn_plans.clear();
    for (void* p : owned) cudaFree(p);
    owned.clear();
    if (arena_base) cudaFree(arena_base);
    arena_base = nullptr;
    ops.clear(); named.clear(); stem_split.clear();
    arena = Arena();
    finalized = false; dev_bytes = 0; n_launches = 0; n_gn = 0; ch_stats = nullptr; ch_stats_floats = 0;
    e0 = l1 = emb = tb = w_emb_cat = b_emb_cat = nullptr; gn_sums = nullptr;
    for (auto& ws : wspecs) ws.priv = nullptr;
  }

  // ------------------------------------------------------------------ topology
  void add_w(const std::string& name, std::vector<int64_t> shape) {
    windex[name] = (int)wspecs.size();
    wspecs.push_back({name, std::move(shape), nullptr});
  }
  void add_conv_w(const std::string& p, int cout, int cin, int k) {
    add_w(p + "weight", {cout, cin, k, k});
    add_w(p + "bias", {cout});
  }
  void add_res_w(Layer& L) {
    const std::string& p = L.prefix;
    add_w(p + "in_layers.0.weight", {L.cin}); add_w(p + "in_layers.0.bias", {L.cin});
    add_conv_w(p + "in_layers.2.", L.cout, L.cin, 3);
    const int erows = (cfg.use_scale_shift_norm ? 2 : 1) * L.cout;      // FiLM: (scale | shift)
    add_w(p + "emb_layers.1.weight", {erows, ted}); add_w(p + "emb_layers.1.bias", {erows});
    add_w(p + "out_layers.0.weight", {L.cout}); add_w(p + "out_layers.0.bias", {L.cout});
    add_conv_w(p + "out_layers.3.", L.cout, L.cout, 3);
    if (L.cin != L.cout) add_conv_w(p + "skip_connection.", L.cout, L.cin, 1);
    L.tb_off = tb_total;
    tb_total += erows;
  }
  void add_attn_w(Layer& L) {
    const std::string& p = L.prefix;
    add_w(p + "norm.weight", {L.cin}); add_w(p + "norm.bias", {L.cin});
    add_w(p + "qkv.weight", {3 * L.cin, L.cin, 1}); add_w(p + "qkv.bias", {3 * L.cin});
    add_w(p + "proj_out.weight", {L.cin, L.cin, 1}); add_w(p + "proj_out.bias", {L.cin});
  }

  int build_topology() {
    const eo_unet_cfg& c = cfg;
    ted = c.model_channels * c.time_emb_factor;
    add_w("time_embed.freqs", {c.model_channels / 2});   // host-computed, see unet.py
    add_w("time_embed.0.weight", {ted, c.model_channels}); add_w("time_embed.0.bias", {ted});
    add_w("time_embed.2.weight", {ted, ted}); add_w("time_embed.2.bias", {ted});
    if (c.num_classes > 0) add_w("label_emb.weight", {c.num_classes, ted});
    auto is_attn = [&](int ds) {
      for (int i = 0; i < c.n_attention_resolutions; ++i) if (c.attention_resolutions[i] == ds) return true;
      return false;
    };
    auto nheads = [&](int ch, int h) { return c.num_head_channels == -1 ? h : ch / c.num_head_channels; };
    const int heads_up = c.num_heads_upsample == -1 ? c.num_heads : c.num_heads_upsample;
    int ch = c.channel_mult[0] * c.model_channels;
    std::vector<int> chans;
    {
      Block b; b.name = "input_blocks.0";
      Layer L; L.kind = L_CONV_IN; L.prefix = "input_blocks.0.0."; L.cin = c.in_channels; L.cout = ch;
      add_conv_w(L.prefix, ch, c.in_channels, 3);
      b.layers.push_back(L); in_blocks.push_back(b); chans.push_back(ch);
    }
    int ds = 1;
    for (int level = 0; level < c.n_channel_mult; ++level) {
      int mult = c.channel_mult[level];
      for (int r = 0; r < c.num_res_blocks; ++r) {
        Block b; b.name = "input_blocks." + std::to_string(in_blocks.size());
        Layer L; L.kind = L_RES; L.prefix = b.name + ".0."; L.cin = ch; L.cout = mult * c.model_channels;
        add_res_w(L); b.layers.push_back(L);
        ch = L.cout;
        if (is_attn(ds)) {
          Layer A; A.kind = L_ATTN; A.prefix = b.name + ".1."; A.cin = A.cout = ch; A.heads = nheads(ch, c.num_heads);
          add_attn_w(A); b.layers.push_back(A);
        }
        in_blocks.push_back(b); chans.push_back(ch);
      }
      if (level != c.n_channel_mult - 1) {
        Block b; b.name = "input_blocks." + std::to_string(in_blocks.size());
        Layer L; L.prefix = b.name + ".0."; L.cin = L.cout = ch;
        if (c.resblock_updown) { L.kind = L_RES; L.updown = 1; add_res_w(L); }       // unet_openai.py:645-658
        else { L.kind = L_DOWN; add_conv_w(L.prefix + "op.", ch, ch, 3); }
        b.layers.push_back(L); in_blocks.push_back(b); chans.push_back(ch);
        ds *= 2;
      }
    }
    {
      mid.name = "middle_block";
      Layer R0; R0.kind = L_RES; R0.prefix = "middle_block.0."; R0.cin = R0.cout = ch; add_res_w(R0);
      Layer A; A.kind = L_ATTN; A.prefix = "middle_block.1."; A.cin = A.cout = ch; A.heads = nheads(ch, c.num_heads); add_attn_w(A);
      Layer R1; R1.kind = L_RES; R1.prefix = "middle_block.2."; R1.cin = R1.cout = ch; add_res_w(R1);
      mid.layers = {R0, A, R1};
    }
    for (int level = c.n_channel_mult - 1; level >= 0; --level) {
      int mult = c.channel_mult[level];
      for (int i = 0; i < c.num_res_blocks + 1; ++i) {
        int ich = chans.back(); chans.pop_back();
        Block b; b.name = "output_blocks." + std::to_string(out_blocks.size());
        int li = 0;
        Layer L; L.kind = L_RES; L.prefix = b.name + "." + std::to_string(li++) + "."; L.cin = ch + ich; L.cout = c.model_channels * mult;
        add_res_w(L); b.layers.push_back(L);
        ch = L.cout;
        if (is_attn(ds)) {
          Layer A; A.kind = L_ATTN; A.prefix = b.name + "." + std::to_string(li++) + "."; A.cin = A.cout = ch; A.heads = nheads(ch, heads_up);
          add_attn_w(A); b.layers.push_back(A);
        }
        if (level && i == c.num_res_blocks) {
          Layer U; U.prefix = b.name + "." + std::to_string(li++) + "."; U.cin = U.cout = ch;
          if (c.resblock_updown) { U.kind = L_RES; U.updown = 2; add_res_w(U); }     // unet_openai.py:722-735
          else { U.kind = L_UP; add_conv_w(U.prefix + "conv.", ch, ch, 3); }
          b.layers.push_back(U);
          ds /= 2;
        }
        out_blocks.push_back(b);
      }
    }
    final_ch = ch;
    add_w("out.0.weight", {ch}); add_w("out.0.bias", {ch});
    add_conv_w("out.2.", c.out_channels, c.channel_mult[0] * c.model_channels, 3);
    return EO_OK;
  }

  const float* w(const std::string& name) const {
    auto it = windex.find(name);
    if (it == windex.end()) return nullptr;
    const WSpec& ws = wspecs[it->second];
    return ws.priv ? ws.priv : ws.src;
  }

  // ------------------------------------------------------------------ finalize helpers
  template <typename T> int dmalloc(T** out, size_t count) {
    void* p = nullptr;
    EO_CHECK_CUDA(cudaMalloc(&p, std::max<size_t>(count * sizeof(T), 16)));
    owned.push_back(p);
    dev_bytes += (int64_t)(count * sizeof(T));
    *out = reinterpret_cast<T*>(p);
    return EO_OK;
  }
  Act new_act(int C, int Hh, int Ww, int batch_mult = 1) {
    Act a; a.C = C; a.H = Hh; a.W = Ww;
    a.bytes = (size_t)Bmax * batch_mult * Hh * Ww * C * dtype_size(act_dt);
    a.off = arena.alloc(a.bytes);
    a.valid = true;
    return a;
  }
  size_t new_scratch(size_t bytes) { return arena.alloc(bytes); }
  void free_act(Act& a) { if (a.valid) arena.release(a.off, a.bytes); a.valid = false; }
  template <typename T = void> T* ptr(size_t off) const { return reinterpret_cast<T*>(arena_base + off); }

  struct PackSeg { const float* wsrc; int cin_total; int ksize; int cin_off; int C; };
  float* last_packed_bias = nullptr;   // bias vector of the conv planned last (plan_attn patches the qkv one)
  const float* pack_row_scale = nullptr;   // per-output-row factor applied by pack_tc / pack_bias2 while set (device, [rows])

  // SIMT layout [Ktot][Cout] fp32
  int pack_simt(const std::vector<PackSeg>& segs, int Cout, float** out, int* ktot, cudaStream_t st) {
    int K = 0;
    for (auto& s : segs) K += s.ksize * s.ksize * s.C;
    float* d = nullptr;
    int rc = dmalloc(&d, (size_t)K * Cout);
    if (rc) return rc;
    int koff = 0;
    for (auto& s : segs) {
      rc = launch_pack_conv_weight(s.wsrc, s.cin_total, s.ksize, s.cin_off, s.C, d, DT_F32, 1, Cout, koff, Cout, nullptr, st);
      if (rc) return rc;
      koff += s.ksize * s.ksize * s.C;
    }
    *out = d; *ktot = K;
    return EO_OK;
  }
  // tensor-core layout [Nrows][Ktot] bf16, K contiguous
  int pack_tc(const std::vector<PackSeg>& segs, int Nrows, const int* d_row_map, void** out, int* ktot, cudaStream_t st) {
    int K = 0;
    for (auto& s : segs) K += s.ksize * s.ksize * s.C;
    __nv_bfloat16* d = nullptr;
    int rc = dmalloc(&d, (size_t)K * Nrows);
    if (rc) return rc;
    int koff = 0;
    for (auto& s : segs) {
      rc = launch_pack_conv_weight(s.wsrc, s.cin_total, s.ksize, s.cin_off, s.C, d, DT_BF16, K, 1, koff, Nrows, d_row_map, st, pack_row_scale);
      if (rc) return rc;
      koff += s.ksize * s.ksize * s.C;
    }
    *out = d; *ktot = K;
    return EO_OK;
  }
  int pack_bias2(const float* a, const float* b, int N, const int* d_row_map, float** out, cudaStream_t st) {
    float* d = nullptr;
    int rc = dmalloc(&d, (size_t)N);
    if (rc) return rc;
    rc = launch_pack_bias(a, b, d, N, d_row_map, st, pack_row_scale);
    if (rc) return rc;
    *out = d;
    return EO_OK;
  }

  void push(const std::string& name, std::function<int(int, cudaStream_t)> run, int launches = 1,
            std::function<int()> prepare = nullptr) {
    ops.push_back({name, std::move(prepare), std::move(run)});
    n_launches += launches;
  }
  // annotate the op pushed last
  void note(const char* kernel, double flops, double bytes, double exec_flops = -1.0) {
    ops.back().kernel = kernel; ops.back().flops = flops; ops.back().bytes = bytes;
    ops.back().exec_flops = exec_flops >= 0 ? exec_flops : flops;
  }

  // GroupNorm statistics of (a [, b]) -> scale/shift [Bmax, Ctot] scratch in the arena
  struct GnOut { size_t scale_off, shift_off; int C; size_t bytes; };
  GnOut plan_gn(const std::string& name, const Act& a, const Act* b, const float* gamma, const float* beta) {
    GnOut g;
    g.C = a.C + (b ? b->C : 0);
    g.bytes = (size_t)Bmax * g.C * sizeof(float);
    g.scale_off = new_scratch(g.bytes);
    g.shift_off = new_scratch(g.bytes);
    const int HW = a.H * a.W;
    Act aa = a; Act bb = b ? *b : Act();
    const bool two = b != nullptr;
    const int dt = act_dt;
    if (a.stats >= 0 && (!two || b->stats >= 0)) {
      // every source carries per-channel sums written by its producer's epilogue: no extra read
      push(name + ".gn", [=](int B, cudaStream_t st) -> int {
        return launch_gn_finalize_ch(ch_stats + aa.stats, aa.C, two ? ch_stats + bb.stats : nullptr, two ? bb.C : 0,
                                     gamma, beta, B, HW, ptr<float>(g.scale_off), ptr<float>(g.shift_off), st);
      });
      note("k_gn_finalize_ch", 0, 0);
      return g;
    }
    const int gi = n_gn++;
    push(name + ".gn", [=](int B, cudaStream_t st) -> int {
      GnSrc s[2];
      s[0].ptr = ptr(aa.off); s[0].C = aa.C;
      if (two) { s[1].ptr = ptr(bb.off); s[1].C = bb.C; }
      double* sums = gn_sums + (size_t)gi * Bmax * 64;
      int rc = launch_gn_stats(s, two ? 2 : 1, dt, B, HW, sums, st);
      if (rc) return rc;
      return launch_gn_finalize(sums, gamma, beta, B, g.C, HW, ptr<float>(g.scale_off), ptr<float>(g.shift_off), st);
    }, 2);
    note("k_gn_stats", 0, (double)HW * g.C * dtype_size(act_dt));
    return g;
  }
  void free_gn(GnOut& g) { arena.release(g.scale_off, g.bytes); arena.release(g.shift_off, g.bytes); }

  // ------------------------------------------------------------------ layer planners
  // conv over up to three sources; returns the output activation
  struct SrcSpec {
    Act act; int ksize = 3;
    const GnOut* gn = nullptr; int gn_coff = 0; int silu = 0;   // fp32 mode: folded into the load
  };

  int plan_conv_fp32(const std::string& name, const std::vector<SrcSpec>& srcs, const std::vector<PackSeg>& segs,
                     int Cout, const float* bias_a, const float* bias_b, int tb_off, const Act* residual,
                     int stride, int up, Act* out, cudaStream_t st) {
    float* Wp = nullptr; int K = 0;
    int rc = pack_simt(segs, Cout, &Wp, &K, st);
    if (rc) return rc;
    float* bias = nullptr;
    if (bias_a || bias_b) { rc = pack_bias2(bias_a, bias_b, Cout, nullptr, &bias, st); if (rc) return rc; }
    const Act& s0 = srcs[0].act;
    int Ho = up ? s0.H * 2 : (stride == 2 ? s0.H / 2 : s0.H);
    int Wo = up ? s0.W * 2 : (stride == 2 ? s0.W / 2 : s0.W);
    Act o = new_act(Cout, Ho, Wo);
    std::vector<SrcSpec> sv = srcs;
    std::vector<GnOut> gns;
    for (auto& s : sv) gns.push_back(s.gn ? *s.gn : GnOut());
    std::vector<int> woffs; { int k = 0; for (auto& sg : segs) { woffs.push_back(k); k += sg.ksize * sg.ksize * sg.C; } }
    Act res = residual ? *residual : Act();
    const bool has_res = residual != nullptr;
    push(name, [=](int B, cudaStream_t stx) -> int {
      ConvSimtParams p;
      p.nsrc = (int)sv.size();
      for (int i = 0; i < p.nsrc; ++i) {
        p.src[i].ptr = ptr(sv[i].act.off); p.src[i].C = sv[i].act.C; p.src[i].ksize = sv[i].ksize;
        p.src[i].dt = DT_F32; p.src[i].nchw = 0;
        if (sv[i].gn) {
          p.src[i].gn_scale = ptr<float>(gns[i].scale_off) + sv[i].gn_coff;
          p.src[i].gn_shift = ptr<float>(gns[i].shift_off) + sv[i].gn_coff;
          p.src[i].gn_ld = gns[i].C;
        }
        p.src[i].silu = sv[i].silu;
        p.src[i].w_off = woffs[i];
      }
      p.B = B; p.Hin = sv[0].act.H; p.Win = sv[0].act.W; p.Hout = Ho; p.Wout = Wo;
      p.stride = stride; p.up = up;
      p.W = Wp; p.Ktot = K; p.Cout = Cout; p.bias = bias;
      if (tb_off >= 0) { p.bias_nc = tb + tb_off; p.ld_bias_nc = tb_total; }
      p.residual = has_res ? ptr(res.off) : nullptr;
      p.out = ptr(o.off); p.out_dt = DT_F32;
      return launch_conv_simt(p, stx);
    });
    note("k_conv_simt", 2.0 * Ho * Wo * Cout * K, 0);
    *out = o;
    return EO_OK;
  }

  // bf16 tensor-core conv.  `tsegs` carry activation + taps; weights in `segs` (same order).
  struct TcSegSpec {
    Act act; int batch_mult = 1; int ntaps = 9; int8_t dh[9]; int8_t dw[9]; int plane[9];
    int stride = 1;   // 2: the source grid is twice the output grid (stride-2 conv read through a strided tensor map)
    bool has_gn = false; GnOut gn{}; int gn_coff = 0; int silu = 0;   // GroupNorm (+SiLU) folded into the operand load
  };
  static TcSegSpec with_gn(TcSegSpec s, const GnOut& g, int coff, int silu) {
    s.has_gn = true; s.gn = g; s.gn_coff = coff; s.silu = silu;
    return s;
  }
  // GroupNorm + SiLU can ride on the conv's operand load (persistent kernel; 3x3 windows need halo patches)
  bool can_fuse_gn(int ksize, int Ho, int Wo, int C) const {
    if (C % 64 != 0) return false;
    // a 1x1 conv would redo the transform for each of its N tiles (qkv: 6x) against one K block of MMA work
    return ksize == 3 && tc_conv_patch_supported(Ho, Wo);
  }
  static TcSegSpec seg3x3(const Act& a) {
    TcSegSpec s; s.act = a; s.ntaps = 9;
    for (int t = 0; t < 9; ++t) { s.dh[t] = (int8_t)(t / 3 - 1); s.dw[t] = (int8_t)(t % 3 - 1); s.plane[t] = 0; }
    return s;
  }
  static TcSegSpec seg1x1(const Act& a) {
    TcSegSpec s; s.act = a; s.ntaps = 1; s.dh[0] = 0; s.dw[0] = 0; s.plane[0] = 0;
    return s;
  }
  // a plain 3x3 window over one tensor -- or a 2x2 corner of it, the taps of a sub-pixel convolution of Upsample --
  // is served from halo patches when the output grid allows it: the tc_patch_code of the window, 0 = plain tiles
  static int patch_code(const TcSegSpec& s, int Ho, int Wo) {
    if (s.stride != 1 || !tc_conv_patch_supported(Ho, Wo) || (s.ntaps != 9 && s.ntaps != 4)) return 0;
    const int n = s.ntaps == 9 ? 3 : 2;
    const int r0 = s.dh[0] + 1, c0 = s.dw[0] + 1;
    if (r0 < 0 || c0 < 0 || r0 + n > 3 || c0 + n > 3) return 0;
    for (int t = 0; t < s.ntaps; ++t)
      if (s.dh[t] != r0 + t / n - 1 || s.dw[t] != c0 + t % n - 1 || s.plane[t] != 0) return 0;
    return tc_patch_code(r0, n, c0, n);
  }
  static bool patchable(const TcSegSpec& s, int Ho, int Wo) { return patch_code(s, Ho, Wo) != 0; }
  // weights of a patch segment are K-ordered (64-channel block, tap, channel): one PackSeg per block
  static std::vector<PackSeg> patch_order(const std::vector<PackSeg>& segs, const std::vector<bool>& patch) {
    std::vector<PackSeg> o;
    for (size_t i = 0; i < segs.size(); ++i) {
      if (!patch[i]) { o.push_back(segs[i]); continue; }
      for (int c0 = 0; c0 < segs[i].C; c0 += 64) {
        PackSeg s = segs[i]; s.cin_off += c0; s.C = 64;
        o.push_back(s);
      }
    }
    return o;
  }
  // write into a strided view of an existing activation instead of a new one (sub-pixel convs of Upsample)
  struct OutView { Act target; long long off, sw, sh, sn; double alg_flops; };   // alg_flops: algorithmic FLOPs per image to report
  int plan_conv_tc(const std::string& name, const std::vector<TcSegSpec>& tsegs, const std::vector<PackSeg>& segs_in,
                   int Cout_rows, const int* d_row_map, const float* bias_a, const float* bias_b, int tb_off,
                   const Act* residual, int Ho, int Wo, Act* out, cudaStream_t st, bool want_stats = true,
                   const OutView* view = nullptr, double alg_flops = -1.0, int nchw_C = 0) {
    std::vector<bool> patch;
    std::vector<int> pcode;
    for (auto& ts : tsegs) {
      pcode.push_back(ts.act.C % 64 == 0 ? patch_code(ts, Ho, Wo) : 0);
      patch.push_back(pcode.back() != 0);
    }
    const std::vector<PackSeg> segs = patch_order(segs_in, patch);
    void* Wp = nullptr; int K = 0;
    int rc = pack_tc(segs, Cout_rows, d_row_map, &Wp, &K, st);
    if (rc) return rc;
    float* bias = nullptr;
    if (bias_a || bias_b) { rc = pack_bias2(bias_a, bias_b, Cout_rows, d_row_map, &bias, st); if (rc) return rc; }
    last_packed_bias = bias;
    // nchw_C > 0: the head -- fp32 NCHW straight to the caller's eps buffer (io_out at launch), no activation
    Act o = view ? view->target : (nchw_C ? Act() : new_act(Cout_rows, Ho, Wo));
    if (!view && want_stats && tc_conv_stats_supported(Ho, Wo)) {
      o.stats = (long long)ch_stats_floats;
      ch_stats_floats += (size_t)Bmax * Cout_rows * 2;
    }
    const bool has_view = view != nullptr;
    const OutView vw = view ? *view : OutView();
    Act res = residual ? *residual : Act();
    const bool has_res = residual != nullptr;
    std::vector<TcSegSpec> tv = tsegs;
    const size_t plan_idx = tc_plans.size();
    tc_plans.push_back(nullptr);
    auto prepare = [=]() -> int {
      TcConvParams p;
      p.nseg = (int)tv.size();
      for (int i = 0; i < p.nseg; ++i) {
        p.seg[i].ptr = ptr(tv[i].act.off); p.seg[i].C = tv[i].act.C;
        p.seg[i].Bt = Bmax * tv[i].batch_mult; p.seg[i].ntaps = tv[i].ntaps; p.seg[i].patch = pcode[i];
        p.seg[i].stride = tv[i].stride;
        if (tv[i].has_gn) {
          p.seg[i].gn_scale = ptr<float>(tv[i].gn.scale_off); p.seg[i].gn_shift = ptr<float>(tv[i].gn.shift_off);
          p.seg[i].gn_ld = tv[i].gn.C; p.seg[i].gn_coff = tv[i].gn_coff; p.seg[i].silu = tv[i].silu;
        }
        for (int t = 0; t < tv[i].ntaps; ++t) {
          p.seg[i].dh[t] = tv[i].dh[t]; p.seg[i].dw[t] = tv[i].dw[t]; p.seg[i].dn[t] = tv[i].plane[t] * Bmax;
        }
      }
      p.B = Bmax; p.H = Ho; p.W = Wo; p.Wp = Wp; p.Ktot = K; p.Cout = Cout_rows; p.bias = bias;
      if (tb_off >= 0) { p.bias_nc = tb + tb_off; p.ld_bias_nc = tb_total; }
      p.residual = has_res ? ptr(res.off) : nullptr;
      p.out = nchw_C ? nullptr : ptr(o.off);
      p.out_nchw_C = nchw_C;
      if (has_view) {
        p.out = ptr<__nv_bfloat16>(o.off) + vw.off;
        p.out_sw = vw.sw; p.out_sh = vw.sh; p.out_sn = vw.sn;
      }
      p.stats = o.stats >= 0 ? ch_stats + o.stats : nullptr;
      return tc_conv_plan_create(p, &tc_plans[plan_idx]);
    };
    push(name, [=](int B, cudaStream_t stx) -> int { return tc_conv_launch(tc_plans[plan_idx], B, stx, nchw_C ? io_out : nullptr); },
         1, prepare);
    // algorithmic FLOPs (2 * out_elems * Cin * k of the reference's layer) next to what the launch
    // executes (padded qkv / head / stem rows, the 4/9 sub-pixel form of Upsample)
    note("k_conv_tc3",
         has_view ? vw.alg_flops : alg_flops >= 0 ? alg_flops : 2.0 * Ho * Wo * Cout_rows * K, 0,
         2.0 * Ho * Wo * Cout_rows * K);
    *out = o;
    return EO_OK;
  }

  // bf16: materialise act(GN(a [, b])) as one concatenated bf16 tensor
  Act plan_gn_apply(const std::string& name, const Act& a, const Act* b, const GnOut& g, int silu) {
    Act o = new_act(g.C, a.H, a.W);
    Act aa = a; Act bb = b ? *b : Act();
    const bool two = b != nullptr;
    const int HW = a.H * a.W;
    GnOut gg = g;
    push(name + ".gn_apply", [=](int B, cudaStream_t st) -> int {
      GnSrc s[2];
      s[0].ptr = ptr(aa.off); s[0].C = aa.C;
      if (two) { s[1].ptr = ptr(bb.off); s[1].C = bb.C; }
      return launch_gn_apply(s, two ? 2 : 1, B, HW, ptr<float>(gg.scale_off), ptr<float>(gg.shift_off), silu, ptr(o.off), st);
    });
    note("k_gn_apply", 0, 2.0 * HW * g.C * dtype_size(act_dt));
    return o;
  }

  // ResBlock._forward for the variants of the reference's UNet / UNetBig / UNetSmall factories: scale-shift (FiLM)
  // conditioning (unet_openai.py:377-381) and up/down-sampling blocks (:366-371).  SiLU(GroupNorm(x)) is
  // materialised (resampled for an up/down block, like x itself), the first convolution reads it plainly; the second
  // GroupNorm's folded affine is modulated by the block's (scale | shift) row before it rides on the second conv.
  int plan_res_x(const Layer& L, const Act& a, const Act* b, Act* out, cudaStream_t st) {
    const std::string& p = L.prefix;
    const bool film = cfg.use_scale_shift_norm != 0;
    const int rmode = L.updown == 2 ? 1 : L.updown == 1 ? 2 : 0;
    const int Ca = a.C, Cb = b ? b->C : 0;
    if (Ca + Cb != L.cin) { set_error("plan_res %s: channel mismatch", p.c_str()); return EO_ERR_STATE; }
    if (rmode == 2 && (a.H % 2 || a.W % 2)) { set_error("down ResBlock %s: odd feature map %dx%d", p.c_str(), a.H, a.W); return EO_ERR_ARG; }
    if (rmode == 1 && a.H == 3 && a.W == 3) { set_error("up ResBlock: the reference's 3x3 -> 7x7 pad quirk is not implemented"); return EO_ERR_ARG; }
    const int Ho = rmode == 1 ? a.H * 2 : rmode == 2 ? a.H / 2 : a.H, Wo = rmode == 1 ? a.W * 2 : rmode == 2 ? a.W / 2 : a.W;
    const bool has_skip = L.cin != L.cout;
    const int dt = act_dt;
    int rc;
    GnOut g1 = plan_gn(p + "in_layers.0", a, b, w(p + "in_layers.0.weight"), w(p + "in_layers.0.bias"));
    Act hin = new_act(L.cin, Ho, Wo);
    Act aa = a, bb = b ? *b : Act();
    const bool two = b != nullptr;
    const int cin = L.cin;
    push(p + "in_layers.1.resample", [=](int B, cudaStream_t s) -> int {
      int r = launch_resample(ptr(aa.off), aa.C, ptr(hin.off), cin, 0, dt, B, aa.H, aa.W, rmode, ptr<float>(g1.scale_off),
                              ptr<float>(g1.shift_off), g1.C, 0, 1, s);
      if (r || !two) return r;
      return launch_resample(ptr(bb.off), bb.C, ptr(hin.off), cin, aa.C, dt, B, bb.H, bb.W, rmode, ptr<float>(g1.scale_off),
                             ptr<float>(g1.shift_off), g1.C, aa.C, 1, s);
    }, two ? 2 : 1);
    note("k_resample", 0, (double)(a.H * a.W + Ho * Wo) * L.cin * dtype_size(act_dt));
    free_gn(g1);
    // the skip input: resampled, or concatenated when the block has no skip convolution to read two sources
    const bool xmat = rmode != 0 || (two && !has_skip);
    Act xin;
    if (xmat) {
      xin = new_act(L.cin, Ho, Wo);
      push(p + "x_upd", [=](int B, cudaStream_t s) -> int {
        int r = launch_resample(ptr(aa.off), aa.C, ptr(xin.off), cin, 0, dt, B, aa.H, aa.W, rmode, nullptr, nullptr, 0, 0, 0, s);
        if (r || !two) return r;
        return launch_resample(ptr(bb.off), bb.C, ptr(xin.off), cin, aa.C, dt, B, bb.H, bb.W, rmode, nullptr, nullptr, 0, 0, 0, s);
      }, two ? 2 : 1);
      note("k_resample", 0, (double)(a.H * a.W + Ho * Wo) * L.cin * dtype_size(act_dt));
    }
    const float* w1 = w(p + "in_layers.2.weight");
    const float* w2 = w(p + "out_layers.3.weight");
    const float* wsk = has_skip ? w(p + "skip_connection.weight") : nullptr;
    const float* b1 = film ? w(p + "in_layers.2.bias") : nullptr;     // else folded into the embedding table
    const int tbo = film ? -1 : L.tb_off;
    Act h1;
    if (mode == EO_MODE_FP32) {
      SrcSpec s0; s0.act = hin;
      rc = plan_conv_fp32(p + "in_layers.2", {s0}, {{w1, L.cin, 3, 0, L.cin}}, L.cout, b1, nullptr, tbo, nullptr, 1, 0, &h1, st);
    } else {
      rc = plan_conv_tc(p + "in_layers.2", {seg3x3(hin)}, {{w1, L.cin, 3, 0, L.cin}}, L.cout, nullptr, b1, nullptr, tbo, nullptr,
                        Ho, Wo, &h1, st);
    }
    if (rc) return rc;
    free_act(hin);
    GnOut g2 = plan_gn(p + "out_layers.0", h1, nullptr, w(p + "out_layers.0.weight"), w(p + "out_layers.0.bias"));
    if (film) {
      const int toff = L.tb_off, cout = L.cout;
      push(p + "out_layers.0.film", [=](int B, cudaStream_t s) -> int {
        return launch_gn_modulate(ptr<float>(g2.scale_off), ptr<float>(g2.shift_off), tb, tb_total, toff, B, cout, s);
      });
      note("k_gn_modulate", 0, 0);
    }
    const float* b2 = w(p + "out_layers.3.bias");
    const float* bsk = has_skip ? w(p + "skip_connection.bias") : nullptr;
    const Act* resid = has_skip ? nullptr : (xmat ? &xin : &a);
    if (mode == EO_MODE_FP32) {
      std::vector<SrcSpec> s2; std::vector<PackSeg> sg2;
      SrcSpec t0; t0.act = h1; t0.gn = &g2; t0.silu = 1; s2.push_back(t0);
      sg2.push_back({w2, L.cout, 3, 0, L.cout});
      if (has_skip) {
        if (xmat) { SrcSpec t1; t1.act = xin; t1.ksize = 1; s2.push_back(t1); sg2.push_back({wsk, L.cin, 1, 0, L.cin}); }
        else {
          SrcSpec t1; t1.act = a; t1.ksize = 1; s2.push_back(t1); sg2.push_back({wsk, L.cin, 1, 0, Ca});
          if (b) { SrcSpec t2; t2.act = *b; t2.ksize = 1; s2.push_back(t2); sg2.push_back({wsk, L.cin, 1, Ca, Cb}); }
        }
      }
      rc = plan_conv_fp32(p + "out_layers.3", s2, sg2, L.cout, b2, bsk, -1, resid, 1, 0, out, st);
      if (rc) return rc;
      free_gn(g2);
      free_act(h1);
    } else {
      const bool fuse2 = can_fuse_gn(3, Ho, Wo, L.cout);
      Act hn;
      std::vector<TcSegSpec> ts; std::vector<PackSeg> sg;
      if (fuse2) ts.push_back(with_gn(seg3x3(h1), g2, 0, 1));
      else { hn = plan_gn_apply(p + "out_layers.0", h1, nullptr, g2, 1); free_gn(g2); free_act(h1); ts.push_back(seg3x3(hn)); }
      sg.push_back({w2, L.cout, 3, 0, L.cout});
      if (has_skip) {
        if (xmat) { ts.push_back(seg1x1(xin)); sg.push_back({wsk, L.cin, 1, 0, L.cin}); }
        else {
          ts.push_back(seg1x1(a)); sg.push_back({wsk, L.cin, 1, 0, Ca});
          if (b) { ts.push_back(seg1x1(*b)); sg.push_back({wsk, L.cin, 1, Ca, Cb}); }
        }
      }
      rc = plan_conv_tc(p + "out_layers.3", ts, sg, L.cout, nullptr, b2, bsk, -1, resid, Ho, Wo, out, st);
      if (rc) return rc;
      if (fuse2) { free_gn(g2); free_act(h1); } else free_act(hn);
    }
    if (xmat) free_act(xin);
    return EO_OK;
  }

  // ResBlock._forward (unet_openai.py:365-385); x = cat(a, b) when b != nullptr (:773)
  int plan_res(const Layer& L, const Act& a, const Act* b, Act* out, cudaStream_t st) {
    if (cfg.use_scale_shift_norm || L.updown) return plan_res_x(L, a, b, out, st);
    const std::string& p = L.prefix;
    const int Ca = a.C, Cb = b ? b->C : 0;
    if (Ca + Cb != L.cin) { set_error("plan_res %s: channel mismatch", p.c_str()); return EO_ERR_STATE; }
    const bool has_skip = L.cin != L.cout;
    int rc;
    GnOut g1 = plan_gn(p + "in_layers.0", a, b, w(p + "in_layers.0.weight"), w(p + "in_layers.0.bias"));
    Act h1;
    const float* w1 = w(p + "in_layers.2.weight");
    const float* w2 = w(p + "out_layers.3.weight");
    const float* wsk = has_skip ? w(p + "skip_connection.weight") : nullptr;
    if (mode == EO_MODE_FP32) {
      std::vector<SrcSpec> srcs; std::vector<PackSeg> segs;
      SrcSpec s0; s0.act = a; s0.gn = &g1; s0.gn_coff = 0; s0.silu = 1; srcs.push_back(s0);
      segs.push_back({w1, L.cin, 3, 0, Ca});
      if (b) { SrcSpec s1; s1.act = *b; s1.gn = &g1; s1.gn_coff = Ca; s1.silu = 1; srcs.push_back(s1); segs.push_back({w1, L.cin, 3, Ca, Cb}); }
      // bias of conv1 is folded into the per-step embedding table (tb)
      rc = plan_conv_fp32(p + "in_layers.2", srcs, segs, L.cout, nullptr, nullptr, L.tb_off, nullptr, 1, 0, &h1, st);
      if (rc) return rc;
      free_gn(g1);
      GnOut g2 = plan_gn(p + "out_layers.0", h1, nullptr, w(p + "out_layers.0.weight"), w(p + "out_layers.0.bias"));
      std::vector<SrcSpec> s2; std::vector<PackSeg> sg2;
      SrcSpec t0; t0.act = h1; t0.gn = &g2; t0.silu = 1; s2.push_back(t0);
      sg2.push_back({w2, L.cout, 3, 0, L.cout});
      if (has_skip) {
        SrcSpec t1; t1.act = a; t1.ksize = 1; s2.push_back(t1); sg2.push_back({wsk, L.cin, 1, 0, Ca});
        if (b) { SrcSpec t2; t2.act = *b; t2.ksize = 1; s2.push_back(t2); sg2.push_back({wsk, L.cin, 1, Ca, Cb}); }
      }
      rc = plan_conv_fp32(p + "out_layers.3", s2, sg2, L.cout, w(p + "out_layers.3.bias"),
                          has_skip ? w(p + "skip_connection.bias") : nullptr, -1, has_skip ? nullptr : &a, 1, 0, out, st);
      if (rc) return rc;
      free_gn(g2);
      free_act(h1);
    } else {
      const bool fuse1 = can_fuse_gn(3, a.H, a.W, Ca) && (!b || Cb % 64 == 0);
      if (fuse1) {
        std::vector<TcSegSpec> ts; std::vector<PackSeg> sg;
        ts.push_back(with_gn(seg3x3(a), g1, 0, 1)); sg.push_back({w1, L.cin, 3, 0, Ca});
        if (b) { ts.push_back(with_gn(seg3x3(*b), g1, Ca, 1)); sg.push_back({w1, L.cin, 3, Ca, Cb}); }
        rc = plan_conv_tc(p + "in_layers.2", ts, sg, L.cout, nullptr, nullptr, nullptr, L.tb_off, nullptr, a.H, a.W, &h1, st);
        if (rc) return rc;
        free_gn(g1);
      } else {
        Act xn = plan_gn_apply(p + "in_layers.0", a, b, g1, 1);
        free_gn(g1);
        rc = plan_conv_tc(p + "in_layers.2", {seg3x3(xn)}, {{w1, L.cin, 3, 0, L.cin}}, L.cout, nullptr, nullptr, nullptr,
                          L.tb_off, nullptr, a.H, a.W, &h1, st);
        if (rc) return rc;
        free_act(xn);
      }
      GnOut g2 = plan_gn(p + "out_layers.0", h1, nullptr, w(p + "out_layers.0.weight"), w(p + "out_layers.0.bias"));
      const bool fuse2 = can_fuse_gn(3, a.H, a.W, L.cout);
      Act hn;
      std::vector<TcSegSpec> ts; std::vector<PackSeg> sg;
      if (fuse2) {
        ts.push_back(with_gn(seg3x3(h1), g2, 0, 1));
      } else {
        hn = plan_gn_apply(p + "out_layers.0", h1, nullptr, g2, 1);
        free_gn(g2);
        free_act(h1);
        ts.push_back(seg3x3(hn));
      }
      sg.push_back({w2, L.cout, 3, 0, L.cout});
      if (has_skip) {
        ts.push_back(seg1x1(a)); sg.push_back({wsk, L.cin, 1, 0, Ca});
        if (b) { ts.push_back(seg1x1(*b)); sg.push_back({wsk, L.cin, 1, Ca, Cb}); }
      }
      rc = plan_conv_tc(p + "out_layers.3", ts, sg, L.cout, nullptr, w(p + "out_layers.3.bias"),
                        has_skip ? w(p + "skip_connection.bias") : nullptr, -1, has_skip ? nullptr : &a, a.H, a.W, out, st);
      if (rc) return rc;
      if (fuse2) { free_gn(g2); free_act(h1); } else free_act(hn);
    }
    return EO_OK;
  }

  // AttentionBlock._forward (unet_openai.py:427-433)
  int plan_attn(const Layer& L, const Act& x, Act* out, cudaStream_t st) {
    const std::string& p = L.prefix;
    const int C = L.cin, heads = L.heads, ch = C / heads, T = x.H * x.W;
    if (C % heads != 0) { set_error("attention %s: %d channels not divisible by %d heads", p.c_str(), C, heads); return EO_ERR_ARG; }
    const bool new_order = cfg.use_new_attention_order != 0;
    int rc;
    GnOut g = plan_gn(p + "norm", x, nullptr, w(p + "norm.weight"), w(p + "norm.bias"));
    const float* wq = w(p + "qkv.weight"); const float* wp = w(p + "proj_out.weight");
    if (mode == EO_MODE_FP32) {
      SrcSpec s; s.act = x; s.ksize = 1; s.gn = &g; s.silu = 0;
      Act qkv;
      rc = plan_conv_fp32(p + "qkv", {s}, {{wq, C, 1, 0, C}}, 3 * C, w(p + "qkv.bias"), nullptr, -1, nullptr, 1, 0, &qkv, st);
      if (rc) return rc;
      free_gn(g);
      Act a = new_act(C, x.H, x.W);
      const int hs = new_order ? ch : 3 * ch, ps = new_order ? C : ch;
      push(p + "attention", [=](int B, cudaStream_t stx) -> int {
        return launch_attention_simt(ptr<float>(qkv.off), ptr<float>(a.off), B, T, heads, ch, 3 * C, hs, ps, stx);
      });
      note(ch > 64 ? "k_attention_wide" : "k_attention_simt", 4.0 * heads * (double)T * T * ch, 0);
      free_act(qkv);
      SrcSpec sa; sa.act = a; sa.ksize = 1;
      rc = plan_conv_fp32(p + "proj_out", {sa}, {{wp, C, 1, 0, C}}, C, w(p + "proj_out.bias"), nullptr, -1, &x, 1, 0, out, st);
      if (rc) return rc;
      free_act(a);
    } else {
      if (ch > 64 || ch % 8 != 0) {
        // head dimensions the tensor-core kernel does not take (the reference's own scripts use num_heads = 1:
        // 512 / 1024 channels per head, train.py:50, inference.py:59): qkv in the reference's channel order from
        // the tensor-core 1x1 conv, attention by the one-warp-per-query kernel in fp32 arithmetic
        Act xn = plan_gn_apply(p + "norm", x, nullptr, g, 0);
        free_gn(g);
        Act qkv;
        rc = plan_conv_tc(p + "qkv", {seg1x1(xn)}, {{wq, C, 1, 0, C}}, 3 * C, nullptr, w(p + "qkv.bias"), nullptr, -1, nullptr,
                          x.H, x.W, &qkv, st, /*want_stats=*/false);
        if (rc) return rc;
        free_act(xn);
        Act a = new_act(C, x.H, x.W);
        const int hs = new_order ? ch : 3 * ch, ps = new_order ? C : ch;
        push(p + "attention", [=](int B, cudaStream_t stx) -> int {
          return launch_attention_wide(ptr(qkv.off), ptr(a.off), DT_BF16, B, T, heads, ch, 3 * C, hs, ps, stx);
        });
        note("k_attention_wide", 4.0 * heads * (double)T * T * ch, 0);
        free_act(qkv);
        rc = plan_conv_tc(p + "proj_out", {seg1x1(a)}, {{wp, C, 1, 0, C}}, C, nullptr, w(p + "proj_out.bias"), nullptr, -1, &x,
                          x.H, x.W, out, st);
        if (rc) return rc;
        free_act(a);
        return EO_OK;
      }
      const bool fuse = can_fuse_gn(1, x.H, x.W, C);
      Act xn;
      if (!fuse) { xn = plan_gn_apply(p + "norm", x, nullptr, g, 0); free_gn(g); }
      // qkv rows re-ordered to [head][q|k|v][64] with zero rows padding each part to 64
      const int rows = heads * 3 * 64;
      std::vector<int> rmap(rows, -1);
      for (int h = 0; h < heads; ++h)
        for (int part = 0; part < 3; ++part)
          for (int c2 = 0; c2 < ch; ++c2)
            rmap[(h * 3 + part) * 64 + c2] = new_order ? part * C + h * ch + c2 : h * 3 * ch + part * ch + c2;
      int* d_rmap = nullptr;
      rc = dmalloc(&d_rmap, (size_t)rows);
      if (rc) return rc;
      EO_CHECK_CUDA(cudaMemcpyAsync(d_rmap, rmap.data(), rows * sizeof(int), cudaMemcpyHostToDevice, st));
      EO_CHECK_CUDA(cudaStreamSynchronize(st));   // rmap is a stack-lifetime host buffer
      // head dimension <= 48: the logit scale ch^-1/2 (and log2 e: the softmax runs on exp2) is folded into the q rows of
      // the projection, weights and bias, in fp32 before they are rounded -- the attention kernel gets S * scale out
      // of the tensor core and q is still rounded to bf16 exactly once
      const bool k_one = ch <= 48;
      float* d_rscale = nullptr;
      if (k_one) {
        std::vector<float> rs(rows, 1.0f);
        const float sl2 = (1.0f / sqrtf((float)ch)) * 1.4426950408889634f;
        for (int h = 0; h < heads; ++h)
          for (int c2 = 0; c2 < 64; ++c2) rs[(h * 3 + 0) * 64 + c2] = sl2;
        rc = dmalloc(&d_rscale, (size_t)rows);
        if (rc) return rc;
        EO_CHECK_CUDA(cudaMemcpyAsync(d_rscale, rs.data(), rows * sizeof(float), cudaMemcpyHostToDevice, st));
        EO_CHECK_CUDA(cudaStreamSynchronize(st));   // rs is a stack-lifetime host buffer
      }
      Act qkv;
      pack_row_scale = d_rscale;
      rc = plan_conv_tc(p + "qkv", {fuse ? with_gn(seg1x1(x), g, 0, 0) : seg1x1(xn)}, {{wq, C, 1, 0, C}}, rows, d_rmap,
                        w(p + "qkv.bias"), nullptr, -1, nullptr, x.H, x.W, &qkv, st, /*want_stats=*/false, nullptr,
                        /*alg_flops=*/2.0 * x.H * x.W * 3.0 * C * C);
      pack_row_scale = nullptr;
      if (rc) return rc;
      if (fuse) free_gn(g);
      // ... and the first padded channel behind round16(ch) of every head's k becomes 1.0 (zero weight row, bias 1):
      // the attention kernel keeps -m (its running reference maximum) in the same channel of q, so the tensor core
      // hands it S * scale - m (tc_attn.cu)
      if (k_one) {
        float* qb = last_packed_bias;
        const float one = 1.0f;
        const int on = (ch + 15) & ~15;
        for (int h = 0; h < heads; ++h)
          EO_CHECK_CUDA(cudaMemcpyAsync(qb + (h * 3 + 1) * 64 + on, &one, sizeof(float), cudaMemcpyHostToDevice, st));
        EO_CHECK_CUDA(cudaStreamSynchronize(st));   // `one` is a stack variable
      }
      if (!fuse) free_act(xn);
      Act a = new_act(C, x.H, x.W);
      const size_t ai = attn_plans.size();
      attn_plans.push_back(nullptr);
      auto prepare = [=]() -> int {
        TcAttnParams ap; ap.qkv = ptr(qkv.off); ap.out = ptr(a.off); ap.B = Bmax; ap.T = T; ap.heads = heads; ap.ch = ch;
        ap.k_one = k_one ? 1 : 0;
        return tc_attn_plan_create(ap, &attn_plans[ai]);
      };
      push(p + "attention", [=](int B, cudaStream_t stx) -> int { return tc_attn_launch(attn_plans[ai], B, stx); }, 1, prepare);
      // executed: S over round16(ch) (+ 16 with the baked maximum) channels, P V over round16(ch) (64 for wide heads)
      note("k_attn_tc6", 4.0 * heads * (double)T * T * ch, 0,
           k_one ? 2.0 * heads * (double)T * T * (2 * ((ch + 15) & ~15) + 16) : 4.0 * heads * (double)T * T * 64);
      free_act(qkv);
      rc = plan_conv_tc(p + "proj_out", {seg1x1(a)}, {{wp, C, 1, 0, C}}, C, nullptr, w(p + "proj_out.bias"), nullptr, -1, &x,
                        x.H, x.W, out, st);
      if (rc) return rc;
      free_act(a);
    }
    return EO_OK;
  }

  // Downsample.forward (unet_openai.py:269-271): conv 3x3 stride 2 pad 1
  int plan_down(const Layer& L, const Act& x, Act* out, cudaStream_t st) {
    const std::string p = L.prefix + "op.";
    const float* wd = w(p + "weight");
    if (x.H % 2 || x.W % 2) { set_error("downsample: odd feature map %dx%d", x.H, x.W); return EO_ERR_ARG; }
    if (mode == EO_MODE_FP32) {
      SrcSpec s; s.act = x; s.ksize = 3;
      return plan_conv_fp32(L.prefix + "op", {s}, {{wd, L.cin, 3, 0, L.cin}}, L.cout, w(p + "bias"), nullptr, -1, nullptr, 2, 0, out, st);
    }
    // output pixel (h, w) reads source pixels (2h + kh - 1, 2w + kw - 1): nine taps fetched straight from the
    // full-resolution tensor by a tensor map with traversal stride 2 (out-of-image rows / columns zero-filled by the
    // TMA unit = the conv's padding) -- no space-to-depth copy of the input
    TcSegSpec s = seg3x3(x);
    s.stride = 2;
    return plan_conv_tc(L.prefix + "op", {s}, {{wd, L.cin, 3, 0, L.cin}}, L.cout, nullptr, w(p + "bias"), nullptr, -1, nullptr,
                        x.H / 2, x.W / 2, out, st);
  }

  // Upsample.forward (unet_openai.py:229-242): nearest x2 then conv 3x3
  int plan_up(const Layer& L, const Act& x, Act* out, cudaStream_t st) {
    const std::string p = L.prefix + "conv.";
    const float* wu = w(p + "weight");
    if (x.H == 3 && x.W == 3) { set_error("upsample: the reference's 3x3 -> 7x7 pad quirk is not implemented"); return EO_ERR_ARG; }
    if (mode == EO_MODE_FP32) {
      SrcSpec s; s.act = x; s.ksize = 3;
      return plan_conv_fp32(L.prefix + "conv", {s}, {{wu, L.cin, 3, 0, L.cin}}, L.cout, w(p + "bias"), nullptr, -1, nullptr, 1, 1, out, st);
    }
    if (tc_conv_stats_supported(x.H, x.W)) {
      // Four sub-pixel convolutions over the low-resolution input, one per output parity (a, b), with the
      // 3x3 taps that read the same source pixel summed (k_fold_upsample_weight): 4/9 of the FLOPs and no
      // materialised upsampled tensor.  Each writes its quarter of the output through a strided view.
      const int C = x.C, Ho = x.H * 2, Wo = x.W * 2;
      Act o = new_act(L.cout, Ho, Wo);
      o.stats = (long long)ch_stats_floats;
      ch_stats_floats += (size_t)Bmax * L.cout * 2;
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
          float* wf = nullptr;
          int rc = dmalloc(&wf, (size_t)L.cout * C * 4);
          if (rc) return rc;
          if ((rc = launch_fold_upsample_weight(wu, L.cout, C, a, b, wf, st))) return rc;
          TcSegSpec sgm; sgm.act = x; sgm.ntaps = 4;
          for (int t = 0; t < 4; ++t) { sgm.dh[t] = (int8_t)(a - 1 + t / 2); sgm.dw[t] = (int8_t)(b - 1 + t % 2); sgm.plane[t] = 0; }
          // reported FLOPs stay the algorithmic ones of the 3x3 conv on the upsampled grid (a quarter per launch);
          // the launch executes 4/9 of them
          OutView vw{o, ((long long)a * Wo + b) * L.cout, 2LL * L.cout, 2LL * Wo * L.cout, (long long)Ho * Wo * L.cout,
                     2.0 * x.H * x.W * L.cout * 9.0 * C};
          Act dummy;
          rc = plan_conv_tc(L.prefix + "conv.p" + std::to_string(a * 2 + b), {sgm}, {{wf, C, 2, 0, C}}, L.cout, nullptr,
                            w(p + "bias"), nullptr, -1, nullptr, x.H, x.W, &dummy, st, true, &vw);
          if (rc) return rc;
        }
      *out = o;
      return EO_OK;
    }
    Act up = new_act(x.C, x.H * 2, x.W * 2);
    Act xx = x;
    push(L.prefix + "interp", [=](int B, cudaStream_t stx) -> int {
      return launch_upsample2x(ptr(xx.off), ptr(up.off), B, xx.H, xx.W, xx.C, stx);
    });
    note("k_upsample2x", 0, 5.0 * x.H * x.W * x.C * 2);
    int rc = plan_conv_tc(L.prefix + "conv", {seg3x3(up)}, {{wu, L.cin, 3, 0, L.cin}}, L.cout, nullptr, w(p + "bias"), nullptr, -1,
                          nullptr, x.H * 2, x.W * 2, out, st);
    if (rc) return rc;
    free_act(up);
    return EO_OK;
  }

  // Plans the layers of one block.  `h` is the block input; `skip` (decoder only) is the
  // tensor th.cat appends to it (unet_openai.py:773).  `free_input`: the block input dies with
  // its last reader (false while it still sits on the skip stack `hs`).  `skip`, when given,
  // was popped from the stack and always dies here.
  int run_layers(const Block& blk, Act h, const Act* skip, bool free_input, Act* out, cudaStream_t st) {
    Act cur = h;
    bool first = true;
    for (const Layer& L : blk.layers) {
      Act nxt; int rc = EO_OK;
      switch (L.kind) {
        case L_RES: rc = plan_res(L, cur, first ? skip : nullptr, &nxt, st); break;
        case L_ATTN: rc = plan_attn(L, cur, &nxt, st); break;
        case L_DOWN: rc = plan_down(L, cur, &nxt, st); break;
        case L_UP: rc = plan_up(L, cur, &nxt, st); break;
        default: set_error("unexpected layer kind"); rc = EO_ERR_STATE;
      }
      if (rc) return rc;
      if (!first || free_input) free_act(cur);
      if (first && skip) { Act s = *skip; free_act(s); }
      cur = nxt; first = false;
    }
    *out = cur;
    return EO_OK;
  }

  // stem weights re-packed for an (x, cond) channel split: K order (x taps, cond taps)
  std::map<int, float*> stem_split;
  float* stem_split_weights(int cx, int cc, cudaStream_t st) {
    auto it = stem_split.find(cx);
    if (it != stem_split.end()) return it->second;
    const Layer& L = in_blocks[0].layers[0];
    float* Wp = nullptr; int K = 0;
    int rc = pack_simt({{w(L.prefix + "weight"), L.cin, 3, 0, cx}, {w(L.prefix + "weight"), L.cin, 3, cx, cc}},
                       L.cout, &Wp, &K, st);
    if (rc) return nullptr;
    stem_split[cx] = Wp;
    return Wp;
  }

  int finalize(int mode_, int Bmax_, int H_, int W_, cudaStream_t st);
  int forward(const float* x, int Cx, const float* cond, int Cc, const int64_t* t, const int64_t* y, float* out, int B, cudaStream_t st,
              float* ms_per_op = nullptr);
};

int eo_unet::finalize(int mode_, int Bmax_, int H_, int W_, cudaStream_t st) {
  release_plan();
  for (auto& ws : wspecs)
    EO_REQUIRE(ws.src != nullptr, EO_ERR_STATE, "finalize: weight '%s' was never set", ws.name.c_str());
  EO_REQUIRE(mode_ == EO_MODE_FP32 || mode_ == EO_MODE_BF16, EO_ERR_ARG, "finalize: unknown mode %d", mode_);
  EO_REQUIRE(Bmax_ > 0 && H_ > 0 && W_ > 0, EO_ERR_ARG, "finalize: bad geometry");
  const int levels = cfg.n_channel_mult;
  EO_REQUIRE(H_ % (1 << (levels - 1)) == 0 && W_ % (1 << (levels - 1)) == 0, EO_ERR_ARG,
             "finalize: %dx%d is not divisible by 2^%d", H_, W_, levels - 1);
  if (mode_ == EO_MODE_BF16)
    EO_REQUIRE(cfg.model_channels % 64 == 0, EO_ERR_ARG,
               "bf16 mode needs model_channels %% 64 == 0 (got %d)", cfg.model_channels);
  mode = mode_; Bmax = Bmax_; H = H_; W = W_;
  act_dt = mode == EO_MODE_FP32 ? DT_F32 : DT_BF16;
#ifdef EO_DEVTOOLS
  arena.keep = std::getenv("EO_DEBUG_KEEP") != nullptr;
#endif
  int rc;
  // private copies of every parameter the kernels read directly at run time (GroupNorm affine,
  // biases, Linear / embedding matrices); convolution weights are only read by the packers below
  for (auto& ws : wspecs) {
    ws.priv = nullptr;
    // (the stem weight too: its (x, cond) split is packed at the first forward that uses it, when the caller's
    // staging copy may be gone)
    if (ws.shape.size() > 2 && ws.name != in_blocks[0].layers[0].prefix + "weight") continue;
    size_t n = 1;
    for (int64_t d : ws.shape) n *= (size_t)d;
    float* p = nullptr;
    if ((rc = dmalloc(&p, n))) return rc;
    EO_CHECK_CUDA(cudaMemcpyAsync(p, ws.src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    ws.priv = p;
  }

  // ---- timestep-embedding path (always fp32; unet_openai.py:763, :374)
  const int mc = cfg.model_channels;
  if ((rc = dmalloc(&e0, (size_t)Bmax * mc))) return rc;
  if ((rc = dmalloc(&l1, (size_t)Bmax * ted))) return rc;
  if ((rc = dmalloc(&emb, (size_t)Bmax * ted))) return rc;
  if ((rc = dmalloc(&tb, (size_t)Bmax * tb_total))) return rc;
  if ((rc = dmalloc(&w_emb_cat, (size_t)tb_total * ted))) return rc;
  if ((rc = dmalloc(&b_emb_cat, (size_t)tb_total))) return rc;
  auto cat_emb = [&](const Block& blk) -> int {
    for (const Layer& L : blk.layers) {
      if (L.kind != L_RES) continue;
      const bool film = cfg.use_scale_shift_norm != 0;
      const int erows = (film ? 2 : 1) * L.cout;
      EO_CHECK_CUDA(cudaMemcpyAsync(w_emb_cat + (size_t)L.tb_off * ted, w(L.prefix + "emb_layers.1.weight"),
                                    (size_t)erows * ted * sizeof(float), cudaMemcpyDeviceToDevice, st));
      // emb bias + the bias of the conv that the embedding is added to (in_layers.2); with FiLM the row is
      // (scale | shift) and the conv keeps its own bias
      int r = launch_pack_bias(w(L.prefix + "emb_layers.1.bias"), film ? nullptr : w(L.prefix + "in_layers.2.bias"),
                               b_emb_cat + L.tb_off, erows, nullptr, st);
      if (r) return r;
    }
    return EO_OK;
  };
  for (auto& b : in_blocks) if ((rc = cat_emb(b))) return rc;
  if ((rc = cat_emb(mid))) return rc;
  for (auto& b : out_blocks) if ((rc = cat_emb(b))) return rc;
  {
    const float* freqs = w("time_embed.freqs");
    const float *w0 = w("time_embed.0.weight"), *b0 = w("time_embed.0.bias");
    const float *w2 = w("time_embed.2.weight"), *b2 = w("time_embed.2.bias");
    const float* lab = cfg.num_classes > 0 ? w("label_emb.weight") : nullptr;
    push("time_embed", [=](int B, cudaStream_t s) -> int {
      // a precomputed table serves every forward without class labels (label_emb(y) is added BEFORE the
      // projections, :764-766, so those rows depend on (t, y))
      if (tt_on && tt_table && !io_y) return launch_gather_rows(tt_table, tt_n, tb_total, io_t, B, tb, s);
      int r = launch_sinusoid(io_t, freqs, B, mc / 2, e0, s);
      if (r) return r;
      if ((r = launch_linear(e0, w0, b0, nullptr, nullptr, nullptr, 0, B, mc, ted, l1, s))) return r;
      if ((r = launch_linear(l1, w2, b2, nullptr, lab, io_y, 1, B, ted, ted, emb, s))) return r;
      return launch_linear(emb, w_emb_cat, b_emb_cat, nullptr, nullptr, nullptr, 1, B, ted, tb_total, tb, s);
    }, 4);
    note("k_linear", 2.0 * ((double)mc * ted + (double)ted * ted + (double)ted * tb_total), 0);
  }

  // ---- stem: conv3x3 over cat(x, cond), NCHW fp32 -> NHWC (unet_openai.py:754-756, :608)
  Act h;
  const Layer& L0 = in_blocks[0].layers[0];
  if (mode == EO_MODE_BF16 && 18 * L0.cin <= 64 && L0.cout % 64 == 0) {
    // few input channels: the 3x3 windows (bf16 value + rounding residual of every element, exact to 2^-17)
    // become one 64-channel pixel, and the stem a 64-deep 1x1 convolution on the tensor cores, GroupNorm
    // statistics of its output included
    Act col = new_act(64, H, W);
    const int cin = L0.cin;
    push("input_blocks.0.im2col", [=](int B, cudaStream_t s) -> int {
      if (io_cx + io_cc != cin) { set_error("stem: %d + %d input channels, expected %d", io_cx, io_cc, cin); return EO_ERR_ARG; }
      return launch_stem_im2col(io_x, io_cx, io_cond, io_cc, ptr(col.off), B, H, W, s);
    });
    note("k_stem_im2col", 0, (double)H * W * (64 * 2 + cin * 4));
    float* w2 = nullptr;
    if ((rc = dmalloc(&w2, (size_t)L0.cout * 64))) return rc;
    if ((rc = launch_stem_weight(w(L0.prefix + "weight"), L0.cout, cin, w2, st))) return rc;
    if ((rc = plan_conv_tc("input_blocks.0", {seg1x1(col)}, {{w2, 64, 1, 0, 64}}, L0.cout, nullptr, w(L0.prefix + "bias"), nullptr,
                           -1, nullptr, H, W, &h, st, true, nullptr, 2.0 * H * W * L0.cout * 9 * cin)))
      return rc;
    free_act(col);
  } else if (mode == EO_MODE_BF16 && 2 * L0.cin <= 64 && L0.cout % 64 == 0) {
    // up to 32 input channels (the multispectral concat configuration: 13 + 15): the input becomes one 64-channel
    // bf16 NHWC block (values + rounding residuals) and the stem an ordinary 3x3 tensor-core convolution
    Act xin = new_act(64, H, W);
    const int cin = L0.cin;
    push("input_blocks.0.nhwc", [=](int B, cudaStream_t s) -> int {
      if (io_cx + io_cc != cin) { set_error("stem: %d + %d input channels, expected %d", io_cx, io_cc, cin); return EO_ERR_ARG; }
      return launch_stem_nhwc(io_x, io_cx, io_cond, io_cc, ptr(xin.off), B, H, W, s);
    });
    note("k_stem_nhwc", 0, (double)H * W * (64 * 2 + cin * 4));
    float* w2 = nullptr;
    if ((rc = dmalloc(&w2, (size_t)L0.cout * 64 * 9))) return rc;
    if ((rc = launch_stem_weight3(w(L0.prefix + "weight"), L0.cout, cin, w2, st))) return rc;
    if ((rc = plan_conv_tc("input_blocks.0", {seg3x3(xin)}, {{w2, 64, 3, 0, 64}}, L0.cout, nullptr, w(L0.prefix + "bias"), nullptr,
                           -1, nullptr, H, W, &h, st, true, nullptr, 2.0 * H * W * L0.cout * 9 * cin)))
      return rc;
    free_act(xin);
  } else
  {
    const Layer& L = in_blocks[0].layers[0];
    float* Wp = nullptr; int K = 0;
    // one K segment per possible (x, cond) split is not known yet: pack as a single segment
    // of in_channels and address the two NCHW sources by channel offset at launch.
    if ((rc = pack_simt({{w(L.prefix + "weight"), L.cin, 3, 0, L.cin}}, L.cout, &Wp, &K, st))) return rc;
    const float* bias = w(L.prefix + "bias");
    h = new_act(L.cout, H, W);
    const int cin = L.cin, cout = L.cout;
    const int odt = act_dt;
    // the packed K order is (tap, channel); with two sources the kernel needs per-source
    // segments, so a second packing ordered (x taps, cond taps) is built lazily per split
    const bool stem_kernel = (size_t)9 * cin * 64 * sizeof(float) <= 200 * 1024;
    push("input_blocks.0", [=](int B, cudaStream_t s) -> int {
      const float* Wuse = Wp;
      if (io_cc != 0) {
        Wuse = stem_split_weights(io_cx, io_cc, s);
        if (!Wuse) return EO_ERR_CUDA;
      }
      if (stem_kernel)
        return launch_conv_stem(io_x, io_cx, io_cond, io_cc, Wuse, bias, ptr(h.off), odt, B, H, W, cout, s);
      ConvSimtParams p;
      p.B = B; p.Hin = H; p.Win = W; p.Hout = H; p.Wout = W; p.stride = 1; p.up = 0;
      p.Cout = cout; p.bias = bias; p.out = ptr(h.off); p.out_dt = odt;
      p.W = Wuse; p.Ktot = K;
      p.nsrc = io_cc == 0 ? 1 : 2;
      p.src[0].ptr = io_x; p.src[0].C = io_cx; p.src[0].ksize = 3; p.src[0].dt = DT_F32; p.src[0].nchw = 1; p.src[0].w_off = 0;
      if (io_cc != 0) {
        p.src[1].ptr = io_cond; p.src[1].C = io_cc; p.src[1].ksize = 3; p.src[1].dt = DT_F32; p.src[1].nchw = 1; p.src[1].w_off = 9 * io_cx;
      }
      return launch_conv_simt(p, s);
    });
    (void)cin;
    note(stem_kernel ? "k_conv_stem" : "k_conv_simt", 2.0 * H * W * cout * K, 0);
    if (mode == EO_MODE_BF16 && cout % 8 == 0) {
      h.stats = (long long)ch_stats_floats;
      ch_stats_floats += (size_t)Bmax * cout * 2;
      Act hc = h;
      push("input_blocks.0.stats", [=](int B, cudaStream_t s) -> int {
        return launch_chan_stats(ptr(hc.off), B, hc.H * hc.W, hc.C, ch_stats + hc.stats, s);
      });
      note("k_chan_stats", 0, (double)H * W * cout * 2);
    }
  }
  named[in_blocks[0].name] = h;

  // ---- encoder
  std::vector<Act> hs;
  hs.push_back(h);
  for (size_t i = 1; i < in_blocks.size(); ++i) {
    Act o;
    if ((rc = run_layers(in_blocks[i], h, nullptr, false, &o, st))) return rc;
    h = o; hs.push_back(h); named[in_blocks[i].name] = h;
  }
  // ---- middle: its input is hs.back(), still needed by the decoder
  {
    Act o;
    if ((rc = run_layers(mid, h, nullptr, false, &o, st))) return rc;
    h = o; named[mid.name] = h;
  }
  // ---- decoder: h = cat(h, hs.pop()) (:773), both consumed by the block's ResBlock
  for (size_t i = 0; i < out_blocks.size(); ++i) {
    Act skip = hs.back(); hs.pop_back();
    Act o;
    if ((rc = run_layers(out_blocks[i], h, &skip, true, &o, st))) return rc;
    h = o; named[out_blocks[i].name] = h;
  }
  // ---- head: conv3x3(SiLU(GN(h))) -> NCHW fp32 (unet_openai.py:739-743, :780)
  if (mode == EO_MODE_BF16 && cfg.out_channels <= 64 && can_fuse_gn(3, H, W, final_ch)) {
    // on the tensor cores with the output channels padded to one 64-wide tile (zero weight rows) and
    // GroupNorm + SiLU folded into the operand load; the epilogue writes the real channels as fp32 NCHW
    GnOut g = plan_gn("out.0", h, nullptr, w("out.0.weight"), w("out.0.bias"));
    const int cout = cfg.out_channels, C = final_ch;
    std::vector<int> rmap(64, -1);
    for (int n = 0; n < cout; ++n) rmap[n] = n;
    int* d_rmap = nullptr;
    if ((rc = dmalloc(&d_rmap, (size_t)64))) return rc;
    EO_CHECK_CUDA(cudaMemcpyAsync(d_rmap, rmap.data(), 64 * sizeof(int), cudaMemcpyHostToDevice, st));
    EO_CHECK_CUDA(cudaStreamSynchronize(st));   // rmap is a stack-lifetime host buffer
    Act none;
    if ((rc = plan_conv_tc("out", {with_gn(seg3x3(h), g, 0, 1)}, {{w("out.2.weight"), C, 3, 0, C}}, 64, d_rmap, w("out.2.bias"),
                           nullptr, -1, nullptr, H, W, &none, st, false, nullptr, 2.0 * H * W * cout * 9 * C, cout)))
      return rc;
    free_gn(g);
  } else {
    GnOut g = plan_gn("out.0", h, nullptr, w("out.0.weight"), w("out.0.bias"));
    float* Wp = nullptr; int K = 0;
    if ((rc = pack_simt({{w("out.2.weight"), final_ch, 3, 0, final_ch}}, cfg.out_channels, &Wp, &K, st))) return rc;
    const float* bias = w("out.2.bias");
    const int cout = cfg.out_channels, C = final_ch;
    const int adt = act_dt;
    const bool small = (mode == EO_MODE_BF16) && cout <= 16 && (size_t)(9 * C * (cout <= 4 ? 4 : 16) + 2 * C) * 4 <= 200 * 1024;
    // bf16 mode: GroupNorm + SiLU applied once per element by the bandwidth pass (the conv would
    // otherwise redo it for each of its 9 taps), then a plain small-N conv
    Act hh = small ? plan_gn_apply("out.0", h, nullptr, g, 1) : h;
    push("out", [=](int B, cudaStream_t s) -> int {
      if (small) {
        ConvSmallNParams p;
        p.x = ptr(hh.off); p.dt = DT_BF16; p.gn_scale = nullptr; p.gn_shift = nullptr;
        p.B = B; p.H = H; p.W = W; p.C = C; p.Cout = cout; p.Wp = Wp; p.bias = bias; p.out_nchw = io_out;
        return launch_conv_small_n(p, s);
      }
      ConvSimtParams p;
      p.nsrc = 1;
      p.src[0].ptr = ptr(hh.off); p.src[0].C = C; p.src[0].ksize = 3; p.src[0].dt = adt;
      p.src[0].gn_scale = ptr<float>(g.scale_off); p.src[0].gn_shift = ptr<float>(g.shift_off); p.src[0].gn_ld = C;
      p.src[0].silu = 1; p.src[0].w_off = 0;
      p.B = B; p.Hin = H; p.Win = W; p.Hout = H; p.Wout = W; p.stride = 1; p.up = 0;
      p.W = Wp; p.Ktot = K; p.Cout = cout; p.bias = bias; p.out = io_out; p.out_dt = DT_F32; p.out_nchw = 1;
      return launch_conv_simt(p, s);
    });
    note(small ? "k_conv_small_n" : "k_conv_simt", 2.0 * H * W * cout * K, 0);
  }

  // ---- allocate persistent GN accumulators and the arena, then prepare TC plans
  if ((rc = dmalloc(&gn_sums, (size_t)std::max(n_gn, 1) * Bmax * 64))) return rc;
  if ((rc = dmalloc(&ch_stats, std::max<size_t>(ch_stats_floats, 4)))) return rc;
  {
    void* p = nullptr;
    EO_CHECK_CUDA(cudaMalloc(&p, std::max<size_t>(arena.peak, 1024)));
    arena_base = reinterpret_cast<uint8_t*>(p);
    dev_bytes += (int64_t)arena.peak;
  }
  for (auto& op : ops)
    if (op.prepare) { rc = op.prepare(); if (rc) return rc; }
  EO_CHECK_CUDA(cudaStreamSynchronize(st));   // weight packing done before callers may free sources
  finalized = true;
  return EO_OK;
}

// The embedding path of `n` timestep values 0 .. n-1 with the SAME kernels the per-step path uses (k_linear
// reduces each output element on its own, so a row does not depend on what else is in the batch: table rows are
// bit-identical to the rows the per-step path would produce).
int eo_unet::build_time_tables(int n, cudaStream_t st) {
  EO_REQUIRE(finalized, EO_ERR_STATE, "eo_unet_build_time_tables before eo_unet_finalize");
  EO_REQUIRE(n >= 1 && n <= (1 << 20), EO_ERR_ARG, "eo_unet_build_time_tables: %d timesteps", n);
  EO_REQUIRE(tb_total % 4 == 0, EO_ERR_ARG, "eo_unet_build_time_tables: table width %d", tb_total);
  if (tt_table && tt_n >= n) { tt_on = true; return EO_OK; }
  release_time_tables();
  release_graphs_only();            // graphs captured with the old table's address
  const int mc = cfg.model_channels;
  float *t_e0 = nullptr, *t_l1 = nullptr, *t_emb = nullptr;
  int64_t* t_idx = nullptr;
  auto cleanup = [&]() { cudaFree(t_e0); cudaFree(t_l1); cudaFree(t_emb); cudaFree(t_idx); };
  cudaError_t e = cudaMalloc(&tt_table, (size_t)n * tb_total * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&t_e0, (size_t)n * mc * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&t_l1, (size_t)n * ted * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&t_emb, (size_t)n * ted * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&t_idx, (size_t)n * sizeof(int64_t));
  if (e != cudaSuccess) {
    cleanup(); cudaFree(tt_table); tt_table = nullptr;
    set_error("eo_unet_build_time_tables: %s", cudaGetErrorString(e));
    return EO_ERR_CUDA;
  }
  tt_n = n; tt_on = true;
  dev_bytes += (int64_t)n * tb_total * sizeof(float);
  int rc = launch_iota64(t_idx, n, st);
  if (!rc) rc = launch_sinusoid(t_idx, w("time_embed.freqs"), n, mc / 2, t_e0, st);
  if (!rc) rc = launch_linear(t_e0, w("time_embed.0.weight"), w("time_embed.0.bias"), nullptr, nullptr, nullptr, 0, n, mc, ted, t_l1, st);
  if (!rc) rc = launch_linear(t_l1, w("time_embed.2.weight"), w("time_embed.2.bias"), nullptr, nullptr, nullptr, 1, n, ted, ted, t_emb, st);
  if (!rc) rc = launch_linear(t_emb, w_emb_cat, b_emb_cat, nullptr, nullptr, nullptr, 1, n, ted, tb_total, tt_table, st);
  e = cudaStreamSynchronize(st);          // the scratch buffers die here
  cleanup();
  if (!rc && e != cudaSuccess) { set_error("eo_unet_build_time_tables: %s", cudaGetErrorString(e)); rc = EO_ERR_CUDA; }
  if (rc) release_time_tables();
  return rc;
}

int eo_unet::forward(const float* x, int Cx, const float* cond, int Cc, const int64_t* t, const int64_t* y,
                     float* out, int B, cudaStream_t st, float* ms_per_op) {
  EO_REQUIRE(finalized, EO_ERR_STATE, "eo_unet_forward before eo_unet_finalize");
  EO_REQUIRE(x && t && out, EO_ERR_ARG, "eo_unet_forward: null pointer");
  EO_REQUIRE(B > 0 && B <= Bmax, EO_ERR_ARG, "eo_unet_forward: batch %d exceeds planned batch %d", B, Bmax);
  EO_REQUIRE((cond != nullptr) == (Cc > 0), EO_ERR_ARG, "eo_unet_forward: cond/Cc mismatch");
  EO_REQUIRE(Cx + Cc == cfg.in_channels, EO_ERR_ARG,
             "eo_unet_forward: x has %d channels + cond %d != in_channels %d", Cx, Cc, cfg.in_channels);
  // reference assert, unet_openai.py:758-760
  EO_REQUIRE((y != nullptr) == (cfg.num_classes > 0), EO_ERR_ARG,
             "must specify y if and only if the model is class-conditional");
  if (!ms_per_op) {
    bool handled = false;
    int rc = forward_graph(x, Cx, cond, Cc, t, y, out, B, st, &handled);
    if (rc || handled) return rc;
  }
  io_x = x; io_cx = Cx; io_cond = cond; io_cc = Cc; io_t = t; io_y = y; io_out = out;
  if (!ms_per_op) return run_ops(B, st);
  EO_CHECK_CUDA(cudaMemsetAsync(gn_sums, 0, (size_t)std::max(n_gn, 1) * Bmax * 64 * sizeof(double), st));
  if (ch_stats_floats) EO_CHECK_CUDA(cudaMemsetAsync(ch_stats, 0, ch_stats_floats * sizeof(double), st));
  // profiling variant: one CUDA event between consecutive ops, on `st`
  std::vector<cudaEvent_t> ev(ops.size() + 1);
  for (auto& e : ev) EO_CHECK_CUDA(cudaEventCreate(&e));
  EO_CHECK_CUDA(cudaEventRecord(ev[0], st));
  int rc = EO_OK;
  for (size_t i = 0; i < ops.size() && !rc; ++i) {
    rc = ops[i].run(B, st);
    if (rc) {
      std::string msg = std::string("op '") + ops[i].name + "': " + get_error();
      set_error("%s", msg.c_str());
    } else {
      cudaEventRecord(ev[i + 1], st);
    }
  }
  if (!rc && cudaEventSynchronize(ev.back()) != cudaSuccess) {
    set_error("eo_unet_forward_timed: %s", cudaGetErrorString(cudaGetLastError()));
    rc = EO_ERR_CUDA;
  }
  for (size_t i = 0; i < ops.size() && !rc; ++i) cudaEventElapsedTime(&ms_per_op[i], ev[i], ev[i + 1]);
  for (auto& e : ev) cudaEventDestroy(e);
  return rc;
}

// the forward's launches, reading io_* (zeroing of the GroupNorm accumulators included)
int eo_unet::run_ops(int B, cudaStream_t st) {
  EO_CHECK_CUDA(cudaMemsetAsync(gn_sums, 0, (size_t)std::max(n_gn, 1) * Bmax * 64 * sizeof(double), st));
  if (ch_stats_floats) EO_CHECK_CUDA(cudaMemsetAsync(ch_stats, 0, ch_stats_floats * sizeof(double), st));
  // the two memsets above are not kernels: the first op launches the ordinary way, every later kernel of the chain may be
  // scheduled behind its predecessor programmatically (common.cuh)
  for (size_t i = 0; i < ops.size(); ++i) {
    pdl_set_allowed(i > 0);
    int rc = ops[i].run(B, st);
    if (rc) {
      pdl_set_allowed(false);
      std::string msg = std::string("op '") + ops[i].name + "': " + get_error();
      set_error("%s", msg.c_str());
      return rc;
    }
  }
  pdl_set_allowed(false);
  return EO_OK;
}

// Launch-bound sizes (the reference's 64 x 64, batch 1 case: ~175 launches of a few microseconds of work each):
// the first call of a shape runs eagerly (lazy packing happens there), the second is captured into a CUDA graph
// over staging buffers, later ones are four small copies + one graph launch.  *handled = false: run eagerly.
int eo_unet::forward_graph(const float* x, int Cx, const float* cond, int Cc, const int64_t* t, const int64_t* y,
                           float* out, int B, cudaStream_t st, bool* handled) {
  *handled = false;
  const long long key = ((((long long)B * 64 + Cx) * 64 + Cc) * 2 + (y ? 1 : 0)) * 2 + (tt_on ? 1 : 0);
  GraphSlot& g = graphs[key];
  if (g.failed || ++g.seen == 1) return EO_OK;
  const size_t hw = (size_t)H * W;
  if (!gs_x) {
    int rc;
    if ((rc = dmalloc(&gs_x, (size_t)Bmax * cfg.in_channels * hw))) return rc;
    if ((rc = dmalloc(&gs_cond, (size_t)Bmax * cfg.in_channels * hw))) return rc;
    if ((rc = dmalloc(&gs_out, (size_t)Bmax * cfg.out_channels * hw))) return rc;
    if ((rc = dmalloc(&gs_t, (size_t)Bmax))) return rc;
    if ((rc = dmalloc(&gs_y, (size_t)Bmax))) return rc;
    EO_CHECK_CUDA(cudaStreamCreateWithFlags(&graph_stream, cudaStreamNonBlocking));
  }
  if (!g.exec) {
    io_x = gs_x; io_cx = Cx; io_cond = cond ? gs_cond : nullptr; io_cc = Cc; io_t = gs_t; io_y = y ? gs_y : nullptr; io_out = gs_out;
    cudaGraph_t graph = nullptr;
    bool ok = cudaStreamBeginCapture(graph_stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
      const int rc = run_ops(B, graph_stream);
      const cudaError_t e = cudaStreamEndCapture(graph_stream, &graph);
      ok = rc == EO_OK && e == cudaSuccess && graph != nullptr;
    }
    if (ok) ok = cudaGraphInstantiate(&g.exec, graph, 0) == cudaSuccess;
    if (graph) cudaGraphDestroy(graph);
    if (!ok) {            // same kernels, launched one by one
      cudaGetLastError();
      g.failed = true; g.exec = nullptr;
      return EO_OK;
    }
  }
  EO_CHECK_CUDA(cudaMemcpyAsync(gs_x, x, (size_t)B * Cx * hw * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (cond) EO_CHECK_CUDA(cudaMemcpyAsync(gs_cond, cond, (size_t)B * Cc * hw * sizeof(float), cudaMemcpyDeviceToDevice, st));
  EO_CHECK_CUDA(cudaMemcpyAsync(gs_t, t, (size_t)B * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  if (y) EO_CHECK_CUDA(cudaMemcpyAsync(gs_y, y, (size_t)B * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  EO_CHECK_CUDA(cudaGraphLaunch(g.exec, st));
  EO_CHECK_CUDA(cudaMemcpyAsync(out, gs_out, (size_t)B * cfg.out_channels * hw * sizeof(float), cudaMemcpyDeviceToDevice, st));
  *handled = true;
  return EO_OK;
}

// ---------------------------------------------------------------------------------------
// C ABI (include/eo_b200.h)
// ---------------------------------------------------------------------------------------
namespace eo {
static thread_local std::string g_err;
static thread_local bool g_pdl = false;
bool pdl_allowed() { return g_pdl; }
#ifdef EO_NO_PDL      // A/B builds (tools/ab.sh): every launch the ordinary way
void pdl_set_allowed(bool) { g_pdl = false; }
#else
void pdl_set_allowed(bool on) { g_pdl = on; }
#endif
void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}
const char* get_error() { return g_err.c_str(); }

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}
}  // namespace eo

extern "C" {

const char* eo_last_error(void) { return eo::get_error(); }
int eo_version(void) { return 100; }

int eo_device_check(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    set_error("no CUDA device available (libeo_b200 has no CPU fallback)");
    return EO_ERR_DEVICE;
  }
  if (major != 10) {
    set_error("compute capability %d.x device: libeo_b200 is built for sm_100a only", major);
    return EO_ERR_DEVICE;
  }
  return EO_OK;
}

int eo_unet_create(const eo_unet_cfg* cfg, eo_unet** out) {
  EO_REQUIRE(cfg && out, EO_ERR_ARG, "eo_unet_create: null argument");
  EO_REQUIRE(cfg->dims == 2, EO_ERR_ARG, "eo_unet_create: only dims=2 is implemented (got %d)", cfg->dims);
  EO_REQUIRE(cfg->conv_resample == 1, EO_ERR_ARG, "eo_unet_create: conv_resample=False is not implemented");
  EO_REQUIRE(cfg->n_channel_mult >= 1 && cfg->n_channel_mult <= 8, EO_ERR_ARG, "eo_unet_create: channel_mult length");
  EO_REQUIRE(cfg->n_attention_resolutions >= 0 && cfg->n_attention_resolutions <= 8, EO_ERR_ARG,
             "eo_unet_create: attention_resolutions length");
  EO_REQUIRE(cfg->in_channels > 0 && cfg->out_channels > 0 && cfg->num_res_blocks > 0, EO_ERR_ARG,
             "eo_unet_create: channel / block counts must be positive");
  EO_REQUIRE(cfg->model_channels > 0 && cfg->model_channels % 32 == 0, EO_ERR_ARG,
             "eo_unet_create: model_channels %d must be a positive multiple of 32 (GroupNorm32)", cfg->model_channels);
  EO_REQUIRE(cfg->model_channels % 2 == 0 && cfg->time_emb_factor > 0, EO_ERR_ARG, "eo_unet_create: time embedding");
  eo_unet* u = new eo_unet();
  u->cfg = *cfg;
  int rc = u->build_topology();
  if (rc) { delete u; return rc; }
  *out = u;
  return EO_OK;
}

void eo_unet_destroy(eo_unet* u) { delete u; }

int eo_unet_num_weights(const eo_unet* u) { return u ? (int)u->wspecs.size() : EO_ERR_ARG; }

const char* eo_unet_weight_name(const eo_unet* u, int index) {
  if (!u || index < 0 || index >= (int)u->wspecs.size()) return nullptr;
  return u->wspecs[index].name.c_str();
}

int eo_unet_weight_shape(const eo_unet* u, int index, int64_t shape[4]) {
  if (!u || index < 0 || index >= (int)u->wspecs.size()) { set_error("eo_unet_weight_shape: bad index"); return EO_ERR_ARG; }
  const auto& s = u->wspecs[index].shape;
  for (size_t i = 0; i < s.size() && i < 4; ++i) shape[i] = s[i];
  return (int)s.size();
}

int eo_unet_set_weight(eo_unet* u, const char* key, const float* dev_ptr, const int64_t* shape, int ndim) {
  EO_REQUIRE(u && key && dev_ptr && shape, EO_ERR_ARG, "eo_unet_set_weight: null argument");
  auto it = u->windex.find(key);
  EO_REQUIRE(it != u->windex.end(), EO_ERR_KEY, "eo_unet_set_weight: unknown key '%s'", key);
  WSpec& ws = u->wspecs[it->second];
  bool same = (int)ws.shape.size() == ndim;
  for (int i = 0; same && i < ndim; ++i) same = ws.shape[i] == shape[i];
  EO_REQUIRE(same, EO_ERR_KEY, "eo_unet_set_weight: shape mismatch for '%s'", key);
  ws.src = dev_ptr;
  u->finalized = false;   // packed copies are stale until the next finalize
  return EO_OK;
}

int eo_unet_finalize(eo_unet* u, int mode, int max_batch, int H, int W, void* stream) {
  EO_REQUIRE(u, EO_ERR_ARG, "eo_unet_finalize: null handle");
  int rc = eo_device_check();
  if (rc) return rc;
  rc = u->finalize(mode, max_batch, H, W, (cudaStream_t)stream);
  if (rc) u->release_plan();
  return rc;
}

int eo_unet_forward(eo_unet* u, const float* x, int Cx, const float* cond, int Cc, const int64_t* timesteps,
                    const int64_t* y, float* eps_out, int B, void* stream) {
  EO_REQUIRE(u, EO_ERR_ARG, "eo_unet_forward: null handle");
  return u->forward(x, Cx, cond, Cc, timesteps, y, eps_out, B, (cudaStream_t)stream);
}

int eo_unet_forward_timed(eo_unet* u, const float* x, int Cx, const float* cond, int Cc, const int64_t* timesteps,
                          const int64_t* y, float* eps_out, int B, void* stream, float* ms_per_op) {
  EO_REQUIRE(u && ms_per_op, EO_ERR_ARG, "eo_unet_forward_timed: null argument");
  return u->forward(x, Cx, cond, Cc, timesteps, y, eps_out, B, (cudaStream_t)stream, ms_per_op);
}

int eo_unet_num_ops(const eo_unet* u) { return u && u->finalized ? (int)u->ops.size() : 0; }

int eo_unet_op_info(const eo_unet* u, int index, const char** name, const char** kernel, double* flops_per_image,
                    double* bytes_per_image) {
  EO_REQUIRE(u && u->finalized && index >= 0 && index < (int)u->ops.size(), EO_ERR_ARG, "eo_unet_op_info: bad index");
  const Op& op = u->ops[index];
  if (name) *name = op.name.c_str();
  if (kernel) *kernel = op.kernel;
  if (flops_per_image) *flops_per_image = op.flops;
  if (bytes_per_image) *bytes_per_image = op.bytes;
  return EO_OK;
}

double eo_unet_op_executed_flops(const eo_unet* u, int index) {
  if (!u || !u->finalized || index < 0 || index >= (int)u->ops.size()) return -1.0;
  return u->ops[index].exec_flops;
}

int64_t eo_unet_device_bytes(const eo_unet* u) { return u ? u->dev_bytes : 0; }
// + the two GN memsets; with timestep tables installed the embedding path is one gather instead of four launches
int eo_unet_launches_per_forward(const eo_unet* u) { return u ? u->n_launches + 2 - (u->tt_on ? 3 : 0) : 0; }

int eo_unet_build_time_tables(eo_unet* u, int n_timesteps, void* stream) {
  EO_REQUIRE(u, EO_ERR_ARG, "eo_unet_build_time_tables: null handle");
  return u->build_time_tables(n_timesteps, (cudaStream_t)stream);
}

int eo_unet_clear_time_tables(eo_unet* u) {
  EO_REQUIRE(u, EO_ERR_ARG, "eo_unet_clear_time_tables: null handle");
  // Only the switch: the rows depend on the timestep value alone, so the table stays valid (and its graphs with it)
  // for the next sampling loop.  Freeing it here cost a device-wide cudaFree per sampling() call -- measured at up to
  // 800 ms on a B200 with graphs alive (tools/prof_sampling.py) -- and a graph re-capture per call.
  u->tt_on = false;
  return EO_OK;
}

int64_t eo_unet_read_activation(eo_unet* u, const char* name, float* out_dev, int64_t capacity, int B, void* stream) {
  EO_REQUIRE(u && name && out_dev, EO_ERR_ARG, "eo_unet_read_activation: null argument");
  EO_REQUIRE(u->finalized, EO_ERR_STATE, "eo_unet_read_activation before finalize");
  EO_REQUIRE(u->arena.keep, EO_ERR_STATE,
             "eo_unet_read_activation needs a -DEO_DEVTOOLS build and EO_DEBUG_KEEP=1 at finalize (workspace reuse "
             "overwrites activations otherwise)");
  auto it = u->named.find(name);
  EO_REQUIRE(it != u->named.end(), EO_ERR_KEY, "eo_unet_read_activation: unknown activation '%s'", name);
  const Act& a = it->second;
  int64_t n = (int64_t)B * a.C * a.H * a.W;
  EO_REQUIRE(n <= capacity, EO_ERR_ARG, "eo_unet_read_activation: capacity %lld < %lld", (long long)capacity, (long long)n);
  int rc = launch_nhwc_to_nchw_f32(u->ptr(a.off), u->act_dt, out_dev, B, a.H * a.W, a.C, (cudaStream_t)stream);
  return rc ? rc : n;
}

// ------------------------------------------------------------------ kernel self-tests
int eo_test_conv_tc(const void* x_bf16, const float* w, const float* bias, const void* residual, void* y_bf16,
                    int B, int H, int W, int Cin, int Cout, int k, void* stream) {
  int rc = eo_device_check();
  if (rc) return rc;
  EO_REQUIRE(k == 1 || k == 3, EO_ERR_ARG, "eo_test_conv_tc: k must be 1 or 3");
  cudaStream_t st = (cudaStream_t)stream;
  const int K = k * k * Cin;
  __nv_bfloat16* Wp = nullptr;
  EO_CHECK_CUDA(cudaMalloc(&Wp, (size_t)K * Cout * sizeof(__nv_bfloat16)));
  const bool patch = k == 3 && Cin % 64 == 0 && tc_conv_patch_supported(H, W);
  if (patch) {   // K order (64-channel block, tap, channel)
    for (int c0 = 0; c0 < Cin && !rc; c0 += 64)
      rc = launch_pack_conv_weight(w, Cin, k, c0, 64, Wp, DT_BF16, K, 1, (c0 / 64) * 9 * 64, Cout, nullptr, st);
  } else {
    rc = launch_pack_conv_weight(w, Cin, k, 0, Cin, Wp, DT_BF16, K, 1, 0, Cout, nullptr, st);
  }
  TcConvPlan* plan = nullptr;
  float* gn_buf = nullptr;
  double* stats_buf = nullptr;
  if (!rc) {
    TcConvParams p;
    p.nseg = 1;
    p.seg[0].ptr = x_bf16; p.seg[0].C = Cin; p.seg[0].Bt = B; p.seg[0].ntaps = k * k;
    for (int t = 0; t < k * k; ++t) {
      p.seg[0].dh[t] = (int8_t)(k == 3 ? t / 3 - 1 : 0);
      p.seg[0].dw[t] = (int8_t)(k == 3 ? t % 3 - 1 : 0);
      p.seg[0].dn[t] = 0;
    }
    p.seg[0].patch = patch ? TC_PATCH_3X3 : 0;
    // development aid (tools/conv_check.py): EO_TEST_GN=1 folds an identity GroupNorm affine (scale 1,
    // shift 0) into the operand load, =2 adds the SiLU (the caller then compares against conv(silu(x)))
#ifdef EO_DEVTOOLS
    const char* tg = std::getenv("EO_TEST_GN");
    if (tg && (tg[0] == '1' || tg[0] == '2') && patch) {
      std::vector<float> ones((size_t)B * Cin, 1.0f);
      EO_CHECK_CUDA(cudaMalloc(&gn_buf, 2 * ones.size() * sizeof(float)));
      EO_CHECK_CUDA(cudaMemsetAsync(gn_buf, 0, 2 * ones.size() * sizeof(float), st));
      EO_CHECK_CUDA(cudaMemcpyAsync(gn_buf, ones.data(), ones.size() * sizeof(float), cudaMemcpyHostToDevice, st));
      EO_CHECK_CUDA(cudaStreamSynchronize(st));
      p.seg[0].gn_scale = gn_buf; p.seg[0].gn_shift = gn_buf + ones.size(); p.seg[0].gn_ld = Cin;
      p.seg[0].silu = tg[0] == '2';
    }
#endif
    p.B = B; p.H = H; p.W = W; p.Wp = Wp; p.Ktot = K; p.Cout = Cout; p.bias = bias;
    p.residual = residual; p.out = y_bf16;
    // development aid (tools/conv3_trace.py): EO_TEST_STATS=1 also emits the fused GroupNorm statistics
#ifdef EO_DEVTOOLS
    const char* tsx = std::getenv("EO_TEST_STATS");
    if (tsx && tsx[0] == '1' && tc_conv_stats_supported(H, W)) {
      EO_CHECK_CUDA(cudaMalloc(&stats_buf, (size_t)B * Cout * 2 * sizeof(double)));
      EO_CHECK_CUDA(cudaMemsetAsync(stats_buf, 0, (size_t)B * Cout * 2 * sizeof(double), st));
      p.stats = stats_buf;
    }
#endif
    rc = tc_conv_plan_create(p, &plan);
  }
  if (!rc) rc = tc_conv_launch(plan, B, st);
  cudaError_t e = cudaStreamSynchronize(st);
  tc_conv_plan_destroy(plan);
  cudaFree(Wp);
  if (gn_buf) cudaFree(gn_buf);
  if (stats_buf) cudaFree(stats_buf);
  if (!rc && e != cudaSuccess) { set_error("eo_test_conv_tc: %s", cudaGetErrorString(e)); rc = EO_ERR_CUDA; }
  return rc;
}

int eo_debug_conv_trace(void* dev_buf, int n_ctas) {
  tc_conv_set_trace(reinterpret_cast<long long*>(dev_buf), n_ctas);
  tc_attn_set_trace(reinterpret_cast<long long*>(dev_buf), n_ctas);
  return EO_OK;
}

int eo_test_attention_tc(const void* qkv_bf16, void* out_bf16, int B, int T, int heads, int ch, void* stream) {
  int rc = eo_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (ch > 64 || ch % 8 != 0) {      // the engine's route for wide / odd heads: one warp per query, fp32 arithmetic
    rc = launch_attention_wide(qkv_bf16, out_bf16, DT_BF16, B, T, heads, ch, heads * 3 * ch, 3 * ch, ch, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (!rc && e != cudaSuccess) { set_error("eo_test_attention_tc: %s", cudaGetErrorString(e)); rc = EO_ERR_CUDA; }
    return rc;
  }
  // re-pack [B,T,heads*3*ch] (legacy order) into the kernel's [B,T,heads*3*64] padded layout
  const int ld_in = heads * 3 * ch, ld = heads * 3 * 64;
  __nv_bfloat16* padded = nullptr;
  EO_CHECK_CUDA(cudaMalloc(&padded, (size_t)B * T * ld * sizeof(__nv_bfloat16)));
  EO_CHECK_CUDA(cudaMemsetAsync(padded, 0, (size_t)B * T * ld * sizeof(__nv_bfloat16), st));
  for (int hp = 0; hp < heads * 3; ++hp)
    EO_CHECK_CUDA(cudaMemcpy2DAsync(padded + hp * 64, (size_t)ld * 2,
                                    reinterpret_cast<const __nv_bfloat16*>(qkv_bf16) + hp * ch, (size_t)ld_in * 2,
                                    (size_t)ch * 2, (size_t)B * T, cudaMemcpyDeviceToDevice, st));
  TcAttnParams p; p.qkv = padded; p.out = out_bf16; p.B = B; p.T = T; p.heads = heads; p.ch = ch;
  __nv_bfloat16* ones = nullptr;
  if (ch <= 48) {   // channel round16(ch) of every head's k = 1.0 (what the qkv convolution's bias does inside the UNet)
    std::vector<__nv_bfloat16> h1((size_t)B * T, __float2bfloat16(1.0f));
    EO_CHECK_CUDA(cudaMalloc(&ones, h1.size() * sizeof(__nv_bfloat16)));
    EO_CHECK_CUDA(cudaMemcpyAsync(ones, h1.data(), h1.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice, st));
    for (int hh = 0; hh < heads; ++hh)
      EO_CHECK_CUDA(cudaMemcpy2DAsync(padded + (hh * 3 + 1) * 64 + ((ch + 15) & ~15), (size_t)ld * 2, ones, 2, 2, (size_t)B * T,
                                      cudaMemcpyDeviceToDevice, st));
    // ... and q multiplied by ch^-1/2 * log2(e), re-rounded to bf16 (inside the UNet the factor sits in the projection's weights)
    const float sl2 = (1.0f / sqrtf((float)ch)) * 1.4426950408889634f;
    for (int hh = 0; hh < heads && !rc; ++hh) rc = launch_scale_cols_bf16(padded, (long long)B * T, ld, hh * 3 * 64, ch, sl2, st);
    EO_CHECK_CUDA(cudaStreamSynchronize(st));
    p.k_one = 1;
  }
  TcAttnPlan* plan = nullptr;
  rc = tc_attn_plan_create(p, &plan);
  if (!rc) rc = tc_attn_launch(plan, B, st);
#ifdef EO_DEVTOOLS
  if (!rc) {     // development builds: the kernel alone, best of three (stderr)
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int i = 0; i < 3 && !rc; ++i) {
      cudaEventRecord(e0, st);
      rc = tc_attn_launch(plan, B, st);
      cudaEventRecord(e1, st);
      cudaEventSynchronize(e1);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      best = ms < best ? ms : best;
    }
    fprintf(stderr, "[eo devtools] attention kernel B=%d T=%d heads=%d ch=%d: %.4f ms, %.1f TFLOP/s algorithmic\n", B, T, heads, ch,
            best, 4.0 * B * heads * (double)T * T * ch / best / 1e9);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
#endif
  cudaError_t e = cudaStreamSynchronize(st);
  tc_attn_plan_destroy(plan);
  cudaFree(padded);
  if (ones) cudaFree(ones);
  if (!rc && e != cudaSuccess) { set_error("eo_test_attention_tc: %s", cudaGetErrorString(e)); rc = EO_ERR_CUDA; }
  return rc;
}

// EODiffusion.sampling's loop (diffusion/model.py:46-75) over the per-step entry points
int eo_sample_ddpm(eo_unet* u, float* x, const float* noise_tape, const float* gt, const float* mask, const float* cond,
                   int Cc, const int64_t* y, const int64_t* timestep_rows, const float* table, float* eps_scratch,
                   int T, int B, int Cx, int H, int W, int clip, void* stream) {
  EO_REQUIRE(u && x && noise_tape && timestep_rows && table && eps_scratch, EO_ERR_ARG, "eo_sample_ddpm: null argument");
  EO_REQUIRE(T >= 1 && B >= 1 && Cx >= 1 && H >= 1 && W >= 1, EO_ERR_ARG, "eo_sample_ddpm: bad extents");
  EO_REQUIRE(u->finalized, EO_ERR_STATE, "eo_sample_ddpm before eo_unet_finalize");
  EO_REQUIRE(H == u->H && W == u->W && B <= u->Bmax && Cx == u->cfg.out_channels, EO_ERR_ARG,
             "eo_sample_ddpm: [%d, %d, %d, %d] does not match the finalized geometry (batch <= %d, %d channels, %d x %d)", B, Cx,
             H, W, u->Bmax, u->cfg.out_channels, u->H, u->W);
  EO_REQUIRE((gt == nullptr) == (mask == nullptr), EO_ERR_ARG, "eo_sample_ddpm: 'sum' conditioning needs gt and mask");
  const int HW = H * W;
  const size_t n = (size_t)B * Cx * HW;
  int rc;
  // the loop visits the timestep values T-1 .. 0: their embedding rows are computed once, not once per step
  // (the table stays cached in the handle; forwards outside this loop use it only if the caller installed it)
  const bool tt_prev = u->tt_on;
  if (!y && (rc = u->build_time_tables(T, (cudaStream_t)stream))) return rc;
  struct Restore { eo_unet* u; bool prev; ~Restore() { u->tt_on = prev; } } restore{u, tt_prev};
  if (gt && (rc = eo_ddpm_sum_mix(x, gt, mask, noise_tape, timestep_rows + (size_t)(T - 1) * B, table, x, B, Cx, HW, stream)))
    return rc;
  for (int i = T - 1; i >= 0; --i) {
    const int k = T - 1 - i;
    const int64_t* t = timestep_rows + (size_t)i * B;
    if ((rc = eo_unet_forward(u, x, Cx, cond, Cc, t, y, eps_scratch, B, stream))) return rc;
    const float* noise = noise_tape + (size_t)k * n;
    if (gt && i > 0)
      rc = eo_ddpm_step_mix(x, eps_scratch, noise, t, gt, mask, noise + n, timestep_rows + (size_t)(i - 1) * B, table, x, B, Cx,
                            HW, clip, 1, stream);
    else
      rc = eo_ddpm_step(x, eps_scratch, noise, t, table, x, B, Cx, HW, clip, i > 0 ? 1 : 0, stream);
    if (rc) return rc;
  }
  return EO_OK;
}

// DDIMSampler.ddim_sampling's loop (diffusion/ddim.py:114-164) over the per-step entry points
int eo_sample_ddim(eo_unet* u, float* x, const float* noise_tape, const float* cond, int Cc, const int64_t* y,
                   const int64_t* timestep_rows, const float* scalars, float* eps_scratch, float* pred_x0, int S, int B, int Cx,
                   int H, int W, void* stream) {
  EO_REQUIRE(u && x && timestep_rows && scalars && eps_scratch && pred_x0, EO_ERR_ARG, "eo_sample_ddim: null argument");
  EO_REQUIRE(S >= 1 && B >= 1 && Cx >= 1 && H >= 1 && W >= 1, EO_ERR_ARG, "eo_sample_ddim: bad extents");
  EO_REQUIRE(u->finalized, EO_ERR_STATE, "eo_sample_ddim before eo_unet_finalize");
  EO_REQUIRE(H == u->H && W == u->W && B <= u->Bmax && Cx == u->cfg.out_channels, EO_ERR_ARG,
             "eo_sample_ddim: [%d, %d, %d, %d] does not match the finalized geometry (batch <= %d, %d channels, %d x %d)", B, Cx,
             H, W, u->Bmax, u->cfg.out_channels, u->H, u->W);
  const int64_t n = (int64_t)B * Cx * H * W;
  for (int k = 0; k < S; ++k) {
    const int index = S - 1 - k;
    const float* sc = scalars + (size_t)index * 6;
    EO_REQUIRE(noise_tape || sc[4] == 0.0f, EO_ERR_ARG, "eo_sample_ddim: sigma[%d] != 0 needs a noise tape", index);
    int rc = eo_unet_forward(u, x, Cx, cond, Cc, timestep_rows + (size_t)index * B, y, eps_scratch, B, stream);
    if (rc) return rc;
    rc = eo_ddim_step(x, eps_scratch, noise_tape ? noise_tape + (size_t)k * n : nullptr, x, pred_x0, sc[0], sc[1], sc[2], sc[3],
                      sc[4], sc[5], n, stream);
    if (rc) return rc;
  }
  return EO_OK;
}

}  // extern "C"
