// Host-side launch chains of the MV-Former head and the projection MLP + the C ABI (include/mvf_b200.h).
// No device memory is allocated here: every buffer is a named region of the caller's `save`, `ws` or `gpack`
// allocations, laid out deterministically from the descriptor.
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "kernels.cuh"

namespace mvf {

// ------------------------------------------------------------------------------------------------------------
// error string
// ------------------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MVF_PDL");
    on = (e && atoi(e) == 0) ? 0 : 1;
  }
  return on != 0;
}

// ------------------------------------------------------------------------------------------------------------
// launch counter + optional CUDA-event timing of the dominant kernels (bench.py's roofline numbers)
// ------------------------------------------------------------------------------------------------------------
static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }

constexpr int PROF_TAGS = 10, PROF_CAP = 512;
static int g_prof_on = 0;
static cudaEvent_t g_prof_ev[PROF_TAGS][PROF_CAP][2];
static int g_prof_n[PROF_TAGS] = {0};
static bool g_prof_init[PROF_TAGS][PROF_CAP] = {{false}};
struct ProfScope {
  int tag, slot;
  cudaStream_t st;
  ProfScope(int tag_, cudaStream_t st_) : tag(tag_), slot(-1), st(st_) {
    // level 1: the streaming pooling kernels only (two brackets per step); level 2: every tagged kernel group.  Inside a
    // captured graph every bracket is a pair of event-record nodes, which also cuts the programmatic-dependent-launch overlap
    // of the kernels around it -- so the timed graph of a ~1.6 ms step carries level 1 and the shares are read at level 2
    if (!g_prof_on || tag < 0 || tag >= PROF_TAGS || g_prof_n[tag] >= PROF_CAP || (g_prof_on == 1 && tag > 1)) return;
    slot = g_prof_n[tag];
    if (!g_prof_init[tag][slot]) {
      if (cudaEventCreate(&g_prof_ev[tag][slot][0]) != cudaSuccess || cudaEventCreate(&g_prof_ev[tag][slot][1]) != cudaSuccess) {
        slot = -1;
        return;
      }
      g_prof_init[tag][slot] = true;
    }
    record(g_prof_ev[tag][slot][0]);
  }
  ~ProfScope() {
    if (slot < 0) return;
    record(g_prof_ev[tag][slot][1]);
    g_prof_n[tag] = slot + 1;
  }
  // inside a stream capture a plain record is only a dependency edge; an EXTERNAL record becomes an event-record node
  // that fires on every replay, so the pair can be read back with cudaEventElapsedTime after a graph launch
  void record(cudaEvent_t e) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusActive)
      cudaEventRecordWithFlags(e, st, cudaEventRecordExternal);
    else
      cudaEventRecord(e, st);
  }
};

// ------------------------------------------------------------------------------------------------------------
// side stream: weight-gradient GEMMs and bias column-sums only feed the gradient buffer, so in backward they are
// forked onto a second stream and overlap the dX chain on the caller's stream (joined before every return)
// ------------------------------------------------------------------------------------------------------------
struct SideCtl {
  cudaStream_t owner = nullptr;   // caller stream this side stream is paired with
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[64];
  int n_ev = 0, next = 0;
};
constexpr int MAX_DEVICES = 64, SIDE_PER_DEV = 4;
// one side stream + event ring per (device, caller stream): concurrent micro-batches on different caller streams fork onto
// different side streams (streams and events are device-bound).  More than SIDE_PER_DEV caller streams share the last slot.
static SideCtl g_side[MAX_DEVICES][SIDE_PER_DEV];
static std::mutex g_side_mu;
static int side_acquire(cudaStream_t caller, cudaStream_t* s, int* dev_out) {
  static int enabled = -1;  // MVF_SIDE_STREAM=0 keeps everything on the caller's stream (debugging / A-B timing)
  if (enabled < 0) {
    const char* e = getenv("MVF_SIDE_STREAM");
    enabled = (e && atoi(e) == 0) ? 0 : 1;
  }
  *s = nullptr;
  *dev_out = 0;
  if (!enabled) return MVF_OK;
  int dev = 0;
  MVF_CHECK_CUDA(cudaGetDevice(&dev));
  MVF_REQUIRE(dev >= 0 && dev < MAX_DEVICES, MVF_ERR_UNSUPPORTED, "device ordinal %d out of range", dev);
  std::lock_guard<std::mutex> lk(g_side_mu);
  int slot = SIDE_PER_DEV - 1;
  for (int i = 0; i < SIDE_PER_DEV; ++i) {
    if (g_side[dev][i].stream == nullptr || g_side[dev][i].owner == caller) { slot = i; break; }
  }
  SideCtl& c = g_side[dev][slot];
  if (c.stream == nullptr) {
    MVF_CHECK_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    for (int i = 0; i < 64; ++i) MVF_CHECK_CUDA(cudaEventCreateWithFlags(&c.ev[i], cudaEventDisableTiming));
    c.n_ev = 64;
    c.owner = caller;
  }
  *s = c.stream;
  *dev_out = dev * SIDE_PER_DEV + slot;
  return MVF_OK;
}
static cudaEvent_t side_event(int handle) {
  std::lock_guard<std::mutex> lk(g_side_mu);
  SideCtl& c = g_side[handle / SIDE_PER_DEV][handle % SIDE_PER_DEV];
  cudaEvent_t e = c.ev[c.next];
  c.next = (c.next + 1) % c.n_ev;
  return e;
}

// ------------------------------------------------------------------------------------------------------------
// parameter table (canonical order == reference state_dict order of embed.* then ssl_projection.*)
// ------------------------------------------------------------------------------------------------------------
struct ParamInfo {
  std::string name;
  int64_t rows, cols;  // vectors: rows = 1
};

struct Model {
  mvf_head_desc d;
  int E_oh;    // one-hot columns appended to the pooled entities
  int W0;      // logical width of the first FC input (SPC + E_oh)
  int ld0;     // padded
  int Hin;     // video_emb output width (== H unless one_hot == enc)
  int64_t F, R, S, rows, N;  // frames, entity rows, tokens per view, encoder rows, frame rows
  int act;     // dtype of every activation downstream of the pooling: always fp32 (exact FMA on the SIMT backend,
               // tf32 tensor-core GEMMs on the tcgen05 backend -- bf16 there costs ~10 % on the gradient through the 1/tau
               // amplification of SCL, tf32 ~2 %)
  int kvt;     // dtype of tokens, W_k|W_v, K|V and dK|dV (the 96 % of the FLOPs): d.dtype
  bool tc;     // chain GEMMs on tcgen05 (bf16 tokens or an explicit TCGEN05 backend): forward weights are also packed pre-split
  bool fold;   // rank-E folded pooling (pool_fold.cu): no K|V tensors, one streaming pass over the tokens per direction
  bool fwb;    // FIXED_WIDTH_BASELINE: FWBPooling (one Linear of the CLS embedding) instead of the entity cross-attention
  int iWlc, iblc;
  std::vector<ParamInfo> params;
  // indices into params
  int iQs, iQb, iWk, ibk, iWv, ibv;
  int iFcW[MVF_MAX_FC], iFcB[MVF_MAX_FC], iFcG[MVF_MAX_FC], iFcBeta[MVF_MAX_FC];
  int iWe, ibe;
  std::vector<int> iLayer;  // base index per encoder layer (16 params each)
  int iWemb, ibemb, iWlin, iblin;
  int iWp1, ibp1, iGp, iBp, iWp2, ibp2;
};
// per-layer param offsets
enum { L_LN0W = 0, L_LN0B, L_LN1W, L_LN1B, L_WQ, L_BQ, L_WK, L_BK, L_WV, L_BV, L_WO, L_BO, L_W1, L_B1, L_W2, L_B2, L_COUNT };

static int add_param(Model& m, const std::string& name, int64_t rows, int64_t cols) {
  m.params.push_back({name, rows, cols});
  return (int)m.params.size() - 1;
}

static int build_model(const mvf_head_desc* dp, Model& m) {
  MVF_REQUIRE(dp != nullptr, MVF_ERR_BAD_ARG, "null descriptor");
  const mvf_head_desc& d = *dp;
  m.d = d;
  MVF_REQUIRE(d.BV > 0 && d.T > 0 && d.P > 0 && d.C_in > 0, MVF_ERR_BAD_ARG, "bad input shape BV=%d T=%d P=%d C_in=%d",
              d.BV, d.T, d.P, d.C_in);
  MVF_REQUIRE(d.E >= 1 && d.E <= MVF_MAX_ENTITIES, MVF_ERR_UNSUPPORTED, "SMART_TOKENS=%d outside [1,%d]", d.E,
              MVF_MAX_ENTITIES);
  MVF_REQUIRE(d.SPC > 0 && d.H > 0 && d.DFF > 0 && d.D > 0 && d.PS > 0, MVF_ERR_BAD_ARG, "bad widths");
  MVF_REQUIRE(d.n_fc >= 0 && d.n_fc <= MVF_MAX_FC, MVF_ERR_UNSUPPORTED, "FC_LAYERS: %d layers (max %d)", d.n_fc,
              MVF_MAX_FC);
  MVF_REQUIRE(d.L >= 0 && d.L <= 64, MVF_ERR_UNSUPPORTED, "NUM_LAYERS=%d", d.L);
  MVF_REQUIRE(d.heads > 0 && d.H % d.heads == 0, MVF_ERR_BAD_ARG, "HIDDEN_SIZE %d not divisible by NUM_HEADS %d", d.H,
              d.heads);
  MVF_REQUIRE(d.dtype == MVF_F32 || d.dtype == MVF_BF16, MVF_ERR_BAD_ARG, "dtype %d", d.dtype);
  MVF_REQUIRE(d.one_hot == MVF_ONEHOT_NONE || d.one_hot == MVF_ONEHOT_POOL, MVF_ERR_UNSUPPORTED,
              "SMART_ONE_HOT=enc is not implemented in this build");
  MVF_REQUIRE(d.final_mode >= MVF_FINAL_MAX && d.final_mode <= MVF_FINAL_LIN, MVF_ERR_BAD_ARG, "final_mode %d",
              d.final_mode);
  MVF_REQUIRE(d.drop_p >= 0.f && d.drop_p < 1.f, MVF_ERR_BAD_ARG, "drop_p %f", d.drop_p);
  if (d.dtype == MVF_BF16 || d.gemm_backend == MVF_GEMM_TCGEN05) {
    bool ok = d.C_in % 8 == 0 && d.SPC % 8 == 0 && d.H % 4 == 0 && d.DFF % 4 == 0 && d.D % 4 == 0 && d.PS % 4 == 0;
    for (int i = 0; i < d.n_fc; ++i) ok = ok && d.fc[i] % 4 == 0;
    MVF_REQUIRE(ok, MVF_ERR_ALIGN,
                "the tensor-core path needs C_in and SMART_POOL_CHANNELS to be multiples of 8 and every other channel "
                "width a multiple of 4 (TMA 16-byte rows)");
  }
  m.E_oh = d.one_hot == MVF_ONEHOT_POOL ? d.E : 0;
  m.W0 = d.SPC + m.E_oh;
  m.ld0 = (int)round_up(m.W0, 8);
  m.Hin = d.H;
  m.F = (int64_t)d.BV * d.T;
  m.R = m.F * d.E;
  m.S = (int64_t)d.E * d.T;
  m.rows = (int64_t)d.BV * m.S;
  m.N = m.F;
  m.act = MVF_F32;
  m.kvt = d.dtype;
  m.tc = d.gemm_backend == MVF_GEMM_TCGEN05 || (d.gemm_backend == MVF_GEMM_AUTO && d.dtype == MVF_BF16);
  {
    int pm = d.pool_mode;
    MVF_REQUIRE(pm >= MVF_POOL_AUTO && pm <= MVF_POOL_FOLDED, MVF_ERR_BAD_ARG, "pool_mode %d", pm);
    if (pm == MVF_POOL_AUTO) {
      const char* e = getenv("MVF_POOL_MODE");   // A/B measurements only
      if (e && strcmp(e, "dense") == 0) pm = MVF_POOL_DENSE;
      else if (e && strcmp(e, "folded") == 0) pm = MVF_POOL_FOLDED;
    }
    const bool ok = pool_fold_supported(d.dtype, d.C_in, d.E, d.P);
    MVF_REQUIRE(pm != MVF_POOL_FOLDED || ok, MVF_ERR_UNSUPPORTED,
                "folded pooling needs C_in (%d) to be a multiple of 8 and at most 5120", d.C_in);
    m.fold = pm == MVF_POOL_FOLDED || (pm == MVF_POOL_AUTO && ok);
  }
  m.params.clear();
  MVF_REQUIRE(d.pool_kind == MVF_POOLKIND_LSTP || d.pool_kind == MVF_POOLKIND_FWB, MVF_ERR_BAD_ARG, "pool_kind %d", d.pool_kind);
  m.fwb = d.pool_kind == MVF_POOLKIND_FWB;
  m.iQs = m.iQb = m.iWk = m.ibk = m.iWv = m.ibv = m.iWlc = m.iblc = -1;
  if (m.fwb) {
    MVF_REQUIRE(d.cls_dim > 0 && d.cls_dim % 4 == 0, MVF_ERR_BAD_ARG, "FIXED_WIDTH_BASELINE needs cls_dim (%d) > 0, a multiple of 4", d.cls_dim);
    m.fold = false;
    m.iWlc = add_param(m, "embed.pooling.lin_conv.weight", (int64_t)d.SPC * d.E, d.cls_dim);
    m.iblc = add_param(m, "embed.pooling.lin_conv.bias", 1, (int64_t)d.SPC * d.E);
  } else {
    const std::string ca = "embed.pooling.cross_att.";
    m.iQs = add_param(m, ca + "Q_s", d.E, d.SPC);
    m.iQb = add_param(m, ca + "Q_s_b", 1, d.SPC);
    m.iWk = add_param(m, ca + "linear_K2d.weight", d.SPC, d.C_in);
    m.ibk = add_param(m, ca + "linear_K2d.bias", 1, d.SPC);
    m.iWv = add_param(m, ca + "linear_V2d.weight", d.SPC, d.C_in);
    m.ibv = add_param(m, ca + "linear_V2d.bias", 1, d.SPC);
  }
  int cin = m.W0;
  for (int i = 0; i < d.n_fc; ++i) {
    MVF_REQUIRE(d.fc[i] > 0, MVF_ERR_BAD_ARG, "fc[%d] = %d", i, d.fc[i]);
    const std::string lin = "embed.fc_layers." + std::to_string(4 * i + 1), bn = "embed.fc_layers." + std::to_string(4 * i + 2);
    m.iFcW[i] = add_param(m, lin + ".weight", d.fc[i], cin);
    m.iFcB[i] = add_param(m, lin + ".bias", 1, d.fc[i]);
    m.iFcG[i] = add_param(m, bn + ".weight", 1, d.fc[i]);
    m.iFcBeta[i] = add_param(m, bn + ".bias", 1, d.fc[i]);
    cin = d.fc[i];
  }
  m.iWe = add_param(m, "embed.video_emb.weight", m.Hin, cin);
  m.ibe = add_param(m, "embed.video_emb.bias", 1, m.Hin);
  m.iLayer.clear();
  for (int l = 0; l < d.L; ++l) {
    const std::string q = "embed.video_encoder.enc_layers." + std::to_string(l) + ".";
    int base = add_param(m, q + "res_layer0.norm.weight", 1, d.H);
    add_param(m, q + "res_layer0.norm.bias", 1, d.H);
    add_param(m, q + "res_layer1.norm.weight", 1, d.H);
    add_param(m, q + "res_layer1.norm.bias", 1, d.H);
    const char* nm[4] = {"linear_Q2d", "linear_K2d", "linear_V2d", "linear_d2Q"};
    for (int k = 0; k < 4; ++k) {
      add_param(m, q + "self_att." + nm[k] + ".weight", d.H, d.H);
      add_param(m, q + "self_att." + nm[k] + ".bias", 1, d.H);
    }
    add_param(m, q + "feed_forward.fc1.weight", d.DFF, d.H);
    add_param(m, q + "feed_forward.fc1.bias", 1, d.DFF);
    add_param(m, q + "feed_forward.fc2.weight", d.H, d.DFF);
    add_param(m, q + "feed_forward.fc2.bias", 1, d.H);
    m.iLayer.push_back(base);
  }
  m.iWemb = add_param(m, "embed.embedding_layer.weight", d.D, d.H);
  m.ibemb = add_param(m, "embed.embedding_layer.bias", 1, d.D);
  m.iWlin = m.iblin = -1;
  if (d.final_mode == MVF_FINAL_LIN) {
    m.iWlin = add_param(m, "embed.lin_final.weight", d.H, (int64_t)d.E * d.H);
    m.iblin = add_param(m, "embed.lin_final.bias", 1, d.H);
  }
  m.iWp1 = add_param(m, "ssl_projection.net.0.weight", d.PS, d.D);
  m.ibp1 = add_param(m, "ssl_projection.net.0.bias", 1, d.PS);
  m.iGp = add_param(m, "ssl_projection.net.1.weight", 1, d.PS);
  m.iBp = add_param(m, "ssl_projection.net.1.bias", 1, d.PS);
  m.iWp2 = add_param(m, "ssl_projection.net.3.weight", d.D, d.PS);
  m.ibp2 = add_param(m, "ssl_projection.net.3.bias", 1, d.D);
  return MVF_OK;
}

// ------------------------------------------------------------------------------------------------------------
// region layouts
// ------------------------------------------------------------------------------------------------------------
enum { RT_F32 = 0, RT_BF16 = 1, RT_F64 = 2, RT_I32 = 3 };
static size_t rt_size(int t) { return t == RT_BF16 ? 2 : (t == RT_F64 ? 8 : 4); }

struct Region {
  std::string name;
  size_t off;
  int64_t rows, cols, ld;
  int dtype;
};
struct Layout {
  std::vector<Region> regs;
  size_t total = 0;
  void add(const std::string& name, int64_t rows, int64_t cols, int dtype, int64_t ld = -1) {
    if (ld < 0) ld = cols;
    Region r{name, total, rows, cols, ld, dtype};
    regs.push_back(r);
    size_t bytes = (size_t)rows * (size_t)ld * rt_size(dtype);
    total += (bytes + 255) / 256 * 256;
  }
  const Region* find(const std::string& name) const {
    for (const Region& r : regs)
      if (r.name == name) return &r;
    return nullptr;
  }
};

// pre-split (bf16 hi|lo blocks) copy of a forward weight [rows, cols]: container of round_up(cols, 32) floats per row
static void add_split(Layout& L, const std::string& name, int64_t rows, int64_t cols) {
  L.add(name + ".s", rows, round_up(cols, 32), RT_F32);
}

static std::string lname(int l, const char* s) { return "l" + std::to_string(l) + "." + s; }
static std::string fname(int i, const char* s) { return "fc" + std::to_string(i) + "." + s; }

// activations kept for backward by the head
static void head_save_layout(const Model& m, Layout& L) {
  const mvf_head_desc& d = m.d;
  const int A = m.act;
  if (m.fwb) {
    // lin_conv with its output rows regrouped entity-major: row e*SPC + c of "w.lc" = row c*E + e of the parameter
    L.add("w.lc", (int64_t)d.E * d.SPC, d.cls_dim, A);
    L.add("b.lc", 1, (int64_t)d.E * d.SPC, RT_F32);
    if (m.tc) add_split(L, "w.lc", (int64_t)d.E * d.SPC, d.cls_dim);
  } else if (m.fold) {
    L.add("wq", d.E, d.C_in, RT_F32);
    if (m.tc) add_split(L, "w.v", d.SPC, d.C_in);
  } else {
    L.add("w.kv", 2 * d.SPC, d.C_in, m.kvt, round_up(d.C_in, 8));
    L.add("b.kv", 1, 2 * d.SPC, RT_F32);
  }
  int cin_ld = m.ld0;
  for (int i = 0; i < d.n_fc; ++i) {
    L.add(fname(i, "w"), d.fc[i], i == 0 ? m.W0 : d.fc[i - 1], A, cin_ld);
    if (m.tc) add_split(L, fname(i, "w"), d.fc[i], cin_ld);
    cin_ld = d.fc[i];
  }
  L.add("w.e", m.Hin, d.n_fc ? d.fc[d.n_fc - 1] : m.W0, A, cin_ld);
  if (m.tc) add_split(L, "w.e", m.Hin, cin_ld);
  for (int l = 0; l < d.L; ++l) {
    L.add(lname(l, "w.qkv"), 3 * d.H, d.H, A);
    L.add(lname(l, "b.qkv"), 1, 3 * d.H, RT_F32);
    L.add(lname(l, "w.o"), d.H, d.H, A);
    L.add(lname(l, "w.1"), d.DFF, d.H, A);
    L.add(lname(l, "w.2"), d.H, d.DFF, A);
    if (m.tc) {
      add_split(L, lname(l, "w.qkv"), 3 * d.H, d.H);
      add_split(L, lname(l, "w.o"), d.H, d.H);
      add_split(L, lname(l, "w.1"), d.DFF, d.H);
      add_split(L, lname(l, "w.2"), d.H, d.DFF);
    }
  }
  L.add("w.emb", d.D, d.H, A);
  if (m.tc) add_split(L, "w.emb", d.D, d.H);
  if (d.final_mode == MVF_FINAL_LIN) {
    L.add("w.lin", d.H, (int64_t)d.E * d.H, A);
    if (m.tc) add_split(L, "w.lin", d.H, (int64_t)d.E * d.H);
  }
  if (!m.fwb) {
    if (m.fold) L.add("px", m.R, d.C_in, RT_F32);      // attention-pooled tokens: the only C_in-wide activation kept
    else L.add("kv", m.F * d.P, 2 * d.SPC, m.kvt);
    L.add("attn", m.F * d.E, d.P, RT_F32);
  }
  L.add("h0", m.R, m.W0, A, m.ld0);
  L.add("ent32", m.R, d.SPC, RT_F32);
  // the BatchNorm sum buffers of every layer are adjacent: one memset per forward call zeroes them all
  for (int i = 0; i < d.n_fc; ++i) {
    L.add(fname(i, "sum"), 1, 2 * d.fc[i], RT_F64);
    L.add(fname(i, "bsum"), 1, 2 * d.fc[i], RT_F64);
  }
  for (int i = 0; i < d.n_fc; ++i) {
    L.add(fname(i, "x"), m.R, d.fc[i], RT_F32);
    L.add(fname(i, "mi"), 1, 2 * d.fc[i], RT_F32);
    L.add(fname(i, "a"), m.R, d.fc[i], A);
  }
  L.add("h3", m.R, m.Hin, RT_F32);
  L.add("pe", d.T, d.H, RT_F32);
  for (int k = 0; k <= 2 * d.L; ++k) L.add("z" + std::to_string(k), m.rows, d.H, RT_F32);
  for (int l = 0; l < d.L; ++l) {
    L.add(lname(l, "ln0"), 2, m.rows, RT_F32);
    L.add(lname(l, "r0"), m.rows, d.H, A);
    L.add(lname(l, "qkv"), m.rows, 3 * d.H, A);
    L.add(lname(l, "lse"), (int64_t)d.BV * d.heads, m.S, RT_F32);
    L.add(lname(l, "ctx"), m.rows, d.H, A);
    L.add(lname(l, "ln1"), 2, m.rows, RT_F32);
    L.add(lname(l, "r1"), m.rows, d.H, A);
    L.add(lname(l, "f"), m.rows, d.DFF, A);
  }
  if (d.final_mode == MVF_FINAL_LIN) {
    L.add("zl", m.N, (int64_t)d.E * d.H, A);
    L.add("ylin", m.N, d.H, RT_F32);
  }
  if (d.final_mode == MVF_FINAL_MAX) L.add("argmax", m.N, d.H, RT_I32);
  L.add("y", m.N, d.H, A);
}

static void head_ws_layout(const Model& m, Layout& L) {
  const mvf_head_desc& d = m.d;
  const int A = m.act;
  int maxfc = m.W0;
  for (int i = 0; i < d.n_fc; ++i) maxfc = d.fc[i] > maxfc ? d.fc[i] : maxfc;
  // forward scratch
  L.add("o", m.rows, d.H, RT_F32);
  // backward scratch
  L.add("dEact", m.N, d.D, A);
  L.add("dy", m.N, d.H, RT_F32);
  if (d.final_mode == MVF_FINAL_LIN) {
    L.add("dylin", m.N, d.H, A);
    L.add("dzl", m.N, (int64_t)d.E * d.H, RT_F32);
  }
  L.add("dzA", m.rows, d.H, RT_F32);
  L.add("dzB", m.rows, d.H, RT_F32);
  // gradients that a forked weight-gradient GEMM reads get one region per use (no reuse before the join)
  for (int l = 0; l < d.L; ++l) {
    L.add(lname(l, "dgF"), m.rows, d.H, A);
    L.add(lname(l, "dgA"), m.rows, d.H, A);
    L.add(lname(l, "df"), m.rows, d.DFF, A);
    L.add(lname(l, "dqkv"), m.rows, 3 * d.H, A);
  }
  L.add("dr", m.rows, d.H, RT_F32);
  L.add("dctx", m.rows, d.H, A);
  L.add("delta", (int64_t)d.BV * d.heads, m.S, RT_F32);
  if (m.tc && d.heads > 0 && d.H / d.heads == 32) {
    // split-operand scratch of the tcgen05 attention kernels (attention_fa.cu); 1024-byte aligned inside the region
    const size_t fa = attention_ws_bytes(d.BV, (int)m.S, d.heads, 32);
    if (fa > 0) L.add("fa", 1, (int64_t)(fa + 1024) / 4, RT_F32);
  }
  L.add("dh3", m.R, m.Hin, A);
  L.add("da", m.R, maxfc, RT_F32);
  for (int i = 0; i < d.n_fc; ++i) L.add(fname(i, "dx"), m.R, d.fc[i], A);
  L.add("dh0", m.R, m.W0, A, m.ld0);
  if (m.fwb) {
    L.add("dent", m.R, d.SPC, RT_F32);
  } else if (m.fold) {
    L.add("dent", m.R, d.SPC, RT_F32);
    L.add("G", m.R, d.C_in, RT_F32);
    L.add("dwq", d.E, d.C_in, RT_F32);
    L.add("pdelta", 1, m.R, RT_F32);
  } else {
    L.add("dkv", m.F * d.P, 2 * d.SPC, m.kvt);
  }
}

static void proj_save_layout(const Model& m, Layout& L) {
  const mvf_head_desc& d = m.d;
  const int A = m.act;
  L.add("w.p1", d.PS, d.D, A);
  L.add("w.p2", d.D, d.PS, A);
  if (m.tc) {
    add_split(L, "w.p1", d.PS, d.D);
    add_split(L, "w.p2", d.D, d.PS);
  }
  L.add("emb", m.N, d.D, A);
  L.add("u1", m.N, d.PS, RT_F32);
  L.add("p.sum", 1, 2 * d.PS, RT_F64);
  L.add("p.bsum", 1, 2 * d.PS, RT_F64);
  L.add("p.mi", 1, 2 * d.PS, RT_F32);
  L.add("a3", m.N, d.PS, A);
  L.add("u", m.N, d.D, RT_F32);
  L.add("ehat", m.N, d.D, RT_F32);
  L.add("norm", 1, m.N, RT_F32);
}
static void proj_ws_layout(const Model& m, Layout& L) {
  const mvf_head_desc& d = m.d;
  const int A = m.act;
  L.add("du", m.N, d.D, A);
  L.add("da3", m.N, d.PS, RT_F32);
  L.add("du1", m.N, d.PS, A);
}

// flat gradient buffer: one region per packed parameter group (fp32), head + projection together
static void gpack_layout(const Model& m, Layout& L) {
  const mvf_head_desc& d = m.d;
  if (m.fwb) {
    L.add("g.w.lc", (int64_t)d.E * d.SPC, d.cls_dim, RT_F32);   // entity-major row order, un-permuted when scattered
    L.add("g.b.lc", 1, (int64_t)d.E * d.SPC, RT_F32);
  } else {
    L.add("g.Qs", d.E, d.SPC, RT_F32);
    L.add("g.Qb", 1, d.SPC, RT_F32);
    L.add("g.w.kv", 2 * d.SPC, d.C_in, RT_F32, round_up(d.C_in, 8));
    L.add("g.b.kv", 1, 2 * d.SPC, RT_F32);
  }
  int cin_ld = m.ld0;
  for (int i = 0; i < d.n_fc; ++i) {
    L.add("g." + fname(i, "w"), d.fc[i], i == 0 ? m.W0 : d.fc[i - 1], RT_F32, cin_ld);
    L.add("g." + fname(i, "b"), 1, d.fc[i], RT_F32);
    L.add("g." + fname(i, "gamma"), 1, d.fc[i], RT_F32);
    L.add("g." + fname(i, "beta"), 1, d.fc[i], RT_F32);
    cin_ld = d.fc[i];
  }
  L.add("g.w.e", m.Hin, d.n_fc ? d.fc[d.n_fc - 1] : m.W0, RT_F32, cin_ld);
  L.add("g.b.e", 1, m.Hin, RT_F32);
  for (int l = 0; l < d.L; ++l) {
    L.add("g." + lname(l, "ln0w"), 1, d.H, RT_F32);
    L.add("g." + lname(l, "ln0b"), 1, d.H, RT_F32);
    L.add("g." + lname(l, "ln1w"), 1, d.H, RT_F32);
    L.add("g." + lname(l, "ln1b"), 1, d.H, RT_F32);
    L.add("g." + lname(l, "w.qkv"), 3 * d.H, d.H, RT_F32);
    L.add("g." + lname(l, "b.qkv"), 1, 3 * d.H, RT_F32);
    L.add("g." + lname(l, "w.o"), d.H, d.H, RT_F32);
    L.add("g." + lname(l, "b.o"), 1, d.H, RT_F32);
    L.add("g." + lname(l, "w.1"), d.DFF, d.H, RT_F32);
    L.add("g." + lname(l, "b.1"), 1, d.DFF, RT_F32);
    L.add("g." + lname(l, "w.2"), d.H, d.DFF, RT_F32);
    L.add("g." + lname(l, "b.2"), 1, d.H, RT_F32);
  }
  L.add("g.w.emb", d.D, d.H, RT_F32);
  L.add("g.b.emb", 1, d.D, RT_F32);
  if (d.final_mode == MVF_FINAL_LIN) {
    L.add("g.w.lin", d.H, (int64_t)d.E * d.H, RT_F32);
    L.add("g.b.lin", 1, d.H, RT_F32);
  }
  L.add("g.w.p1", d.PS, d.D, RT_F32);
  L.add("g.b.p1", 1, d.PS, RT_F32);
  L.add("g.p.gamma", 1, d.PS, RT_F32);
  L.add("g.p.beta", 1, d.PS, RT_F32);
  L.add("g.w.p2", d.D, d.PS, RT_F32);
  L.add("g.b.p2", 1, d.D, RT_F32);
}

struct Buf {
  char* base;
  const Layout* L;
  void* p(const std::string& n) const {
    const Region* r = L->find(n);
    return r ? (void*)(base + r->off) : nullptr;
  }
  float* f(const std::string& n) const { return (float*)p(n); }
  double* dbl(const std::string& n) const { return (double*)p(n); }
  int64_t ld(const std::string& n) const { return L->find(n)->ld; }
};

// ------------------------------------------------------------------------------------------------------------
// chains
// ------------------------------------------------------------------------------------------------------------
struct Ctx {
  Model m;
  Layout Ls, Lw, Lg;
  Buf S, W, G;
  const float* const* P;
  cudaStream_t st;
  cudaStream_t side = nullptr;  // non-null: weight-gradient work is forked onto it
  int side_dev = 0;             // device the side stream / its events belong to
  bool forked = false;
  std::vector<ColsumEntry> bias_sums;   // bias gradients (column sums of dY) deferred to one batched launch at the join
  int backend;
  float p;  // effective dropout rate
  int gemm(int dtype_c, int akm, int bkm, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B,
           int64_t ldb, void* C, int64_t ldc, const float* bias, const void* relu_src = nullptr, int64_t ld_relu = 0,
           int flags = 0, int split_k = 1, cudaStream_t on = nullptr) const {
    return gemm_dispatch(backend, m.act, dtype_c, akm, bkm, M, N, K, A, lda, B, ldb, C, ldc, bias, relu_src, ld_relu,
                         flags, split_k, on ? on : st);
  }
  // the two K|V contractions: operands in the token dtype
  int gemm_kv(int dtype_c, int akm, int bkm, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B,
              int64_t ldb, void* C, int64_t ldc, const float* bias, int flags, int split_k) const {
    return gemm_dispatch(backend, m.kvt, dtype_c, akm, bkm, M, N, K, A, lda, B, ldb, C, ldc, bias, nullptr, 0, flags,
                         split_k, st);
  }
  // 1024-byte aligned scratch of the tcgen05 attention kernels inside the "fa" region of the scratch buffer (or null)
  void* fa_ws() const {
    const Region* r = Lw.find("fa");
    if (r == nullptr) return nullptr;
    return (void*)(((uintptr_t)(W.base + r->off) + 1023) & ~(uintptr_t)1023);
  }
  size_t fa_ws_bytes() const {
    const Region* r = Lw.find("fa");
    if (r == nullptr) return 0;
    const uintptr_t b0 = (uintptr_t)(W.base + r->off), b1 = (b0 + 1023) & ~(uintptr_t)1023;
    return (size_t)r->cols * 4 - (size_t)(b1 - b0);
  }
  int fork() {
    if (!side) return MVF_OK;
    cudaEvent_t e = side_event(side_dev);
    MVF_CHECK_CUDA(cudaEventRecord(e, st));
    MVF_CHECK_CUDA(cudaStreamWaitEvent(side, e, 0));
    forked = true;
    return MVF_OK;
  }
  // bias gradients collected so far: one batched column-sum launch on the side stream, without joining (the dY regions they
  // read are never rewritten before the final join)
  int flush_bias_sums() {
    if (bias_sums.empty()) return MVF_OK;
    MVF_TRY(fork());
    MVF_TRY(colsum_batched(bias_sums.data(), (int)bias_sums.size(), side && forked ? side : st));
    bias_sums.clear();
    return MVF_OK;
  }
  int join() {
    if (!bias_sums.empty()) {   // every dY region is still intact here (callers give each its own scratch region)
      MVF_TRY(colsum_batched(bias_sums.data(), (int)bias_sums.size(), side && forked ? side : st));
      bias_sums.clear();
    }
    if (!side || !forked) return MVF_OK;
    cudaEvent_t e = side_event(side_dev);
    MVF_CHECK_CUDA(cudaEventRecord(e, side));
    MVF_CHECK_CUDA(cudaStreamWaitEvent(st, e, 0));
    forked = false;
    return MVF_OK;
  }
  // y = x W^T + b.  Forward GEMMs feed the softmax/temperature non-linearities of SCL: on the tensor-core backend they
  // run as bf16x3, with the weight taken from its pre-split copy `wsplit` (region "<name>.s") when there is one, and
  // K >= 512 contractions that cannot fill the machine are split along K (TMA reduce-add epilogue).
  // one memset over the adjacent save regions first .. last (the BatchNorm sum buffers of a call)
  int zero_span(const std::string& first, const std::string& last) const {
    const Region* a = Ls.find(first);
    const Region* b = Ls.find(last);
    MVF_REQUIRE(a && b && b->off >= a->off, MVF_ERR_BAD_ARG, "zero_span: %s .. %s", first.c_str(), last.c_str());
    const size_t esz = b->dtype == RT_F64 ? 8 : 4;
    MVF_CHECK_CUDA(cudaMemsetAsync(S.base + a->off, 0, b->off - a->off + (size_t)b->rows * b->ld * esz, st));
    return MVF_OK;
  }
  int linear(int dtype_c, int64_t M, int64_t N, int64_t K, const void* x, int64_t ldx, const void* Wp, int64_t ldw,
             const float* bias, void* y, int64_t ldy, int flags = 0, const char* wsplit = nullptr) const {
    const int sk = (m.tc && !(flags & MVF_GEMM_RELU) && dtype_c == MVF_F32) ? 0 : 1;
    if (m.tc && wsplit != nullptr) {
      const Region* r = Ls.find(std::string(wsplit) + ".s");
      if (r != nullptr)
        return gemm(dtype_c, 1, 1, M, N, K, x, ldx, S.base + r->off, r->ld, y, ldy, bias, nullptr, 0,
                    flags | MVF_GEMM_SPLIT3 | MVF_GEMM_B_PRESPLIT, sk);
    }
    return gemm(dtype_c, 1, 1, M, N, K, x, ldx, Wp, ldw, y, ldy, bias, nullptr, 0, flags | MVF_GEMM_SPLIT3, sk);
  }
  // dX = dY W (optionally masked by relu_src > 0)
  int linear_dx(int dtype_c, int64_t M, int64_t Nout, int64_t Kin, const void* dY, int64_t lddy, const void* Wp,
                int64_t ldw, void* dX, int64_t lddx, const void* relu_src = nullptr, int64_t ld_relu = 0) const {
    return gemm(dtype_c, 1, 0, M, Kin, Nout, dY, lddy, Wp, ldw, dX, lddx, nullptr, relu_src, ld_relu,
                relu_src ? MVF_GEMM_RELUMASK : 0);
  }
  // dW += dY^T X ; db += colsum(dY)
  // (dY and X must stay untouched until join(): callers give every dY its own scratch region)
  int linear_dw(int64_t M, int64_t Nout, int64_t Kin, const void* dY, int64_t lddy, const void* X, int64_t ldx,
                float* dWp, int64_t lddw, float* db, int split_k = 0) {
    MVF_TRY(fork());
    cudaStream_t on = side ? side : st;
    MVF_TRY(gemm(MVF_F32, 0, 0, Nout, Kin, M, dY, lddy, X, ldx, dWp, lddw, nullptr, nullptr, 0, MVF_GEMM_ACCUM, split_k, on));
    if (db) bias_sums.push_back(ColsumEntry{(const float*)dY, db, M, lddy, (int)Nout});   // m.act is fp32
    return MVF_OK;
  }
};

static int make_ctx(const mvf_head_desc* d, Ctx& c, bool proj, void* save, size_t save_bytes, void* ws, size_t ws_bytes,
                    float* gpack, const float* const* params, cudaStream_t st) {
  MVF_TRY(build_model(d, c.m));
  if (proj) { proj_save_layout(c.m, c.Ls); proj_ws_layout(c.m, c.Lw); }
  else { head_save_layout(c.m, c.Ls); head_ws_layout(c.m, c.Lw); }
  gpack_layout(c.m, c.Lg);
  MVF_REQUIRE(save != nullptr && save_bytes >= c.Ls.total, MVF_ERR_WORKSPACE, "save buffer %zu B < required %zu B",
              save_bytes, c.Ls.total);
  MVF_REQUIRE(ws != nullptr && ws_bytes >= c.Lw.total, MVF_ERR_WORKSPACE, "scratch buffer %zu B < required %zu B",
              ws_bytes, c.Lw.total);
  MVF_REQUIRE((((uintptr_t)save) & 255) == 0 && (((uintptr_t)ws) & 255) == 0, MVF_ERR_ALIGN,
              "save / scratch buffers must be 256-byte aligned");
  MVF_REQUIRE(params != nullptr, MVF_ERR_BAD_ARG, "null parameter table");
  c.S = Buf{(char*)save, &c.Ls};
  c.W = Buf{(char*)ws, &c.Lw};
  c.G = Buf{(char*)gpack, &c.Lg};
  c.P = params;
  c.st = st;
  c.backend = d->gemm_backend;
  // AUTO: bf16 tokens -> tensor cores everywhere (bf16 K|V contractions, tf32 for the fp32 activations behind the
  // pooling); fp32 tokens -> exact fp32 FMA everywhere (the 1e-5 parity mode).  TCGEN05 with fp32 tokens = all tf32.
  if (c.backend == MVF_GEMM_AUTO) c.backend = d->dtype == MVF_BF16 ? MVF_GEMM_TCGEN05 : MVF_GEMM_SIMT;
  c.p = d->training ? d->drop_p : 0.f;
  // the head needs embed.*; the projection needs ssl_projection.* (entries of the other part may be null)
  for (int i = 0; i < (int)c.m.params.size(); ++i) {
    const bool is_proj = i >= c.m.iWp1;
    if (is_proj == proj)
      MVF_REQUIRE(params[i] != nullptr, MVF_ERR_BAD_ARG, "parameter %d (%s) is null", i, c.m.params[i].name.c_str());
  }
  return MVF_OK;
}

static int pack_head_weights(Ctx& c, cudaStream_t on) {
  const Model& m = c.m;
  const mvf_head_desc& d = m.d;
  const int bf = m.act == MVF_BF16;
  std::vector<PackEntry> e;
  auto mat = [&](const std::string& reg, int pidx, int row0 = 0) {
    const Region* r = c.Ls.find(reg);
    const ParamInfo& pi = m.params[pidx];
    char* dst = c.S.base + r->off + (size_t)row0 * r->ld * rt_size(r->dtype);
    e.push_back(PackEntry{c.P[pidx], dst, (int)pi.rows, (int)pi.cols, (int)r->ld, r->dtype == RT_BF16});
  };
  auto vec = [&](const std::string& reg, int pidx, int col0) {
    const Region* r = c.Ls.find(reg);
    const ParamInfo& pi = m.params[pidx];
    e.push_back(PackEntry{c.P[pidx], c.S.base + r->off + (size_t)col0 * 4, 1, (int)pi.cols, (int)pi.cols, 0});
  };
  (void)bf;
  auto split = [&](const std::string& reg, int pidx, int row0 = 0) {   // pre-split copy for the forward bf16x3 GEMMs
    if (!m.tc) return;
    const Region* r = c.Ls.find(reg + ".s");
    if (r == nullptr) return;
    const ParamInfo& pi = m.params[pidx];
    char* dst = c.S.base + r->off + (size_t)row0 * r->ld * 4;
    e.push_back(PackEntry{c.P[pidx], dst, (int)pi.rows, (int)pi.cols, (int)r->ld, 2});
  };
  if (m.fwb) {
    const Region* rw = c.Ls.find("w.lc");
    const Region* rb = c.Ls.find("b.lc");
    const ParamInfo& pw = m.params[m.iWlc];
    e.push_back(PackEntry{c.P[m.iWlc], c.S.base + rw->off, (int)pw.rows, (int)pw.cols, (int)rw->ld, 0, d.SPC});
    e.push_back(PackEntry{c.P[m.iblc], c.S.base + rb->off, 1, (int)pw.rows, (int)pw.rows, 0, d.SPC});
    if (m.tc) {
      const Region* rs = c.Ls.find("w.lc.s");
      e.push_back(PackEntry{c.P[m.iWlc], c.S.base + rs->off, (int)pw.rows, (int)pw.cols, (int)rs->ld, 2, d.SPC});
    }
  }
  if (m.fold) split("w.v", m.iWv);
  if (!m.fold && !m.fwb) {
    mat("w.kv", m.iWk, 0);
    mat("w.kv", m.iWv, d.SPC);
    vec("b.kv", m.ibk, 0);
    vec("b.kv", m.ibv, d.SPC);
  }
  for (int i = 0; i < d.n_fc; ++i) { mat(fname(i, "w"), m.iFcW[i]); split(fname(i, "w"), m.iFcW[i]); }
  mat("w.e", m.iWe);
  split("w.e", m.iWe);
  for (int l = 0; l < d.L; ++l) {
    const int b = m.iLayer[l];
    mat(lname(l, "w.qkv"), b + L_WQ, 0);
    mat(lname(l, "w.qkv"), b + L_WK, d.H);
    mat(lname(l, "w.qkv"), b + L_WV, 2 * d.H);
    split(lname(l, "w.qkv"), b + L_WQ, 0);
    split(lname(l, "w.qkv"), b + L_WK, d.H);
    split(lname(l, "w.qkv"), b + L_WV, 2 * d.H);
    split(lname(l, "w.o"), b + L_WO);
    split(lname(l, "w.1"), b + L_W1);
    split(lname(l, "w.2"), b + L_W2);
    vec(lname(l, "b.qkv"), b + L_BQ, 0);
    vec(lname(l, "b.qkv"), b + L_BK, d.H);
    vec(lname(l, "b.qkv"), b + L_BV, 2 * d.H);
    mat(lname(l, "w.o"), b + L_WO);
    mat(lname(l, "w.1"), b + L_W1);
    mat(lname(l, "w.2"), b + L_W2);
  }
  mat("w.emb", m.iWemb);
  split("w.emb", m.iWemb);
  if (d.final_mode == MVF_FINAL_LIN) { mat("w.lin", m.iWlin); split("w.lin", m.iWlin); }
  return pack_params(e.data(), (int)e.size(), on);
}

static double bn_n_global(const Model& m, int64_t local_rows) {
  return (double)local_rows * (double)(m.d.world_size > 1 ? m.d.world_size : 1);
}

static int head_forward_impl(Ctx& c, float* const* bn_running, int64_t* const* bn_tracked, const void* tokens,
                             const float* mask, float* out_emb, float* attn_out, int ph0, int ph1) {
  const Model& m = c.m;
  const mvf_head_desc& d = m.d;
  const int A = m.act;
  cudaStream_t st = c.st;
  const int n_ph = d.n_fc + 1;
  if (ph1 > n_ph) ph1 = n_ph;
  if (ph0 == 0 && d.training && d.n_fc > 0) MVF_TRY(c.zero_span(fname(0, "sum"), fname(d.n_fc - 1, "bsum")));
  for (int ph = ph0; ph < ph1; ++ph) {
    if (ph == 0) {
      // the packed / pre-split copies of the weights are first needed behind the streaming pooling pass: with folded pooling
      // they are written on the side stream while that HBM-bound pass runs (joined below)
      const bool pack_aside = m.fold && c.side != nullptr;
      if (pack_aside) {
        MVF_TRY(c.fork());
        MVF_TRY(pack_head_weights(c, c.side));
      } else {
        MVF_TRY(pack_head_weights(c, st));
      }
      float* attn = m.fwb ? nullptr : c.S.f("attn");
      if (m.fwb) {
        // FWBPooling (mvformer.py:455-462): ent[(f, e), c] = lin_conv(cls[f])[c*E + e]; with the weight rows regrouped
        // entity-major the [F, E*SPC] product IS the [F*E, SPC] entity matrix
        MVF_REQUIRE(d.cls_emb != nullptr, MVF_ERR_BAD_ARG, "FIXED_WIDTH_BASELINE needs the CLS embeddings (desc.cls_emb)");
        const int64_t NE = (int64_t)d.E * d.SPC;
        MVF_TRY(c.linear(MVF_F32, m.F, NE, d.cls_dim, d.cls_emb, d.cls_dim, c.S.p("w.lc"), d.cls_dim, c.S.f("b.lc"), c.S.p("ent32"), NE, 0, "w.lc"));
        MVF_TRY(ent_finish_fwd(c.S.f("ent32"), c.S.f("h0"), m.ld0, m.R, d.SPC, d.E, d.one_hot == MVF_ONEHOT_POOL, c.p,
                               DropSeed(d.seed, d.seed_dev), st));
      } else if (m.fold) {
        // a3-a5 folded: Wq = Q Wk / sqrt(SPC); one streaming pass over the tokens (online softmax + pooling of X);
        // the value projection is applied to the E pooled rows per frame instead of the P token rows
        {
          ProfScope ps(2, st);
          MVF_TRY(fold_prep(c.P[m.iQs], c.P[m.iQb], c.P[m.iWk], d.E, d.SPC, d.C_in, c.S.f("wq"), st));
        }
        {
          ProfScope ps(0, st);
          MVF_TRY(pool_fold_fwd(m.kvt, (int)m.F, d.P, d.E, d.C_in, tokens, c.S.f("wq"), attn, c.S.f("px"), st));
        }
        if (pack_aside) MVF_TRY(c.join());
        {
          ProfScope ps(4, st);
          MVF_TRY(c.linear(MVF_F32, m.R, d.SPC, d.C_in, c.S.p("px"), d.C_in, c.P[m.iWv], d.C_in, c.P[m.ibv], c.S.p("ent32"),
                           d.SPC, 0, "w.v"));
          MVF_TRY(ent_finish_fwd(c.S.f("ent32"), c.S.f("h0"), m.ld0, m.R, d.SPC, d.E, d.one_hot == MVF_ONEHOT_POOL, c.p,
                                 DropSeed(d.seed, d.seed_dev), st));
        }
      } else {
        // a4 as written: K|V projection of every patch token -- the dominant contraction
        {
          ProfScope ps(0, st);
          MVF_TRY(c.gemm_kv(m.kvt, 1, 1, m.F * d.P, 2 * d.SPC, d.C_in, tokens, d.C_in, c.S.p("w.kv"), c.S.ld("w.kv"),
                            c.S.p("kv"), 2 * d.SPC, c.S.f("b.kv"), 0, 1));
        }
        {
          ProfScope ps(2, st);
          MVF_TRY(xattn_pool_fwd(m.kvt, (int)m.F, d.P, d.E, d.SPC, c.S.p("kv"), c.P[m.iQs], c.P[m.iQb], attn, c.S.p("h0"),
                                 m.ld0, c.S.f("ent32"), d.one_hot == MVF_ONEHOT_POOL, c.p, DropSeed(d.seed, d.seed_dev), st));
        }
      }
      if (attn_out && attn)
        MVF_CHECK_CUDA(cudaMemcpyAsync(attn_out, attn, (size_t)m.F * d.E * d.P * 4, cudaMemcpyDeviceToDevice, st));
    }
    // input of FC layer `ph` (or of video_emb when ph == n_fc)
    const void* xin;
    int64_t ldin, kin;
    if (ph == 0) { xin = c.S.p("h0"); ldin = m.ld0; kin = m.ld0; }
    else {
      const int i = ph - 1;
      const int C = d.fc[i];
      float* rm = bn_running ? bn_running[2 * i] : nullptr;
      float* rv = bn_running ? bn_running[2 * i + 1] : nullptr;
      MVF_REQUIRE(d.training || (rm && rv), MVF_ERR_BAD_ARG, "eval mode needs BatchNorm running statistics");
      // dropout of the NEXT FC layer is applied to this activation (fc_layers.{4i}: Dropout before Linear)
      const float pn = (i + 1 < d.n_fc) ? c.p : 0.f;
      const int fused = bn_finalize_apply(c.S.dbl(fname(i, "sum")), C, bn_n_global(m, m.R), d.bn_eps, d.training, d.bn_momentum, rm,
                                          rv, bn_tracked ? bn_tracked[i] : nullptr, c.S.f(fname(i, "mi")), A, c.S.f(fname(i, "x")),
                                          m.R, c.P[m.iFcG[i]], c.P[m.iFcBeta[i]], 1, c.S.p(fname(i, "a")), C, pn,
                                          DropSeed(d.seed, d.seed_dev), SITE_FC0 + i + 1, st);
      if (fused == MVF_ERR_UNSUPPORTED) {
        MVF_TRY(bn_finalize(c.S.dbl(fname(i, "sum")), C, bn_n_global(m, m.R), d.bn_eps, d.training, d.bn_momentum, rm, rv,
                            bn_tracked ? bn_tracked[i] : nullptr, c.S.f(fname(i, "mi")), st));
        MVF_TRY(bn_apply(A, c.S.f(fname(i, "x")), m.R, C, c.S.f(fname(i, "mi")), c.P[m.iFcG[i]], c.P[m.iFcBeta[i]], 1,
                         c.S.p(fname(i, "a")), C, pn, DropSeed(d.seed, d.seed_dev), SITE_FC0 + i + 1, st));
      } else {
        MVF_TRY(fused);
      }
      xin = c.S.p(fname(i, "a")); ldin = C; kin = C;
    }
    if (ph < d.n_fc) {
      const int C = d.fc[ph];
      MVF_TRY(c.linear(MVF_F32, m.R, C, kin, xin, ldin, c.S.p(fname(ph, "w")), c.S.ld(fname(ph, "w")), c.P[m.iFcB[ph]],
                       c.S.p(fname(ph, "x")), C, 0, fname(ph, "w").c_str()));
      if (d.training) MVF_TRY(bn_stats(c.S.f(fname(ph, "x")), m.R, C, c.S.dbl(fname(ph, "sum")), st));   // zeroed at phase 0
      continue;
    }
    // ---- last phase: video_emb, positional encoding, temporal encoder, entity reduction, embedding ----
    MVF_TRY(c.linear(MVF_F32, m.R, m.Hin, kin, xin, ldin, c.S.p("w.e"), c.S.ld("w.e"), c.P[m.ibe], c.S.p("h3"), m.Hin, 0, "w.e"));
    MVF_TRY(posenc_table(c.S.f("pe"), d.T, d.H, d.train_frames, st));
    MVF_TRY(posenc_add(c.S.f("h3"), c.S.f("pe"), c.S.f("z0"), d.BV, d.T, d.E, d.H, c.p, DropSeed(d.seed, d.seed_dev), st));
    const float* keymask_src = d.has_mask ? mask : nullptr;
    // key mask over s = e*T + t replicates the frame mask per entity (mvformer.py:174-177)
    float* keymask = nullptr;
    if (keymask_src) {
      keymask = c.W.f("delta");  // scratch reuse: [BV*heads*S] >= [BV*S]
      for (int e = 0; e < d.E; ++e)
        MVF_CHECK_CUDA(cudaMemcpy2DAsync(keymask + (size_t)e * d.T, (size_t)m.S * 4, keymask_src, (size_t)d.T * 4,
                                         (size_t)d.T * 4, d.BV, cudaMemcpyDeviceToDevice, st));
    }
    const float* pend_o = nullptr;
    int pend_site = 0;
    int zi = 0;
    float* o = c.W.f("o");
    const int dk = d.H / d.heads;
    for (int l = 0; l < d.L; ++l) {
      const int b = m.iLayer[l];
      const std::string zin = "z" + std::to_string(zi), zmid = "z" + std::to_string(2 * l + 1);
      const std::string z0n = "z" + std::to_string(2 * l);
      float* ln0 = c.S.f(lname(l, "ln0"));
      // z[2l] = z_prev + drop(pending o); r0 = LN(z[2l])
      MVF_TRY(ln_fwd(A, c.S.f(zin), pend_o, c.S.f(z0n), c.S.p(lname(l, "r0")), ln0, ln0 + m.rows, c.P[b + L_LN0W],
                     c.P[b + L_LN0B], m.rows, d.H, d.ln_eps, c.p, DropSeed(d.seed, d.seed_dev), pend_site, st));
      MVF_TRY(c.linear(A, m.rows, 3 * d.H, d.H, c.S.p(lname(l, "r0")), d.H, c.S.p(lname(l, "w.qkv")), d.H,
                       c.S.f(lname(l, "b.qkv")), c.S.p(lname(l, "qkv")), 3 * d.H, 0, lname(l, "w.qkv").c_str()));
      {
        ProfScope ps(6, st);
        MVF_TRY(attention_fwd(A, d.BV, (int)m.S, d.heads, dk, c.S.p(lname(l, "qkv")), keymask, c.S.p(lname(l, "ctx")),
                              c.S.f(lname(l, "lse")), st, m.tc, c.fa_ws(), c.fa_ws_bytes()));
      }
      MVF_TRY(c.linear(MVF_F32, m.rows, d.H, d.H, c.S.p(lname(l, "ctx")), d.H, c.S.p(lname(l, "w.o")), d.H, c.P[b + L_BO],
                       o, d.H, 0, lname(l, "w.o").c_str()));
      float* ln1 = c.S.f(lname(l, "ln1"));
      MVF_TRY(ln_fwd(A, c.S.f(z0n), o, c.S.f(zmid), c.S.p(lname(l, "r1")), ln1, ln1 + m.rows, c.P[b + L_LN1W],
                     c.P[b + L_LN1B], m.rows, d.H, d.ln_eps, c.p, DropSeed(d.seed, d.seed_dev), SITE_ENC0 + 2 * l, st));
      MVF_TRY(c.linear(A, m.rows, d.DFF, d.H, c.S.p(lname(l, "r1")), d.H, c.S.p(lname(l, "w.1")), d.H, c.P[b + L_B1],
                       c.S.p(lname(l, "f")), d.DFF, MVF_GEMM_RELU, lname(l, "w.1").c_str()));
      MVF_TRY(c.linear(MVF_F32, m.rows, d.H, d.DFF, c.S.p(lname(l, "f")), d.DFF, c.S.p(lname(l, "w.2")), d.DFF,
                       c.P[b + L_B2], o, d.H, 0, lname(l, "w.2").c_str()));
      pend_o = o;
      pend_site = SITE_ENC0 + 2 * l + 1;
      zi = 2 * l + 1;
    }
    const std::string zlast = "z" + std::to_string(2 * d.L);
    if (d.L > 0)
      MVF_TRY(ln_fwd(A, c.S.f("z" + std::to_string(zi)), pend_o, c.S.f(zlast), nullptr, nullptr, nullptr, nullptr, nullptr,
                     m.rows, d.H, d.ln_eps, c.p, DropSeed(d.seed, d.seed_dev), pend_site, st));
    // a9: entity reduction + embedding layer
    if (d.final_mode == MVF_FINAL_LIN) {
      MVF_TRY(entity_gather_lin(A, c.S.f(zlast), c.S.p("zl"), d.BV, d.T, d.E, d.H, st));
      MVF_TRY(c.linear(MVF_F32, m.N, d.H, (int64_t)d.E * d.H, c.S.p("zl"), (int64_t)d.E * d.H, c.S.p("w.lin"),
                       (int64_t)d.E * d.H, c.P[m.iblin], c.S.p("ylin"), d.H, 0, "w.lin"));
      MVF_TRY(cast_f32(A, c.S.f("ylin"), c.S.p("y"), m.N * d.H, st));
    } else {
      MVF_TRY(entity_reduce_fwd(A, c.S.f(zlast), c.S.p("y"), d.final_mode == MVF_FINAL_MAX ? (int32_t*)c.S.p("argmax") : nullptr,
                                d.BV, d.T, d.E, d.H, d.final_mode, st));
    }
    MVF_TRY(c.linear(MVF_F32, m.N, d.D, d.H, c.S.p("y"), d.H, c.S.p("w.emb"), d.H, c.P[m.ibemb], out_emb, d.D, 0, "w.emb"));
  }
  return MVF_OK;
}

static int head_backward_impl(Ctx& c, const void* tokens, const float* mask, const float* d_emb, int ph0, int ph1) {
  const Model& m = c.m;
  const mvf_head_desc& d = m.d;
  const int A = m.act;
  cudaStream_t st = c.st;
  MVF_REQUIRE(d.training, MVF_ERR_BAD_ARG, "backward needs training = 1 (batch statistics were not saved in eval mode)");
  // phases 0 .. n_fc are cut at the BatchNorm statistics (as in forward); the pooling backward is a phase of its own
  // (n_fc + 1) so that a multi-GPU caller can start all-reducing the chain's gradients while it runs
  const int n_ph = d.n_fc + 2;
  if (ph1 > n_ph) ph1 = n_ph;
  if (ph0 == 0 && d.n_fc > 0) MVF_TRY(c.zero_span(fname(0, "sum"), fname(d.n_fc - 1, "bsum")));   // the forward sums are spent
  for (int ph = ph0; ph < ph1; ++ph) {
    const void* d_in;   // gradient w.r.t. the output of the FC Linear handled in this phase (act dtype)
    int64_t ld_din;
    if (ph == 0) {
      // ---- embedding layer + entity reduction ----
      MVF_TRY(cast_f32(A, d_emb, c.W.p("dEact"), m.N * d.D, st));
      MVF_TRY(c.linear_dw(m.N, d.D, d.H, c.W.p("dEact"), d.D, c.S.p("y"), d.H, c.G.f("g.w.emb"), d.H, c.G.f("g.b.emb")));
      MVF_TRY(c.linear_dx(MVF_F32, m.N, d.D, d.H, c.W.p("dEact"), d.D, c.S.p("w.emb"), d.H, c.W.p("dy"), d.H));
      float* dz = c.W.f("dzA");
      float* dz_other = c.W.f("dzB");
      if (d.final_mode == MVF_FINAL_LIN) {
        MVF_TRY(cast_f32(A, c.W.f("dy"), c.W.p("dylin"), m.N * d.H, st));
        MVF_TRY(c.linear_dw(m.N, d.H, (int64_t)d.E * d.H, c.W.p("dylin"), d.H, c.S.p("zl"), (int64_t)d.E * d.H,
                            c.G.f("g.w.lin"), (int64_t)d.E * d.H, c.G.f("g.b.lin")));
        MVF_TRY(c.linear_dx(MVF_F32, m.N, d.H, (int64_t)d.E * d.H, c.W.p("dylin"), d.H, c.S.p("w.lin"),
                            (int64_t)d.E * d.H, c.W.p("dzl"), (int64_t)d.E * d.H));
        MVF_TRY(entity_scatter_lin(c.W.f("dzl"), dz, d.BV, d.T, d.E, d.H, st));
      } else {
        MVF_TRY(entity_reduce_bwd(c.W.f("dy"), d.final_mode == MVF_FINAL_MAX ? (const int32_t*)c.S.p("argmax") : nullptr,
                                  dz, d.BV, d.T, d.E, d.H, d.final_mode, st));
      }
      // ---- temporal encoder, last layer first ----
      float* keymask = nullptr;
      if (d.has_mask && mask) {
        keymask = c.W.f("o");  // scratch: forward's sublayer buffer is free during backward
        for (int e = 0; e < d.E; ++e)
          MVF_CHECK_CUDA(cudaMemcpy2DAsync(keymask + (size_t)e * d.T, (size_t)m.S * 4, mask, (size_t)d.T * 4,
                                           (size_t)d.T * 4, d.BV, cudaMemcpyDeviceToDevice, st));
      }
      const int dk = d.H / d.heads;
      for (int l = d.L - 1; l >= 0; --l) {
        const int b = m.iLayer[l];
        const std::string g = "g." + std::string("l") + std::to_string(l) + ".";
        // FFN branch: z[2l+2] = z[2l+1] + drop(W2 f + b2)
        void* dgF = c.W.p(lname(l, "dgF"));
        void* dgA = c.W.p(lname(l, "dgA"));
        void* df = c.W.p(lname(l, "df"));
        void* dqkv = c.W.p(lname(l, "dqkv"));
        // dgF = drop'(dz): written by the ln_bwd of the layer above (fused), by a stand-alone launch for the top layer
        if (l == d.L - 1) MVF_TRY(dropout_cast(A, dz, dgF, m.rows, d.H, d.H, c.p, DropSeed(d.seed, d.seed_dev), SITE_ENC0 + 2 * l + 1, st));
        MVF_TRY(c.linear_dw(m.rows, d.H, d.DFF, dgF, d.H, c.S.p(lname(l, "f")), d.DFF, c.G.f(g + "w.2"), d.DFF,
                            c.G.f(g + "b.2")));
        MVF_TRY(c.linear_dx(A, m.rows, d.H, d.DFF, dgF, d.H, c.S.p(lname(l, "w.2")), d.DFF, df, d.DFF,
                            c.S.p(lname(l, "f")), d.DFF));
        MVF_TRY(c.linear_dw(m.rows, d.DFF, d.H, df, d.DFF, c.S.p(lname(l, "r1")), d.H, c.G.f(g + "w.1"), d.H,
                            c.G.f(g + "b.1")));
        MVF_TRY(c.linear_dx(MVF_F32, m.rows, d.DFF, d.H, df, d.DFF, c.S.p(lname(l, "w.1")), d.H, c.W.p("dr"), d.H));
        const float* ln1 = c.S.f(lname(l, "ln1"));
        // ... and the masked copy the attention branch consumes (dgA = drop'(dz_out), site 2l) comes out of the same launch
        MVF_TRY(ln_bwd(c.W.f("dr"), c.S.f("z" + std::to_string(2 * l + 1)), ln1, ln1 + m.rows, c.P[b + L_LN1W], dz,
                       dz_other, c.G.f(g + "ln1w"), c.G.f(g + "ln1b"), m.rows, d.H, st, (float*)dgA, c.p, DropSeed(d.seed, d.seed_dev),
                       SITE_ENC0 + 2 * l));
        std::swap(dz, dz_other);
        // attention branch: z[2l+1] = z[2l] + drop(Wo ctx + bo)
        MVF_TRY(c.linear_dw(m.rows, d.H, d.H, dgA, d.H, c.S.p(lname(l, "ctx")), d.H, c.G.f(g + "w.o"), d.H,
                            c.G.f(g + "b.o")));
        MVF_TRY(c.linear_dx(A, m.rows, d.H, d.H, dgA, d.H, c.S.p(lname(l, "w.o")), d.H, c.W.p("dctx"), d.H));
        {
          ProfScope ps(7, st);
          MVF_TRY(attention_bwd(A, d.BV, (int)m.S, d.heads, dk, c.S.p(lname(l, "qkv")), keymask, c.S.p(lname(l, "ctx")),
                                c.S.f(lname(l, "lse")), c.W.p("dctx"), dqkv, c.W.f("delta"), st, m.tc, c.fa_ws(), c.fa_ws_bytes()));
        }
        MVF_TRY(c.linear_dw(m.rows, 3 * d.H, d.H, dqkv, 3 * d.H, c.S.p(lname(l, "r0")), d.H, c.G.f(g + "w.qkv"),
                            d.H, c.G.f(g + "b.qkv")));
        MVF_TRY(c.linear_dx(MVF_F32, m.rows, 3 * d.H, d.H, dqkv, 3 * d.H, c.S.p(lname(l, "w.qkv")), d.H,
                            c.W.p("dr"), d.H));
        const float* ln0 = c.S.f(lname(l, "ln0"));
        // the layer below consumes drop'(dz_out) with the site of ITS FFN branch (2(l-1)+1)
        float* next_dgF = l > 0 ? (float*)c.W.p(lname(l - 1, "dgF")) : nullptr;
        MVF_TRY(ln_bwd(c.W.f("dr"), c.S.f("z" + std::to_string(2 * l)), ln0, ln0 + m.rows, c.P[b + L_LN0W], dz, dz_other,
                       c.G.f(g + "ln0w"), c.G.f(g + "ln0b"), m.rows, d.H, st, next_dgF, c.p, DropSeed(d.seed, d.seed_dev),
                       SITE_ENC0 + 2 * (l - 1) + 1));
        std::swap(dz, dz_other);
      }
      // ---- positional encoding (+dropout) and video_emb ----
      MVF_TRY(posenc_bwd(A, dz, c.W.p("dh3"), d.BV, d.T, d.E, d.H, c.p, DropSeed(d.seed, d.seed_dev), st));
      const void* xin = d.n_fc ? c.S.p(fname(d.n_fc - 1, "a")) : c.S.p("h0");
      const int64_t ldin = d.n_fc ? d.fc[d.n_fc - 1] : m.ld0;
      MVF_TRY(c.linear_dw(m.R, m.Hin, ldin, c.W.p("dh3"), m.Hin, xin, ldin, c.G.f("g.w.e"), c.Lg.find("g.w.e")->ld,
                          c.G.f("g.b.e")));
      if (d.n_fc == 0) {
        MVF_TRY(c.linear_dx(A, m.R, m.Hin, m.ld0, c.W.p("dh3"), m.Hin, c.S.p("w.e"), c.S.ld("w.e"), c.W.p("dh0"), m.ld0));
      } else {
        const int i = d.n_fc - 1, C = d.fc[i];
        MVF_TRY(c.linear_dx(MVF_F32, m.R, m.Hin, C, c.W.p("dh3"), m.Hin, c.S.p("w.e"), c.S.ld("w.e"), c.W.p("da"), C));
        MVF_TRY(bn_bwd_stats(c.W.f("da"), C, c.S.f(fname(i, "x")), m.R, C, c.S.f(fname(i, "mi")), c.P[m.iFcG[i]],
                             c.P[m.iFcBeta[i]], 1, 0.f, DropSeed(d.seed, d.seed_dev), SITE_FC0 + i + 1, c.S.dbl(fname(i, "bsum")),
                             c.G.f("g." + fname(i, "gamma")), c.G.f("g." + fname(i, "beta")), st));
      }
    }
    if (ph >= 1 && ph <= d.n_fc) {
      // BatchNorm i = n_fc - ph: finish its backward with the (all-reduced) sums, then the Linear in front of it
      const int i = d.n_fc - ph, C = d.fc[i];
      const float pn = (i + 1 < d.n_fc) ? c.p : 0.f;  // dropout that followed this activation
      MVF_TRY(bn_bwd_apply(A, c.W.f("da"), C, c.S.f(fname(i, "x")), m.R, C, c.S.f(fname(i, "mi")), c.P[m.iFcG[i]],
                           c.P[m.iFcBeta[i]], 1, pn, DropSeed(d.seed, d.seed_dev), SITE_FC0 + i + 1, c.S.dbl(fname(i, "bsum")),
                           bn_n_global(m, m.R), c.W.p(fname(i, "dx")), C, st));
      d_in = c.W.p(fname(i, "dx"));
      ld_din = C;
      const void* xin = i > 0 ? c.S.p(fname(i - 1, "a")) : c.S.p("h0");
      const int64_t ldin = i > 0 ? d.fc[i - 1] : m.ld0;
      const std::string gw = "g." + fname(i, "w");
      MVF_TRY(c.linear_dw(m.R, C, ldin, d_in, ld_din, xin, ldin, c.G.f(gw), c.Lg.find(gw)->ld, c.G.f("g." + fname(i, "b"))));
      if (i > 0) {
        const int Cp = d.fc[i - 1];
        MVF_TRY(c.linear_dx(MVF_F32, m.R, C, Cp, d_in, ld_din, c.S.p(fname(i, "w")), c.S.ld(fname(i, "w")), c.W.p("da"), Cp));
        const float pp = c.p;  // activation i-1 was followed by the dropout of FC layer i
        MVF_TRY(bn_bwd_stats(c.W.f("da"), Cp, c.S.f(fname(i - 1, "x")), m.R, Cp, c.S.f(fname(i - 1, "mi")),
                             c.P[m.iFcG[i - 1]], c.P[m.iFcBeta[i - 1]], 1, pp, DropSeed(d.seed, d.seed_dev), SITE_FC0 + i,
                             c.S.dbl(fname(i - 1, "bsum")), c.G.f("g." + fname(i - 1, "gamma")),
                             c.G.f("g." + fname(i - 1, "beta")), st));
      } else {
        MVF_TRY(c.linear_dx(A, m.R, C, m.ld0, d_in, ld_din, c.S.p(fname(0, "w")), c.S.ld(fname(0, "w")), c.W.p("dh0"), m.ld0));
      }
    }
    if (ph == n_ph - 1) {
      // ---- entity cross-attention pooling and the K|V projection weight gradient ----
      if (m.fwb) {
        MVF_REQUIRE(d.cls_emb != nullptr, MVF_ERR_BAD_ARG, "FIXED_WIDTH_BASELINE needs the CLS embeddings (desc.cls_emb)");
        const int64_t NE = (int64_t)d.E * d.SPC;
        MVF_TRY(ent_finish_bwd(c.W.f("dh0"), m.ld0, c.W.f("dent"), m.R, d.SPC, m.W0, c.p, DropSeed(d.seed, d.seed_dev), st));
        // d(lin_conv) in entity-major row order: dW' = dEnt[F, E*SPC]^T cls, db' = column sums (cls is frozen: no dX)
        MVF_TRY(c.linear_dw(m.F, NE, d.cls_dim, c.W.p("dent"), NE, d.cls_emb, d.cls_dim, c.G.f("g.w.lc"), d.cls_dim, c.G.f("g.b.lc")));
        continue;
      }
      const int o_spc = d.SPC;
      float* gbkv = c.G.f("g.b.kv");
      if (m.fold) {
        const int64_t ldg = c.Lg.find("g.w.kv")->ld;
        float* gwk = c.G.f("g.w.kv");
        float* gwv = gwk + (size_t)d.SPC * ldg;
        // the bias column sums of the chain so far go to the side stream now and overlap the GEMMs below; the side stream is
        // joined BEFORE the streaming pass, which is HBM-bound on every SM and loses more to a concurrent kernel than it hides
        MVF_TRY(c.flush_bias_sums());
        {
          ProfScope ps(3, st);
          // dEnt and, from the same pass, delta = <dEnt, ent - bv> (= <G, px>) for the streaming kernel
          MVF_TRY(ent_finish_bwd(c.W.f("dh0"), m.ld0, c.W.f("dent"), m.R, d.SPC, m.W0, c.p, DropSeed(d.seed, d.seed_dev), st,
                                 c.S.f("ent32"), c.P[m.ibv], c.W.f("pdelta")));
          // dWv = dEnt^T px, dbv = colsum(dEnt) (forked); d(bk) is analytically zero (a per-entity constant under softmax)
          MVF_TRY(c.linear_dw(m.R, d.SPC, d.C_in, c.W.p("dent"), d.SPC, c.S.p("px"), d.C_in, gwv, ldg, gbkv + o_spc));
          // G = dEnt Wv
          MVF_TRY(c.linear_dx(MVF_F32, m.R, d.SPC, d.C_in, c.W.p("dent"), d.SPC, c.P[m.iWv], d.C_in, c.W.p("G"), d.C_in));
          MVF_CHECK_CUDA(cudaMemsetAsync(c.W.p("dwq"), 0, (size_t)d.E * d.C_in * 4, st));
        }
        MVF_TRY(c.join());
        {
          ProfScope ps(1, st);
          MVF_TRY(pool_fold_bwd(m.kvt, (int)m.F, d.P, d.E, d.C_in, tokens, c.W.f("G"), c.S.f("px"), c.S.f("attn"),
                                c.W.f("dwq"), st, c.W.f("pdelta")));
        }
        {
          ProfScope ps(5, st);
          MVF_TRY(fold_finish(c.W.f("dwq"), c.P[m.iQs], c.P[m.iQb], c.P[m.iWk], d.E, d.SPC, d.C_in, gwk, ldg, c.G.f("g.Qs"),
                              c.G.f("g.Qb"), st));
        }
        continue;
      }
      {
        ProfScope ps(3, st);
        MVF_TRY(xattn_pool_bwd(m.kvt, (int)m.F, d.P, d.E, d.SPC, c.S.p("kv"), c.P[m.iQs], c.P[m.iQb], c.S.f("attn"),
                               c.W.p("dh0"), m.ld0, c.S.f("ent32"), d.one_hot == MVF_ONEHOT_POOL, c.p, DropSeed(d.seed, d.seed_dev), c.W.p("dkv"),
                               c.G.f("g.Qs"), c.G.f("g.Qb"), gbkv, gbkv + o_spc, st));
      }
      // dW_kv = dKV^T X : 2*SPC x C_in outputs, K = frames*tokens -> split-K across the machine
      {
        ProfScope ps(1, st);
        MVF_TRY(c.gemm_kv(MVF_F32, 0, 0, 2 * d.SPC, d.C_in, m.F * d.P, c.W.p("dkv"), 2 * d.SPC, tokens, d.C_in,
                          c.G.f("g.w.kv"), c.Lg.find("g.w.kv")->ld, nullptr, MVF_GEMM_ACCUM, 0));
      }
    }
  }
  return c.join();
}

// bn_bwd_stats with dropout: the dropout that follows activation i belongs to site SITE_FC0+i+1.
// (kept consistent between bn_apply in forward and bn_bwd_* in backward.)

static int proj_forward_impl(Ctx& c, float* const* bn_running, int64_t* const* bn_tracked, const float* emb, int project,
                             float* out, int ph0, int ph1) {
  const Model& m = c.m;
  const mvf_head_desc& d = m.d;
  const int A = m.act;
  cudaStream_t st = c.st;
  if (!project) {
    // MODEL.L2_NORMALIZE without the projection head (evaluate.py path, transformer.py:229-230)
    if (ph0 == 0) {
      MVF_TRY(l2norm_fwd(emb, c.S.f("ehat"), c.S.f("norm"), m.N, d.D, st, out));
    }
    return MVF_OK;
  }
  const int ibn = d.n_fc;  // index of the projection BatchNorm in the bn tables
  if (ph1 > 2) ph1 = 2;
  for (int ph = ph0; ph < ph1; ++ph) {
    if (ph == 0) {
      PackEntry e[4];
      int ne = 2;
      const Region* r1 = c.Ls.find("w.p1");
      const Region* r2 = c.Ls.find("w.p2");
      e[0] = PackEntry{c.P[m.iWp1], c.S.base + r1->off, d.PS, d.D, (int)r1->ld, r1->dtype == RT_BF16};
      e[1] = PackEntry{c.P[m.iWp2], c.S.base + r2->off, d.D, d.PS, (int)r2->ld, r2->dtype == RT_BF16};
      if (m.tc) {
        const Region* s1 = c.Ls.find("w.p1.s");
        const Region* s2 = c.Ls.find("w.p2.s");
        e[ne++] = PackEntry{c.P[m.iWp1], c.S.base + s1->off, d.PS, d.D, (int)s1->ld, 2};
        e[ne++] = PackEntry{c.P[m.iWp2], c.S.base + s2->off, d.D, d.PS, (int)s2->ld, 2};
      }
      MVF_TRY(pack_params(e, ne, st));
      MVF_TRY(cast_f32(A, emb, c.S.p("emb"), m.N * d.D, st));
      MVF_TRY(c.linear(MVF_F32, m.N, d.PS, d.D, c.S.p("emb"), d.D, c.S.p("w.p1"), d.D, c.P[m.ibp1], c.S.p("u1"), d.PS, 0, "w.p1"));
      if (d.training) {
        MVF_TRY(c.zero_span("p.sum", "p.bsum"));
        MVF_TRY(bn_stats(c.S.f("u1"), m.N, d.PS, c.S.dbl("p.sum"), st));
      }
    } else {
      float* rm = bn_running ? bn_running[2 * ibn] : nullptr;
      float* rv = bn_running ? bn_running[2 * ibn + 1] : nullptr;
      MVF_REQUIRE(d.training || (rm && rv), MVF_ERR_BAD_ARG, "eval mode needs BatchNorm running statistics");
      const int fused = bn_finalize_apply(c.S.dbl("p.sum"), d.PS, bn_n_global(m, m.N), d.bn_eps, d.training, d.bn_momentum, rm, rv,
                                          bn_tracked ? bn_tracked[ibn] : nullptr, c.S.f("p.mi"), A, c.S.f("u1"), m.N, c.P[m.iGp],
                                          c.P[m.iBp], 1, c.S.p("a3"), d.PS, 0.f, 0, 0, st);
      if (fused == MVF_ERR_UNSUPPORTED) {
        MVF_TRY(bn_finalize(c.S.dbl("p.sum"), d.PS, bn_n_global(m, m.N), d.bn_eps, d.training, d.bn_momentum, rm, rv,
                            bn_tracked ? bn_tracked[ibn] : nullptr, c.S.f("p.mi"), st));
        MVF_TRY(bn_apply(A, c.S.f("u1"), m.N, d.PS, c.S.f("p.mi"), c.P[m.iGp], c.P[m.iBp], 1, c.S.p("a3"), d.PS, 0.f, 0, 0, st));
      } else {
        MVF_TRY(fused);
      }
      MVF_TRY(c.linear(MVF_F32, m.N, d.D, d.PS, c.S.p("a3"), d.PS, c.S.p("w.p2"), d.PS, c.P[m.ibp2], c.S.p("u"), d.D, 0, "w.p2"));
      if (project == 2) {  // MLPHead.forward on its own (resnet_c2d.py:122-126): no normalisation
        MVF_CHECK_CUDA(cudaMemcpyAsync(out, c.S.p("u"), (size_t)m.N * d.D * 4, cudaMemcpyDeviceToDevice, st));
      } else {
        MVF_TRY(l2norm_fwd(c.S.f("u"), c.S.f("ehat"), c.S.f("norm"), m.N, d.D, st, out));
      }
    }
  }
  return MVF_OK;
}

static int proj_backward_impl(Ctx& c, const float* d_out, int project, float* d_emb, const float* y_out, int ph0, int ph1) {
  const Model& m = c.m;
  const mvf_head_desc& d = m.d;
  const int A = m.act;
  cudaStream_t st = c.st;
  if (!project) {
    // y_out: the normalised output of forward (caller keeps it)
    if (ph0 == 0) MVF_TRY(l2norm_bwd(d_out, y_out, c.S.f("norm"), d_emb, MVF_F32, m.N, d.D, st));
    return MVF_OK;
  }
  MVF_REQUIRE(d.training, MVF_ERR_BAD_ARG, "backward needs training = 1");
  if (ph1 > 2) ph1 = 2;
  for (int ph = ph0; ph < ph1; ++ph) {
    if (ph == 0) {
      if (project == 2) MVF_TRY(cast_f32(A, d_out, c.W.p("du"), m.N * d.D, st));
      else MVF_TRY(l2norm_bwd(d_out, c.S.f("ehat"), c.S.f("norm"), c.W.p("du"), A, m.N, d.D, st));
      MVF_TRY(c.linear_dw(m.N, d.D, d.PS, c.W.p("du"), d.D, c.S.p("a3"), d.PS, c.G.f("g.w.p2"), d.PS, c.G.f("g.b.p2")));
      MVF_TRY(c.linear_dx(MVF_F32, m.N, d.D, d.PS, c.W.p("du"), d.D, c.S.p("w.p2"), d.PS, c.W.p("da3"), d.PS));
      MVF_CHECK_CUDA(cudaMemsetAsync(c.S.p("p.bsum"), 0, (size_t)2 * d.PS * 8, st));
      MVF_TRY(bn_bwd_stats(c.W.f("da3"), d.PS, c.S.f("u1"), m.N, d.PS, c.S.f("p.mi"), c.P[m.iGp], c.P[m.iBp], 1, 0.f, 0, 0,
                           c.S.dbl("p.bsum"), c.G.f("g.p.gamma"), c.G.f("g.p.beta"), st));
    } else {
      MVF_TRY(bn_bwd_apply(A, c.W.f("da3"), d.PS, c.S.f("u1"), m.N, d.PS, c.S.f("p.mi"), c.P[m.iGp], c.P[m.iBp], 1, 0.f, 0,
                           0, c.S.dbl("p.bsum"), bn_n_global(m, m.N), c.W.p("du1"), d.PS, st));
      MVF_TRY(c.linear_dw(m.N, d.PS, d.D, c.W.p("du1"), d.PS, c.S.p("emb"), d.D, c.G.f("g.w.p1"), d.D, c.G.f("g.b.p1")));
      MVF_TRY(c.linear_dx(MVF_F32, m.N, d.PS, d.D, c.W.p("du1"), d.PS, c.S.p("w.p1"), d.D, d_emb, d.D));
    }
  }
  return c.join();
}

static int unpack_impl(const Model& m, const Layout& Lg, const float* gpack, float* const* grads, float scale,
                       cudaStream_t st) {
  const mvf_head_desc& d = m.d;
  std::vector<UnpackEntry> e;
  auto add = [&](const std::string& reg, int pidx, int64_t row0 = 0, int64_t col0 = 0, int perm_n = 0) {
    if (pidx < 0 || grads[pidx] == nullptr) return;
    const Region* r = Lg.find(reg);
    const ParamInfo& pi = m.params[pidx];
    const float* src = (const float*)((const char*)gpack + r->off) + row0 * r->ld + col0;
    e.push_back(UnpackEntry{src, grads[pidx], (int)pi.rows, (int)pi.cols, (int)r->ld, perm_n});
  };
  if (m.fwb) {
    add("g.w.lc", m.iWlc, 0, 0, d.E);     // parameter row c*E + e <- entity-major row e*SPC + c
    add("g.b.lc", m.iblc, 0, 0, d.E);
  } else {
    add("g.Qs", m.iQs);
    add("g.Qb", m.iQb);
    add("g.w.kv", m.iWk, 0);
    add("g.w.kv", m.iWv, d.SPC);
    add("g.b.kv", m.ibk, 0, 0);
    add("g.b.kv", m.ibv, 0, d.SPC);
  }
  for (int i = 0; i < d.n_fc; ++i) {
    add("g." + fname(i, "w"), m.iFcW[i]);
    add("g." + fname(i, "b"), m.iFcB[i]);
    add("g." + fname(i, "gamma"), m.iFcG[i]);
    add("g." + fname(i, "beta"), m.iFcBeta[i]);
  }
  add("g.w.e", m.iWe);
  add("g.b.e", m.ibe);
  for (int l = 0; l < d.L; ++l) {
    const int b = m.iLayer[l];
    const std::string g = "g.l" + std::to_string(l) + ".";
    add(g + "ln0w", b + L_LN0W);
    add(g + "ln0b", b + L_LN0B);
    add(g + "ln1w", b + L_LN1W);
    add(g + "ln1b", b + L_LN1B);
    add(g + "w.qkv", b + L_WQ, 0);
    add(g + "w.qkv", b + L_WK, d.H);
    add(g + "w.qkv", b + L_WV, 2 * d.H);
    add(g + "b.qkv", b + L_BQ, 0, 0);
    add(g + "b.qkv", b + L_BK, 0, d.H);
    add(g + "b.qkv", b + L_BV, 0, 2 * d.H);
    add(g + "w.o", b + L_WO);
    add(g + "b.o", b + L_BO);
    add(g + "w.1", b + L_W1);
    add(g + "b.1", b + L_B1);
    add(g + "w.2", b + L_W2);
    add(g + "b.2", b + L_B2);
  }
  add("g.w.emb", m.iWemb);
  add("g.b.emb", m.ibemb);
  if (d.final_mode == MVF_FINAL_LIN) {
    add("g.w.lin", m.iWlin);
    add("g.b.lin", m.iblin);
  }
  add("g.w.p1", m.iWp1);
  add("g.b.p1", m.ibp1);
  add("g.p.gamma", m.iGp);
  add("g.p.beta", m.iBp);
  add("g.w.p2", m.iWp2);
  add("g.b.p2", m.ibp2);
  return unpack_grads(e.data(), (int)e.size(), scale, st);
}

}  // namespace mvf

// =================================================== C ABI ===================================================
using namespace mvf;

extern "C" {

int mvf_version(void) { return MVF_ABI_VERSION; }
const char* mvf_last_error(void) { return get_error(); }
int mvf_has_tcgen05(void) { return tc_available() ? 1 : 0; }
uint64_t mvf_launch_count(void) { return (uint64_t)__atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
int mvf_profile_enable(int on) {
  g_prof_on = on < 0 ? 0 : (on > 2 ? 2 : on);
  for (int t = 0; t < PROF_TAGS; ++t) g_prof_n[t] = 0;
  return MVF_OK;
}
int mvf_profile_read(int tag, float* ms, int cap, int* n) {
  MVF_REQUIRE(tag >= 0 && tag < PROF_TAGS && ms != nullptr && n != nullptr, MVF_ERR_BAD_ARG, "profile_read: bad argument");
  int cnt = g_prof_n[tag] < cap ? g_prof_n[tag] : cap;
  for (int i = 0; i < cnt; ++i) {
    MVF_CHECK_CUDA(cudaEventSynchronize(g_prof_ev[tag][i][1]));
    MVF_CHECK_CUDA(cudaEventElapsedTime(&ms[i], g_prof_ev[tag][i][0], g_prof_ev[tag][i][1]));
  }
  *n = cnt;
  return MVF_OK;
}

int mvf_num_params(const mvf_head_desc* d) {
  Model m;
  if (build_model(d, m) != MVF_OK) return -1;
  return (int)m.params.size();
}
int mvf_param_info(const mvf_head_desc* d, int idx, char* name, size_t name_cap, int64_t* rows, int64_t* cols) {
  Model m;
  MVF_TRY(build_model(d, m));
  MVF_REQUIRE(idx >= 0 && idx < (int)m.params.size(), MVF_ERR_BAD_ARG, "parameter index %d out of range", idx);
  if (name && name_cap) snprintf(name, name_cap, "%s", m.params[idx].name.c_str());
  if (rows) *rows = m.params[idx].rows;
  if (cols) *cols = m.params[idx].cols;
  return MVF_OK;
}
int mvf_num_bn(const mvf_head_desc* d) {
  Model m;
  if (build_model(d, m) != MVF_OK) return -1;
  return m.d.n_fc + 1;
}
int mvf_bn_info(const mvf_head_desc* d, int idx, char* name, size_t name_cap, int64_t* channels) {
  Model m;
  MVF_TRY(build_model(d, m));
  MVF_REQUIRE(idx >= 0 && idx <= m.d.n_fc, MVF_ERR_BAD_ARG, "BatchNorm index %d out of range", idx);
  if (idx < m.d.n_fc) {
    if (name && name_cap) snprintf(name, name_cap, "embed.fc_layers.%d", 4 * idx + 2);
    if (channels) *channels = m.d.fc[idx];
  } else {
    if (name && name_cap) snprintf(name, name_cap, "ssl_projection.net.1");
    if (channels) *channels = m.d.PS;
  }
  return MVF_OK;
}

static size_t layout_total(const mvf_head_desc* d, int which) {
  Model m;
  if (build_model(d, m) != MVF_OK) return 0;
  Layout L;
  switch (which) {
    case 0: head_save_layout(m, L); break;
    case 1: head_ws_layout(m, L); break;
    case 2: gpack_layout(m, L); break;
    case 3: proj_save_layout(m, L); break;
    default: proj_ws_layout(m, L); break;
  }
  return L.total;
}
size_t mvf_save_bytes(const mvf_head_desc* d) { return layout_total(d, 0); }
size_t mvf_ws_bytes(const mvf_head_desc* d) { return layout_total(d, 1); }
size_t mvf_gpack_elems(const mvf_head_desc* d) { return layout_total(d, 2) / 4; }
size_t mvf_gpack_pool_elems(const mvf_head_desc* d) {
  // the regions written by the pooling phase of backward (g.Qs, g.Qb, g.w.kv, g.b.kv) lead the buffer
  Model m;
  if (build_model(d, m) != MVF_OK) return 0;
  Layout L;
  gpack_layout(m, L);
  const auto* r = L.find(d->n_fc > 0 ? ("g." + fname(0, "w")).c_str() : "g.w.e");
  return r ? r->off / 4 : 0;
}
int mvf_pool_bwd_reserve_sms(int32_t n) { return pool_fold_reserve_sms(n); }
size_t mvf_proj_save_bytes(const mvf_head_desc* d) { return layout_total(d, 3); }
size_t mvf_proj_ws_bytes(const mvf_head_desc* d) { return layout_total(d, 4); }

static int lookup(const Layout& L, const char* name, size_t* offset, int64_t* rows, int64_t* cols, int64_t* ld,
                  int32_t* dtype) {
  const Region* r = L.find(name);
  MVF_REQUIRE(r != nullptr, MVF_ERR_BAD_ARG, "no region named '%s'", name);
  if (offset) *offset = r->off;
  if (rows) *rows = r->rows;
  if (cols) *cols = r->cols;
  if (ld) *ld = r->ld;
  if (dtype) *dtype = r->dtype;
  return MVF_OK;
}
int mvf_save_lookup(const mvf_head_desc* d, const char* name, size_t* offset, int64_t* rows, int64_t* cols, int64_t* ld,
                    int32_t* dtype) {
  Model m;
  MVF_TRY(build_model(d, m));
  MVF_REQUIRE(name != nullptr, MVF_ERR_BAD_ARG, "null name");
  Layout L;
  if (strncmp(name, "proj:", 5) == 0) {
    proj_save_layout(m, L);
    return lookup(L, name + 5, offset, rows, cols, ld, dtype);
  }
  if (strncmp(name, "g.", 2) == 0) {
    gpack_layout(m, L);
    return lookup(L, name, offset, rows, cols, ld, dtype);
  }
  head_save_layout(m, L);
  return lookup(L, name, offset, rows, cols, ld, dtype);
}
int mvf_bn_stat_lookup(const mvf_head_desc* d, int bn_idx, int backward, size_t* offset, int64_t* n_doubles) {
  Model m;
  MVF_TRY(build_model(d, m));
  MVF_REQUIRE(bn_idx >= 0 && bn_idx <= m.d.n_fc, MVF_ERR_BAD_ARG, "BatchNorm index %d out of range", bn_idx);
  Layout L;
  std::string nm;
  if (bn_idx < m.d.n_fc) {
    head_save_layout(m, L);
    nm = fname(bn_idx, backward ? "bsum" : "sum");
  } else {
    proj_save_layout(m, L);
    nm = backward ? "p.bsum" : "p.sum";
  }
  const Region* r = L.find(nm);
  if (offset) *offset = r->off;
  if (n_doubles) *n_doubles = r->cols;
  return MVF_OK;
}

int mvf_head_forward(const mvf_head_desc* d, const float* const* params, float* const* bn_running,
                     int64_t* const* bn_tracked, const void* tokens, const float* mask, void* save, size_t save_bytes,
                     void* ws, size_t ws_bytes, float* out_emb, float* attn_out, int phase_begin, int phase_end,
                     mvf_stream_t stream) {
  Ctx c;
  MVF_TRY(make_ctx(d, c, false, save, save_bytes, ws, ws_bytes, nullptr, params, (cudaStream_t)stream));
  MVF_REQUIRE((tokens != nullptr || d->pool_kind == MVF_POOLKIND_FWB) && out_emb != nullptr, MVF_ERR_BAD_ARG, "null tokens / output");
  MVF_REQUIRE(!d->has_mask || mask != nullptr, MVF_ERR_BAD_ARG, "has_mask set but mask is null");
  MVF_TRY(side_acquire(c.st, &c.side, &c.side_dev));
  int rc = head_forward_impl(c, bn_running, bn_tracked, tokens, mask, out_emb, attn_out, phase_begin, phase_end);
  int rj = c.join();   // never leave forked work un-joined
  return rc != MVF_OK ? rc : rj;
}

int mvf_head_backward(const mvf_head_desc* d, const float* const* params, const void* tokens, const float* mask,
                      const float* d_emb, void* save, size_t save_bytes, void* ws, size_t ws_bytes, float* gpack,
                      int phase_begin, int phase_end, mvf_stream_t stream) {
  Ctx c;
  MVF_TRY(make_ctx(d, c, false, save, save_bytes, ws, ws_bytes, gpack, params, (cudaStream_t)stream));
  MVF_REQUIRE((tokens != nullptr || d->pool_kind == MVF_POOLKIND_FWB) && d_emb != nullptr && gpack != nullptr, MVF_ERR_BAD_ARG,
              "null tokens / d_emb / gpack");
  MVF_TRY(side_acquire(c.st, &c.side, &c.side_dev));
  int rc = head_backward_impl(c, tokens, mask, d_emb, phase_begin, phase_end);
  if (rc != MVF_OK) c.join();  // never leave forked work un-joined, even on an error path
  return rc;
}

int mvf_proj_forward(const mvf_head_desc* d, const float* const* params, float* const* bn_running,
                     int64_t* const* bn_tracked, const float* emb, int project, void* save, size_t save_bytes, void* ws,
                     size_t ws_bytes, float* out, int phase_begin, int phase_end, mvf_stream_t stream) {
  Ctx c;
  MVF_TRY(make_ctx(d, c, true, save, save_bytes, ws, ws_bytes, nullptr, params, (cudaStream_t)stream));
  MVF_REQUIRE(emb != nullptr && out != nullptr, MVF_ERR_BAD_ARG, "null emb / out");
  return proj_forward_impl(c, bn_running, bn_tracked, emb, project, out, phase_begin, phase_end);
}

int mvf_proj_backward(const mvf_head_desc* d, const float* const* params, const float* d_out, int project, void* save,
                      size_t save_bytes, void* ws, size_t ws_bytes, float* gpack, float* d_emb, int phase_begin,
                      int phase_end, mvf_stream_t stream) {
  Ctx c;
  MVF_TRY(make_ctx(d, c, true, save, save_bytes, ws, ws_bytes, gpack, params, (cudaStream_t)stream));
  MVF_REQUIRE(d_out != nullptr && d_emb != nullptr, MVF_ERR_BAD_ARG, "null d_out / d_emb");
  MVF_REQUIRE(!project || gpack != nullptr, MVF_ERR_BAD_ARG, "null gpack");
  // the forward stores the normalised rows in `ehat` for both modes
  MVF_TRY(side_acquire(c.st, &c.side, &c.side_dev));
  int rc = proj_backward_impl(c, d_out, project, d_emb, c.S.f("ehat"), phase_begin, phase_end);
  if (rc != MVF_OK) c.join();
  return rc;
}

int mvf_unpack_grads(const mvf_head_desc* d, const float* gpack, float* const* grads, float scale, mvf_stream_t stream) {
  Model m;
  MVF_TRY(build_model(d, m));
  MVF_REQUIRE(gpack != nullptr && grads != nullptr, MVF_ERR_BAD_ARG, "null gpack / grads");
  Layout Lg;
  gpack_layout(m, Lg);
  return unpack_impl(m, Lg, gpack, grads, scale, (cudaStream_t)stream);
}

size_t mvf_scl_ws_bytes(int32_t Bv, int32_t T, int32_t D) { return scl_ws_bytes(Bv, T, D); }
int mvf_scl_fwd_bwd(const float* embs, const int64_t* seq_lens, const int64_t* steps, const float* masks, int32_t Bv,
                    int32_t T, int32_t D, float temperature, float label_variance, int32_t negative_type, int32_t quirk,
                    float* loss_out, float* d_embs, void* ws, size_t ws_bytes, mvf_stream_t stream) {
  ProfScope ps(8, (cudaStream_t)stream);
  return scl_fwd_bwd(embs, seq_lens, steps, masks, Bv, T, D, temperature, label_variance, negative_type, quirk, loss_out,
                     d_embs, ws, ws_bytes, (cudaStream_t)stream);
}

int mvf_gemm(int backend, int dtype_ab, int dtype_c, int a_kmajor, int b_kmajor, int64_t M, int64_t N, int64_t K,
             const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, const float* bias,
             const void* relu_src, int64_t ld_relu, int flags, int split_k, mvf_stream_t stream) {
  MVF_REQUIRE(A && B && C, MVF_ERR_BAD_ARG, "gemm: null operand");
  return gemm_dispatch(backend, dtype_ab, dtype_c, a_kmajor, b_kmajor, M, N, K, A, lda, B, ldb, C, ldc, bias, relu_src,
                       ld_relu, flags, split_k, (cudaStream_t)stream);
}

int mvf_xattn_pool_fwd(int dtype, int32_t F, int32_t P, int32_t E, int32_t SPC, const void* kv, const float* q_s,
                       const float* q_b, float* attn, void* ent, int64_t ld_ent, float* ent_f32, int one_hot, float drop_p,
                       uint64_t seed, mvf_stream_t stream) {
  MVF_REQUIRE(kv && q_s && q_b && ent, MVF_ERR_BAD_ARG, "xattn fwd: null pointer");
  return xattn_pool_fwd(dtype, F, P, E, SPC, kv, q_s, q_b, attn, ent, ld_ent, ent_f32, one_hot, drop_p, seed,
                        (cudaStream_t)stream);
}
int mvf_xattn_pool_bwd(int dtype, int32_t F, int32_t P, int32_t E, int32_t SPC, const void* kv, const float* q_s,
                       const float* q_b, const float* attn, const void* d_ent, int64_t ld_ent, const float* ent_f32,
                       int one_hot, float drop_p, uint64_t seed, void* d_kv, float* d_q_s, float* d_q_b, float* d_bk,
                       float* d_bv, mvf_stream_t stream) {
  MVF_REQUIRE(kv && q_s && q_b && attn && d_ent && d_kv && d_q_s && d_q_b && d_bk && d_bv, MVF_ERR_BAD_ARG,
              "xattn bwd: null pointer");
  return xattn_pool_bwd(dtype, F, P, E, SPC, kv, q_s, q_b, attn, d_ent, ld_ent, ent_f32, one_hot, drop_p, seed, d_kv, d_q_s,
                        d_q_b, d_bk, d_bv, (cudaStream_t)stream);
}

int mvf_pool_fold_prep(const float* q_s, const float* q_b, const float* w_k, int32_t E, int32_t SPC, int32_t C_in, float* wq,
                       mvf_stream_t stream) {
  MVF_REQUIRE(q_s && q_b && w_k && wq, MVF_ERR_BAD_ARG, "pool_fold_prep: null pointer");
  MVF_REQUIRE(E >= 1 && E <= MVF_MAX_ENTITIES && SPC > 0 && C_in > 0, MVF_ERR_BAD_ARG, "pool_fold_prep: bad shape");
  return fold_prep(q_s, q_b, w_k, E, SPC, C_in, wq, (cudaStream_t)stream);
}
int mvf_pool_fold_fwd(int dtype, int32_t F, int32_t P, int32_t E, int32_t C_in, const void* tokens, const float* wq,
                      float* attn, float* px, mvf_stream_t stream) {
  MVF_REQUIRE(tokens && wq && attn && px, MVF_ERR_BAD_ARG, "pool_fold_fwd: null pointer");
  return pool_fold_fwd(dtype, F, P, E, C_in, tokens, wq, attn, px, (cudaStream_t)stream);
}
int mvf_pool_fold_bwd(int dtype, int32_t F, int32_t P, int32_t E, int32_t C_in, const void* tokens, const float* g,
                      const float* px, const float* attn, float* d_wq, mvf_stream_t stream) {
  MVF_REQUIRE(tokens && g && px && attn && d_wq, MVF_ERR_BAD_ARG, "pool_fold_bwd: null pointer");
  return pool_fold_bwd(dtype, F, P, E, C_in, tokens, g, px, attn, d_wq, (cudaStream_t)stream);
}
int mvf_pool_fold_bwd_delta(int dtype, int32_t F, int32_t P, int32_t E, int32_t C_in, const void* tokens, const float* g,
                            const float* px, const float* attn, const float* delta, float* d_wq, mvf_stream_t stream) {
  MVF_REQUIRE(tokens && g && px && attn && d_wq, MVF_ERR_BAD_ARG, "pool_fold_bwd_delta: null pointer");
  return pool_fold_bwd(dtype, F, P, E, C_in, tokens, g, px, attn, d_wq, (cudaStream_t)stream, delta);
}
int mvf_pool_fold_finish(const float* d_wq, const float* q_s, const float* q_b, const float* w_k, int32_t E, int32_t SPC,
                         int32_t C_in, float* d_wk, int64_t ld_dwk, float* d_q_s, float* d_q_b, mvf_stream_t stream) {
  MVF_REQUIRE(d_wq && q_s && q_b && w_k && d_wk && d_q_s && d_q_b, MVF_ERR_BAD_ARG, "pool_fold_finish: null pointer");
  MVF_REQUIRE(E >= 1 && E <= MVF_MAX_ENTITIES && SPC > 0 && C_in > 0 && ld_dwk >= C_in, MVF_ERR_BAD_ARG,
              "pool_fold_finish: bad shape");
  return fold_finish(d_wq, q_s, q_b, w_k, E, SPC, C_in, d_wk, ld_dwk, d_q_s, d_q_b, (cudaStream_t)stream);
}

size_t mvf_opt_ws_bytes(int32_t n_tensors) { return opt_ws_bytes(n_tensors); }
int mvf_opt_adam_step(int32_t n_tensors, float* const* params, const float* const* grads, float* const* m, float* const* v,
                      const int64_t* numel, const float* lr_dev, int64_t* step_dev, double beta1, double beta2, float eps,
                      float weight_decay, int32_t adamw, float max_norm, float inv_scale, float* norm_out, void* ws,
                      size_t ws_bytes, mvf_stream_t stream) {
  return opt_adam_step(n_tensors, params, grads, m, v, numel, lr_dev, step_dev, beta1, beta2, eps, weight_decay, adamw, max_norm,
                       inv_scale, norm_out, ws, ws_bytes, (cudaStream_t)stream);
}
size_t mvf_peer_buffer_bytes(void) { return peer_buffer_bytes(); }
int mvf_peer_sum_f64(double* local, int64_t n, void* const* bufs_dev, int32_t rank, int32_t world, uint32_t* counter,
                     mvf_stream_t stream) {
  return peer_sum_f64(local, n, bufs_dev, rank, world, counter, (cudaStream_t)stream);
}
size_t mvf_peer_allreduce_flag_bytes(void) { return peer_allreduce_flag_bytes(); }
int mvf_peer_allreduce_f32(void* mc_base, void* const* bufs_dev, size_t data_off, size_t flag_off, int64_t n, int32_t rank,
                           int32_t world, uint32_t* counters, int32_t ctas, mvf_stream_t stream) {
  return peer_allreduce_f32(mc_base, bufs_dev, data_off, flag_off, n, rank, world, counters, ctas, (cudaStream_t)stream);
}
size_t mvf_attention_ws_bytes(int32_t B, int32_t S, int32_t heads, int32_t dk) { return attention_ws_bytes(B, S, heads, dk); }
int mvf_attention_fwd(int dtype, int32_t B, int32_t S, int32_t heads, int32_t dk, const void* qkv, const float* keymask,
                      void* ctx, float* lse, void* ws, size_t ws_bytes, mvf_stream_t stream) {
  MVF_REQUIRE(qkv && ctx && lse, MVF_ERR_BAD_ARG, "attention fwd: null pointer");
  return attention_fwd(dtype, B, S, heads, dk, qkv, keymask, ctx, lse, (cudaStream_t)stream, ws != nullptr, ws, ws_bytes);
}
int mvf_attention_bwd(int dtype, int32_t B, int32_t S, int32_t heads, int32_t dk, const void* qkv, const float* keymask,
                      const void* ctx, const float* lse, const void* d_ctx, void* d_qkv, float* ws_delta, void* ws,
                      size_t ws_bytes, mvf_stream_t stream) {
  MVF_REQUIRE(qkv && ctx && lse && d_ctx && d_qkv && ws_delta, MVF_ERR_BAD_ARG, "attention bwd: null pointer");
  return attention_bwd(dtype, B, S, heads, dk, qkv, keymask, ctx, lse, d_ctx, d_qkv, ws_delta, (cudaStream_t)stream,
                       ws != nullptr, ws, ws_bytes);
}

int mvf_dropout_mask(uint64_t seed, int32_t site, int64_t rows, int64_t cols, float p, float* out, mvf_stream_t stream) {
  MVF_REQUIRE(out != nullptr, MVF_ERR_BAD_ARG, "dropout mask: null output");
  return dropout_mask_export(seed, site, rows, cols, p, out, (cudaStream_t)stream);
}

}  // extern "C"
