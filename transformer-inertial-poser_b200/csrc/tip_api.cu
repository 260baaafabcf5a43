// C ABI of the TIP hot path (include/tip_b200.h): handle, weight packing, forward orchestration,
// host-buffer entry point and the streaming window path.
#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

#include "tip_common.cuh"
#include "tip_simt.cuh"
#include "tip_umma.cuh"
#include "tip_rnn_umma.cuh"
#include "tip_attn_mma.cuh"
#include "tip_qkv_attn.cuh"
#include "tip_ffn_fused.cuh"
#include "tip_umma_atm.cuh"
#include "tip_umma_ln2.cuh"

using namespace tip;

namespace {
unsigned long long* g_tbuf = nullptr;
unsigned long long* g_rnn_tbuf = nullptr;
std::string g_create_error;
constexpr int CHUNK_WINDOWS = 1024;   // windows per pass through the workspace
}  // namespace

// CUDA graphs of whole forwards, keyed by everything a captured launch sequence depends on.  A (shape,
// pointer) combination is captured the second time it is seen (the first call runs eagerly and does all the
// lazy initialisation: workspace, tensor maps, occupancy queries), then replayed: one graph launch instead
// of ~25 kernel launches.
struct FwdGraphKey {
    const void *x_imu, *x_s, *y, *keep;
    int B, L, engine;
    float past_scale;
    float p_in, p_past, p_enc;          // dropout rates are baked into the captured launches (the SEED is not: device memory)
    bool operator==(const FwdGraphKey& o) const {
        return x_imu == o.x_imu && x_s == o.x_s && y == o.y && keep == o.keep && B == o.B && L == o.L &&
               engine == o.engine && past_scale == o.past_scale && p_in == o.p_in && p_past == o.p_past && p_enc == o.p_enc;
    }
};
struct FwdGraph {
    FwdGraphKey key{};
    cudaGraphExec_t exec = nullptr;     // null: seen once, not captured yet
    int launches = 0;
    uint64_t last_use = 0;
};
constexpr int FWD_GRAPH_SLOTS = 64;     // captured forwards kept per handle (LRU); a lane rotating 16 input sets in two dropout modes needs 32

// Packed weights: owned by the handle tip_create made, shared (read-only) by the lanes tip_create_lane derives from it.
struct SharedWeights {
    float* blob = nullptr;
    int refs = 1;
    bool packed = false;
    uint64_t pack_seq = 0;              // bumped by every tip_pack_weights
    cudaEvent_t ev_pack = nullptr;      // recorded on the packing stream at the end of the pack
};

struct DropP {                          // dropout rates of a call (all 0 = deterministic)
    float p_in = 0.f, p_past = 0.f, p_enc = 0.f;
    uint64_t seed = 0;
    bool any() const { return p_in > 0.f || p_past > 0.f || p_enc > 0.f; }
    bool same_p(const DropP& o) const { return p_in == o.p_in && p_past == o.p_past && p_enc == o.p_enc; }
};
static DropP drop_of(const tip_dropout* d) {
    DropP r;
    if (d) { r.p_in = d->in_dropout; r.p_past = d->past_state_dropout; r.p_enc = d->encoder_dropout; r.seed = d->seed; }
    return r;
}

constexpr int NARROW_CTAS = 40;                                          // width of a throughput-mode launch: about a quarter of the GPU (measured best at B = 256)
constexpr int SCHED_SLOTS = 128;                                         // one per GEMM launch of a forward part (2 parts x 64)
constexpr size_t SEED_SCHED_BYTES = sizeof(uint64_t) + SCHED_SLOTS * 2 * sizeof(int);

struct tip_model {
    SharedWeights* sw = nullptr;
    bool is_lane = false;
    bool laned = false;                 // this handle has lanes or is one: several forwards share the GPU (throughput mode)
    // tuning knobs (tip_set_tuning; initial values from the TIP_* environment)
    int tune_atm = -1;                  // GEMMs on the A-in-TMEM kernel: bit mask 1 in_linear, 2 qkv, 4 ff1, 8 rnn_ih; -1 = auto (all four when laned)
    int tune_atm_pair = 0;              // ... on CTA pairs (cta_group::2: each CTA stages half of every W k-block)
    int tune_atm_grid = 0;              // CTAs per A-in-TMEM launch (0 = one per two 128-row tiles)
    int tune_atm_min_tiles = 64;        // ... for forwards of at least this many row tiles
    int tune_dyn_sched = 0;             // dynamic tile scheduler of the plain GEMMs
    int tune_ln_pair = 0;               // LayerNorm GEMMs with K >= this on CTA pairs (0 = never)
    int tune_ln_share = -1;             // LayerNorm GEMMs on the two-row-tile kernel (W k-blocks shared): 1 on, 0 off, -1 auto (laned)
    int tune_ln_grid = -1;              // CTAs per fused-LayerNorm GEMM launch (0 = one per 128-row tile; -1 = auto: one per two row tiles when laned)
    int tune_attn_grid = 0;             // attention: 0 = one CTA per (window, 8 heads), one wave; N > 0 = N persistent double-buffered CTAs; -1 = two per SM
    int tune_rnn_clusters = 0;          // 8-CTA clusters per tensor-core recurrence launch (0 = as many as the batch needs / the GPU co-schedules)
    uint64_t pack_ordered_seq = 0;      // last pack this handle's streams are known to be ordered after (pack complete)
    uint64_t* d_seed = nullptr;         // base seed of the current stochastic call (device memory; graphs read it); followed by
                                        // the {next tile, CTAs done} counter pairs of the GEMMs' dynamic tile scheduler (SCHED_SLOTS)
    tip_dims cdims{};
    Dims d{};
    PackOff off{};
    int device = 0;
    float* blob = nullptr;    // = sw->blob
    int engine = 0;           // 0 auto, 1 FFMA, 2 tcgen05
    int use_graphs = 1;
    int launches = 0;
    int64_t last_rows = 0;
    int rnn_clusters = -1;          // co-schedulable 8-CTA clusters (queried on first use)
    int rnn_umma_clusters = -1;
    bool attn_attr_set = false, qa_attr_set = false, ffn_attr_set = false;
    int rnn_stream_fallback = 0;    // TIP_RNN_STREAM=1: L2-streaming kernel (debug / comparison)
    std::string err;

    // workspace (capacity in rows)
    int cap_rows = 0;
    float* ws = nullptr;
    float *xin = nullptr, *xa = nullptr, *xb = nullptr, *qkv = nullptr, *att = nullptr,
          *hid = nullptr, *gi = nullptr, *hs = nullptr, *pre = nullptr;   // pre: fp32 [rows][256] scratch of the un-fused LayerNorm path
    size_t plane_xin = 0, plane_e = 0, plane_f = 0, plane_r = 0;   // elements per plane (= hi->lo stride in halves)
    UmmaMaps maps;            // TMA descriptors of the workspace + weights (tcgen05 engine)
    bool maps_ready = false;

    // host-entry staging
    size_t host_cap = 0;      // windows*L capacity in rows
    float *h_in = nullptr, *h_out = nullptr, *d_ximu = nullptr, *d_xs = nullptr, *d_y = nullptr;

    // streaming state
    int n_streams = 0, stream_len = 0;
    float *win_imu = nullptr, *win_s = nullptr, *st_rows = nullptr, *st_ximu = nullptr,
          *st_xs = nullptr, *st_y = nullptr, *st_ylast = nullptr, *h_rows = nullptr, *h_ylast = nullptr;
    float *raw_ring = nullptr, *st_raw = nullptr, *h_raw = nullptr;    // N1: raw-IMU pre-processing state
    double* acc_ring = nullptr;
    int n_raw = 0, n_rows = 0;
    // N3: post-model step state (closed loop)
    double *pp_ring = nullptr, *pp_last = nullptr, *pp_out = nullptr, *h_pp_out = nullptr;   // the fed-back x_s row lives in st_rows
    int n_post = 0;
    bool fb_set = false;
    cudaGraphExec_t st_graph = nullptr;   // captured steady-state step (L == MAXL)
    int st_graph_launches = 0;
    DropP st_graph_p;                     // dropout rates baked into st_graph
    FwdGraph fwd_graphs[FWD_GRAPH_SLOTS];
    uint64_t fwd_tick = 0;
    // host entry, two-part pipeline: copy / compute streams, events, and one captured forward per batch part
    cudaStream_t s_in = nullptr, s_out = nullptr, s_part[2] = {nullptr, nullptr};
    cudaEvent_t ev_start = nullptr, ev_in[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    struct PartGraph { cudaGraphExec_t exec = nullptr; int B = 0, L = 0, w0 = 0, nw = 0, launches = 0, seen = 0; DropP p; } part_graph[2];
    // host entry, job pipeline (tip_forward_host_submit / _wait): per-slot device staging + completion events; one upload,
    // one forward and one download stream shared by all slots (the forwards share the workspace, so they serialise)
    struct HostSlot {
        size_t cap = 0;           // rows
        float *d_ximu = nullptr, *d_xs = nullptr, *d_y = nullptr;
        cudaEvent_t ev_in = nullptr, ev_fwd = nullptr, ev_out = nullptr;
        bool busy = false;
        int launches = 0;
    } hslot[TIP_HOST_SLOTS];
    cudaStream_t hs_in = nullptr, hs_fwd = nullptr, hs_out = nullptr;
    cudaEvent_t ev_user = nullptr;      // last forward a caller queued on its own stream (the pipeline's forwards wait for it)
    bool ev_user_set = false;

    // per-stage profiling (tip_set_profile)
    int profile = 0;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<std::string> st_names;
    std::vector<int> st_layers;

    void set_error(const std::string& s) { err = s; }
};

// record "a stage named `name` starts here" (name == nullptr: end of the forward)
static void mark(tip_model* m, cudaStream_t st, const char* name, int layer = -1) {
    if (!m->profile) return;
    const size_t ev_i = m->st_names.size();   // event i opens stage i and closes stage i-1
    while (m->ev_pool.size() <= ev_i) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        m->ev_pool.push_back(e);
    }
    cudaEventRecord(m->ev_pool[ev_i], st);
    if (name) { m->st_names.push_back(name); m->st_layers.push_back(layer); }
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static int quiesce_host_jobs(tip_model* m);      // wait until no pipelined host job still uses the workspace / weights

// every captured graph bakes in workspace addresses, tensor maps and the engine choice
static void init_tuning(tip_model* m);
static void drop_graphs(tip_model* m) {
    if (m->st_graph) { cudaGraphExecDestroy(m->st_graph); m->st_graph = nullptr; }
    for (FwdGraph& g : m->fwd_graphs) {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        g = FwdGraph{};
    }
    for (auto& pg : m->part_graph) {
        if (pg.exec) cudaGraphExecDestroy(pg.exec);
        pg = tip_model::PartGraph{};
    }
}

static void compute_offsets(tip_model* m) {
    const Dims& d = m->d;
    PackOff& o = m->off;
    size_t p = 0;
    auto take = [&](size_t n) { size_t r = p; p = align_up(p + n, 64); return r; };   // 256-byte aligned
    o.win = take((size_t)E * d.kin_pad);
    o.bin = take(E);
    o.win_hi = take((size_t)E * d.kin_pad);
    o.win_lo = take((size_t)E * d.kin_pad);
    for (int l = 0; l < d.layers; ++l) {
        LayerOff& L = o.layer[l];
        L.wqkv = take((size_t)3 * E * E); L.bqkv = take(3 * E);
        L.wo = take((size_t)E * E);       L.bo = take(E);
        L.w1 = take((size_t)F * E);       L.b1 = take(F);
        L.w2 = take((size_t)E * F);       L.b2 = take(E);
        L.g1 = take(E); L.be1 = take(E); L.g2 = take(E); L.be2 = take(E);
        L.wqkv_hi = take((size_t)3 * E * E); L.wqkv_lo = take((size_t)3 * E * E);
        L.wo_hi = take((size_t)E * E);       L.wo_lo = take((size_t)E * E);
        L.w1_hi = take((size_t)F * E);       L.w1_lo = take((size_t)F * E);
        L.w2_hi = take((size_t)E * F);       L.w2_lo = take((size_t)E * F);
        L.wqkvr_hi = take((size_t)3 * E * E / 2); L.wqkvr_lo = take((size_t)3 * E * E / 2); L.bqkvr = take(3 * E);
    }
    if (d.with_rnn) {
        o.wih = take((size_t)R * E); o.brnn = take(R);
        o.whh = take((size_t)R * R); o.whh_t = take((size_t)R * R);
        o.wih_hi = take((size_t)R * E); o.wih_lo = take((size_t)R * E);
        o.whh_hi = take((size_t)R * R); o.whh_lo = take((size_t)R * R);
    }
    o.wl = take((size_t)HEAD_NPAD * d.khead); o.bl = take(HEAD_NPAD);
    o.wl_hi = take((size_t)HEAD_NPAD * d.khead); o.wl_lo = take((size_t)HEAD_NPAD * d.khead);
    o.scales = take(2 * SC_COUNT);            // [SC_COUNT] epilogue factors, then [SC_COUNT] weight scales
    o.total = p;
}

extern "C" int tip_abi_version(void) { return TIP_ABI_VERSION; }

extern "C" const char* tip_last_error(const tip_model* m) {
    return m ? m->err.c_str() : g_create_error.c_str();
}

extern "C" int tip_create(const tip_dims* dims, tip_model** out) {
    if (!dims || !out) { g_create_error = "null argument"; return TIP_ERR_INVALID_ARG; }
    *out = nullptr;
    if (dims->tf_in_dim != E || dims->n_heads != NH || dims->tf_hid_size != F ||
        dims->rnn_hid_size != R || dims->tf_layers < 1 || dims->tf_layers > MAX_LAYERS ||
        dims->input_size_imu != 72 || dims->size_s < 111 || dims->size_s > HEAD_NPAD) {
        g_create_error =
            "unsupported hyper-parameters: kernels are specialised for tf_in_dim=256, n_heads=16, "
            "tf_hid_size=1024, rnn_hid_size=512, input_size_imu=72, 111<=size_s<=144, tf_layers<=8";
        return TIP_ERR_INVALID_ARG;
    }
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { g_create_error = "no CUDA device"; return TIP_ERR_NO_DEVICE; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess || prop.major != 10) {
        g_create_error = "tip_b200 kernels are built for sm_100a only (device is not compute 10.x)";
        return TIP_ERR_NO_DEVICE;
    }
    tip_model* m = new tip_model();
    m->cdims = *dims;
    m->device = dev;
    m->d.n_imu = dims->input_size_imu + (dims->with_acc_sum ? 18 : 0);
    m->d.size_s = dims->size_s;
    m->d.d_in = m->d.n_imu + dims->size_s;
    m->d.kin_pad = (int)align_up(m->d.d_in, 64);
    m->d.layers = dims->tf_layers;
    m->d.with_rnn = dims->with_rnn ? 1 : 0;
    m->d.khead = dims->with_rnn ? R : E;
    m->rnn_stream_fallback = getenv("TIP_RNN_STREAM") ? atoi(getenv("TIP_RNN_STREAM")) : 0;
    init_tuning(m);
    compute_offsets(m);
    m->sw = new SharedWeights();
    cudaError_t e = cudaMalloc(&m->sw->blob, m->off.total * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&m->d_seed, SEED_SCHED_BYTES);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->sw->ev_pack, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaMalloc(weights): ") + cudaGetErrorString(e);
        if (m->sw->blob) cudaFree(m->sw->blob);
        if (m->d_seed) cudaFree(m->d_seed);
        delete m->sw;
        delete m;
        return TIP_ERR_OOM;
    }
    m->blob = m->sw->blob;
    cudaMemset(m->blob, 0, m->off.total * sizeof(float));
    cudaMemset(m->d_seed, 0, SEED_SCHED_BYTES);
    *out = m;
    return TIP_OK;
}

// A lane: a second handle on the same device that SHARES the owner's packed weights (read-only; no second copy, no
// second pack) but has its own workspace, tensor maps, captured graphs, job slots and seed -- forwards of different
// lanes may overlap on different streams.  The weights stay alive until the last handle sharing them is destroyed.
extern "C" int tip_create_lane(tip_model* owner, tip_model** out) {
    if (!owner || !out) { g_create_error = "null argument"; return TIP_ERR_INVALID_ARG; }
    *out = nullptr;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != owner->device) { g_create_error = "tip_create_lane: the owner lives on another device"; return TIP_ERR_INVALID_ARG; }
    tip_model* m = new tip_model();
    m->cdims = owner->cdims;
    m->d = owner->d;
    m->off = owner->off;
    m->device = owner->device;
    m->engine = owner->engine;
    m->rnn_stream_fallback = owner->rnn_stream_fallback;
    m->is_lane = true;
    init_tuning(m);
    m->laned = true;
    if (!owner->laned) { owner->laned = true; drop_graphs(owner); }      // (its captured forwards predate the throughput-mode kernel choice)
    if (cudaMalloc(&m->d_seed, SEED_SCHED_BYTES) != cudaSuccess) {
        g_create_error = "cudaMalloc(seed) failed";
        delete m;
        return TIP_ERR_OOM;
    }
    cudaMemset(m->d_seed, 0, SEED_SCHED_BYTES);
    m->sw = owner->sw;
    m->sw->refs++;
    m->blob = m->sw->blob;
    *out = m;
    return TIP_OK;
}

// Make `s` wait for the last weight pack (it may have run on another stream: another lane's, or the caller's stream at
// load_state_dict time).  Once the pack is known to be complete the check is one integer compare.
static int order_after_pack(tip_model* m, cudaStream_t s) {
    SharedWeights* w = m->sw;
    if (m->pack_ordered_seq == w->pack_seq) return TIP_OK;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return TIP_OK;   // (never with a pending pack)
    const uint64_t seq = w->pack_seq;
    TIP_CUDA_TRY(m, cudaStreamWaitEvent(s, w->ev_pack, 0));
    if (cudaEventQuery(w->ev_pack) == cudaSuccess) m->pack_ordered_seq = seq; else cudaGetLastError();
    return TIP_OK;
}

static void free_stream_state(tip_model* m) {
    drop_graphs(m);
    if (m->acc_ring) { cudaFree(m->acc_ring); m->acc_ring = nullptr; }
    if (m->h_raw) { cudaFreeHost(m->h_raw); m->h_raw = nullptr; }
    for (double** p : {&m->pp_ring, &m->pp_last, &m->pp_out})
        if (*p) { cudaFree(*p); *p = nullptr; }
    if (m->h_pp_out) { cudaFreeHost(m->h_pp_out); m->h_pp_out = nullptr; }
    m->n_raw = m->n_rows = m->n_post = 0;
    m->fb_set = false;
    for (float** p : {&m->win_imu, &m->win_s, &m->st_rows, &m->st_ximu, &m->st_xs, &m->st_y, &m->st_ylast, &m->raw_ring, &m->st_raw})
        if (*p) { cudaFree(*p); *p = nullptr; }
    for (float** p : {&m->h_rows, &m->h_ylast})
        if (*p) { cudaFreeHost(*p); *p = nullptr; }
    m->n_streams = 0;
    m->stream_len = 0;
}

extern "C" void tip_destroy(tip_model* m) {
    if (!m) return;
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);              // garbage collection must not change the caller's current device
    const int m_device_of_destroyed = m->device;
    cudaSetDevice(m->device);
    free_stream_state(m);
    for (cudaEvent_t e : m->ev_pool) cudaEventDestroy(e);
    if (m->s_in) {
        cudaStreamDestroy(m->s_in); cudaStreamDestroy(m->s_out); cudaStreamDestroy(m->s_part[0]); cudaStreamDestroy(m->s_part[1]);
        cudaEventDestroy(m->ev_start);
        for (int i = 0; i < 2; ++i) { cudaEventDestroy(m->ev_in[i]); cudaEventDestroy(m->ev_out[i]); }
    }
    for (auto& hsl : m->hslot) if (hsl.busy) cudaEventSynchronize(hsl.ev_out);
    if (m->hs_in) { cudaStreamDestroy(m->hs_in); cudaStreamDestroy(m->hs_fwd); cudaStreamDestroy(m->hs_out); cudaEventDestroy(m->ev_user); }
    for (auto& hsl : m->hslot) {
        for (float* p : {hsl.d_ximu, hsl.d_xs, hsl.d_y}) if (p) cudaFree(p);
        for (cudaEvent_t e : {hsl.ev_in, hsl.ev_fwd, hsl.ev_out}) if (e) cudaEventDestroy(e);
    }
    if (m->sw && --m->sw->refs == 0) {
        cudaDeviceSynchronize();           // (another lane's forward may have been the last reader)
        if (m->sw->blob) cudaFree(m->sw->blob);
        if (m->sw->ev_pack) cudaEventDestroy(m->sw->ev_pack);
        delete m->sw;
    }
    if (m->d_seed) cudaFree(m->d_seed);
    if (m->ws) cudaFree(m->ws);
    for (float* p : {m->d_ximu, m->d_xs, m->d_y}) if (p) cudaFree(p);
    for (float* p : {m->h_in, m->h_out}) if (p) cudaFreeHost(p);
    delete m;
    if (prev_dev >= 0 && prev_dev != m_device_of_destroyed) cudaSetDevice(prev_dev);
}

extern "C" int tip_num_weight_tensors(const tip_model* m) {
    return m ? 2 + 12 * m->d.layers + (m->d.with_rnn ? 4 : 0) + 2 : 0;
}

extern "C" int tip_pack_weights(tip_model* m, const float* const* t, const int64_t* numels, int n,
                                void* stream_) {
    if (!m || !t || !numels) return TIP_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream_;
    const Dims& d = m->d;
    if (n != tip_num_weight_tensors(m)) { m->set_error("wrong number of weight tensors"); return TIP_ERR_INVALID_ARG; }
    if (m->is_lane) { m->set_error("tip_pack_weights: a lane shares its owner's weights; pack the owner"); return TIP_ERR_INVALID_ARG; }
    { const int qrc = quiesce_host_jobs(m); if (qrc != TIP_OK) return qrc; }
    // expected element counts in state-dict order
    std::vector<int64_t> exp;
    exp.push_back((int64_t)E * d.d_in); exp.push_back(E);
    for (int l = 0; l < d.layers; ++l)
        for (int64_t v : {(int64_t)3 * E * E, (int64_t)3 * E, (int64_t)E * E, (int64_t)E, (int64_t)F * E,
                          (int64_t)F, (int64_t)E * F, (int64_t)E, (int64_t)E, (int64_t)E, (int64_t)E, (int64_t)E})
            exp.push_back(v);
    if (d.with_rnn) { exp.push_back((int64_t)R * E); exp.push_back((int64_t)R * R); exp.push_back(R); exp.push_back(R); }
    exp.push_back((int64_t)d.size_s * d.khead); exp.push_back(d.size_s);
    for (int i = 0; i < n; ++i)
        if (numels[i] != exp[i] || t[i] == nullptr) {
            m->set_error("weight tensor " + std::to_string(i) + " has " + std::to_string(numels[i]) +
                         " elements, expected " + std::to_string(exp[i]));
            return TIP_ERR_INVALID_ARG;
        }
    TIP_CUDA_TRY(m, cudaSetDevice(m->device));
    // forwards still in flight on OTHER streams (lanes sharing these weights, the job pipeline, the two-part host entry)
    // read the blob this call overwrites; a caller that only ever uses one stream keeps the pack fully asynchronous
    if (m->sw->refs > 1 || m->hs_fwd || m->s_in) TIP_CUDA_TRY(m, cudaDeviceSynchronize());
    float* B = m->blob;
    const PackOff& o = m->off;
    auto copy = [&](size_t dst, const float* src, size_t cnt) {
        return cudaMemcpyAsync(B + dst, src, cnt * sizeof(float), cudaMemcpyDeviceToDevice, st);
    };
    // FP16 hi/lo planes of s_w * W (s_w = power of two from max|W|) + the epilogue's 1/(s_w * s_a)
    auto split = [&](size_t src, size_t hi, size_t lo, size_t cnt, int sc_idx, float act_scale) {
        float* inv = B + o.scales + sc_idx;
        float* wsc = B + o.scales + SC_COUNT + sc_idx;
        pack_scale_kernel<<<1, 1024, 0, st>>>(B + src, (int64_t)cnt, act_scale, inv, wsc);
        pack_split_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(B + src, reinterpret_cast<__half*>(B + hi),
                                                                        reinterpret_cast<__half*>(B + lo), (int64_t)cnt, wsc);
    };
    int i = 0;
    pack_in_linear_kernel<<<(E * d.kin_pad + 255) / 256, 256, 0, st>>>(t[0], t[1], B + o.win, B + o.bin,
                                                                     d.d_in, d.kin_pad, d.n_imu);
    split(o.win, o.win_hi, o.win_lo, (size_t)E * d.kin_pad, SC_IN, 1.f);    // raw input planes carry scale 1
    i = 2;
    for (int l = 0; l < d.layers; ++l) {
        const LayerOff& L = o.layer[l];
        // 1/sqrt(head_dim) = 0.25 folded into the q rows (exact: power of two)
        pack_scale_rows_kernel<<<(3 * E * E + 255) / 256, 256, 0, st>>>(t[i], B + L.wqkv, (int64_t)3 * E * E,
                                                                       (int64_t)E * E, 0.25f);
        pack_scale_rows_kernel<<<(3 * E + 255) / 256, 256, 0, st>>>(t[i + 1], B + L.bqkv, 3 * E, E, 0.25f);
        TIP_CUDA_TRY(m, copy(L.wo, t[i + 2], (size_t)E * E));
        TIP_CUDA_TRY(m, copy(L.bo, t[i + 3], E));
        TIP_CUDA_TRY(m, copy(L.w1, t[i + 4], (size_t)F * E));
        TIP_CUDA_TRY(m, copy(L.b1, t[i + 5], F));
        TIP_CUDA_TRY(m, copy(L.w2, t[i + 6], (size_t)E * F));
        TIP_CUDA_TRY(m, copy(L.b2, t[i + 7], E));
        TIP_CUDA_TRY(m, copy(L.g1, t[i + 8], E));
        TIP_CUDA_TRY(m, copy(L.be1, t[i + 9], E));
        TIP_CUDA_TRY(m, copy(L.g2, t[i + 10], E));
        TIP_CUDA_TRY(m, copy(L.be2, t[i + 11], E));
        split(L.wqkv, L.wqkv_hi, L.wqkv_lo, (size_t)3 * E * E, SC_LAYER0 + 4 * l + 0, ACT_SCALE);
        split(L.wo, L.wo_hi, L.wo_lo, (size_t)E * E, SC_LAYER0 + 4 * l + 1, ACT_SCALE);
        split(L.w1, L.w1_hi, L.w1_lo, (size_t)F * E, SC_LAYER0 + 4 * l + 2, ACT_SCALE);
        split(L.w2, L.w2_hi, L.w2_lo, (size_t)E * F, SC_LAYER0 + 4 * l + 3, ACT_SCALE);
        // the same in_proj planes with their rows grouped per 4 heads (q | k | v of the group contiguous): fused QKV + attention
        pack_qkv_reorder_kernel<<<(3 * E * (E / 8) + 255) / 256, 256, 0, st>>>(
            reinterpret_cast<const __half*>(B + L.wqkv_hi), reinterpret_cast<const __half*>(B + L.wqkv_lo), B + L.bqkv,
            reinterpret_cast<__half*>(B + L.wqkvr_hi), reinterpret_cast<__half*>(B + L.wqkvr_lo), B + L.bqkvr);
        i += 12;
    }
    if (d.with_rnn) {
        TIP_CUDA_TRY(m, copy(o.wih, t[i], (size_t)R * E));
        TIP_CUDA_TRY(m, copy(o.whh, t[i + 1], (size_t)R * R));
        pack_transpose_kernel<<<dim3(R / 32, R / 32), dim3(32, 32), 0, st>>>(t[i + 1], B + o.whh_t, R);
        pack_add_kernel<<<(R + 255) / 256, 256, 0, st>>>(t[i + 2], t[i + 3], B + o.brnn, R);
        split(o.wih, o.wih_hi, o.wih_lo, (size_t)R * E, SC_IH, ACT_SCALE);
        split(o.whh, o.whh_hi, o.whh_lo, (size_t)R * R, SC_HH, ACT_SCALE);
        i += 4;
    }
    pack_pad_rows_kernel<<<(HEAD_NPAD * d.khead + 255) / 256, 256, 0, st>>>(t[i], B + o.wl, d.size_s, HEAD_NPAD, d.khead);
    pack_pad_rows_kernel<<<1, 256, 0, st>>>(t[i + 1], B + o.bl, d.size_s, HEAD_NPAD, 1);
    split(o.wl, o.wl_hi, o.wl_lo, (size_t)HEAD_NPAD * d.khead, SC_HEAD, ACT_SCALE);
    TIP_CUDA_TRY(m, cudaGetLastError());
    // every stream that reads the weights later (lanes, the job pipeline's forward stream, another caller stream)
    // waits for this event first (order_after_pack)
    TIP_CUDA_TRY(m, cudaEventRecord(m->sw->ev_pack, st));
    m->sw->pack_seq++;
    m->sw->packed = true;
    return TIP_OK;        // (addresses unchanged: tensor maps and captured graphs of every handle stay valid)
}
static void init_tuning(tip_model* m) {
    auto env = [](const char* k, int dflt) { const char* v = getenv(k); return v ? atoi(v) : dflt; };
    m->tune_atm = env("TIP_ATM", -1);
    m->tune_atm_grid = env("TIP_ATM_GRID", 0);
    m->tune_atm_pair = env("TIP_ATM_PAIR", 0);
    m->tune_atm_min_tiles = env("TIP_ATM_MIN_TILES", 64);
    m->tune_dyn_sched = env("TIP_DYN_SCHED", 0);
    m->tune_ln_pair = env("TIP_LN_PAIR", 0);
    m->tune_ln_grid = env("TIP_LN_GRID", -1);
    m->tune_ln_share = env("TIP_LN_SHARE", -1);
    m->tune_rnn_clusters = env("TIP_RNN_UMMA_CLUSTERS", 0);
    m->tune_attn_grid = env("TIP_ATTN_GRID", 0);
}
extern "C" int tip_set_tuning(tip_model* m, const char* key, int value) {
    if (!m || !key) return TIP_ERR_INVALID_ARG;
    const std::string k(key);
    if (k == "atm") m->tune_atm = value;
    else if (k == "atm_grid") m->tune_atm_grid = value;
    else if (k == "atm_pair") m->tune_atm_pair = value;
    else if (k == "atm_min_tiles") m->tune_atm_min_tiles = value;
    else if (k == "dyn_sched") m->tune_dyn_sched = value;
    else if (k == "ln_pair") m->tune_ln_pair = value;
    else if (k == "ln_grid") m->tune_ln_grid = value;
    else if (k == "ln_share") m->tune_ln_share = value;
    else if (k == "rnn_clusters") m->tune_rnn_clusters = value;
    else if (k == "attn_grid") m->tune_attn_grid = value;
    else { m->set_error("tip_set_tuning: unknown key '" + k + "'"); return TIP_ERR_INVALID_ARG; }
    drop_graphs(m);                     // captured forwards have the old kernel choice baked in
    return TIP_OK;
}
extern "C" int tip_set_gemm_engine(tip_model* m, int engine) {
    if (!m || engine < 0 || engine > 2) return TIP_ERR_INVALID_ARG;
    m->engine = engine;
    drop_graphs(m);
    return TIP_OK;
}
extern "C" int tip_set_use_graphs(tip_model* m, int enable) {
    if (!m) return TIP_ERR_INVALID_ARG;
    m->use_graphs = enable ? 1 : 0;
    drop_graphs(m);
    return TIP_OK;
}
extern "C" int tip_last_launch_count(const tip_model* m) { return m ? m->launches : 0; }

extern "C" int tip_debug_rnn_timestamps(unsigned long long* host_out) {
    if (!g_rnn_tbuf) { cudaMalloc(&g_rnn_tbuf, 64); cudaMemset(g_rnn_tbuf, 0, 64); return TIP_OK; }
    return cudaMemcpy(host_out, g_rnn_tbuf, 64, cudaMemcpyDeviceToHost) == cudaSuccess ? TIP_OK : TIP_ERR_CUDA;
}
extern "C" int tip_debug_timestamps(unsigned long long* host_out, int n) {
    if (!g_tbuf || n > 64 * 32) return TIP_ERR_INVALID_ARG;
    return cudaMemcpy(host_out, g_tbuf, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess ? TIP_OK : TIP_ERR_CUDA;
}
extern "C" int tip_set_profile(tip_model* m, int enable) {
    if (!m) return TIP_ERR_INVALID_ARG;
    m->profile = enable ? 1 : 0;
    m->st_names.clear();
    m->st_layers.clear();
    drop_graphs(m);
    return TIP_OK;
}
extern "C" int tip_profile_stages(const tip_model* m) { return m ? (int)m->st_names.size() : 0; }
extern "C" int tip_profile_get(tip_model* m, int i, char* name, int name_cap, int* layer, float* ms) {
    if (!m || i < 0 || i >= (int)m->st_names.size() || (size_t)i + 1 >= m->ev_pool.size()) return TIP_ERR_INVALID_ARG;
    if (name && name_cap > 0) { strncpy(name, m->st_names[i].c_str(), name_cap - 1); name[name_cap - 1] = 0; }
    if (layer) *layer = m->st_layers[i];
    if (ms) {
        TIP_CUDA_TRY(m, cudaEventSynchronize(m->ev_pool[i + 1]));
        TIP_CUDA_TRY(m, cudaEventElapsedTime(ms, m->ev_pool[i], m->ev_pool[i + 1]));
    }
    return TIP_OK;
}

extern "C" int tip_debug_tensor(tip_model* m, const char* name, float* dst, int64_t capacity,
                                int64_t* numel, void* stream_) {
    if (!m || !name || !numel) return TIP_ERR_INVALID_ARG;
    const std::string n(name);
    const int64_t rows = m->last_rows;
    const float* src = nullptr;
    if (n == "embed")    { src = m->xa;  *numel = rows * E; }
    else if (n == "qkv") { src = m->qkv; *numel = rows * 3 * E; }
    else if (n == "gi")  { src = m->gi;  *numel = rows * R; }
    else if (n == "hs")  { src = m->hs;  *numel = rows * R; }
    else if (n == "win_imu") { src = m->win_imu; *numel = (int64_t)m->n_streams * MAXL * m->d.n_imu; }
    else if (n == "win_s")   { src = m->win_s;   *numel = (int64_t)m->n_streams * MAXL * m->d.size_s; }
    else { m->set_error("tip_debug_tensor: unknown buffer " + n); return TIP_ERR_INVALID_ARG; }
    if (dst) {
        if (capacity < *numel || !src) { m->set_error("tip_debug_tensor: capacity too small"); return TIP_ERR_INVALID_ARG; }
        TIP_CUDA_TRY(m, cudaMemcpyAsync(dst, src, *numel * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream_));
    }
    return TIP_OK;
}

extern "C" int tip_algorithmic_cost(const tip_model* m, int B, int L, double* bytes, double* flops) {
    if (!m || B < 1 || L < 1) return TIP_ERR_INVALID_ARG;
    const Dims& d = m->d;
    double wparams = (double)E * d.d_in + E;
    wparams += d.layers * ((double)3 * E * E + 3 * E + (double)E * E + E + (double)F * E + F + (double)E * F + E + 4.0 * E);
    if (d.with_rnn) wparams += (double)R * E + (double)R * R + 2.0 * R;
    wparams += (double)d.size_s * d.khead + d.size_s;
    double macs_row = (double)E * d.d_in + d.layers * ((double)3 * E * E + (double)E * E + 2.0 * E * F) +
                      (d.with_rnn ? (double)R * E + (double)R * R : 0.0) + (double)d.size_s * d.khead;
    double macs_attn = d.layers * 2.0 * (double)L * L * E;   // QK^T + PV, dense (unmasked) count
    if (bytes) *bytes = wparams * 4.0 + (double)B * L * 4.0 * (d.d_in + d.size_s);
    if (flops) *flops = 2.0 * B * (macs_row * L + macs_attn);
    return TIP_OK;
}

// ------------------------------------------------------------------------------------------------
static int ensure_workspace(tip_model* m, int rows) {
    if (rows <= m->cap_rows) return TIP_OK;
    const Dims& d = m->d;
    int cap = (int)align_up((size_t)std::max(rows, 128), 128);
    if (m->ws) { cudaFree(m->ws); m->ws = nullptr; m->cap_rows = 0; }
    m->plane_xin = (size_t)cap * d.kin_pad;
    m->plane_e = (size_t)cap * E;
    m->plane_f = (size_t)cap * F;
    m->plane_r = (size_t)cap * R;
    size_t p = 0;
    auto take = [&](size_t n) { size_t r = p; p = align_up(p + n, 256); return r; };
    // one fp32 plane per activation; the tcgen05 engine uses the same bytes as two fp16 planes (hi, lo)
    const size_t o_xin = take(m->plane_xin), o_xa = take(m->plane_e), o_xb = take(m->plane_e),
                 o_qkv = take((size_t)cap * 3 * E), o_att = take(m->plane_e),
                 o_hid = take(m->plane_f), o_gi = take(m->plane_r), o_hs = take(m->plane_r), o_pre = take(m->plane_e);
    cudaError_t e = cudaMalloc(&m->ws, p * sizeof(float));
    if (e != cudaSuccess) {
        m->set_error(std::string("cudaMalloc(workspace): ") + cudaGetErrorString(e));
        return TIP_ERR_OOM;
    }
    cudaMemset(m->ws, 0, p * sizeof(float));
    m->xin = m->ws + o_xin; m->xa = m->ws + o_xa; m->xb = m->ws + o_xb; m->qkv = m->ws + o_qkv;
    m->att = m->ws + o_att; m->hid = m->ws + o_hid; m->gi = m->ws + o_gi; m->hs = m->ws + o_hs; m->pre = m->ws + o_pre;
    m->cap_rows = cap;
    m->maps_ready = false;
    drop_graphs(m);
    return TIP_OK;
}

static void launch_sgemm(tip_model* m, cudaStream_t st, const float* A, int lda, const float* W, int ldw,
                         int M, int N, int K, const Epi& ep, bool ln) {
    if (ln) {
        sgemm_nt_kernel<32, 256, 4, true><<<dim3(1, (M + 31) / 32), 256, 0, st>>>(A, lda, W, ldw, M, N, K, ep);
    } else if (M >= 1024) {
        sgemm_nt_kernel<128, 128, 8, false><<<dim3((N + 127) / 128, (M + 127) / 128), 256, 0, st>>>(A, lda, W, ldw, M, N, K, ep);
    } else {
        sgemm_nt_kernel<32, 256, 4, false><<<dim3((N + 255) / 256, (M + 31) / 32), 256, 0, st>>>(A, lda, W, ldw, M, N, K, ep);
    }
    m->launches++;
}

static void launch_attention(tip_model* m, cudaStream_t st, const float* qkv, float* out, float* out_lo,
                             int B, int L, float drop_p, uint64_t seed, int row0 = 0) {
    const uint64_t* sp = m->d_seed;            // base seed (device); `seed` = site offset
    const int b0 = row0 / L;                   // first window of this launch (dropout indices are batch-global)
    pdl_kind() = 2;
    if (out_lo) {
        // tcgen05 engine: qkv and the output are FP16 hi/lo planes; warp-level tensor-core kernel
        static const int hpb = getenv("TIP_ATTN_HPB") ? atoi(getenv("TIP_ATTN_HPB")) : 8;
        if (!m->attn_attr_set) {      // function attributes are per device context: once per handle
            cudaFuncSetAttribute(attention_mma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg<8>::SMEM_BYTES);
            cudaFuncSetAttribute(attention_mma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg<4>::SMEM_BYTES);
            cudaFuncSetAttribute(attention_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg<2>::SMEM_BYTES);
            m->attn_attr_set = true;
        }
        const __half* qh = reinterpret_cast<const __half*>(qkv) + (size_t)row0 * 3 * E;
        const __half* ql = qh + (size_t)m->cap_rows * 3 * E;
        __half* oh = reinterpret_cast<__half*>(out) + (size_t)row0 * E;
        __half* ol = reinterpret_cast<__half*>(out_lo) + (size_t)row0 * E;
        if (m->tune_attn_grid != 0 && hpb == 8) {
            // persistent, double-buffered variant: tune_attn_grid CTAs (-1: two per SM) walk the (window, head group) units
            static bool pipe_attr = false;
            if (!pipe_attr) { cudaFuncSetAttribute(attention_mma_pipe_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * AttnCfg<8>::SMEM_BYTES); pipe_attr = true; }
            const int units = B * (NH / 8);
            const int grid = std::min(units, m->tune_attn_grid > 0 ? m->tune_attn_grid : 2 * m->maps.num_sms);
            launch_k(attention_mma_pipe_kernel<8>, dim3(grid), dim3(256), 2 * AttnCfg<8>::SMEM_BYTES, st, qh, ql, oh, ol, L, drop_p, sp, seed, b0, B);
            m->launches++;
            return;
        }
        // smaller CTAs (fewer heads each) quantise better over the SMs; the qkv pieces stay >= 64 bytes
        static const int attn_ts = getenv("TIP_TS") ? atoi(getenv("TIP_TS")) : 0;
        if (hpb == 8)      launch_k(attention_mma_kernel<8>, dim3(dim3(B, NH / 8)), dim3(256), AttnCfg<8>::SMEM_BYTES, st, qh, ql, oh, ol, L, drop_p, sp, seed, b0,
                                    (attn_ts && g_tbuf) ? g_tbuf + 1500 : (unsigned long long*)nullptr);
        else if (hpb == 2) launch_k(attention_mma_kernel<2>, dim3(dim3(B, NH / 2)), dim3(64), AttnCfg<2>::SMEM_BYTES, st, qh, ql, oh, ol, L, drop_p, sp, seed, b0, (unsigned long long*)nullptr);
        else               launch_k(attention_mma_kernel<4>, dim3(dim3(B, NH / 4)), dim3(128), AttnCfg<4>::SMEM_BYTES, st, qh, ql, oh, ol, L, drop_p, sp, seed, b0, (unsigned long long*)nullptr);
        m->launches++;
        return;
    }
    static const int akind = getenv("TIP_ATTN") ? atoi(getenv("TIP_ATTN")) : 0;
    if (B >= 32 && akind == 1) attention_kernel<4, 4><<<dim3(B, NH / 4), 320, 0, st>>>(qkv, out, out_lo, L, drop_p, sp, seed, b0);
    else if (B >= 32) attention_kernel<2, 4><<<dim3(B, NH / 4), 160, 0, st>>>(qkv, out, out_lo, L, drop_p, sp, seed, b0);
    else         attention_kernel<4, 2><<<dim3(B, NH / 2), 160, 0, st>>>(qkv, out, out_lo, L, drop_p, sp, seed, b0);
    m->launches++;
}

// fused QKV projection + attention of layer l for the windows [w0, w0 + nw): xa -> att (FP16 hi/lo planes)
// TIP_FUSED_ATTN: 0 never, 1 always, 2 (default) when the forward's work units fit ONE wave (<= 148 units = 111 windows at
// L = 40).  Measured (stage times incl. ~4 us of event overhead per kernel): B = 1: 20.4 vs 11.0 + 10.4 us for the QKV GEMM +
// attention kernels, B = 64: 21.0 vs 24.4, B = 256: 44.9 vs 22.5 + 18.4 -- with several rounds per SM the attention
// epilogue (CUDA cores, 8 warps per SM: 49 % issue utilisation, ncu) is slower than the dedicated one-wave mma.sync kernel.
static int fused_attn_kind() {
    static const int v = getenv("TIP_FUSED_ATTN") ? atoi(getenv("TIP_FUSED_ATTN")) : 2;
    return v;
}
static void launch_qkv_attn(tip_model* m, cudaStream_t st, int l, int w0, int nw, int L, float drop_p, uint64_t seed_off, int pdl_early) {
    pdl_kind() = 1;
    if (!m->qa_attr_set) {
        cudaFuncSetAttribute(qkv_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, QA_SMEM_BYTES);
        m->qa_attr_set = true;
    }
    const LayerOff& Lo = m->off.layer[l];
    const int wpt = UM_BM / L;
    const int units = ((nw + wpt - 1) / wpt) * QA_GROUPS;
    __half* oh = reinterpret_cast<__half*>(m->att);
    __half* ol = reinterpret_cast<__half*>(m->att + m->plane_e / 2);
    launch_k(qkv_attn_kernel, dim3(std::min(units, m->maps.num_sms)), dim3(UM_THREADS), QA_SMEM_BYTES, st,
             m->maps.a_xa32.hi, m->maps.a_xa32.lo, m->maps.w_qkvr[l].hi, m->maps.w_qkvr[l].lo,
             (const float*)(m->blob + Lo.bqkvr), oh, ol, (const float*)(m->blob + m->off.scales + SC_LAYER0 + 4 * l + 0),
             w0, nw, L, drop_p, drop_threshold(drop_p), (const uint64_t*)m->d_seed, seed_off, pdl_early);
    m->launches++;
}

// fused feed-forward block of layer l (ff1 + ReLU [+ dropout] + ff2 [+ dropout2] + residual + LayerNorm2) for the row tiles
// [T0, T0 + TN): xb -> xa; the hidden activation stays on the SM.  TIP_FFN_FUSED: 0 (default) never, 1 always, 2: when the
// forward has more row tiles than the un-fused LayerNorm path takes (TIP_SKINNY_TILES, 74).  Parity-green (it runs the
// B = 256 goldens and the mask-for-mask stochastic tests when switched on) but not faster: 63.3 us per layer at B = 256
// against 25.7 + 30.6 us for ff1 and ff2 + LayerNorm, 3 lanes 415 vs 411 us per forward.  ncu: the tensor pipe of the 80 busy
// SMs is 88 % active yet delivers 44 % of its peak -- the kernel is SHARED-MEMORY-BANDWIDTH bound: every k-step of a
// 128 x N tile reads 3 x (128 + N) x 32 B of operands (three MMAs of the split), i.e. 128 B/clk at N = 128, all the SM has,
// while TMA writes the next stages and the epilogue writes hidA into the same memory.  (The same arithmetic explains the
// 40-45 % tensor-pipe ceiling of the 128 x 128 GEMMs.)
static int ffn_fused_kind() {
    static const int v = getenv("TIP_FFN_FUSED") ? atoi(getenv("TIP_FFN_FUSED")) : 0;
    return v;
}
static void launch_ffn(tip_model* m, cudaStream_t st, int l, int M, int T0, int TN, float drop_p, uint64_t seed_base, int pdl_early) {
    pdl_kind() = 1;
    if (!m->ffn_attr_set) {
        cudaFuncSetAttribute(ffn_ln_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FF_SMEM_BYTES);
        cudaFuncSetAttribute(ffn_ln_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FF_SMEM_BYTES);
        m->ffn_attr_set = true;
    }
    const LayerOff& Lo = m->off.layer[l];
    const float* W = m->blob;
    FfnArgs a{};
    a.b1 = W + Lo.b1; a.b2 = W + Lo.b2; a.gamma = W + Lo.g2; a.beta = W + Lo.be2;
    a.sc1 = W + m->off.scales + SC_LAYER0 + 4 * l + 2;
    a.sc2 = W + m->off.scales + SC_LAYER0 + 4 * l + 3;
    a.M = M; a.m_tile0 = T0; a.m_tiles = TN;
    a.drop_thr = drop_threshold(drop_p); a.drop_inv = drop_inv_keep(drop_p);
    a.seed_ptr = m->d_seed; a.seed_ff1 = seed_base + seed_ff1(l); a.seed_ff2 = seed_base + seed_ff2(l);
    a.pdl_early = pdl_early;
    const UmmaMaps& mp = m->maps;
    const dim3 grid(std::min(TN, mp.num_sms));
    if (a.drop_thr)
        launch_k(ffn_ln_kernel<true>, grid, dim3(UM_THREADS), FF_SMEM_BYTES, st, mp.a_xb32.hi, mp.a_xb32.lo, mp.w_1k32[l].hi, mp.w_1k32[l].lo,
                 mp.w_2k32[l].hi, mp.w_2k32[l].lo, mp.a_xb.hi, mp.a_xb.lo, mp.o_xa.c0, mp.o_xa.c1, a);
    else
        launch_k(ffn_ln_kernel<false>, grid, dim3(UM_THREADS), FF_SMEM_BYTES, st, mp.a_xb32.hi, mp.a_xb32.lo, mp.w_1k32[l].hi, mp.w_1k32[l].lo,
                 mp.w_2k32[l].hi, mp.w_2k32[l].lo, mp.a_xb.hi, mp.a_xb.lo, mp.o_xa.c0, mp.o_xa.c1, a);
    m->launches++;
}

static void launch_rnn(tip_model* m, cudaStream_t st, const float* gi, float* hs, float* hs_lo, int B, int L) {
    pdl_kind() = 4;
    if (m->rnn_clusters < 0) {
        // how many 8-CTA clusters the device can co-schedule (GPC layout dependent; 16..18 on B200)
        cudaFuncSetAttribute(rnn_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RC_SMEM_BYTES);
        cudaLaunchConfig_t q{};
        q.gridDim = dim3(RC_CTAS * 18); q.blockDim = dim3(256); q.dynamicSmemBytes = RC_SMEM_BYTES;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = RC_CTAS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        q.attrs = nullptr; q.numAttrs = 0;   // the kernel carries __cluster_dims__ itself
        (void)at;
        int n = 0;
        const cudaError_t pending = cudaPeekAtLastError();     // an earlier launch error must survive the query
        cudaError_t qe = cudaOccupancyMaxActiveClusters(&n, rnn_cluster_kernel, &q);
        if (getenv("TIP_VERBOSE")) fprintf(stderr, "[tip] cudaOccupancyMaxActiveClusters -> %d (%s)\n", n, cudaGetErrorString(qe));
        if (qe != cudaSuccess || n < 1) { n = 0; if (pending == cudaSuccess) cudaGetLastError(); }
        if (getenv("TIP_RNN_CLUSTERS")) n = atoi(getenv("TIP_RNN_CLUSTERS"));
        if (getenv("TIP_VERBOSE")) fprintf(stderr, "[tip] rnn clusters co-schedulable: %d\n", n);
        m->rnn_clusters = n;
    }
    static const int rnn_kind = getenv("TIP_RNN") ? atoi(getenv("TIP_RNN")) : 0;   // 1: force the FFMA cluster kernel, 2: force the tensor-core one
    // small batches (one group of <= 8 windows per cluster): the FFMA cluster kernel with W_hh in registers has the
    // shorter step (2.5 vs 2.8 us: no MMA issue phase) -- measured 98 vs 119 us at B <= 8, 104 vs 122 us at B = 64, forward 416 vs 434 us at B = 96;
    // from ~100 windows on it becomes FFMA-bound and the tensor-core recurrence wins (588 vs 127 us at B = 256)
    static const int ffma_max_b = getenv("TIP_RNN_FFMA_MAX_B") ? atoi(getenv("TIP_RNN_FFMA_MAX_B")) : 120;
    const bool small_b = m->rnn_clusters > 0 && B <= std::min(ffma_max_b, RC_GROUP * m->rnn_clusters) && rnn_kind != 2;
    if (hs_lo && m->maps_ready && rnn_kind != 1 && !small_b && !m->rnn_stream_fallback) {
        // tensor-core recurrence (tcgen05 engine): clusters of 8 CTAs, RU_N windows each
        if (m->rnn_umma_clusters < 0) {
            cudaFuncSetAttribute(rnn_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RU_SMEM_BYTES);
            cudaFuncSetAttribute(rnn_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RU_SMEM_BYTES);
            cudaFuncSetAttribute(rnn_umma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RU_SMEM_BYTES);
            cudaFuncSetAttribute(rnn_umma_kernel<false, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, RU_SMEM_BYTES);
            cudaFuncSetAttribute(rnn_umma_kernel<false, false, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RU_SMEM_BYTES);
            cudaFuncSetAttribute(rnn_umma_kernel<false, false, 1, true, 20>, cudaFuncAttributeMaxDynamicSharedMemorySize, RU_SMEM_BYTES);
            cudaLaunchConfig_t q{};
            q.gridDim = dim3(RU_CTAS * 18); q.blockDim = dim3(RU_THREADS); q.dynamicSmemBytes = RU_SMEM_BYTES;
            int n = 0;
            const cudaError_t pending = cudaPeekAtLastError();
            if (cudaOccupancyMaxActiveClusters(&n, rnn_umma_kernel<false>, &q) != cudaSuccess || n < 1) { n = 0; if (pending == cudaSuccess) cudaGetLastError(); }
            if (getenv("TIP_VERBOSE")) fprintf(stderr, "[tip] rnn_umma clusters co-schedulable: %d\n", n);
            m->rnn_umma_clusters = n;
        }
        if (m->rnn_umma_clusters > 0) {
            const int blocks = (B + RU_N - 1) / RU_N;
            const int nc = std::min(blocks, m->rnn_umma_clusters);
            static const int a_tmem = getenv("TIP_RNN_TMEMA") ? atoi(getenv("TIP_RNN_TMEMA")) : 0;
            const __half* wh = reinterpret_cast<const __half*>(m->blob + m->off.whh_hi);
            const __half* wl = reinterpret_cast<const __half*>(m->blob + m->off.whh_lo);
            static const int st_async = getenv("TIP_RNN_STASYNC") ? atoi(getenv("TIP_RNN_STASYNC")) : 0;
            static const int iss2 = getenv("TIP_RNN_ISS2") ? atoi(getenv("TIP_RNN_ISS2")) : 0;
            static const int nw20 = getenv("TIP_RNN_NW20") ? atoi(getenv("TIP_RNN_NW20")) : 1;
            static const int pipe = getenv("TIP_RNN_PIPE") ? atoi(getenv("TIP_RNN_PIPE")) : 1;     // pipelined all-gather (default)
            if (a_tmem)
                launch_k(rnn_umma_kernel<true>, dim3(nc * RU_CTAS), dim3(RU_THREADS), RU_SMEM_BYTES, st,
                    m->maps.w_hh.hi, m->maps.w_hh.lo, gi, reinterpret_cast<__half*>(hs), reinterpret_cast<__half*>(hs_lo),
                    m->blob + m->off.scales + SC_HH, B, L, g_rnn_tbuf, wh, wl);
            else if (pipe && nw20 && ((B + 19) / 20 <= m->rnn_umma_clusters || m->tune_rnn_clusters > 0))
                // the batch fits the co-resident clusters with 20 windows each: smaller all-gather per step.  (rnn_clusters
                // knob: fewer clusters, each walking several 20-window groups -- a narrower, longer kernel for laned runs)
                launch_k(rnn_umma_kernel<false, false, 1, true, 20>,
                    dim3((m->tune_rnn_clusters > 0 ? std::min((B + 19) / 20, std::min(m->tune_rnn_clusters, m->rnn_umma_clusters)) : (B + 19) / 20) * RU_CTAS),
                    dim3(RU_THREADS), RU_SMEM_BYTES, st,
                    m->maps.w_hh.hi, m->maps.w_hh.lo, gi, reinterpret_cast<__half*>(hs), reinterpret_cast<__half*>(hs_lo),
                    m->blob + m->off.scales + SC_HH, B, L, g_rnn_tbuf, wh, wl);
            else if (pipe)
                launch_k(rnn_umma_kernel<false, false, 1, true>, dim3(nc * RU_CTAS), dim3(RU_THREADS), RU_SMEM_BYTES, st,
                    m->maps.w_hh.hi, m->maps.w_hh.lo, gi, reinterpret_cast<__half*>(hs), reinterpret_cast<__half*>(hs_lo),
                    m->blob + m->off.scales + SC_HH, B, L, g_rnn_tbuf, wh, wl);
            else if (iss2)
                launch_k(rnn_umma_kernel<false, false, 2>, dim3(nc * RU_CTAS), dim3(RU_THREADS), RU_SMEM_BYTES, st,
                    m->maps.w_hh.hi, m->maps.w_hh.lo, gi, reinterpret_cast<__half*>(hs), reinterpret_cast<__half*>(hs_lo),
                    m->blob + m->off.scales + SC_HH, B, L, g_rnn_tbuf, wh, wl);
            else if (st_async)
                launch_k(rnn_umma_kernel<false, true>, dim3(nc * RU_CTAS), dim3(RU_THREADS), RU_SMEM_BYTES, st,
                    m->maps.w_hh.hi, m->maps.w_hh.lo, gi, reinterpret_cast<__half*>(hs), reinterpret_cast<__half*>(hs_lo),
                    m->blob + m->off.scales + SC_HH, B, L, g_rnn_tbuf, wh, wl);
            else
                launch_k(rnn_umma_kernel<false>, dim3(nc * RU_CTAS), dim3(RU_THREADS), RU_SMEM_BYTES, st,
                    m->maps.w_hh.hi, m->maps.w_hh.lo, gi, reinterpret_cast<__half*>(hs), reinterpret_cast<__half*>(hs_lo),
                    m->blob + m->off.scales + SC_HH, B, L, g_rnn_tbuf, wh, wl);
            m->launches++;
            return;
        }
    }
    static const int small_on = getenv("TIP_RNN_SMALL") ? atoi(getenv("TIP_RNN_SMALL")) : 1;
    if (small_on && m->rnn_clusters > 0 && !m->rnn_stream_fallback && B <= m->rnn_clusters && rnn_kind != 1) {
        // one window per cluster: latency-optimised kernel (st.async exchange straight from registers).  Measured:
        // recurrence 98 -> 76 us, B = 1 forward 293 -> 264 us.  With 2+ windows per cluster its serial tail (16 lanes
        // doing tanh + 8 remote stores per window) loses to the group kernel below (B = 16: 127 vs 104 us), so it is
        // only used while every window gets its own cluster.
        launch_k(rnn_small_kernel<1>, dim3(B * RC_CTAS), dim3(256), 0, st, gi, m->blob + m->off.whh, hs, hs_lo, B, L, g_rnn_tbuf);
        m->launches++;
        return;
    }
    if (m->rnn_clusters > 0 && !m->rnn_stream_fallback) {
        // windows per cluster pass: spread the batch over all co-resident clusters, in whole groups of 8
        const int ncl = m->rnn_clusters;
        int rpp = (B + ncl - 1) / ncl;
        rpp = std::min(RC_ROWS, std::max(1, rpp));
        if (rpp > RC_GROUP) rpp = (rpp + RC_GROUP - 1) / RC_GROUP * RC_GROUP;
        const int nc = std::min((B + rpp - 1) / rpp, ncl);
        rnn_cluster_kernel<<<nc * RC_CTAS, 256, RC_SMEM_BYTES, st>>>(gi, m->blob + m->off.whh, hs, hs_lo, B, L, rpp);
        m->launches++;
        return;
    }
    const float* whh_t = m->blob + m->off.whh_t;
    if (B >= 8 * 64) rnn_stream_kernel<8><<<(B + 7) / 8, R, 0, st>>>(gi, whh_t, hs, hs_lo, B, L);
    else if (B >= 2 * 74) rnn_stream_kernel<2><<<(B + 1) / 2, R, 0, st>>>(gi, whh_t, hs, hs_lo, B, L);
    else rnn_stream_kernel<1><<<B, R, 0, st>>>(gi, whh_t, hs, hs_lo, B, L);
    m->launches++;
}

// One pass over <= CHUNK_WINDOWS windows.
static int forward_chunk(tip_model* m, const float* x_imu, const float* x_s, float* y, int B, int L,
                         const float* keep_mask, float past_scale, const DropP& drop,
                         cudaStream_t st, int w0 = 0, int nw = -1, uint64_t seed_chunk = 0) {
    // [w0, w0 + nw): the windows of the batch this call processes (default: all).  x_imu / x_s / y / keep_mask point at
    // window 0; the part uses workspace rows [R0, M).  A part that does not start at window 0 must start on a 128-row
    // tile boundary (tcgen05 engine only); parts of one batch may run concurrently on different streams.
    const Dims& d = m->d;
    const PackOff& o = m->off;
    const float* W = m->blob;
    if (nw < 0) nw = B - w0;
    const int R0 = w0 * L;
    const int M = (w0 + nw) * L;
    const int T0 = R0 / UM_BM, TN = (M - R0 + UM_BM - 1) / UM_BM;
    m->last_rows = (int64_t)B * L;
    pdl_rows() = M - R0;                      // programmatic dependent launch only pays for small forwards
    int rc = ensure_workspace(m, B * L);
    if (rc != TIP_OK) return rc;
    const float p_in = drop.p_in, p_past = drop.p_past, p_enc = drop.p_enc;
    const uint64_t seed = seed_chunk;          // site offsets are added to this; the call's base seed is read from m->d_seed on the device
    const bool umma = (m->engine == 2) || (m->engine == 0 && UMMA_AVAILABLE);
    if (umma) {
        if (!m->maps_ready) {
            rc = umma_build_maps(m->maps, m->blob, o, d, m->xin, m->plane_xin, m->xa, m->xb, m->att,
                                 m->plane_e, m->hid, m->plane_f, m->hs, m->plane_r, m->qkv, m->gi, m->pre, m->cap_rows, m->err);
            if (rc != TIP_OK) return rc;
            m->maps_ready = true;
        }
    }
    float* lo_xin = umma ? m->xin + m->plane_xin / 2 : nullptr;
    float* lo_xa = umma ? m->xa + m->plane_e / 2 : nullptr;
    float* lo_xb = umma ? m->xb + m->plane_e / 2 : nullptr;
    float* lo_att = umma ? m->att + m->plane_e / 2 : nullptr;
    float* lo_hid = umma ? m->hid + m->plane_f / 2 : nullptr;
    float* lo_hs = umma ? m->hs + m->plane_r / 2 : nullptr;

    // TIP_SKIP (diagnostic, results become garbage): bit mask of stages NOT launched -- 1 condition, 2 in_linear, 4 qkv,
    // 8 attention, 16 out_proj+LN, 32 ff1, 64 ff2+LN, 128 rnn_ih, 256 rnn, 512 head.  tools/lane_ablation.py times the
    // laned forward with one stage removed at a time: the stage's MARGINAL cost when other lanes fill the machine.
    static const int skip = getenv("TIP_SKIP") ? atoi(getenv("TIP_SKIP")) : 0;
    m->st_names.clear();
    m->st_layers.clear();
    mark(m, st, "condition");
    {
        const int nr = M - R0;
        const int64_t total = (int64_t)nr * (d.kin_pad / 8);
        const int blocks = (skip & 1) ? 1 : (int)std::min<int64_t>((total + 255) / 256, 148 * 8);
        float* xo = umma ? reinterpret_cast<float*>(reinterpret_cast<__half*>(m->xin) + (size_t)R0 * d.kin_pad) : m->xin + (size_t)R0 * d.kin_pad;
        float* xl = umma ? reinterpret_cast<float*>(reinterpret_cast<__half*>(lo_xin) + (size_t)R0 * d.kin_pad) : nullptr;
        pdl_kind() = 8;
        launch_k(condition_kernel, dim3(blocks), dim3(256), 0, st, x_imu + (size_t)R0 * d.n_imu, x_s + (size_t)R0 * d.size_s,
                 keep_mask ? keep_mask + (size_t)R0 * d.size_s : nullptr, past_scale, xo, xl, nr,
                 d.n_imu, d.size_s, d.kin_pad, p_in, p_past, (const uint64_t*)m->d_seed, seed, R0);
        m->launches++;
    }
    auto gemm = [&](int which, int layer, const float* A, int K, const float* Wp, int N, Epi ep, bool ln) {
        if (skip & (which == UG_IN ? 2 : which == UG_QKV ? 4 : which == UG_OUT ? 16 : which == UG_FF1 ? 32 : which == UG_FF2 ? 64
                    : which == UG_IH ? 128 : 512)) return;
        ep.seed_ptr = m->d_seed;
        ep.drop_thr = drop_threshold(ep.drop_p);
        ep.drop_inv = drop_inv_keep(ep.drop_p);
        if (umma) {
            const int sc = which == UG_IN ? SC_IN : which == UG_IH ? SC_IH : (which == UG_HEAD_R || which == UG_HEAD_E) ? SC_HEAD
                           : SC_LAYER0 + 4 * layer + (which == UG_QKV ? 0 : which == UG_OUT ? 1 : which == UG_FF1 ? 2 : 3);
            ep.acc_scale = W + o.scales + sc;
            static const int dbg = getenv("TIP_DBG") ? atoi(getenv("TIP_DBG")) : 0;
            ep.dbg = dbg;
            ep.pdl_early = (M - R0 <= 1024) ? 1 : 0;
            // dynamic tile scheduler of the plain GEMMs: one counter pair per launch of the forward (the host entry's second
            // batch part, which runs concurrently on its own stream, uses the upper half of the slots)
            ep.sched = (m->tune_dyn_sched && !ln) ? reinterpret_cast<int*>(m->d_seed + 1) + 2 * ((T0 > 0 ? 64 : 0) + (sc % 64)) : nullptr;
            ep.tbuf = nullptr;
            static const int ts_on = getenv("TIP_TS") ? atoi(getenv("TIP_TS")) : 0;
            if ((dbg & 4) || ts_on) {
                static unsigned long long* tb = nullptr;
                if (!tb) cudaMalloc(&tb, 64 * 32 * sizeof(unsigned long long));
                ep.tbuf = tb + 8 * (which + (layer > 0 ? 16 : 0));
                g_tbuf = tb;
            }
            // The fused LayerNorm GEMM needs whole rows per CTA, i.e. ONE CTA per 128-row tile: with few row tiles most SMs
            // idle while each busy one streams the whole weight matrix (ff2: 17 us per tile).  Up to 74 row tiles
            // (4 x 74 = two full waves of 148 CTAs) it is faster to split the columns over 4 CTAs (64-column tiles into
            // the fp32 scratch) and normalise in a second, small kernel.  Measured forward, un-fused vs fused: B = 1
            // 328 vs 401 us, 32: 303 vs 373, 64: 321 vs 387, 96: 350 vs 412, 128: 393 vs 412, 192: 444 vs 459;
            // B = 256 (80 tiles = 320 CTAs = three waves): 549 vs 514, so the fused kernel keeps the big batches.
            static const int skinny_tiles = getenv("TIP_SKINNY_TILES") ? atoi(getenv("TIP_SKINNY_TILES")) : 74;
            // A operand resident in tensor memory (tip_umma_atm.cuh): a CTA owns a contiguous run of (row tile, n-tile) pairs,
            // keeps the row tile's A in TMEM and streams only W.  Knobs (tip_set_tuning / TIP_ATM, TIP_ATM_GRID,
            // TIP_ATM_MIN_TILES): which GEMMs use it (bit mask 1 in_linear, 2 qkv, 4 ff1, 8 rnn_ih; auto = all four on handles
            // that run as execution lanes, where its smaller footprint -- one CTA per row tile by default -- leaves SMs to the
            // other lanes' kernels; a lone forward is faster on the plain kernel's 148 CTAs), CTAs per launch, minimum row tiles.
            const int atm_mask = m->tune_atm >= 0 ? m->tune_atm : (m->laned ? 15 : 0);
            const int atm_min_tiles = m->tune_atm_min_tiles, atm_grid = m->tune_atm_grid > 0 ? m->tune_atm_grid : std::min((TN + 1) / 2, NARROW_CTAS);
            const int atm_bit = which == UG_IN ? 1 : which == UG_QKV ? 2 : which == UG_FF1 ? 4 : which == UG_IH ? 8 : 0;
            if (!ln && (atm_mask & atm_bit) && K == E && (N % AT_BN) == 0 && N <= AT_MAX_N_PER_UNIT && TN >= atm_min_tiles && !(dbg & 7)) {
                const UmmaMaps& mp = m->maps;
                const UmmaOperand& Am = which == UG_IN ? mp.a_xin : which == UG_FF1 ? mp.a_xb : mp.a_xa;
                const UmmaOperand& Bm = which == UG_IN ? mp.w_in : which == UG_QKV ? mp.w_qkv[layer] : which == UG_FF1 ? mp.w_1[layer] : mp.w_ih;
                const UmmaOutput& Cm = which == UG_IN ? mp.o_xa : which == UG_QKV ? mp.o_qkv : which == UG_FF1 ? mp.o_hid : mp.o_gi;
                if (Cm.valid) {
                    if (ep.tbuf) ep.tbuf = g_tbuf + 1024 + 64 * (which == UG_IN ? 0 : which == UG_QKV ? 1 : which == UG_FF1 ? 2 : 3);
                    const UmmaOperand* B64 = which == UG_QKV ? &mp.w_qkv64[layer] : which == UG_FF1 ? &mp.w_164[layer] : which == UG_IH ? &mp.w_ih64 : nullptr;
                    if (m->tune_atm_pair && B64 && (TN % 2) == 0 && (T0 % 2) == 0)
                        launch_atm_pair_gemm(mp, Am, *B64, Cm, M, N, T0, TN, atm_grid, ep, st);
                    else
                        launch_atm_gemm(mp, Am, Bm, Cm, M, N, T0, TN, atm_grid, ep, st);
                    m->launches++;
                    return;
                }
            }
            m->maps.ln_pair_min_k = m->tune_ln_pair;
            m->maps.ln_grid = m->tune_ln_grid > 0 ? m->tune_ln_grid : (m->tune_ln_grid < 0 && m->laned ? std::min((TN + 1) / 2, NARROW_CTAS) : 0);
            // LayerNorm GEMMs in throughput mode: a CTA takes PAIRS of row tiles that share every W k-block (tip_umma_ln2.cuh).
            // Knob "ln_share": 1 on, 0 off, -1 (default) = on for handles that run as lanes.
            if (ln && TN > skinny_tiles && (which == UG_OUT || which == UG_FF2) && m->maps.o_xa.valid && m->maps.o_xb.valid &&
                (m->tune_ln_share > 0 || (m->tune_ln_share < 0 && m->laned)) && !(dbg & 7) && !ts_on) {
                const int cap = m->tune_ln_grid > 0 ? m->tune_ln_grid : NARROW_CTAS;
                launch_ln2_gemm(m->maps, which, layer, M, N, K, T0, TN, cap, ep, st);
                m->launches++;
                return;
            }
            if (ln && TN <= skinny_tiles) {
                Epi gp = ep;
                gp.resid = gp.resid_lo = nullptr; gp.gamma = gp.beta = nullptr;
                gp.out = m->pre; gp.out_lo = nullptr; gp.ldc = E;              // fp32 [rows][256] scratch (its own buffer: batch parts may overlap)
                umma_gemm(m->maps, which, layer, M, N, K, gp, false, st, T0, TN, true);
                pdl_kind() = 8;
                launch_k(resid_ln_kernel, dim3((M - R0 + 7) / 8), dim3(256), 0, st, m->pre, reinterpret_cast<const __half*>(ep.resid),
                                                             reinterpret_cast<const __half*>(ep.resid_lo), ep.gamma, ep.beta,
                                                             reinterpret_cast<__half*>(ep.out), reinterpret_cast<__half*>(ep.out_lo), R0, M - R0);
                m->launches += 2;
            } else {
                umma_gemm(m->maps, which, layer, M, N, K, ep, ln, st, T0, TN);
                m->launches++;
            }
        } else {
            launch_sgemm(m, st, A, K, Wp, K, M, N, K, ep, ln);      // FFMA engine: whole batches only (R0 == 0)
        }
    };
    auto head = [&](int which, const float* A, int K) {
        Epi hp{}; hp.bias = W + o.bl; hp.out = y; hp.ldc = d.size_s;
        gemm(which, 0, A, K, W + o.wl, d.size_s, hp, false);
    };
    Epi ep{};
    // in_linear (+ folded head permutation)                                      reference :79-89
    mark(m, st, "in_linear");
    ep = Epi{}; ep.bias = W + o.bin; ep.out = m->xa; ep.out_lo = lo_xa; ep.ldc = E;
    gemm(UG_IN, 0, m->xin, d.kin_pad, W + o.win, E, ep, false);
    for (int l = 0; l < d.layers; ++l) {                                        // reference :91
        const LayerOff& Lo = o.layer[l];
        const int qa_units = ((nw + UM_BM / L - 1) / (UM_BM / L)) * QA_GROUPS;
        const bool fused = umma && (L % 2) == 0 &&                 // (the fused kernel pairs adjacent queries of a window: even L)
                           (fused_attn_kind() == 1 || (fused_attn_kind() == 2 && qa_units <= 148));
        if (fused) {
            mark(m, st, "qkv_attn", l);             // projection + attention in one kernel: q | k | v never leave the SM
            launch_qkv_attn(m, st, l, w0, nw, L, p_enc, seed + seed_attn(l), (M - R0 <= 1024) ? 1 : 0);
        } else {
        mark(m, st, "qkv", l);
        ep = Epi{}; ep.bias = W + Lo.bqkv; ep.out = m->qkv; ep.ldc = 3 * E;
        if (umma) ep.out_lo = m->qkv + (size_t)m->cap_rows * 3 * E / 2;     // FP16 planes for the mma attention
        gemm(UG_QKV, l, m->xa, E, W + Lo.wqkv, 3 * E, ep, false);
        mark(m, st, "attention", l);
        if (!(skip & 8)) launch_attention(m, st, m->qkv, m->att, lo_att, nw, L, p_enc, seed + seed_attn(l), R0);
        }
        mark(m, st, "out_proj_ln", l);
        ep = Epi{}; ep.bias = W + Lo.bo; ep.resid = m->xa; ep.resid_lo = lo_xa; ep.ldr = E;
        ep.gamma = W + Lo.g1; ep.beta = W + Lo.be1; ep.out = m->xb; ep.out_lo = lo_xb; ep.ldc = E;
        ep.drop_p = p_enc; ep.seed = seed + seed_out(l);
        gemm(UG_OUT, l, m->att, E, W + Lo.wo, E, ep, true);
        static const int skinny_tiles_ffn = getenv("TIP_SKINNY_TILES") ? atoi(getenv("TIP_SKINNY_TILES")) : 74;
        if (umma && m->maps.o_xa.valid && (ffn_fused_kind() == 1 || (ffn_fused_kind() == 2 && TN > skinny_tiles_ffn))) {
            mark(m, st, "ffn_ln", l);               // ff1 + ff2 + residual + LayerNorm2 in one kernel: `hid` never leaves the SM
            launch_ffn(m, st, l, M, T0, TN, p_enc, seed, (M - R0 <= 1024) ? 1 : 0);
            continue;
        }
        mark(m, st, "ff1", l);
        ep = Epi{}; ep.bias = W + Lo.b1; ep.relu = 1; ep.out = m->hid; ep.out_lo = lo_hid; ep.ldc = F;
        ep.drop_p = p_enc; ep.seed = seed + seed_ff1(l);
        gemm(UG_FF1, l, m->xb, E, W + Lo.w1, F, ep, false);
        mark(m, st, "ff2_ln", l);
        ep = Epi{}; ep.bias = W + Lo.b2; ep.resid = m->xb; ep.resid_lo = lo_xb; ep.ldr = E;
        ep.gamma = W + Lo.g2; ep.beta = W + Lo.be2; ep.out = m->xa; ep.out_lo = lo_xa; ep.ldc = E;
        ep.drop_p = p_enc; ep.seed = seed + seed_ff2(l);
        gemm(UG_FF2, l, m->hid, F, W + Lo.w2, E, ep, true);
    }
    if (d.with_rnn) {                                                           // reference :95-99
        mark(m, st, "rnn_ih");
        ep = Epi{}; ep.bias = W + o.brnn; ep.out = m->gi; ep.ldc = R;
        gemm(UG_IH, 0, m->xa, E, W + o.wih, R, ep, false);
        mark(m, st, "rnn");
        if (!(skip & 256)) launch_rnn(m, st, m->gi + (size_t)R0 * R, umma ? reinterpret_cast<float*>(reinterpret_cast<__half*>(m->hs) + (size_t)R0 * R) : m->hs,
                   lo_hs ? reinterpret_cast<float*>(reinterpret_cast<__half*>(lo_hs) + (size_t)R0 * R) : nullptr, nw, L);
        mark(m, st, "head");
        head(UG_HEAD_R, m->hs, R);                                              // reference :102
    } else {
        mark(m, st, "head");
        head(UG_HEAD_E, m->xa, E);
    }
    mark(m, st, nullptr);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { m->set_error(std::string("kernel launch: ") + cudaGetErrorString(e)); return TIP_ERR_CUDA; }
    return TIP_OK;
}

// The job pipeline's forwards (tip_forward_host_submit, stream hs_fwd) and forwards a caller queues on its own stream
// share the handle's workspace.  Entry points that are not part of the pipeline first wait (host side) until no
// submitted job still needs the workspace; the pipeline's next forward waits (device side) for the caller's last one.
static int quiesce_host_jobs(tip_model* m) {
    for (auto& s : m->hslot)
        if (s.busy) TIP_CUDA_TRY(m, cudaEventSynchronize(s.ev_fwd));
    return TIP_OK;
}

static int forward_impl(tip_model* m, const float* x_imu, const float* x_s, float* y, int B, int L,
                        const float* keep_mask, float past_scale, const tip_dropout* drop, void* stream_);

extern "C" int tip_forward(tip_model* m, const float* x_imu, const float* x_s, float* y, int B, int L,
                           const float* keep_mask, float past_scale, const tip_dropout* drop, void* stream_) {
    if (!m) return TIP_ERR_INVALID_ARG;
    int rc = order_after_pack(m, (cudaStream_t)stream_);
    if (rc != TIP_OK) return rc;
    if (!m->hs_fwd) return forward_impl(m, x_imu, x_s, y, B, L, keep_mask, past_scale, drop, stream_);
    // the job pipeline has been used on this handle: order this forward against it
    rc = quiesce_host_jobs(m);
    if (rc != TIP_OK) return rc;
    rc = forward_impl(m, x_imu, x_s, y, B, L, keep_mask, past_scale, drop, stream_);
    if (rc != TIP_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream_;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusNone) {
        TIP_CUDA_TRY(m, cudaEventRecord(m->ev_user, st));
        m->ev_user_set = true;
    }
    return TIP_OK;
}

static int forward_impl(tip_model* m, const float* x_imu, const float* x_s, float* y, int B, int L,
                        const float* keep_mask, float past_scale, const tip_dropout* drop, void* stream_) {
    if (!m) return TIP_ERR_INVALID_ARG;
    if (!m->sw->packed) { m->set_error("tip_forward before tip_pack_weights"); return TIP_ERR_NOT_PACKED; }
    if (!x_imu || !x_s || !y || B < 1 || L < 1 || L > MAXL) {
        m->set_error("tip_forward: need non-null tensors, B >= 1 and 1 <= L <= 40");
        return TIP_ERR_INVALID_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream_;
    TIP_CUDA_TRY(m, cudaSetDevice(m->device));
    m->launches = 0;
    const Dims& d = m->d;
    DropP dp = drop_of(drop);
    if (keep_mask) dp.p_past = 0.f;             // an explicit mask replaces the drawn one (:77)
    cudaStreamCaptureStatus cs0 = cudaStreamCaptureStatusNone;
    const bool capturing = !(cudaStreamIsCapturing(st, &cs0) == cudaSuccess && cs0 == cudaStreamCaptureStatusNone);
    if (dp.any() && !capturing) {
        // this call's base seed -> device memory; captured forwards read it there, so a replay draws fresh masks
        // (a caller capturing tip_forward into its own graph sets the seed itself: the stream-step graph does)
        seed_set_kernel<<<1, 1, 0, st>>>(m->d_seed, dp.seed);
    }
    FwdGraph* slot = nullptr;
    const FwdGraphKey key{x_imu, x_s, y, keep_mask, B, L, m->engine, past_scale, dp.p_in, dp.p_past, dp.p_enc};
    if (m->use_graphs && !m->profile && B <= CHUNK_WINDOWS && !getenv("TIP_NO_FWD_GRAPH")) {
        if (!capturing) {
            FwdGraph* lru = &m->fwd_graphs[0];
            for (FwdGraph& g : m->fwd_graphs) {
                if (g.last_use && g.key == key) { slot = &g; break; }
                if (g.last_use < lru->last_use) lru = &g;
            }
            if (slot && slot->exec) {                                    // replay
                slot->last_use = ++m->fwd_tick;
                TIP_CUDA_TRY(m, cudaGraphLaunch(slot->exec, st));
                m->launches = slot->launches;
                m->last_rows = (int64_t)B * L;
                return TIP_OK;
            }
            if (!slot) {                                                 // first sighting: run eagerly, remember the key
                if (lru->exec) cudaGraphExecDestroy(lru->exec);
                *lru = FwdGraph{};
                lru->key = key;
                lru->last_use = ++m->fwd_tick;
            }
        }
    }
    if (slot) {
        // second sighting: capture this forward (all lazy initialisation happened on the first, eager call)
        cudaStream_t cstream;
        TIP_CUDA_TRY(m, cudaStreamCreateWithFlags(&cstream, cudaStreamNonBlocking));
        cudaGraph_t g = nullptr;
        TIP_CUDA_TRY(m, cudaStreamBeginCapture(cstream, cudaStreamCaptureModeThreadLocal));
        int rc = forward_chunk(m, x_imu, x_s, y, B, L, keep_mask, past_scale, dp, cstream);
        cudaError_t ce = cudaStreamEndCapture(cstream, &g);
        cudaGraphExec_t exec = nullptr;
        if (rc == TIP_OK && ce == cudaSuccess) ce = cudaGraphInstantiate(&exec, g, 0);
        if (g) cudaGraphDestroy(g);
        cudaStreamDestroy(cstream);
        if (rc != TIP_OK) return rc;
        if (ce != cudaSuccess) { m->set_error(std::string("forward graph capture: ") + cudaGetErrorString(ce)); return TIP_ERR_CUDA; }
        // forward_chunk may have re-allocated the workspace (drop_graphs) -- then `slot` was reset; re-take it
        slot->key = key;
        slot->exec = exec;
        slot->launches = m->launches;
        slot->last_use = ++m->fwd_tick;
        TIP_CUDA_TRY(m, cudaGraphLaunch(exec, st));
        return TIP_OK;
    }
    for (int b0 = 0; b0 < B; b0 += CHUNK_WINDOWS) {
        const int nb = std::min(CHUNK_WINDOWS, B - b0);
        const size_t r0 = (size_t)b0 * L;
        int rc = forward_chunk(m, x_imu + r0 * d.n_imu, x_s + r0 * d.size_s, y + r0 * d.size_s, nb, L,
                               keep_mask ? keep_mask + r0 * d.size_s : nullptr, past_scale, dp, st, 0, -1,
                               SEED_CHUNK * (uint64_t)(b0 / CHUNK_WINDOWS));
        if (rc != TIP_OK) return rc;
    }
    return TIP_OK;
}

// window index at which a batch of B windows of L rows can be cut so that the second part starts on a 128-row tile
// boundary, as close to the middle as possible (0: no such cut)
static int split_window(int B, int L) {
    int step = UM_BM;                       // smallest w with (w * L) % 128 == 0
    for (int w = 1; w <= UM_BM; ++w) if ((w * L) % UM_BM == 0) { step = w; break; }
    const int w = (B / 2) / step * step;
    return (w > 0 && w < B) ? w : 0;
}

// forward of the windows [w0, w0 + nw) of the batch staged in d_ximu / d_xs -> d_y, on `st`; replayed from a CUDA graph
// once the same part has been seen twice
static int run_part(tip_model* m, int p, int B, int L, int w0, int nw, const DropP& dp, cudaStream_t st) {
    tip_model::PartGraph& pg = m->part_graph[p];
    const bool same = pg.B == B && pg.L == L && pg.w0 == w0 && pg.nw == nw && pg.p.same_p(dp);
    const bool graphs = m->use_graphs && !m->profile && !getenv("TIP_NO_FWD_GRAPH");
    if (graphs && same && pg.exec) {
        TIP_CUDA_TRY(m, cudaGraphLaunch(pg.exec, st));
        m->launches += pg.launches;
        return TIP_OK;
    }
    if (graphs && same && pg.seen >= 1) {
        cudaStream_t cs;
        TIP_CUDA_TRY(m, cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        cudaGraph_t g = nullptr;
        const int before = m->launches;
        TIP_CUDA_TRY(m, cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        int rc = forward_chunk(m, m->d_ximu, m->d_xs, m->d_y, B, L, nullptr, 1.f, dp, cs, w0, nw);
        cudaError_t ce = cudaStreamEndCapture(cs, &g);
        cudaGraphExec_t exec = nullptr;
        if (rc == TIP_OK && ce == cudaSuccess) ce = cudaGraphInstantiate(&exec, g, 0);
        if (g) cudaGraphDestroy(g);
        cudaStreamDestroy(cs);
        if (rc != TIP_OK) return rc;
        if (ce != cudaSuccess) { m->set_error(std::string("part graph capture: ") + cudaGetErrorString(ce)); return TIP_ERR_CUDA; }
        tip_model::PartGraph& pg2 = m->part_graph[p];          // (drop_graphs may have run inside forward_chunk)
        pg2.exec = exec; pg2.B = B; pg2.L = L; pg2.w0 = w0; pg2.nw = nw; pg2.launches = m->launches - before; pg2.seen = 2; pg2.p = dp;
        TIP_CUDA_TRY(m, cudaGraphLaunch(exec, st));
        return TIP_OK;
    }
    if (!same) { if (pg.exec) cudaGraphExecDestroy(pg.exec); pg = tip_model::PartGraph{}; pg.B = B; pg.L = L; pg.w0 = w0; pg.nw = nw; pg.p = dp; }
    pg.seen = 1;
    return forward_chunk(m, m->d_ximu, m->d_xs, m->d_y, B, L, nullptr, 1.f, dp, st, w0, nw);
}

static bool is_pinned_host(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

extern "C" int tip_forward_host(tip_model* m, const float* x_imu_h, const float* x_s_h, float* y_h, int B,
                                int L, int last_row_only, const tip_dropout* drop, void* stream_) {
    if (!m) return TIP_ERR_INVALID_ARG;
    if (!x_imu_h || !x_s_h || !y_h || B < 1 || L < 1 || L > MAXL) {
        m->set_error("tip_forward_host: bad arguments");
        return TIP_ERR_INVALID_ARG;
    }
    if (!m->sw->packed) { m->set_error("tip_forward_host before tip_pack_weights"); return TIP_ERR_NOT_PACKED; }
    cudaStream_t st = (cudaStream_t)stream_;
    TIP_CUDA_TRY(m, cudaSetDevice(m->device));
    { const int qrc = quiesce_host_jobs(m); if (qrc != TIP_OK) return qrc; }
    { const int prc = order_after_pack(m, st); if (prc != TIP_OK) return prc; }
    const Dims& d = m->d;
    const size_t rows = (size_t)B * L;
    if (rows > m->host_cap) {
        drop_graphs(m);                    // captured forwards / parts have the old staging addresses baked in
        for (float* p : {m->d_ximu, m->d_xs, m->d_y}) if (p) cudaFree(p);
        for (float* p : {m->h_in, m->h_out}) if (p) cudaFreeHost(p);
        m->host_cap = 0;
        const size_t cap = align_up(rows, 64);
        TIP_CUDA_TRY(m, cudaMalloc(&m->d_ximu, cap * d.n_imu * sizeof(float)));
        TIP_CUDA_TRY(m, cudaMalloc(&m->d_xs, cap * d.size_s * sizeof(float)));
        TIP_CUDA_TRY(m, cudaMalloc(&m->d_y, cap * d.size_s * sizeof(float)));
        TIP_CUDA_TRY(m, cudaMallocHost(&m->h_in, cap * d.d_in * sizeof(float)));
        TIP_CUDA_TRY(m, cudaMallocHost(&m->h_out, cap * d.size_s * sizeof(float)));
        m->host_cap = cap;
    }
    // pinned caller buffers are used in place; pageable ones go through the handle's pinned staging
    const float* src_imu = x_imu_h;
    const float* src_s = x_s_h;
    if (!(is_pinned_host(x_imu_h) && is_pinned_host(x_s_h))) {
        float* h_imu = m->h_in;
        float* h_s = m->h_in + rows * d.n_imu;
        memcpy(h_imu, x_imu_h, rows * d.n_imu * sizeof(float));
        memcpy(h_s, x_s_h, rows * d.size_s * sizeof(float));
        src_imu = h_imu; src_s = h_s;
    }
    const bool out_pinned = is_pinned_host(y_h);
    // Two-part pipeline (default; TIP_HOST_PARTS=1 = whole copies): the batch is cut on a 128-row tile boundary; part 1's
    // upload overlaps part 0's forward, the two forwards run on their own streams and part 0's download overlaps part
    // 1's forward.  Same-box A/B at B = 256: 805 vs 816 us per call -- a small win only, because the staggered half
    // forwards are less efficient than one whole-batch forward (the recurrence costs its 90 us per part, the
    // LayerNorm GEMMs fill 40 SMs), which eats most of the 265 us of copies that now overlap.
    static const int host_parts = getenv("TIP_HOST_PARTS") ? atoi(getenv("TIP_HOST_PARTS")) : 2;
    const DropP dp = drop_of(drop);
    const bool umma_engine = (m->engine == 2) || (m->engine == 0 && UMMA_AVAILABLE);
    const int wsplit = (host_parts == 2 && !last_row_only && umma_engine && B >= 64 && B <= CHUNK_WINDOWS)
                           ? split_window(B, L) : 0;
    if (wsplit > 0) {
        if (!m->s_in) {
            TIP_CUDA_TRY(m, cudaStreamCreateWithFlags(&m->s_in, cudaStreamNonBlocking));
            TIP_CUDA_TRY(m, cudaStreamCreateWithFlags(&m->s_out, cudaStreamNonBlocking));
            for (int i = 0; i < 2; ++i) {
                TIP_CUDA_TRY(m, cudaStreamCreateWithFlags(&m->s_part[i], cudaStreamNonBlocking));
                TIP_CUDA_TRY(m, cudaEventCreateWithFlags(&m->ev_in[i], cudaEventDisableTiming));
                TIP_CUDA_TRY(m, cudaEventCreateWithFlags(&m->ev_out[i], cudaEventDisableTiming));
            }
            TIP_CUDA_TRY(m, cudaEventCreateWithFlags(&m->ev_start, cudaEventDisableTiming));
        }
        int rc = ensure_workspace(m, (int)rows);
        if (rc != TIP_OK) return rc;
        const int pw0[2] = {0, wsplit}, pnw[2] = {wsplit, B - wsplit};
        if (dp.any()) seed_set_kernel<<<1, 1, 0, st>>>(m->d_seed, dp.seed);     // both parts draw from this call's seed
        TIP_CUDA_TRY(m, cudaEventRecord(m->ev_start, st));      // earlier work on `st` may still use the staging buffers
        TIP_CUDA_TRY(m, cudaStreamWaitEvent(m->s_in, m->ev_start, 0));
        TIP_CUDA_TRY(m, cudaStreamWaitEvent(m->s_out, m->ev_start, 0));
        float* hdst = out_pinned ? y_h : m->h_out;
        m->launches = 0;
        for (int p2 = 0; p2 < 2; ++p2) {
            const size_t r0 = (size_t)pw0[p2] * L, nr = (size_t)pnw[p2] * L;
            TIP_CUDA_TRY(m, cudaMemcpyAsync(m->d_ximu + r0 * d.n_imu, src_imu + r0 * d.n_imu, nr * d.n_imu * sizeof(float), cudaMemcpyHostToDevice, m->s_in));
            TIP_CUDA_TRY(m, cudaMemcpyAsync(m->d_xs + r0 * d.size_s, src_s + r0 * d.size_s, nr * d.size_s * sizeof(float), cudaMemcpyHostToDevice, m->s_in));
            TIP_CUDA_TRY(m, cudaEventRecord(m->ev_in[p2], m->s_in));
        }
        for (int p2 = 0; p2 < 2; ++p2) {
            const size_t r0 = (size_t)pw0[p2] * L, nr = (size_t)pnw[p2] * L;
            TIP_CUDA_TRY(m, cudaStreamWaitEvent(m->s_part[p2], m->ev_in[p2], 0));
            rc = run_part(m, p2, B, L, pw0[p2], pnw[p2], dp, m->s_part[p2]);
            if (rc != TIP_OK) return rc;
            TIP_CUDA_TRY(m, cudaEventRecord(m->ev_out[p2], m->s_part[p2]));
            TIP_CUDA_TRY(m, cudaStreamWaitEvent(m->s_out, m->ev_out[p2], 0));
            TIP_CUDA_TRY(m, cudaMemcpyAsync(hdst + r0 * d.size_s, m->d_y + r0 * d.size_s, nr * d.size_s * sizeof(float), cudaMemcpyDeviceToHost, m->s_out));
        }
        TIP_CUDA_TRY(m, cudaStreamSynchronize(m->s_out));
        if (!out_pinned) memcpy(y_h, m->h_out, rows * d.size_s * sizeof(float));
        return TIP_OK;
    }
    // small batches / last-row-only / stochastic calls: one copy in per tensor, the forward (a CUDA-graph replay from the
    // second call on: the staging addresses are stable), one copy out.
    TIP_CUDA_TRY(m, cudaMemcpyAsync(m->d_ximu, src_imu, rows * d.n_imu * sizeof(float), cudaMemcpyHostToDevice, st));
    TIP_CUDA_TRY(m, cudaMemcpyAsync(m->d_xs, src_s, rows * d.size_s * sizeof(float), cudaMemcpyHostToDevice, st));
    int rc = tip_forward(m, m->d_ximu, m->d_xs, m->d_y, B, L, nullptr, 1.f, drop, st);
    if (rc != TIP_OK) return rc;
    const size_t out_rows = last_row_only ? (size_t)B : rows;
    const float* dsrc = m->d_y;
    if (last_row_only) {
        launch_k(last_row_kernel, dim3((B * d.size_s + 255) / 256), dim3(256), 0, st, m->d_y, m->d_xs, B, L, d.size_s);
        m->launches++;
        dsrc = m->d_xs;
    }
    float* hdst = out_pinned ? y_h : m->h_out;
    TIP_CUDA_TRY(m, cudaMemcpyAsync(hdst, dsrc, out_rows * d.size_s * sizeof(float), cudaMemcpyDeviceToHost, st));
    TIP_CUDA_TRY(m, cudaStreamSynchronize(st));
    if (!out_pinned) memcpy(y_h, m->h_out, out_rows * d.size_s * sizeof(float));
    return TIP_OK;
}

// ---- job pipeline over host buffers ----------------------------------------------------------------
// tip_forward_host is one blocking call: upload (167 us at B = 256), forward (512 us), download (98 us) in sequence.  A
// caller with many batches (the offline evaluator, a serving loop) does not need step i's result before it hands
// over step i+1, so the three legs of consecutive jobs can overlap: uploads on hs_in, forwards (whole-batch graph
// replays, serialised: they share the workspace) on hs_fwd, downloads on hs_out, each job with its own device staging.
// In steady state the GPU runs forwards back to back and both PCIe directions are busy underneath them.
extern "C" int tip_forward_host_wait(tip_model* m, int slot) {
    if (!m) return TIP_ERR_INVALID_ARG;
    if (slot < 0 || slot >= TIP_HOST_SLOTS) { m->set_error("tip_forward_host_wait: slot out of range"); return TIP_ERR_INVALID_ARG; }
    tip_model::HostSlot& s = m->hslot[slot];
    if (!s.busy) return TIP_OK;
    TIP_CUDA_TRY(m, cudaSetDevice(m->device));
    s.busy = false;
    TIP_CUDA_TRY(m, cudaEventSynchronize(s.ev_out));
    return TIP_OK;
}

extern "C" int tip_forward_host_submit(tip_model* m, int slot, const float* x_imu_h, const float* x_s_h, float* y_h,
                                       int B, int L, int last_row_only, const tip_dropout* drop) {
    if (!m) return TIP_ERR_INVALID_ARG;
    if (slot < 0 || slot >= TIP_HOST_SLOTS || !x_imu_h || !x_s_h || !y_h || B < 1 || L < 1 || L > MAXL) {
        m->set_error("tip_forward_host_submit: need 0 <= slot < TIP_HOST_SLOTS, non-null buffers, B >= 1 and 1 <= L <= 40");
        return TIP_ERR_INVALID_ARG;
    }
    if (!m->sw->packed) { m->set_error("tip_forward_host_submit before tip_pack_weights"); return TIP_ERR_NOT_PACKED; }
    TIP_CUDA_TRY(m, cudaSetDevice(m->device));
    if (!(is_pinned_host(x_imu_h) && is_pinned_host(x_s_h) && is_pinned_host(y_h))) {
        // a pageable buffer would turn every cudaMemcpyAsync into a staged, host-blocking copy: no overlap, and the
        // caller could not tell when its buffer is free again
        m->set_error("tip_forward_host_submit: the three host buffers must be page-locked (cudaHostAlloc / cudaHostRegister / "
                     "tensor.pin_memory()); use tip_forward_host for pageable memory");
        return TIP_ERR_INVALID_ARG;
    }
    int rc = tip_forward_host_wait(m, slot);            // re-using a slot waits for its previous job
    if (rc != TIP_OK) return rc;
    const Dims& d = m->d;
    tip_model::HostSlot& s = m->hslot[slot];
    if (!m->hs_in) {
        TIP_CUDA_TRY(m, cudaStreamCreateWithFlags(&m->hs_in, cudaStreamNonBlocking));
        TIP_CUDA_TRY(m, cudaStreamCreateWithFlags(&m->hs_fwd, cudaStreamNonBlocking));
        TIP_CUDA_TRY(m, cudaStreamCreateWithFlags(&m->hs_out, cudaStreamNonBlocking));
        TIP_CUDA_TRY(m, cudaEventCreateWithFlags(&m->ev_user, cudaEventDisableTiming));
        TIP_CUDA_TRY(m, cudaDeviceSynchronize());      // once: forwards queued on caller streams before the pipeline existed
    }
    if (!s.ev_in) {
        TIP_CUDA_TRY(m, cudaEventCreateWithFlags(&s.ev_in, cudaEventDisableTiming));
        TIP_CUDA_TRY(m, cudaEventCreateWithFlags(&s.ev_fwd, cudaEventDisableTiming));
        TIP_CUDA_TRY(m, cudaEventCreateWithFlags(&s.ev_out, cudaEventDisableTiming));
    }
    const size_t rows = (size_t)B * L;
    if (rows > s.cap) {
        for (float** p : {&s.d_ximu, &s.d_xs, &s.d_y}) if (*p) { cudaFree(*p); *p = nullptr; }
        s.cap = 0;
        const size_t cap = align_up(rows, 64);
        TIP_CUDA_TRY(m, cudaMalloc(&s.d_ximu, cap * d.n_imu * sizeof(float)));
        TIP_CUDA_TRY(m, cudaMalloc(&s.d_xs, cap * d.size_s * sizeof(float)));
        TIP_CUDA_TRY(m, cudaMalloc(&s.d_y, cap * d.size_s * sizeof(float)));
        s.cap = cap;
    }
    TIP_CUDA_TRY(m, cudaMemcpyAsync(s.d_ximu, x_imu_h, rows * d.n_imu * sizeof(float), cudaMemcpyHostToDevice, m->hs_in));
    TIP_CUDA_TRY(m, cudaMemcpyAsync(s.d_xs, x_s_h, rows * d.size_s * sizeof(float), cudaMemcpyHostToDevice, m->hs_in));
    TIP_CUDA_TRY(m, cudaEventRecord(s.ev_in, m->hs_in));
    TIP_CUDA_TRY(m, cudaStreamWaitEvent(m->hs_fwd, s.ev_in, 0));
    if (m->ev_user_set) { TIP_CUDA_TRY(m, cudaStreamWaitEvent(m->hs_fwd, m->ev_user, 0)); m->ev_user_set = false; }
    rc = order_after_pack(m, m->hs_fwd);               // a (re-)pack queued on a caller stream since the last job
    if (rc != TIP_OK) return rc;
    // graph replay from the slot's third job on (the captured forwards are keyed on the staging addresses)
    rc = forward_impl(m, s.d_ximu, s.d_xs, s.d_y, B, L, nullptr, 1.f, drop, m->hs_fwd);
    if (rc != TIP_OK) return rc;
    const float* dsrc = s.d_y;
    size_t out_rows = rows;
    if (last_row_only) {
        launch_k(last_row_kernel, dim3((B * d.size_s + 255) / 256), dim3(256), 0, m->hs_fwd, s.d_y, s.d_xs, B, L, d.size_s);
        m->launches++;
        dsrc = s.d_xs;
        out_rows = (size_t)B;
    }
    s.launches = m->launches;
    TIP_CUDA_TRY(m, cudaEventRecord(s.ev_fwd, m->hs_fwd));
    TIP_CUDA_TRY(m, cudaStreamWaitEvent(m->hs_out, s.ev_fwd, 0));
    TIP_CUDA_TRY(m, cudaMemcpyAsync(y_h, dsrc, out_rows * d.size_s * sizeof(float), cudaMemcpyDeviceToHost, m->hs_out));
    TIP_CUDA_TRY(m, cudaEventRecord(s.ev_out, m->hs_out));
    s.busy = true;
    return TIP_OK;
}

// ------------------------------------------------------------------------------------------------
extern "C" int tip_stream_reset(tip_model* m, int n_streams) {
    if (!m || n_streams < 1) return TIP_ERR_INVALID_ARG;
    TIP_CUDA_TRY(m, cudaSetDevice(m->device));
    free_stream_state(m);
    const Dims& d = m->d;
    const size_t S = n_streams;
    TIP_CUDA_TRY(m, cudaMalloc(&m->win_imu, S * MAXL * d.n_imu * sizeof(float)));
    TIP_CUDA_TRY(m, cudaMalloc(&m->win_s, S * MAXL * d.size_s * sizeof(float)));
    TIP_CUDA_TRY(m, cudaMalloc(&m->st_rows, S * d.d_in * sizeof(float)));
    TIP_CUDA_TRY(m, cudaMalloc(&m->st_ximu, S * MAXL * d.n_imu * sizeof(float)));
    TIP_CUDA_TRY(m, cudaMalloc(&m->st_xs, S * MAXL * d.size_s * sizeof(float)));
    TIP_CUDA_TRY(m, cudaMalloc(&m->st_y, S * MAXL * d.size_s * sizeof(float)));
    TIP_CUDA_TRY(m, cudaMalloc(&m->st_ylast, S * d.size_s * sizeof(float)));
    TIP_CUDA_TRY(m, cudaMallocHost(&m->h_rows, S * d.d_in * sizeof(float)));
    TIP_CUDA_TRY(m, cudaMallocHost(&m->h_ylast, S * d.size_s * sizeof(float)));
    TIP_CUDA_TRY(m, cudaMalloc(&m->raw_ring, S * IMU_RING * IMU_RAW * sizeof(float)));
    TIP_CUDA_TRY(m, cudaMalloc(&m->st_raw, S * IMU_RAW * sizeof(float)));
    TIP_CUDA_TRY(m, cudaMalloc(&m->acc_ring, S * ACC_WIN * 18 * sizeof(double)));
    TIP_CUDA_TRY(m, cudaMallocHost(&m->h_raw, S * IMU_RAW * sizeof(float)));
    const size_t out_w = 60 + (d.size_s - 111);
    TIP_CUDA_TRY(m, cudaMalloc(&m->pp_ring, S * PP_TAPS * d.size_s * sizeof(double)));
    TIP_CUDA_TRY(m, cudaMalloc(&m->pp_last, S * PP_TAIL * sizeof(double)));
    TIP_CUDA_TRY(m, cudaMalloc(&m->pp_out, S * out_w * sizeof(double)));
    TIP_CUDA_TRY(m, cudaMallocHost(&m->h_pp_out, S * out_w * sizeof(double)));
    TIP_CUDA_TRY(m, cudaMemset(m->win_imu, 0, S * MAXL * d.n_imu * sizeof(float)));
    TIP_CUDA_TRY(m, cudaMemset(m->win_s, 0, S * MAXL * d.size_s * sizeof(float)));
    m->n_streams = n_streams;
    m->stream_len = 0;
    return TIP_OK;
}

extern "C" int tip_stream_length(const tip_model* m) { return m ? m->stream_len : 0; }

// device work of one streaming step; inputs already in st_rows (imu rows then s rows)
static int stream_step_device(tip_model* m, const tip_dropout* drop, cudaStream_t st, int len_before) {
    const Dims& d = m->d;
    const int S = m->n_streams;
    pdl_rows() = S * MAXL;
    const float* imu_rows = m->st_rows;
    const float* s_rows = m->st_rows + (size_t)S * d.n_imu;
    pdl_kind() = 8;
    launch_k(window_push_kernel, dim3(S, 2), dim3(256), 0, st, m->win_imu, imu_rows, d.n_imu, m->win_s, s_rows, d.size_s, len_before);
    m->launches += 1;
    const int L = std::min(len_before + 1, MAXL);
    const float *xi = m->win_imu, *xs = m->win_s;
    int extra = 1;
    if (L < MAXL) {
        const int64_t ti = (int64_t)S * L * d.n_imu, ts = (int64_t)S * L * d.size_s;
        launch_k(window_compact_kernel, dim3((unsigned)std::min<int64_t>((ti + 255) / 256, 1184)), dim3(256), 0, st, m->win_imu, m->st_ximu, S, L, d.n_imu);
        launch_k(window_compact_kernel, dim3((unsigned)std::min<int64_t>((ts + 255) / 256, 1184)), dim3(256), 0, st, m->win_s, m->st_xs, S, L, d.size_s);
        xi = m->st_ximu; xs = m->st_xs;
        extra += 2;
    }
    int rc = tip_forward(m, xi, xs, m->st_y, S, L, nullptr, 1.f, drop, st);   // resets m->launches
    if (rc != TIP_OK) return rc;
    launch_k(last_row_kernel, dim3((S * d.size_s + 255) / 256), dim3(256), 0, st, m->st_y, m->st_ylast, S, L, d.size_s);
    m->launches += extra + 1;
    return TIP_OK;
}

// device + completion part of one streaming step; st_rows already holds (imu rows | s rows)
static int stream_step_core(tip_model* m, float* y_last, int rows_on_host, const tip_dropout* drop, cudaStream_t st) {
    const Dims& d = m->d;
    const size_t S = m->n_streams;
    const size_t n_s = S * d.size_s;
    const int len_before = m->stream_len;
    const DropP dp = drop_of(drop);
    const bool steady = (len_before == MAXL);
    int rc = order_after_pack(m, st);
    if (rc != TIP_OK) return rc;
    if (steady && m->use_graphs) {
        // steady state: the whole frame (2 window shifts + forward + last-row gather) is one graph launch; dropout rates
        // are baked in, the seed is read from device memory (set right before the launch)
        if (m->st_graph && !m->st_graph_p.same_p(dp)) { cudaGraphExecDestroy(m->st_graph); m->st_graph = nullptr; }
        if (!m->st_graph) {
            rc = ensure_workspace(m, (int)S * MAXL);
            if (rc != TIP_OK) return rc;
            if ((m->engine == 2 || (m->engine == 0 && UMMA_AVAILABLE)) && !m->maps_ready) {
                // descriptors must exist before capture (their creation is host work)
                rc = umma_build_maps(m->maps, m->blob, m->off, m->d, m->xin, m->plane_xin, m->xa, m->xb, m->att,
                                     m->plane_e, m->hid, m->plane_f, m->hs, m->plane_r, m->qkv, m->gi, m->pre, m->cap_rows, m->err);
                if (rc != TIP_OK) return rc;
                m->maps_ready = true;
            }
            cudaStream_t cs;
            TIP_CUDA_TRY(m, cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
            cudaGraph_t g = nullptr;
            TIP_CUDA_TRY(m, cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
            rc = stream_step_device(m, drop, cs, MAXL);
            cudaError_t ce = cudaStreamEndCapture(cs, &g);
            if (rc == TIP_OK && ce == cudaSuccess) {
                ce = cudaGraphInstantiate(&m->st_graph, g, 0);
                m->st_graph_launches = m->launches;
                m->st_graph_p = dp;
            }
            if (g) cudaGraphDestroy(g);
            cudaStreamDestroy(cs);
            if (rc != TIP_OK) return rc;
            if (ce != cudaSuccess) { m->set_error(std::string("graph capture: ") + cudaGetErrorString(ce)); return TIP_ERR_CUDA; }
        }
        if (dp.any()) seed_set_kernel<<<1, 1, 0, st>>>(m->d_seed, dp.seed);
        TIP_CUDA_TRY(m, cudaGraphLaunch(m->st_graph, st));
        m->launches = m->st_graph_launches;
    } else {
        rc = stream_step_device(m, drop, st, len_before);
        if (rc != TIP_OK) return rc;
    }
    m->stream_len = std::min(len_before + 1, MAXL);
    if (!y_last) return TIP_OK;               // closed loop: the caller consumes st_ylast on the device
    if (rows_on_host) {
        TIP_CUDA_TRY(m, cudaMemcpyAsync(m->h_ylast, m->st_ylast, n_s * sizeof(float), cudaMemcpyDeviceToHost, st));
        TIP_CUDA_TRY(m, cudaStreamSynchronize(st));
        memcpy(y_last, m->h_ylast, n_s * sizeof(float));
    } else {
        TIP_CUDA_TRY(m, cudaMemcpyAsync(y_last, m->st_ylast, n_s * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    return TIP_OK;
}

extern "C" int tip_stream_step(tip_model* m, const float* imu_row, const float* s_row, float* y_last,
                               int rows_on_host, const tip_dropout* drop, void* stream_) {
    if (!m || !imu_row || !s_row || !y_last) return TIP_ERR_INVALID_ARG;
    if (!m->sw->packed) { m->set_error("tip_stream_step before tip_pack_weights"); return TIP_ERR_NOT_PACKED; }
    if (m->n_streams < 1) { m->set_error("tip_stream_step before tip_stream_reset"); return TIP_ERR_INVALID_ARG; }
    cudaStream_t st = (cudaStream_t)stream_;
    TIP_CUDA_TRY(m, cudaSetDevice(m->device));
    const Dims& d = m->d;
    const size_t S = m->n_streams;
    const size_t n_i = S * d.n_imu, n_s = S * d.size_s;
    if (rows_on_host) {
        memcpy(m->h_rows, imu_row, n_i * sizeof(float));
        memcpy(m->h_rows + n_i, s_row, n_s * sizeof(float));
        TIP_CUDA_TRY(m, cudaMemcpyAsync(m->st_rows, m->h_rows, (n_i + n_s) * sizeof(float), cudaMemcpyHostToDevice, st));
    } else {
        TIP_CUDA_TRY(m, cudaMemcpyAsync(m->st_rows, imu_row, n_i * sizeof(float), cudaMemcpyDeviceToDevice, st));
        TIP_CUDA_TRY(m, cudaMemcpyAsync(m->st_rows + n_i, s_row, n_s * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    return stream_step_core(m, y_last, rows_on_host, drop, st);
}

// Row N1: push RAW IMU frames (S, 72) -- the runner's record_raw_imu + window features run on the device.
// *produced = 0 during the first 5 calls (no smoothed frame yet, the runner returns s_init); then 1.
extern "C" int tip_stream_step_raw(tip_model* m, const float* raw_imu, const float* s_row, float* y_last,
                                   int rows_on_host, const tip_dropout* drop, void* stream_, int* produced) {
    if (!m || !raw_imu || !s_row || !y_last || !produced) return TIP_ERR_INVALID_ARG;
    if (!m->sw->packed) { m->set_error("tip_stream_step_raw before tip_pack_weights"); return TIP_ERR_NOT_PACKED; }
    if (m->n_streams < 1) { m->set_error("tip_stream_step_raw before tip_stream_reset"); return TIP_ERR_INVALID_ARG; }
    cudaStream_t st = (cudaStream_t)stream_;
    TIP_CUDA_TRY(m, cudaSetDevice(m->device));
    const Dims& d = m->d;
    const size_t S = m->n_streams;
    const size_t n_i = S * d.n_imu, n_s = S * d.size_s, n_r = S * IMU_RAW;
    if (rows_on_host) {
        memcpy(m->h_raw, raw_imu, n_r * sizeof(float));
        memcpy(m->h_rows + n_i, s_row, n_s * sizeof(float));
        TIP_CUDA_TRY(m, cudaMemcpyAsync(m->st_raw, m->h_raw, n_r * sizeof(float), cudaMemcpyHostToDevice, st));
        TIP_CUDA_TRY(m, cudaMemcpyAsync(m->st_rows + n_i, m->h_rows + n_i, n_s * sizeof(float), cudaMemcpyHostToDevice, st));
    } else {
        TIP_CUDA_TRY(m, cudaMemcpyAsync(m->st_raw, raw_imu, n_r * sizeof(float), cudaMemcpyDeviceToDevice, st));
        TIP_CUDA_TRY(m, cudaMemcpyAsync(m->st_rows + n_i, s_row, n_s * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    launch_k(imu_push_kernel, dim3((unsigned)S), dim3(32), 0, st, m->st_raw, m->raw_ring, m->acc_ring, m->st_rows, d.n_imu, m->n_raw, m->n_rows);
    TIP_CUDA_TRY(m, cudaGetLastError());
    m->n_raw += (m->n_raw == 0) ? IMU_DELAY + 1 : 1;
    if (m->n_raw < IMU_RING) {
        *produced = 0;
        if (rows_on_host) TIP_CUDA_TRY(m, cudaStreamSynchronize(st));
        return TIP_OK;
    }
    m->n_rows += 1;
    *produced = 1;
    return stream_step_core(m, y_last, rows_on_host, drop, st);
}

// ------------------------------------------------------------------------------------------------
// Row N3: closed loop -- post-model step on the device, state row fed back without leaving the GPU.
extern "C" int tip_stream_state_width(const tip_model* m) { return m ? 60 + (m->d.size_s - 111) : 0; }

extern "C" int tip_stream_set_state(tip_model* m, const float* s_row0, int rows_on_host, void* stream_) {
    if (!m || !s_row0) return TIP_ERR_INVALID_ARG;
    if (m->n_streams < 1) { m->set_error("tip_stream_set_state before tip_stream_reset"); return TIP_ERR_INVALID_ARG; }
    cudaStream_t st = (cudaStream_t)stream_;
    TIP_CUDA_TRY(m, cudaSetDevice(m->device));
    const size_t bytes = (size_t)m->n_streams * m->d.size_s * sizeof(float);
    if (rows_on_host) {
        memcpy(m->h_rows, s_row0, bytes);
        TIP_CUDA_TRY(m, cudaMemcpyAsync(m->st_rows + (size_t)m->n_streams * m->d.n_imu, m->h_rows, bytes, cudaMemcpyHostToDevice, st));
        TIP_CUDA_TRY(m, cudaStreamSynchronize(st));
    } else {
        TIP_CUDA_TRY(m, cudaMemcpyAsync(m->st_rows + (size_t)m->n_streams * m->d.n_imu, s_row0, bytes, cudaMemcpyDeviceToDevice, st));
    }
    m->fb_set = true;
    return TIP_OK;
}

extern "C" int tip_stream_step_closed(tip_model* m, const float* raw_imu, const float* y_override, double* state_out,
                                      int rows_on_host, const tip_dropout* drop, void* stream_, int* produced) {
    if (!m || !raw_imu || !state_out || !produced) return TIP_ERR_INVALID_ARG;
    if (!m->sw->packed) { m->set_error("tip_stream_step_closed before tip_pack_weights"); return TIP_ERR_NOT_PACKED; }
    if (m->n_streams < 1 || !m->fb_set) {
        m->set_error("tip_stream_step_closed needs tip_stream_reset and tip_stream_set_state first");
        return TIP_ERR_INVALID_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream_;
    TIP_CUDA_TRY(m, cudaSetDevice(m->device));
    const Dims& d = m->d;
    const size_t S = m->n_streams;
    const size_t n_i = S * d.n_imu, n_s = S * d.size_s, n_r = S * IMU_RAW;
    const size_t out_w = 60 + (d.size_s - 111);
    if (rows_on_host) {
        memcpy(m->h_raw, raw_imu, n_r * sizeof(float));
        TIP_CUDA_TRY(m, cudaMemcpyAsync(m->st_raw, m->h_raw, n_r * sizeof(float), cudaMemcpyHostToDevice, st));
    } else {
        TIP_CUDA_TRY(m, cudaMemcpyAsync(m->st_raw, raw_imu, n_r * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    // the x_s row of this call is the one the previous post step (or tip_stream_set_state) left in the staging rows
    launch_k(imu_push_kernel, dim3((unsigned)S), dim3(32), 0, st, m->st_raw, m->raw_ring, m->acc_ring, m->st_rows, d.n_imu, m->n_raw, m->n_rows);
    TIP_CUDA_TRY(m, cudaGetLastError());
    m->n_raw += (m->n_raw == 0) ? IMU_DELAY + 1 : 1;
    if (m->n_raw < IMU_RING) {
        *produced = 0;
        if (rows_on_host) TIP_CUDA_TRY(m, cudaStreamSynchronize(st));
        return TIP_OK;
    }
    m->n_rows += 1;
    *produced = 1;
    int rc = stream_step_core(m, nullptr, rows_on_host, drop, st);
    if (rc != TIP_OK) return rc;
    const float* ysrc = m->st_ylast;
    if (y_override) {                       // teacher forcing (parity tests): st_y is free after the last-row gather
        TIP_CUDA_TRY(m, cudaMemcpyAsync(m->st_y, y_override, n_s * sizeof(float),
                                        rows_on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st));
        ysrc = m->st_y;
    }
    launch_k(post_step_kernel, dim3((unsigned)S), dim3(32), 0, st, ysrc, m->st_rows, d.n_imu, m->pp_ring, m->pp_last, m->st_rows + n_i, m->pp_out,
                                                 d.size_s, m->n_post);
    TIP_CUDA_TRY(m, cudaGetLastError());
    m->launches += 2;                       // imu_push + post_step
    m->n_post += 1;
    if (rows_on_host) {
        TIP_CUDA_TRY(m, cudaMemcpyAsync(m->h_pp_out, m->pp_out, S * out_w * sizeof(double), cudaMemcpyDeviceToHost, st));
        TIP_CUDA_TRY(m, cudaStreamSynchronize(st));
        memcpy(state_out, m->h_pp_out, S * out_w * sizeof(double));
    } else {
        TIP_CUDA_TRY(m, cudaMemcpyAsync(state_out, m->pp_out, S * out_w * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    return TIP_OK;
}
