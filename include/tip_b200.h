/*
 * tip_b200.h -- C ABI of the B200-native Transformer-Inertial-Poser hot path.
 *
 * The reference has no FFI / plugin boundary: its boundary is the duck-typed
 * torch.nn.Module `TF_RNN_Past_State` imported by module name
 *   (/root/reference/offline_testing_simple.py:80, live_demo_new.py:16)
 * and called once per frame as `model(x_imu.cuda(), x_s.cuda()).cpu()`
 *   (/root/reference/real_time_runner_minimal.py:149, real_time_runner.py:431).
 * This header is what a ctypes binding for that call binds instead
 * (see INTEGRATION.md); the Python mirror of the module lives in
 * transformer-inertial-poser_b200/simple_transformer_with_state.py.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no C++ / torch types.
 *   - `*_dev` pointers are DEVICE pointers (tensor.data_ptr()); `*_host` are host
 *     pointers.  `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *   - every call returns a tip_status; tip_last_error(handle) gives the message.
 *     Nothing throws, exits or synchronises the device unless documented.
 *   - one handle = one model replica on one GPU (the current device at tip_create).
 *     Calls on one handle must not overlap (the reference caller is single-threaded).
 *   - all tensors are fp32, row-major, contiguous.
 */
#ifndef TIP_B200_H_
#define TIP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TIP_ABI_VERSION 2

typedef enum tip_status {
    TIP_OK = 0,
    TIP_ERR_INVALID_ARG = 1,   /* bad shape / null pointer / unsupported hyper-parameter */
    TIP_ERR_NOT_PACKED = 2,    /* forward before tip_pack_weights */
    TIP_ERR_CUDA = 3,          /* a CUDA runtime call failed (message has the cudaError string) */
    TIP_ERR_NO_DEVICE = 4,     /* no sm_100 device available */
    TIP_ERR_OOM = 5
} tip_status;

/* Constructor arguments of TF_RNN_Past_State
 * (/root/reference/simple_transformer_with_state.py:9-17).  The kernels are specialised for the
 * one architecture the reference ships and constructs (offline_testing_simple.py:87-95,
 * live_demo_new.py:202-210): tf_in_dim=256, n_heads=16, tf_hid_size=1024, rnn_hid_size=512,
 * tf_layers<=8; input_size_imu=72, 111<=size_s<=144, with_rnn and with_acc_sum free.
 * Anything else makes tip_create return TIP_ERR_INVALID_ARG. */
typedef struct tip_dims {
    int32_t input_size_imu;   /* 72 = 6 IMUs x (9 rotation + 3 acceleration) */
    int32_t size_s;           /* 131 (5 SBPs) or 119 (2 SBPs) */
    int32_t rnn_hid_size;     /* 512 */
    int32_t tf_hid_size;      /* 1024 */
    int32_t tf_in_dim;        /* 256 */
    int32_t n_heads;          /* 16 */
    int32_t tf_layers;        /* 4 */
    int32_t with_rnn;         /* 1: tanh RNN + Linear(rnn_hid->size_s); 0: Linear(tf_in_dim->size_s) */
    int32_t with_acc_sum;     /* 1: x_imu carries 18 extra acc-sum columns (90 wide) */
} tip_dims;

/* Per-call stochastic behaviour of the reference forward (lines :73, :77 and the
 * nn.TransformerEncoderLayer dropouts).  All-zero = the deterministic parity mode of
 * SURVEY.md section 8c (eval(), past_state_dropout = 0). */
typedef struct tip_dropout {
    float    in_dropout;          /* p on x_imu            (reference :73; 0 in every shipped config) */
    float    past_state_dropout;  /* p on x_s              (reference :77; ALWAYS live there, 0.8 shipped) */
    float    encoder_dropout;     /* p inside the encoder  (0.1 when the module is in train(), else 0) */
    uint64_t seed;                /* counter-based RNG seed for this call (ignored when all p are 0) */
} tip_dropout;
/* The masks are a pure function of (seed, dropout site, element index): one 64-bit splitmix hash gives four
 * 16-bit uniforms, element idx is dropped when lane (idx & 3) of hash(seed + site, idx >> 2) < round(p * 65536),
 * kept elements are scaled by 1/(1-p).  Sites and element indices (rows are b*L + t, batch-global):
 *   x_imu (:73)  site 0x1111, idx = row * kin_pad + c            (c = column of the concatenated input row,
 *   x_s   (:77)  site 0x2222, idx = row * kin_pad + n_imu + c     kin_pad = (n_imu + size_s) rounded up to 64)
 *   layer l (0-based): attention probabilities 101*(l+1), idx = 4*group + lane with bh = b*16 + h,
 *     group = ((bh*3 + query/16)*8 + query%8)*20 + key/2, lane = 2*((query%16)/8) + key%2 (the four probabilities one
 *     lane of the tensor-core kernel holds -- rows r, r+8 of a 16-row tile x two adjacent keys -- share one hash);
 *   dropout1 (on out_proj, before the residual) 211*(l+1), idx = row*256 + col;  FFN inner dropout (after ReLU)
 *   307*(l+1), idx = row*1024 + col;  dropout2 (on linear2) 401*(l+1), idx = row*256 + col.
 * Batches of more than 1024 windows are processed in chunks of 1024; chunk k adds k * 0x632BE59BD9B4E019 to the
 * seed and its rows / windows restart at 0.  oracle/tip_oracle.py restates the generator, so the stochastic mode
 * is parity-tested mask for mask.  The seed is kept in device memory: stochastic forwards replay from CUDA graphs. */

typedef struct tip_model tip_model;   /* opaque */

/* ---- lifetime ------------------------------------------------------------------------------ */
int  tip_abi_version(void);
/* Mirrors TF_RNN_Past_State.__init__ (:9-54). Binds to the current CUDA device. */
int  tip_create(const tip_dims* dims, tip_model** out);
void tip_destroy(tip_model* m);
/* An execution lane: a second handle on the owner's device that SHARES the owner's packed weights (no copy, no
 * second pack) and owns everything else (workspace, tensor maps, captured graphs, job slots, seed).  Forwards of
 * different handles may overlap on different streams -- one forward at a time leaves SMs idle in its narrow
 * phases.  tip_pack_weights is called on the owner only; it waits for the lanes' in-flight forwards, and every
 * lane orders its next forward after the pack.  The weights live until the last handle sharing them is destroyed.
 * (The reference has no counterpart: it runs one blocking call at a time, offline_testing_simple.py:360-399.) */
int  tip_create_lane(tip_model* owner, tip_model** out);
const char* tip_last_error(const tip_model* m);   /* m may be NULL: last error of tip_create */

/* ---- weights -------------------------------------------------------------------------------- */
/* Number of state-dict tensors the model expects: 2 + 12*tf_layers + (with_rnn ? 4 : 0) + 2
 * (56 for the shipped checkpoints), in the reference's state_dict() order. */
int  tip_num_weight_tensors(const tip_model* m);
/* Mirrors load_state_dict + .cuda() (offline_testing_simple.py:96-97): takes DEVICE pointers to
 * the fp32 tensors in state-dict order, with their element counts (checked against the expected
 * shapes), and builds the private packed copy (head-permutation folded into in_linear rows,
 * root-velocity columns zeroed, 1/sqrt(d) folded into W_q/b_q, RNN biases pre-summed, FP16
 * hi/lo splits).  Asynchronous on `stream`; the source tensors may be freed after the stream
 * has passed this point.  Forwards queued later on ANY stream of this handle or its lanes run after the
 * pack (event); when lanes or the handle's internal streams exist, the call first waits for the device to
 * go idle (their in-flight forwards still read the old weights). */
int  tip_pack_weights(tip_model* m, const float* const* tensors_dev, const int64_t* numels,
                      int n_tensors, void* stream);

/* ---- the hot path --------------------------------------------------------------------------- */
/* Mirrors TF_RNN_Past_State.forward (:60-102):
 *   x_imu_dev (B, L, input_size_imu [+18]), x_s_dev (B, L, size_s) -> y_dev (B, L, size_s).
 * 1 <= L <= 40 (the runner's max_input_l), B >= 1.  Inputs are not modified (reference clones).
 * keep_mask_dev: optional (B, L, size_s) 0/1 mask; when non-NULL x_s is multiplied by
 * keep_mask * past_scale INSTEAD of drawing the past_state_dropout mask (deterministic test of :77).
 * drop may be NULL (= deterministic mode).  Asynchronous on `stream`. */
int  tip_forward(tip_model* m, const float* x_imu_dev, const float* x_s_dev, float* y_dev,
                 int B, int L, const float* keep_mask_dev, float past_scale,
                 const tip_dropout* drop, void* stream);

/* Same call with HOST buffers: H2D of both inputs, forward, D2H of y, and a stream synchronise
 * -- i.e. exactly `model(x_imu.cuda(), x_s.cuda()).cpu()` (real_time_runner_minimal.py:149).
 * If last_row_only != 0, y_host is (B, size_s) = y[:, L-1, :] (what :150 consumes). */
int  tip_forward_host(tip_model* m, const float* x_imu_host, const float* x_s_host, float* y_host,
                      int B, int L, int last_row_only, const tip_dropout* drop, void* stream);

/* The same job split into submit + wait, for callers that have the next batch ready before they need
 * the previous result (offline_testing_simple.py:360-399 walks recorded motions whose inputs are all
 * known up front; the reference runs them one blocking call at a time).  `slot` (0 <=
 * slot < TIP_HOST_SLOTS) names one of the handle's job slots, each with its own device staging:
 * submit queues upload -> forward -> download on the handle's three internal streams and returns at
 * once; uploads, forwards and downloads of different slots overlap (forwards themselves run one
 * after another: they share the workspace).  wait(slot) blocks until y_host of that slot's job is
 * complete; submitting to a busy slot waits for it first.  The three host buffers MUST be
 * page-locked (TIP_ERR_INVALID_ARG otherwise) and must stay untouched until wait returns.  The
 * jobs share the handle's workspace with every other entry point: tip_forward / tip_forward_host /
 * tip_stream_* / tip_pack_weights on the same handle first wait (host side) for the submitted
 * jobs' forwards, and a job submitted after a tip_forward on a caller stream runs after it.  For
 * forwards that overlap each other use one handle per lane (the Python side: make_lane()). */
#define TIP_HOST_SLOTS 4
int  tip_forward_host_submit(tip_model* m, int slot, const float* x_imu_host, const float* x_s_host,
                             float* y_host, int B, int L, int last_row_only, const tip_dropout* drop);
int  tip_forward_host_wait(tip_model* m, int slot);

/* ---- streaming (row a10: the window the runner rebuilds every frame) ------------------------- */
/* Device-resident sliding windows for `n_streams` independent IMU streams
 * (replaces real_time_runner_minimal.py:131-147's per-frame re-assembly of the last <=40 rows).
 * Re-creating resets all streams. */
int  tip_stream_reset(tip_model* m, int n_streams);
/* Push one new row per stream (imu_row (S, 72|90), s_row (S, size_s); host or device pointers per
 * `rows_on_host`), slide each window by one row once it holds 40 (coalesced in-place shift),
 * run the forward on the current L = min(#rows, 40) and return y[:, L-1, :] as (S, size_s).
 * With rows_on_host != 0 the call copies in/out and synchronises `stream` before returning. */
int  tip_stream_step(tip_model* m, const float* imu_row, const float* s_row, float* y_last,
                     int rows_on_host, const tip_dropout* drop, void* stream);
int  tip_stream_length(const tip_model* m);   /* current L (0 before the first push) */
/* Row N1: the same step fed with RAW IMU frames (S, 72) = 6 global rotations (54) + 6 global
 * accelerations (18) as RTRunnerMin.step receives them (real_time_runner_minimal.py:118).  The runner's
 * record_raw_imu (:59-76: 5-frame rotation delay, 11-frame acceleration mean), imu_rotate_to_local
 * (data_utils.py:190-219) and the 40-frame acc-sum / 15 feature (:134-141) run on the device; only the
 * newest window row is computed.  *produced is 0 for the first 5 calls (the runner returns s_init then,
 * :125-128; y_last is untouched) and 1 afterwards. */
int  tip_stream_step_raw(tip_model* m, const float* raw_imu, const float* s_row, float* y_last,
                         int rows_on_host, const tip_dropout* drop, void* stream, int* produced);

/* Row N3: closed-loop streaming.  The model-visible part of RTRunnerMin.step AFTER the model
 * call also runs on the device: the 6-tap 0.6^k output filter and SBP split (real_time_runner_minimal.py:
 * 87-112), 2-axis -> axis-angle (data_utils.py:164-179), the root rotation taken from the IMU (:161-163),
 * the averaging with the previous state (:165-167) and record_state_aa_and_c (:78-85, :196), whose row is
 * fed back as the next x_s row without leaving the GPU.  Per frame the caller pushes one raw IMU row and
 * reads back the pose.  PyBullet FK and the SBP root-translation correction (:169-194) stay on the CPU:
 * they only produce s_t[0:3], which the model never sees.
 *
 * tip_stream_set_state: the first x_s row of every stream, (S, size_s) fp32 = record_state_aa_and_c(s_init,
 * zeros) as the runner's constructor appends it (:47); call after tip_stream_reset.
 * tip_stream_state_width: W = 60 + n_c doubles per stream, n_c = size_s - 111:
 *   state[0:57]       = s_t[3:60] (root axis-angle, 17 joint axis-angles in Nimble order, root velocity as
 *                       stored by :158/:166, i.e. averaged with the previous frame's),
 *   state[57:57+n_c]  = c_t (per SBP: flag in {0,1}, offset xyz in metres),
 *   state[57+n_c:W]   = root_v, the filtered but not yet averaged root velocity that :159 integrates.
 * tip_stream_step_closed: raw_imu (S, 72) fp32 as tip_stream_step_raw; state_out (S, W) float64 (host or
 * device per rows_on_host).  *produced = 0 during the runner's 5 warm-up calls (state_out untouched).
 * y_override (optional, same residence as raw_imu): (S, size_s) fp32 used INSTEAD of the model's last
 * output row by the post step -- teacher forcing, the hook the parity tests use to pin the post step
 * against a trace of the reference runner independently of the model's own fp32 error. */
int  tip_stream_set_state(tip_model* m, const float* s_row0, int rows_on_host, void* stream);
int  tip_stream_state_width(const tip_model* m);
int  tip_stream_step_closed(tip_model* m, const float* raw_imu, const float* y_override, double* state_out,
                            int rows_on_host, const tip_dropout* drop, void* stream, int* produced);

/* ---- introspection -------------------------------------------------------------------------- */
/* Algorithmic bytes / flops of one forward:
 *   bytes = weight_bytes + B*L*4*(d_in + size_s);  flops = 2*MACs. */
int  tip_algorithmic_cost(const tip_model* m, int B, int L, double* bytes, double* flops);
/* Number of kernels the last tip_forward launched (for bench.py's gpu_launches). */
int  tip_last_launch_count(const tip_model* m);
/* Select the GEMM engine: 0 = auto (= 2), 1 = FFMA fp32 kernels (cross-check engine),
 * 2 = tcgen05 3xFP16-split tensor-core kernels. */
int  tip_set_gemm_engine(tip_model* m, int engine);
/* Kernel-selection knobs of the tcgen05 engine (initial values: TIP_* environment; results do not depend on them beyond
 * fp32 round-off, every combination is parity-tested).  Keys:
 *   "atm"           GEMMs that run on the A-operand-in-tensor-memory kernel (csrc/tip_umma_atm.cuh): bit mask 1 in_linear,
 *                   2 qkv, 4 ff1, 8 rnn_ih; -1 (default) = all four on handles that are or own execution lanes, none otherwise
 *   "atm_grid"      CTAs per such launch (0 = one per two 128-row tiles, at most 40)
 *   "atm_min_tiles" ... used for forwards of at least this many 128-row tiles (default 64)
 *   "dyn_sched"     1: the plain GEMMs draw their tiles from a device counter instead of a static round-robin (default 0)
 *   "ln_pair"       LayerNorm GEMMs with K >= value run on CTA pairs (cta_group::2); 0 (default) = never
 *   "ln_grid"       CTAs per fused-LayerNorm GEMM launch: 0 = one per 128-row tile, -1 (default) = that for a lone handle, one
 *                   per two row tiles (at most 40) on handles that are or own execution lanes (narrow kernels pack better across lanes)
 *   "ln_share"      fused-LayerNorm GEMMs on the two-row-tile kernel (csrc/tip_umma_ln2.cuh: a CTA takes pairs of row tiles that
 *                   share every W k-block): 1 on, 0 off, -1 (default) = on for handles that are or own execution lanes
 *   "rnn_clusters"  8-CTA clusters per tensor-core recurrence launch (0 = default: one per 20 windows)
 *   "atm_pair"      1: the A-in-tensor-memory GEMMs run on CTA pairs (cta_group::2, each CTA stages half of W); default 0
 *   "attn_grid"     attention: 0 (default) = one CTA per (window, 8 heads); N > 0 = N persistent CTAs with double-buffered
 *                   K / V tiles; -1 = two such CTAs per SM
 * No reference counterpart (the reference has no kernels of its own). */
int  tip_set_tuning(tip_model* m, const char* key, int value);
/* CUDA-graph the forward for a fixed (B, L) (used by the streaming path); 0 disables. */
int  tip_set_use_graphs(tip_model* m, int enable);
/* Per-stage device timing of the forward (bench.py's roofline leg): when enabled, a CUDA event is
 * recorded on the launch stream before every kernel of tip_forward.  tip_profile_stages returns
 * the number of stages the last forward recorded; tip_profile_get synchronises on the events and
 * returns stage i's name ("in_linear", "qkv", "attention", "out_proj_ln", "ff1", "ff2_ln",
 * "rnn_ih", "rnn", "head", ...), its layer (-1 if none) and its duration in milliseconds. */
int  tip_set_profile(tip_model* m, int enable);
int  tip_profile_stages(const tip_model* m);
int  tip_profile_get(tip_model* m, int i, char* name, int name_cap, int* layer, float* ms);
/* Test hook: copy an internal activation buffer of the LAST forward into dst_dev (capacity in
 * floats; *numel receives the element count): "embed" = encoder output rows (M,256), "qkv" (M,768),
 * "gi" (M,512), "hs" (M,512); rows are b*L + t.  Asynchronous on `stream`. */
int  tip_debug_tensor(tip_model* m, const char* name, float* dst_dev, int64_t capacity,
                      int64_t* numel, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TIP_B200_H_ */
