/*
 * mvf_b200.h -- C ABI of libmvf_b200.so: the MV-Former head + SCL training hot path on B200 (sm_100a).
 *
 * Plain C: device pointers (allocated and owned by the caller, e.g. torch.empty), sizes, POD structs and an
 * explicit cudaStream_t.  The library never allocates persistent device memory, never synchronises the
 * device, holds no thread-local CUDA state (forward runs on the Python main thread, backward on the
 * autograd engine thread) and returns an int status; mvf_last_error() gives the message of the last
 * failure on the calling thread.  There is no CPU fallback anywhere behind this interface.
 *
 * Each entry point names the reference interface it replaces (paths relative to
 * facebookresearch/video_rep_learning/CARL_MVF).  The reference has no native code: what is replaced is
 * the body of the PyTorch modules / functions cited.
 */
#ifndef MVF_B200_H
#define MVF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVF_ABI_VERSION 3

typedef void* mvf_stream_t; /* cudaStream_t */

enum mvf_status {
  MVF_OK = 0,
  MVF_ERR_BAD_ARG = 1,     /* null pointer, negative size, inconsistent descriptor          */
  MVF_ERR_UNSUPPORTED = 2, /* configuration the reference supports but this build does not  */
  MVF_ERR_WORKSPACE = 3,   /* caller-provided save / scratch buffer too small               */
  MVF_ERR_CUDA = 4,        /* a CUDA runtime / driver call failed                           */
  MVF_ERR_ALIGN = 5        /* pointer or leading dimension violates an alignment contract   */
};

enum mvf_dtype { MVF_F32 = 0, MVF_BF16 = 1 };
enum mvf_final { MVF_FINAL_MAX = 0, MVF_FINAL_ONE = 1, MVF_FINAL_AVG = 2, MVF_FINAL_LIN = 3 };
enum mvf_onehot { MVF_ONEHOT_NONE = 0, MVF_ONEHOT_POOL = 1, MVF_ONEHOT_ENC = 2 };
enum mvf_gemm_backend { MVF_GEMM_AUTO = 0, MVF_GEMM_SIMT = 1, MVF_GEMM_TCGEN05 = 2 };
enum mvf_negative { MVF_NEG_SINGLE_NOSELF = 0, MVF_NEG_BATCH_NOSELF = 1 };
/* How the entity cross-attention pooling (a3-a5) is evaluated.  FOLDED: the E static queries are folded into the key
 * projection (Wq = Q Wk / sqrt(SPC), [E, C_in]) and the value projection is applied after the pooling, so K and V are
 * never formed: one HBM-bound streaming pass over the tokens per direction, identical results up to fp32
 * re-association.  DENSE: as written in the reference (K|V projection GEMM on tcgen05, then attention over K|V). */
enum mvf_pool_mode { MVF_POOL_AUTO = 0, MVF_POOL_DENSE = 1, MVF_POOL_FOLDED = 2 };
/* Which pooling module feeds the per-entity MLP.  LSTP: LearnableTokenPooling (entity cross-attention, mvformer.py:207-266).
 * FWB: FIXED_WIDTH_BASELINE (FWBPooling, mvformer.py:421-462; configs_mvf/ablate_dinoB8_fwb{3,5}.yml): the patch tokens are
 * ignored and one Linear(cls_dim -> SPC*E) of each frame's CLS embedding is reshaped [frames, SPC, E]. */
enum mvf_pool_kind { MVF_POOLKIND_LSTP = 0, MVF_POOLKIND_FWB = 1 };

#define MVF_MAX_FC 4
#define MVF_MAX_ENTITIES 16

/* Shape + hyper-parameters of one head invocation.
 * Mirrors the cfg keys read by models/mvformer.py:20-115, models/utils.py:196-242,
 * models/resnet_c2d.py:112-120. */
typedef struct mvf_head_desc {
  int32_t BV;           /* video-views in this call (= 2 * videos during training)                 */
  int32_t T;            /* frames per view                                                         */
  int32_t P;            /* patch tokens per frame (196 for /16 @224, 784 for /8)                   */
  int32_t C_in;         /* token channels = MODEL.BASE_MODEL.OUT_CHANNEL                           */
  int32_t E;            /* SMART_TOKENS (entity queries)                                           */
  int32_t SPC;          /* SMART_POOL_CHANNELS (384 default)                                       */
  int32_t n_fc;         /* len(FC_LAYERS)                                                          */
  int32_t fc[MVF_MAX_FC]; /* FC_LAYERS channels * CAPACITY_SCALAR                                  */
  int32_t H;            /* HIDDEN_SIZE                                                             */
  int32_t DFF;          /* D_FF                                                                    */
  int32_t heads;        /* NUM_HEADS                                                               */
  int32_t L;            /* NUM_LAYERS                                                              */
  int32_t D;            /* EMBEDDING_SIZE                                                          */
  int32_t PS;           /* MODEL.PROJECTION_SIZE (hidden width of MLPHead)                         */
  int32_t one_hot;      /* mvf_onehot                                                              */
  int32_t final_mode;   /* mvf_final                                                               */
  int32_t train_frames; /* TRAIN.NUM_FRAMES: pos-enc uses linspace positions when T differs        */
  int32_t dtype;        /* mvf_dtype of tokens, K|V and their gradients (the 96 % of the FLOPs);   */
                        /* activations behind the pooling are fp32 (tf32 MMAs on the TC backend)   */
  int32_t training;     /* 1: BatchNorm batch statistics + dropout; 0: running statistics          */
  int32_t has_mask;     /* 1: video_masks given ([BV,T] float, 0 = padded frame)                   */
  int32_t gemm_backend; /* mvf_gemm_backend; AUTO = tcgen05 (bf16 + tf32) for bf16 tokens, exact   */
                        /* fp32 FMA (SIMT) for fp32 tokens                                         */
  int32_t world_size;   /* ranks sharing BatchNorm statistics (1 = local statistics)               */
  int32_t pool_mode;    /* mvf_pool_mode; AUTO = FOLDED whenever C_in % 8 == 0 (env MVF_POOL_MODE   */
                        /* = dense|folded overrides AUTO for A/B measurements)                      */
  float drop_p;         /* FC_DROPOUT_RATE (applied only when training)                            */
  float ln_eps, bn_eps, bn_momentum;
  uint64_t seed;        /* dropout stream of this step (counter-based, same in fwd and bwd)        */
  const uint64_t* seed_dev; /* optional DEVICE counter added to `seed` inside the kernels (NULL = none): */
                        /* a CUDA-graph-captured step freezes `seed`, so the caller advances *seed_dev  */
                        /* on the device once per replay to keep drawing fresh dropout masks            */
  int32_t pool_kind;    /* mvf_pool_kind                                                            */
  int32_t cls_dim;      /* FWB: width of the CLS embedding (OUT_CHANNEL / number of SMART_FEATS layers) */
  const float* cls_emb; /* FWB: DEVICE [BV*T, cls_dim] fp32 CLS embeddings of the frames, rows in the order the */
                        /* reference hands them over (transformer.py:200,217); frozen, no gradient         */
} mvf_head_desc;

/* ---- library / bookkeeping ------------------------------------------------------------------------ */
int mvf_version(void);
const char* mvf_last_error(void);
/* 1 when the library was built with the tcgen05/TMA GEMM and the current device is sm_100. */
int mvf_has_tcgen05(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
uint64_t mvf_launch_count(void);
/* Measurement aid: when enabled (1 = tags 0 and 1 only, 2 = every tag), CUDA events are recorded on the launching stream
 * around the dominant kernels.
 * DENSE pooling: tag 0 K|V projection GEMM, 1 its weight-gradient GEMM, 2 cross-attention pooling fwd, 3 its bwd.
 * FOLDED pooling: tag 0 streaming pooling pass fwd, 1 streaming pass bwd, 2 the rest of the forward pooling block
 * (Wq fold, px Wv^T GEMM, dropout/one-hot), 3 the rest of its backward (dEnt, G and dWv GEMMs, dWk/dQ finish).
 * Both modes: tag 6 temporal self-attention forward (one bracket per encoder layer), 7 its backward, 8 SCL (loss + gradient).
 * mvf_profile_read synchronises on the recorded events and returns the durations in milliseconds. */
int mvf_profile_enable(int on);
int mvf_profile_read(int tag, float* ms, int cap, int* n);

/* Canonical parameter table (state_dict order of `embed.*` then `ssl_projection.*`).
 * mvf_param_info fills the reference state_dict key and the logical shape of parameter `idx`. */
int mvf_num_params(const mvf_head_desc* d);
int mvf_param_info(const mvf_head_desc* d, int idx, char* name, size_t name_cap, int64_t* rows, int64_t* cols);
/* BatchNorm layers in call order (fc_layers.2, fc_layers.6, ..., ssl_projection.net.1). */
int mvf_num_bn(const mvf_head_desc* d);
int mvf_bn_info(const mvf_head_desc* d, int idx, char* name, size_t name_cap, int64_t* channels);

/* Sizes of the caller-allocated buffers (bytes / fp32 elements). */
size_t mvf_save_bytes(const mvf_head_desc* d);    /* activations kept from forward for backward         */
size_t mvf_ws_bytes(const mvf_head_desc* d);      /* scratch, free to reuse after each call returns     */
size_t mvf_gpack_elems(const mvf_head_desc* d);   /* flat fp32 gradient buffer (the all-reduce payload)  */
/* Leading elements of the flat gradient buffer that only the LAST backward phase (n_fc + 1: the pooling backward) writes;
 * everything behind them is final when phase n_fc returns, so its all-reduce can overlap the pooling kernel. */
size_t mvf_gpack_pool_elems(const mvf_head_desc* d);
/* Number of SMs the streaming pooling-backward kernel leaves free for a collective running beside it (0 = none);
 * returns the previous setting. */
int mvf_pool_bwd_reserve_sms(int32_t n);
size_t mvf_proj_save_bytes(const mvf_head_desc* d); /* same two, for mvf_proj_forward / mvf_proj_backward */
size_t mvf_proj_ws_bytes(const mvf_head_desc* d);
/* Named views into the head save buffer (stage outputs), into the projection save buffer ("proj:<name>")
 * or into gpack ("g.<name>"), for stage-by-stage parity tests. dtype: 0 f32, 1 bf16, 2 f64, 3 i32. */
int mvf_save_lookup(const mvf_head_desc* d, const char* name, size_t* offset, int64_t* rows, int64_t* cols,
                    int64_t* ld, int32_t* dtype);
/* Per-BatchNorm statistics exchange buffers inside `save` (fp32): forward [sum x, sum x^2] (2*C doubles),
 * backward [sum dy, sum dy*xhat] (2*C doubles).  With world_size > 1 the caller all-reduces (SUM) the
 * buffer (float64) between the two phases that bracket it (see the phase constants below).  Replaces SyncBatchNorm's
 * all_gather / all_reduce (train.py:283). */
int mvf_bn_stat_lookup(const mvf_head_desc* d, int bn_idx, int backward, size_t* offset, int64_t* n_doubles);

/* ---- a2-a9: MultiEntityTransformerEmbModel.forward / backward (models/mvformer.py:128-200) --------- */
/* Phases: the chain is cut at every BatchNorm statistic so that a caller with world_size > 1 can all-reduce
 * the statistics buffer in between.  [0, MVF_PHASE_ALL) runs everything. Head forward has n_fc + 1 phases:
 * phase i ends after the partial statistics of fc BatchNorm i are written.  Head backward mirrors it (phase i ends after
 * the partial sums of BatchNorm n_fc-1-i) and has one more phase, n_fc + 1: the pooling backward (see
 * mvf_gpack_pool_elems). */
#define MVF_PHASE_ALL 255

/* tokens [BV*T*P, C_in] token-major contiguous (dtype = d->dtype); mask [BV,T] fp32 or NULL;
 * params: array (host memory) of mvf_num_params device pointers (fp32) in canonical order;
 * bn_running: array of 2*mvf_num_bn device pointers {running_mean, running_var} (fp32), updated in
 * training; bn_tracked: array of mvf_num_bn device pointers to int64 num_batches_tracked (may be NULL);
 * out_emb [BV*T, D] fp32; attn_out optional [BV*T, E, P] fp32 softmax maps (the attn_holder side channel,
 * mvformer.py:408-411). */
int mvf_head_forward(const mvf_head_desc* d, const float* const* params, float* const* bn_running,
                     int64_t* const* bn_tracked, const void* tokens, const float* mask, void* save,
                     size_t save_bytes, void* ws, size_t ws_bytes, float* out_emb, float* attn_out,
                     int phase_begin, int phase_end, mvf_stream_t stream);

/* d_emb [BV*T, D] fp32 -> parameter gradients accumulated into gpack (caller zero-fills it once per step,
 * before the first backward call that uses it).  Tokens are frozen: no d(tokens) is produced
 * (TRAIN_BASE: frozen; transformer.py:186-189). */
int mvf_head_backward(const mvf_head_desc* d, const float* const* params, const void* tokens,
                      const float* mask, const float* d_emb, void* save, size_t save_bytes, void* ws,
                      size_t ws_bytes, float* gpack, int phase_begin, int phase_end, mvf_stream_t stream);

/* ---- a10: MLPHead + F.normalize (models/resnet_c2d.py:112-126; models/transformer.py:226-230) ------- */
/* emb [N, D] fp32 (N = BV*T) -> out [N, D] fp32.  project = 1: MLPHead then L2-normalise (unit rows);
 * project = 0: L2-normalise only (MODEL.L2_NORMALIZE without the head); project = 2: MLPHead only.
 * Phases: forward 2 (cut at the BatchNorm statistic), backward 2. */
int mvf_proj_forward(const mvf_head_desc* d, const float* const* params, float* const* bn_running,
                     int64_t* const* bn_tracked, const float* emb, int project, void* save, size_t save_bytes,
                     void* ws, size_t ws_bytes, float* out, int phase_begin, int phase_end, mvf_stream_t stream);
int mvf_proj_backward(const mvf_head_desc* d, const float* const* params, const float* d_out, int project,
                      void* save, size_t save_bytes, void* ws, size_t ws_bytes, float* gpack, float* d_emb,
                      int phase_begin, int phase_end, mvf_stream_t stream);

/* Scatter the flat gradient buffer into per-parameter gradient tensors (array of mvf_num_params device
 * pointers, fp32, contiguous logical shapes; NULL entries are skipped). scale multiplies every gradient
 * (1/world_size after a SUM all-reduce -- DDP's mean, train.py:285-286). */
int mvf_unpack_grads(const mvf_head_desc* d, const float* gpack, float* const* grads, float scale,
                     mvf_stream_t stream);

/* ---- a12: SCL.compute_sequence_loss fwd + bwd fused (algos/scl.py:52-105) --------------------------- */
/* embs [Bv,2,T,D] fp32 unit rows; seq_lens [Bv,2] int64; steps [Bv,2,T] int64; masks [Bv,2,T] fp32.
 * loss_out: 1 fp32 (overwritten); d_embs [Bv,2,T,D] fp32 = d loss / d embs (overwritten; may be NULL).
 * quirk = 1 reproduces scl.py:80 exactly (every masked frame of the local batch enters every partition
 * sum with weight 1e-6); quirk = 0 keeps only the own-pair terms. ws: mvf_scl_ws_bytes bytes.
 * T <= 256, D <= 256, D % 4 == 0, embs 16-byte aligned (MVF_ERR_UNSUPPORTED / MVF_ERR_ALIGN otherwise).  The products run on
 * the warp-level tensor cores with bf16 hi/lo operand splits: loss within 1e-6, gradient within 1e-5 of the fp64 evaluation. */
size_t mvf_scl_ws_bytes(int32_t Bv, int32_t T, int32_t D);
int mvf_scl_fwd_bwd(const float* embs, const int64_t* seq_lens, const int64_t* steps, const float* masks,
                    int32_t Bv, int32_t T, int32_t D, float temperature, float label_variance,
                    int32_t negative_type, int32_t quirk, float* loss_out, float* d_embs, void* ws,
                    size_t ws_bytes, mvf_stream_t stream);

/* ---- building blocks, exported for unit parity tests and micro-benchmarks --------------------------- */
/* C[M,N] (+)= opA(A) * opB(B) + bias.  a_kmajor: A stored [M,K] row-major (else [K,M]);
 * b_kmajor: B stored [N,K] row-major like nn.Linear.weight (else [K,N]).  dtype_ab / dtype_c: mvf_dtype.
 * Backend TCGEN05: bf16 operands -> tcgen05 kind::f16, fp32 operands -> kind::tf32 (fp32 output only).
 * flags: bit0 ReLU, bit1 accumulate into C (fp32 C only), bit2 multiply by (relu_src > 0), bit3 (TCGEN05 backend,
 * fp32 operands, both K-major; ignored otherwise) "bf16x3": each fp32 operand is split in shared memory into
 * bf16 hi + lo and A_lo*B_hi + A_hi*B_lo + A_hi*B_hi is accumulated on kind::f16 -- 16 mantissa bits per operand
 * instead of tf32's truncated 10.  The head uses it for every forward GEMM behind the pooling. */
#define MVF_GEMM_RELU 1
#define MVF_GEMM_ACCUM 2
#define MVF_GEMM_RELUMASK 4
#define MVF_GEMM_SPLIT3 8
/* with SPLIT3: B ([N, K] K-major) is already stored "pre-split": each 32-float block of a row holds 32 bf16 hi values
 * followed by 32 bf16 lo values (rows padded with zeros to a multiple of 32 floats; ldb counts floats).  Weights are
 * packed this way once per step, which removes two thirds of the in-kernel conversion work. */
#define MVF_GEMM_B_PRESPLIT 16
int mvf_gemm(int backend, int dtype_ab, int dtype_c, int a_kmajor, int b_kmajor, int64_t M, int64_t N, int64_t K,
             const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, const float* bias,
             const void* relu_src, int64_t ld_relu, int flags, int split_k, mvf_stream_t stream);

/* a3-a5 alone (mvformer.py:243-266, 352-414; utils.py:11-44): kv [F*P, 2*SPC] (K | V), q_s [E,SPC],
 * q_b [SPC] -> attn [F,E,P] fp32, ent [F*E, ld_ent] fp32 (dropout applied), one-hot columns appended when
 * one_hot = 1; d_ent is fp32 as well.  dtype is the element type of kv / d_kv.
 * ent_f32 (optional, [F*E, SPC] fp32): pooled entities before dropout; when given for bf16 and E <= 4 the
 * single-pass kernels run (each K|V row read once); backward needs the same buffer back. */
int mvf_xattn_pool_fwd(int dtype, int32_t F, int32_t P, int32_t E, int32_t SPC, const void* kv, const float* q_s,
                       const float* q_b, float* attn, void* ent, int64_t ld_ent, float* ent_f32, int one_hot,
                       float drop_p, uint64_t seed, mvf_stream_t stream);
int mvf_xattn_pool_bwd(int dtype, int32_t F, int32_t P, int32_t E, int32_t SPC, const void* kv, const float* q_s,
                       const float* q_b, const float* attn, const void* d_ent, int64_t ld_ent, const float* ent_f32,
                       int one_hot, float drop_p, uint64_t seed, void* d_kv, float* d_q_s, float* d_q_b, float* d_bk,
                       float* d_bv, mvf_stream_t stream);

/* a3-a5 folded (see mvf_pool_mode): the four pieces around the two small GEMMs (ent = px Wv^T + bv forward;
 * G = dEnt Wv and dWv = dEnt^T px backward, done with mvf_gemm).  tokens [F*P, C_in] token-major (dtype), everything
 * else fp32: wq [E, C_in]; attn [F, E, P]; px [F*E, C_in] (attention-pooled tokens); g [F*E, C_in];
 * d_wq [E, C_in] (accumulated: zero it first); d_wk [SPC, ld_dwk], d_q_s [E, SPC], d_q_b [SPC] (accumulated). */
int mvf_pool_fold_prep(const float* q_s, const float* q_b, const float* w_k, int32_t E, int32_t SPC, int32_t C_in,
                       float* wq, mvf_stream_t stream);
int mvf_pool_fold_fwd(int dtype, int32_t F, int32_t P, int32_t E, int32_t C_in, const void* tokens, const float* wq,
                      float* attn, float* px, mvf_stream_t stream);
int mvf_pool_fold_bwd(int dtype, int32_t F, int32_t P, int32_t E, int32_t C_in, const void* tokens, const float* g,
                      const float* px, const float* attn, float* d_wq, mvf_stream_t stream);
/* Same as mvf_pool_fold_bwd with the per-row constants delta[F*E] = <g_row, px_row> supplied by the caller (the fused
 * head gets them for free as <dEnt_row, ent_row - b_v> while it builds dEnt); delta == NULL -> reduced inside. */
int mvf_pool_fold_bwd_delta(int dtype, int32_t F, int32_t P, int32_t E, int32_t C_in, const void* tokens, const float* g,
                            const float* px, const float* attn, const float* delta, float* d_wq, mvf_stream_t stream);
int mvf_pool_fold_finish(const float* d_wq, const float* q_s, const float* q_b, const float* w_k, int32_t E, int32_t SPC,
                         int32_t C_in, float* d_wk, int64_t ld_dwk, float* d_q_s, float* d_q_b, mvf_stream_t stream);

/* Fused optimizer tail: gradient unscale (inv_scale, AMP GradScaler train.py:124-127), global-norm
 * clipping (torch.nn.utils.clip_grad_norm_, OPTIMIZER.GRAD_CLIP train.py:126,151; max_norm <= 0 disables it) and Adam
 * (adamw = 0, L2 weight decay folded into the gradient) or AdamW (adamw = 1) as built by utils/optimizer.py:60-73, over
 * n_tensors fp32 tensors.  params / grads / m / v: HOST arrays of device pointers, numel: host array; lr_dev (float) and
 * step_dev (int64, advanced by the call) are DEVICE scalars so that the launches are CUDA-graph replayable; norm_out
 * (device float, may be NULL) receives the pre-clip gradient norm; ws: mvf_opt_ws_bytes(n_tensors) bytes of scratch. */
size_t mvf_opt_ws_bytes(int32_t n_tensors);
int mvf_opt_adam_step(int32_t n_tensors, float* const* params, const float* const* grads, float* const* m, float* const* v,
                      const int64_t* numel, const float* lr_dev, int64_t* step_dev, double beta1, double beta2, float eps,
                      float weight_decay, int32_t adamw, float max_norm, float inv_scale, float* norm_out, void* ws,
                      size_t ws_bytes, mvf_stream_t stream);

/* Cross-rank sum of a small float64 buffer (BatchNorm statistics, replaces the SyncBatchNorm exchange of train.py:283)
 * over NVLink peer memory: bufs_dev = DEVICE array of `world` (<= 16) pointers to the ranks' symmetric buffers of
 * mvf_peer_buffer_bytes() bytes each (2 parities x 16 ranks x slot of 16-byte entries: n <= bytes / 512 values;
 * zero-filled before first use; e.g. torch.distributed._symmetric_memory),
 * counter = this rank's device-resident exchange counter (starts at 0, advanced by the kernel: graph-replayable).
 * Every rank must call it the same number of times in the same order; the result is bitwise identical on all ranks. */
size_t mvf_peer_buffer_bytes(void);
int mvf_peer_sum_f64(double* local, int64_t n, void* const* bufs_dev, int32_t rank, int32_t world, uint32_t* counter,
                     mvf_stream_t stream);

/* In-place SUM over the ranks of the flat fp32 gradient buffer (replaces DDP's all-reduce, train.py:286) through the same
 * kind of symmetric memory: every rank's buffer holds n floats at data_off and mvf_peer_allreduce_flag_bytes() of flags at
 * flag_off (zero before first use); mc_base = NVSwitch multicast address of the buffer (multimem.ld_reduce / multimem.st) or
 * NULL (peer loads / stores); counters = 64 device words (start at 0); ctas <= 64 (0: default).  Each rank reduces and
 * broadcasts its 1/world slice, so all ranks end with the bitwise identical sum.  Graph-replayable. */
size_t mvf_peer_allreduce_flag_bytes(void);
int mvf_peer_allreduce_f32(void* mc_base, void* const* bufs_dev, size_t data_off, size_t flag_off, int64_t n, int32_t rank,
                           int32_t world, uint32_t* counters, int32_t ctas, mvf_stream_t stream);

/* a5/a8 temporal self-attention core (utils.py:11-44 with the [B,1,1,S] key mask): qkv [B*S, 3*H]
 * (Q | K | V, head h at columns h*dk), keymask [B,S] fp32 or NULL -> ctx [B*S, H], lse [B,heads,S] fp32.
 * ws == NULL: exact fp32 arithmetic on CUDA cores (the parity mode).  ws != NULL (mvf_attention_ws_bytes bytes,
 * 1024-byte aligned; fp32 data, dk = 32): tensor-core kernels on bf16 hi|lo operand splits (~2^-16 relative) --
 * mma.sync for S <= 64, tcgen05 / TMEM / TMA flash-attention kernels for longer sequences (the mode the fused head
 * uses on the tensor-core backend). */
size_t mvf_attention_ws_bytes(int32_t B, int32_t S, int32_t heads, int32_t dk);
int mvf_attention_fwd(int dtype, int32_t B, int32_t S, int32_t heads, int32_t dk, const void* qkv,
                      const float* keymask, void* ctx, float* lse, void* ws, size_t ws_bytes, mvf_stream_t stream);
int mvf_attention_bwd(int dtype, int32_t B, int32_t S, int32_t heads, int32_t dk, const void* qkv,
                      const float* keymask, const void* ctx, const float* lse, const void* d_ctx, void* d_qkv,
                      float* ws_delta, void* ws, size_t ws_bytes, mvf_stream_t stream);

/* The dropout keep-mask (pre-scaled by 1/(1-p)) the kernels use at `site` for a [rows, cols] tensor; lets
 * tests feed identical masks to the oracle.  Sites: 0 fc0 input, 1.. fc_i input, 8 pos-enc,
 * 16+2l attention branch of layer l, 17+2l FFN branch. */
int mvf_dropout_mask(uint64_t seed, int32_t site, int64_t rows, int64_t cols, float p, float* out,
                     mvf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MVF_B200_H */
