// Internal launcher declarations (one per kernel family). All launchers are asynchronous on `st`.
#pragma once
#include "common.cuh"

namespace mvf {

// ---- pack / unpack (elementwise.cu) ------------------------------------------------------------------
struct PackEntry {
  const float* src;  // fp32 [rows, cols] contiguous
  void* dst;         // [rows, ld_dst] in dst dtype, zero padded
  int rows, cols, ld_dst, dst_bf16;  // dst_bf16: 0 fp32, 1 bf16, 2 pre-split bf16 hi|lo blocks (ld_dst floats, multiple of 32)
  int perm_n;        // > 1: destination row r (column c for a vector) comes from source row (r % perm_n) * (rows / perm_n) + r / perm_n
};
int pack_params(const PackEntry* entries, int n, cudaStream_t st);
struct UnpackEntry {
  const float* src;  // gpack + offset, [rows, ld_src]
  float* dst;        // [rows, cols] contiguous
  int rows, cols, ld_src;
  int perm_n;        // > 1: destination row r (column c for a vector) comes from source row (r % perm_n) * (rows / perm_n) + r / perm_n
};
int unpack_grads(const UnpackEntry* entries, int n, float scale, cudaStream_t st);

// ---- positional encoding (models/utils.py:113-145) ---------------------------------------------------
int posenc_table(float* table, int T, int H, int train_frames, cudaStream_t st);
// z[b, e*T+t, :] = drop(h3[(b*T+t)*E+e, :] + pe[t, :])
int posenc_add(const float* h3, const float* table, float* z, int BV, int T, int E, int H, float p, DropSeed seed,
               cudaStream_t st);
// dh3[(b*T+t)*E+e, :] = drop'(dz[b, e*T+t, :])   (out dtype)
int posenc_bwd(int dtype_out, const float* dz, void* dh3, int BV, int T, int E, int H, float p, DropSeed seed,
               cudaStream_t st);

// ---- LayerNorm with fused residual-add + dropout (models/utils.py:147-159) ----------------------------
// z_out = z_in + drop(o) (o may be null -> z_out = z_in); r = LN(z_out)*gamma + beta (r may be null -> add only)
int ln_fwd(int dtype_out, const float* z_in, const float* o, float* z_out, void* r, float* mean, float* rstd,
           const float* gamma, const float* beta, int64_t rows, int H, float eps, float p, DropSeed seed, int site,
           cudaStream_t st);
// dz_out = dz_in + LN'(dr); dgamma/dbeta accumulated (atomics). dz_in may be null (treated as 0).
// drop_out (optional): also writes dz_out * dropmask(site) -- the masked gradient the next residual branch consumes.
int ln_bwd(const float* dr, const float* z, const float* mean, const float* rstd, const float* gamma,
           const float* dz_in, float* dz_out, float* dgamma, float* dbeta, int64_t rows, int H, cudaStream_t st,
           float* drop_out = nullptr, float p = 0.f, DropSeed seed = 0, int site = 0);

// ---- BatchNorm1d (+ReLU +dropout) in training / eval mode ---------------------------------------------
// The statistics are produced in three launches so that a multi-GPU caller can all-reduce `sums` in between:
//   bn_stats (local partial sums, float64) -> [all-reduce] -> bn_finalize (mean/invstd + running stats) -> bn_apply
int bn_stats(const float* x, int64_t R, int C, double* sums, cudaStream_t st);  // sums[0:C]+=sum x, [C:2C]+=sum x^2
// mi[0:C] = mean, mi[C:2C] = invstd. training: from sums / n_global (+ running-stat update, momentum, unbiased
// variance, num_batches_tracked += 1); eval: from the running buffers.
int bn_finalize(const double* sums, int C, double n_global, float eps, int training, float momentum, float* rmean,
                float* rvar, int64_t* tracked, float* mi, cudaStream_t st);
// out = drop(relu?(gamma*(x-mean)*invstd+beta))
int bn_apply(int dtype_out, const float* x, int64_t R, int C, const float* mi, const float* gamma, const float* beta,
             int relu, void* out, int64_t ld_out, float p, DropSeed seed, int site, cudaStream_t st);
// bn_finalize + bn_apply in one launch (fp32 output, 16-byte aligned rows, C <= 4096); MVF_ERR_UNSUPPORTED without an error
// string when the shape does not qualify -- the caller then issues the two kernels
int bn_finalize_apply(const double* sums, int C, double n_global, float eps, int training, float momentum, float* rmean,
                      float* rvar, int64_t* tracked, float* mi, int dtype_out, const float* x, int64_t R,
                      const float* gamma, const float* beta, int relu, void* out, int64_t ld_out, float p, DropSeed seed, int site,
                      cudaStream_t st);
// dy = d_out*drop*relu'(y); bsums[0:C]+=sum dy, [C:2C]+=sum dy*xhat; dgamma/dbeta accumulated from the LOCAL sums
int bn_bwd_stats(const float* d_out, int64_t ld_d, const float* x, int64_t R, int C, const float* mi,
                 const float* gamma, const float* beta, int relu, float p, DropSeed seed, int site, double* bsums,
                 float* dgamma, float* dbeta, cudaStream_t st);
// dx = gamma*invstd*(dy - sum(dy)/n - xhat*sum(dy*xhat)/n) with the (all-reduced) bsums and global n
int bn_bwd_apply(int dtype_out, const float* d_out, int64_t ld_d, const float* x, int64_t R, int C, const float* mi,
                 const float* gamma, const float* beta, int relu, float p, DropSeed seed, int site, const double* bsums,
                 double n_global, void* dx, int64_t ld_dx, cudaStream_t st);

// ---- misc ------------------------------------------------------------------------------------------------
// out[c] += sum_r X[r,c] for a table of fp32 matrices in one launch
struct ColsumEntry {
  const float* X;
  float* out;
  int64_t R, ld;
  int C;
};
constexpr int COLSUM_MAX = 32;
int colsum_batched(const ColsumEntry* entries, int n, cudaStream_t st);
// out[c] += sum_r X[r,c]
int colsum(int dtype_in, const void* X, int64_t R, int C, int64_t ld, float* out, cudaStream_t st);
// out = in * dropmask (site) cast to dtype_out; p = 0 -> plain cast
int dropout_cast(int dtype_out, const float* in, void* out, int64_t rows, int cols, int64_t ld_out, float p,
                 DropSeed seed, int site, cudaStream_t st);
int cast_f32(int dtype_out, const float* in, void* out, int64_t n, cudaStream_t st);
// entity reduction (mvformer.py:181-190): z [BV, E*T, H] -> y [BV*T, H]
int entity_reduce_fwd(int dtype_out, const float* z, void* y, int32_t* argmax, int BV, int T, int E, int H, int mode,
                      cudaStream_t st);
int entity_reduce_bwd(const float* dy, const int32_t* argmax, float* dz, int BV, int T, int E, int H, int mode,
                      cudaStream_t st);
// 'lin' gather: zl[b*T+t, e*H + c] = z[b, e*T+t, c] and its transpose-scatter
int entity_gather_lin(int dtype_out, const float* z, void* zl, int BV, int T, int E, int H, cudaStream_t st);
int entity_scatter_lin(const float* dzl, float* dz, int BV, int T, int E, int H, cudaStream_t st);
// F.normalize (eps 1e-12) and its backward; norms saved
int l2norm_fwd(const float* x, float* y, float* norm, int64_t rows, int D, cudaStream_t st, float* y2 = nullptr);
int l2norm_bwd(const float* dy, const float* y, const float* norm, void* dx, int dtype_out, int64_t rows, int D,
               cudaStream_t st);
int dropout_mask_export(DropSeed seed, int site, int64_t rows, int64_t cols, float p, float* out, cudaStream_t st);

// ---- xattn.cu ----------------------------------------------------------------------------------------------
// ent32 (optional, fp32 [F*E, SPC]): the pooled entities before dropout; when given (and bf16, E <= 4) the single-pass
// kernels are used -- backward needs it for rowdot[e] = <dEnt[e], ent[e]>.
int xattn_pool_fwd(int dtype, int F, int P, int E, int SPC, const void* kv, const float* q_s, const float* q_b,
                   float* attn, void* ent, int64_t ld_ent, float* ent32, int one_hot, float drop_p, DropSeed seed,
                   cudaStream_t st);
int xattn_pool_bwd(int dtype, int F, int P, int E, int SPC, const void* kv, const float* q_s, const float* q_b,
                   const float* attn, const void* d_ent, int64_t ld_ent, const float* ent32, int one_hot, float drop_p,
                   DropSeed seed, void* d_kv, float* d_q_s, float* d_q_b, float* d_bk, float* d_bv, cudaStream_t st);

// ---- pool_fold.cu: rank-E folded entity pooling (no K|V tensors; one streaming pass over the tokens per direction) ------
bool pool_fold_supported(int dtype, int C, int E, int P);
// Wq[E,C] = (Q_s + Q_b) Wk / sqrt(SPC)
int fold_prep(const float* q_s, const float* q_b, const float* Wk, int E, int SPC, int C, float* Wq, cudaStream_t st);
// attn[F,E,P] = softmax_p(Wq X^T), px[F*E,C] = attn X        (X: [F*P, C] token-major, dtype bf16 / fp32)
int pool_fold_fwd(int dtype, int F, int P, int E, int C, const void* X, const float* Wq, float* attn, float* px,
                  cudaStream_t st);
// dWq[E,C] += sum_f (attn * (G X^T - <G, px>)) X              (G = dEnt Wv, [F*E, C] fp32)
// delta [F*E] = <G_row, px_row> is optional (nullptr: reduced inside the kernels)
int pool_fold_bwd(int dtype, int F, int P, int E, int C, const void* X, const float* G, const float* px, const float* attn,
                  float* dWq, cudaStream_t st, const float* delta = nullptr);
// dWk += Q^T dWq / sqrt(SPC); dQ_s += dWq Wk^T / sqrt(SPC); dQ_b += column sums of dQ_s
int fold_finish(const float* dWq, const float* q_s, const float* q_b, const float* Wk, int E, int SPC, int C, float* dWk,
                int64_t ld_dwk, float* dQs, float* dQb, cudaStream_t st);
// bf16 tokens, C % 16 == 0: the same two passes on mma.sync tensor cores, warp-specialised (pool_fold_ws.cu): 8-token ring
// slots, mbarrier-only hand-offs; called by the two above
bool pool_fold_ws_supported(int dtype, int C, int P);
int pool_fold_ws_fwd(int F, int P, int E, int C, const void* X, const float* Wq, float* attn, float* px, cudaStream_t st);
int pool_fold_ws_bwd(int F, int P, int E, int C, const void* X, const float* G, const float* px, const float* attn,
                     const float* delta, float* dWq, cudaStream_t st);
// SMs the streaming backward kernel leaves free (a collective running beside it gets them); returns the previous value
int pool_fold_reserve_sms(int n);
// h0 = [drop(ent) | drop(one-hot) | 0]  and its backward (optionally delta[row] = <dEnt[row], ent[row] - bv> = <G_row, px_row>)
int ent_finish_fwd(const float* ent, float* h0, int64_t ld, int64_t R, int SPC, int E, int one_hot, float p, DropSeed seed,
                   cudaStream_t st);
int ent_finish_bwd(const float* d_h0, int64_t ld, float* dEnt, int64_t R, int SPC, int W, float p, DropSeed seed,
                   cudaStream_t st, const float* ent = nullptr, const float* bv = nullptr, float* delta = nullptr);

// ---- optim.cu: fused clip + Adam / AdamW over a list of tensors -----------------------------------------------------------
size_t opt_ws_bytes(int n_tensors);
int opt_adam_step(int n_tensors, float* const* params, const float* const* grads, float* const* m, float* const* v,
                  const int64_t* numel, const float* lr_dev, int64_t* step_dev, double beta1, double beta2, float eps, float wd,
                  int adamw, float max_norm, float inv_scale, float* norm_out, void* ws, size_t ws_bytes, cudaStream_t st);

// ---- peer.cu: sum of small float64 buffers over NVLink peer memory (BatchNorm statistics across ranks) ----------------
size_t peer_buffer_bytes();
int peer_sum_f64(double* local, int64_t n, void* const* bufs_dev, int rank, int world, uint32_t* counter, cudaStream_t st);
// in-place SUM over the ranks of n floats at data_off of every rank's symmetric buffer (mc_base: NVSwitch multicast address
// of the buffer or null); flags at flag_off (peer_allreduce_flag_bytes(), zero before first use); counters: 64 device words
size_t peer_allreduce_flag_bytes();
int peer_allreduce_f32(void* mc_base, void* const* bufs_dev, size_t data_off, size_t flag_off, int64_t n, int rank, int world,
                       uint32_t* counters, int ctas, cudaStream_t st);

// ---- attention.cu --------------------------------------------------------------------------------------------
// allow_split: the caller accepts operands split into bf16 hi + lo (2^-16 relative) -> tensor-core kernels of attention_tc.cu
// for S <= 64, d_k = 32; otherwise (and always for exact-fp32 callers) the CUDA-core kernels
// ws / ws_bytes (optional, attention_ws_bytes): scratch for the split operands of the tcgen05 kernels (attention_fa.cu, S > 64)
size_t attention_ws_bytes(int B, int S, int heads, int dk);
int attention_fwd(int dtype, int B, int S, int heads, int dk, const void* qkv, const float* keymask, void* ctx,
                  float* lse, cudaStream_t st, bool allow_split = false, void* ws = nullptr, size_t ws_bytes = 0);
int attention_bwd(int dtype, int B, int S, int heads, int dk, const void* qkv, const float* keymask, const void* ctx,
                  const float* lse, const void* d_ctx, void* d_qkv, float* delta, cudaStream_t st, bool allow_split = false,
                  void* ws = nullptr, size_t ws_bytes = 0);
// attention_fa.cu: flash-attention forward / dQ / dK,dV on tcgen05 (d_k = 32, fp32 qkv, bf16 hi|lo operand splits)
size_t attention_fa_ws_bytes(int B, int S, int heads);
bool attention_fa_ok(int dtype, int S, int dk, int H, bool allow_split, const void* ws, size_t ws_bytes, int B, int heads);
int attention_fa_fwd(int B, int S, int heads, const void* qkv, const float* keymask, void* ctx, float* lse, void* ws,
                     cudaStream_t st);
int attention_fa_bwd(int B, int S, int heads, const void* qkv, const float* keymask, const void* ctx, const float* lse,
                     const void* d_ctx, void* d_qkv, void* ws, cudaStream_t st);
bool attention_tc_ok(int dtype, int S, int dk, int H, const void* qkv, const void* other, bool allow_split);
int attention_tc_fwd(int B, int S, int heads, const void* qkv, const float* keymask, void* ctx, float* lse, cudaStream_t st);
int attention_tc_bwd(int B, int S, int heads, const void* qkv, const float* keymask, const void* ctx, const float* lse,
                     const void* d_ctx, void* d_qkv, cudaStream_t st);

// ---- scl.cu ----------------------------------------------------------------------------------------------------
size_t scl_ws_bytes(int Bv, int T, int D);
// scl_mma.cu: the per-pair kernel and the cross passes on mma.sync (T <= 256, D <= 256, D % 4 == 0)
struct SclWs;
struct SclCrossJobs;
bool scl_mma_supported(int T, int D);
int scl_pair_mma(const float* embs, const int64_t* seq_lens, const int64_t* steps, const float* masks, int Bv, int T, int D,
                 float tau, float two_var, const SclWs& w, int use_zext, float* loss_out, float* d_embs, cudaStream_t st);
int scl_cross_mma(const float* embs, int N, int T2, int D, float tau, const SclCrossJobs& J, int grad, float* sum_out,
                  float* vec_out, cudaStream_t st);
int scl_fwd_bwd(const float* embs, const int64_t* seq_lens, const int64_t* steps, const float* masks, int Bv, int T,
                int D, float temperature, float label_variance, int negative_type, int quirk, float* loss_out,
                float* d_embs, void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace mvf
