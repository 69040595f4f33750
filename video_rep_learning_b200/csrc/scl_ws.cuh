// Device workspace of the Sequence Contrastive Loss launch sequence (scl.cu, scl_mma.cu).
#pragma once
#include <stddef.h>

namespace mvf {

struct SclWs {
  float* M;       // [1] sum of masks
  float* Z;       // [N] per-row 1 / Z (0: masked row or empty partition sum), exchanged between the CTAs of a pair
  float* g;       // [N] g = sum_j y r
  float* den;     // [N] log2 of the label normaliser (+inf: no valid partner)
  float* c;       // [N]  g / (Z M)
  float* zext;    // [N] partition-sum extras produced by the cross passes (batch_noself negatives)
  int* counts;    // [2] n_valid, n_masked
  int* valid;     // [N] ordered list of valid rows
  int* masked;    // [N] ordered list of masked rows
  int* chunk;     // [2 * (nchunks + 1)] valid / masked counts per 1024-row chunk, then their exclusive scans (large N only)
};

size_t scl_ws_layout(int N, SclWs* w, char* base);

// One cross pass (scl_mma.cu): rows of a compacted list against columns of a compacted list, x_rk = w0 exp(l_rk)
struct SclCrossJob {
  const int* row_idx;
  const int* row_cnt;
  const int* col_idx;
  const int* col_cnt;
  const float* rc;        // per-row factor (null: 1), gradient jobs only
  const float* cc;        // per-column factor (null: 1), gradient jobs only
  float w0;
  int excl_same_video;    // 1: drop (r, k) of the same video (batch_noself)
};
struct SclCrossJobs {
  int n;
  SclCrossJob job[4];
};

}  // namespace mvf
