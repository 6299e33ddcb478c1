// knn_internal.cuh -- structures shared by knn.cu (orchestration, exact kernels) and knn_tc.cu (tensor pass).
#pragma once
#include "common.cuh"

namespace scl {

constexpr int kKeep = 64;        // k': candidates kept per (query, database range) and rescored per query
constexpr int kCandCap = 128;    // capacity of one candidate list (k' + slack between prunes)

// Device-resident description of a database shard's shadow copy (first 256 bytes of the shadow buffer).
struct ShadowHeader {
  long long R;
  int D, Dp;
  unsigned int maxabs_bits;      // max |x| over the shard (float bits; non-negative floats order like uints)
  int scale_exp;                 // fp16 copy holds x * 2^scale_exp
  unsigned long long rmax2_bits; // max squared row norm (double bits)
  int built;
  int pad[55];
};
static_assert(sizeof(ShadowHeader) == 256, "header must stay 256 bytes");

struct TcArgs {
  const float* rn;       // [R] exact squared row norms (fp32-rounded)
  const float* qmul;     // [Q] -2 * 2^-(scale_db + scale_q)
  int Q;
  int R;
  int Dp;
  int num_m_blocks, num_n_tiles, NR, tiles_per_range;
  int group_m;           // query-axis work units processed concurrently (bounds the L2 working set of A)
  float* cand_s;         // [Q, NR, kCandCap]
  uint32_t* cand_i;      // [Q, NR, kCandCap]
  int* cand_cnt;         // [Q, NR]
  unsigned int* q_thr;   // [Q] best (smallest) pruning threshold any range has published for the query, as an
                         //     order-preserving uint (0xffffffff = none yet); shared by all ranges of the query
  float* dbg_scores;     // optional [Q, R]: raw fp16-pass scores (tests; scl_knn_set_debug_scores)
  // Stage 2 ("collect" mode, knn.cu): every query carries a FIXED threshold and every row scoring below it is appended
  // to ONE list per query (no pruning, no published thresholds); a list that overflows coll_cap is detected by its count.
  int collect;
  const float* fixed_thr;   // [Q]
  uint32_t* coll_idx;       // [Q, coll_cap] database rows
  int* coll_cnt;            // [Q] number of rows that passed (may exceed coll_cap)
  int coll_cap;
  // Sliding-window pacing of the TMA producers (performance only, never needed for correctness): producer w bumps
  // sync_ctr[p] once it has issued the loads of its p-th sync point (a fraction of a tile) and does not start point p
  // before every worker has passed point p - sync_window.  Keeps the workers that stream the same database range (and
  // the same query blocks) within a fraction of a tile of each other, so the shared operand is fetched from HBM once
  // and found in L2 by everybody else.  nullptr / sync_total == 0 disables it.
  unsigned int* sync_ctr;   // [kSyncMax], zeroed before the launch
  int sync_total, sync_window, sync_subs;
  // Completion signal per query group (optional, zeroed before the launch): every epilogue warp bumps group_done[g] after
  // it has written the candidate lists of one work item of group g (release: fence, then the atomic), so that a consumer
  // on ANOTHER stream (cuStreamWaitValue32 on the counter) can merge / rescore the queries of a finished group while
  // this kernel is still streaming the database for the next one.  Arrivals per item: 4 warps x CTAs of the worker.
  unsigned int* group_done;
};
constexpr int kMaxGroups = 64;
constexpr int kSyncMax = 1 << 16;
constexpr int kCollectCap = 2048;   // stage-2 list capacity per query

int knn_tc_launch(const TcArgs& a, const void* queries_fp16, const void* db_fp16, cudaStream_t stream);
void knn_tc_tiling(int Q, int64_t R, int Dp, int* num_m_blocks, int* num_n_tiles, int* NR, int* tiles_per_range, int* group_m);

}  // namespace scl
