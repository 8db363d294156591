// Kernels + extern "C" entry points of libriichienv_b200.so (see include/riichienv_b200.h).
// There is NO CPU fallback in this file: every compute entry point launches CUDA kernels
// and returns RV_ERR_CUDA if the device is unavailable.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cub/device/device_scan.cuh>

#include "obs.cuh"
#include "obs_ext.cuh"
#include "obs_ext3.cuh"
#include "seq.cuh"
#include "validate.h"
#include "../../include/rv_synth.h"

using namespace rv;

// ------------------------------------------------------------------ error plumbing
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
int rv_internal_fail(int code, const std::string& msg) { return fail(code, msg); }   // for the host-only sources (replay.cpp)
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(RV_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));               \
  } while (0)

struct HandStage {              // rv_hand_eval_batch: persistent staging (allocated on first use)
  rv_hand_query *d_q[2] = {nullptr, nullptr}, *h_q[2] = {nullptr, nullptr};
  rv_hand_result *d_r[2] = {nullptr, nullptr}, *h_r[2] = {nullptr, nullptr};
  int32_t* d_list[2] = {nullptr, nullptr};   // hands with a winning shape (+ the counter behind them)
  cudaStream_t st[2];
  cudaEvent_t done[2];
};
struct rv_ctx {
  int device;
  std::mutex hands_mu;          // rv_hand_eval_batch shares one pair of staging buffers per context
  int32_t* d_hand_list = nullptr;   // rv_hand_eval_batch_device: list of hands with a winning shape (grow-only)
  size_t hand_list_bytes = 0;
  HandStage hands;
  cudaStream_t stream;
  cudaEvent_t ev[8];
  cudaStream_t aux[3];          // phase pipeline: RESPOND, DEAL and SLOW kernels run beside ACT
  cudaEvent_t fork_ev, join_ev[3];
  uint32_t *suit_info, *honor_info;
  uint64_t *suit_cost, *honor_cost;
  Tables T;
  int sm_count;
};

struct rv_vec {
  rv_ctx* ctx;
  int64_t n;
  int game_mode;
  uint32_t rule_bits;
  G* d_states;
  uint32_t* d_log;  // n * log_cap words or nullptr
  uint32_t log_cap;
  uint64_t* d_seeds;            // scratch for rv_vec_reseed
  int32_t *d_obs_counts, *d_obs_offsets;   // encode: active seats per game and their exclusive scan (n + 1)
  uint32_t* d_idbits;                      // encode: action-id sets [n][4][3]
  int32_t* h_obs_total;                    // pinned: row count of the last encode
  int32_t* d_row_index;                    // observe+step: row -> game*4+seat when the caller does not ask for it
  void* d_scan_tmp;
  unsigned char *d_gather, *h_gather;   // results/counters staging (device, pinned host)
  uint32_t* d_seq_cursor;       // [n][4] event-log word offset of each seat's previous observation (rv_vec_encode_seq)
  uint32_t* d_seq_start;        // [n][4] scratch for caller-supplied cursors
  size_t scan_tmp_bytes;
  // phase pipeline (rv_vec_step_random): per-phase game lists, double buffered, and per-game step budgets
  int32_t* d_lists;             // [2 buffers][3 phases][n]
  uint32_t* d_list_counts;      // [2][4]
  uint32_t* d_budget;           // [n]
  uint32_t* h_counts;           // pinned [16]
  // persistent scheduler (rollout_persistent): ring queue per class + live-game counter
  int32_t* d_q_slots;           // [4][q_cap]
  uint32_t* d_q_ctl;            // head[4], tail[4] (one 128-byte line each), live, error
  uint32_t q_cap;
  cudaGraphExec_t graph_exec;   // one period of the phase pipeline, captured once per agent seed
  uint64_t graph_seed;
  int graph_period;
  unsigned long long* d_steps;  // [0] = env steps executed, [1] = games finished (by step kernels)
  unsigned char* d_obs_snap;    // rv_vec_observe_step_random: packed snapshot of what the encoder reads (n x OBS_STAGE_BYTES)
  void *d_io_a, *d_io_c;        // staging of rv_vec_step / rv_vec_legal_actions (grow-only)
  size_t io_a_bytes, io_c_bytes;
  // replay ingestion with the log resident in HBM (rv_vec_replay_load / rv_vec_replay_advance)
  rv_log_action* d_rp_actions;  // the log actions of every record's kyoku, concatenated
  int64_t* d_rp_first;          // [n + 1] offsets into d_rp_actions
  int32_t* d_rp_cursor;         // [n] actions of the record's kyoku applied so far
  int32_t* d_rp_live;           // records that still had an action in the last advance
};

// ------------------------------------------------------------------ kernels
__global__ void __launch_bounds__(128) synth_hands_kernel(rv_hand_query* q, uint64_t first, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  rv_hand_query h;
  rv_synth_hand(first + (uint64_t)i, &h);
  q[i] = h;
}
// rv_vec_apply_events: thread per game
__global__ void __launch_bounds__(64) apply_events_kernel(Tables T, G* states, int64_t n, const rv_mjai_event* events) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || events[i].type == RV_EV_NONE) return;
  Ctx cx;
  cx.T = T;
  cx.log = nullptr;            // fed events are not appended to the device log (the caller keeps their text)
  cx.log_cap = 0;
  cx.defer_init = cx.defer_tail = false;
  cx.idbits = nullptr;
  apply_mjai_event(cx, states[i], events[i]);
}
// rv_vec_apply_log_actions / rv_vec_replay_begin: thread per game (one record follows one kyoku of a parsed log)
__global__ void __launch_bounds__(64) apply_log_actions_kernel(Tables T, G* states, int64_t n, const rv_log_action* acts) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || acts[i].type == RV_LA_NONE) return;
  Ctx cx;
  cx.T = T;
  cx.log = nullptr;
  cx.log_cap = 0;
  cx.defer_init = cx.defer_tail = false;
  cx.idbits = nullptr;
  apply_log_action(cx, states[i], acts[i]);
}
// rv_vec_replay_advance: every record applies the next action of its own kyoku, read from the log resident in HBM
__global__ void __launch_bounds__(64) replay_advance_kernel(Tables T, G* states, int64_t n, const rv_log_action* __restrict__ acts,
                                                            const int64_t* __restrict__ first, int32_t* cursor, int32_t* live) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool did = false;
  if (i < n) {
    const int32_t c = cursor[i];
    const int64_t at = first[i] + c;
    if (at < first[i + 1]) {
      Ctx cx;
      cx.T = T;
      cx.log = nullptr;
      cx.log_cap = 0;
      cx.defer_init = cx.defer_tail = false;
      cx.idbits = nullptr;
      if (acts[at].type != RV_LA_NONE) apply_log_action(cx, states[i], acts[at]);
      cursor[i] = c + 1;
      did = true;
    }
  }
  const unsigned m = __ballot_sync(0xFFFFFFFFu, did);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(live, __popc(m));
}
__global__ void __launch_bounds__(64) replay_begin_kernel(Tables T, G* states, int64_t n, const rv_log_kyoku* kyokus) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Ctx cx;
  cx.T = T;
  cx.log = nullptr;
  cx.log_cap = 0;
  cx.defer_init = cx.defer_tail = false;
  cx.idbits = nullptr;
  replay_begin_patch(cx, states[i], kyokus[i]);
}
// rv_vec_step_agent: thread per game, every class of work inline (a test / evaluation path, not the throughput path)
__global__ void __launch_bounds__(128) agent_kernel(Tables T, G* states, int64_t n, uint32_t* log, uint32_t cap, int policy,
                                                    uint64_t agent_seed, uint32_t max_steps, unsigned long long* counters);
// Batched hand evaluation in two kernels.  Measured on the one-kernel version (ncu, 10^7 seeded hands of which 10 % have a
// winning shape): 7.5 of 32 lanes active per instruction, instruction-cache hit rate 62 % — the yaku evaluation of the few
// complete hands ran with the rest of their warps masked off, and its code evicted the hot loop.
//   hand_shape_kernel  every hand, uniform work: validation, histograms, wait set, both shanten numbers (table lookups),
//                      and the shape test; hands that have a winning shape are appended to a list (one atomic per warp);
//   hand_yaku_kernel   the listed hands only, densely packed: yaku / han / fu / score (hand_calc).
__global__ void __launch_bounds__(256) hand_shape_kernel(Tables T, const rv_hand_query* __restrict__ q, rv_hand_result* __restrict__ out,
                                                         int64_t n, int32_t* __restrict__ list, unsigned int* __restrict__ count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool shape = false;
  if (i < n) {
    rv_hand_query h = q[i];
    rv_hand_result o;
    shape = hand_eval_shape(T, h, o);
    out[i] = o;
  }
  const unsigned m = __ballot_sync(0xFFFFFFFFu, shape);
  if (m) {
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(count, (unsigned)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (shape) list[base + __popc(m & ((1u << lane) - 1))] = (int32_t)i;
  }
}
constexpr int YAKU_THREADS = 1024;  // one block per SM: 32 warps at 64 registers, paced by BlockPace (hand.cuh)
__global__ void __launch_bounds__(YAKU_THREADS) hand_yaku_kernel(Tables T, const rv_hand_query* __restrict__ q, rv_hand_result* __restrict__ out,
                                                                 const int32_t* __restrict__ list, const unsigned int* __restrict__ count,
                                                                 int paced) {
  __shared__ unsigned arrivals;
  if (threadIdx.x == 0) arrivals = 0;
  __syncthreads();
  const unsigned total = *count, stride = gridDim.x * blockDim.x;
  const unsigned iters = (total + stride - 1) / stride;            // the same for every thread: a generation per iteration
  BlockPace pace{&arrivals, 0u, false};
  for (unsigned it = 0; it < iters; it++) {
    const unsigned k = it * stride + blockIdx.x * blockDim.x + threadIdx.x;
    pace.target = (it + 1) * blockDim.x;
    pace.arrived = false;
    if (paced > 1) __syncthreads();                                 // the division search starts together too (uniform loop: a real barrier)
    if (k < total) {
      const int32_t i = list[k];
      rv_hand_query h = q[i];
      rv_hand_result o = out[i];
      hand_eval_win(T, h, o, paced ? &pace : nullptr);
      out[i] = o;
    }
    if (paced) pace_arrive(&pace);                                 // threads without a hand, and any path that did not arrive
  }
}
// one launch pair on `st`; `list` holds n entries, `count` one word
static cudaError_t launch_hand_eval(const Tables& T, const rv_hand_query* d_q, rv_hand_result* d_out, int64_t n, int32_t* list,
                                    unsigned int* count, int sm_count, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(count, 0, sizeof(unsigned int), st);
  if (e != cudaSuccess) return e;
  hand_shape_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(T, d_q, d_out, n, list, count);
  {
    const char* e = getenv("RV_YAKU_PACED");               // 0: eight 128-thread blocks per SM, no rendezvous (A/B)
    const int paced = e ? atoi(e) : 2;
    if (paced) hand_yaku_kernel<<<sm_count, YAKU_THREADS, 0, st>>>(T, d_q, d_out, list, count, paced);
    else hand_yaku_kernel<<<sm_count * 8, 128, 0, st>>>(T, d_q, d_out, list, count, 0);
  }
  return cudaGetLastError();
}

__device__ __forceinline__ Ctx make_ctx(const Tables& T, uint32_t* log, uint32_t cap, int64_t i);
__global__ void __launch_bounds__(128) agent_kernel(Tables T, G* states, int64_t n, uint32_t* log, uint32_t cap, int policy,
                                                    uint64_t agent_seed, uint32_t max_steps, unsigned long long* counters) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G& g = states[i];
  Ctx cx = make_ctx(T, log, cap, i);
  unsigned long long steps = 0;
  const bool was_done = g.is_done;
  for (uint32_t k = 0; k < max_steps && !g.is_done; k++, steps++) agent_step(cx, g, policy, agent_seed, g.seed);
  if (steps) atomicAdd(&counters[0], steps);
  if (!was_done && g.is_done) atomicAdd(&counters[1], 1ull);
}
__device__ __forceinline__ Ctx make_ctx(const Tables& T, uint32_t* log, uint32_t cap, int64_t i) {
  Ctx cx;
  cx.T = T;
  cx.log = log ? log + (size_t)i * cap : nullptr;
  cx.log_cap = cap;
  cx.defer_init = false;
  cx.defer_tail = false;
  cx.idbits = nullptr;
  return cx;
}

__global__ void create_kernel(G* states, int64_t n, int game_mode, uint32_t rule_bits, const uint64_t* seeds, uint64_t seed_base) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G& g = states[i];
  uint8_t* p = (uint8_t*)&g;
  for (size_t k = 0; k < sizeof(G); k++) p[k] = 0;
  g.game_mode = (uint8_t)game_mode;
  g.rule_bits = (uint8_t)rule_bits;
  g.seed = seeds ? seeds[i] : seed_base + (uint64_t)i;
  g.hand_index = 1;   // GameState::new consumed shuffle #0 (state/mod.rs:165)
  g.last_error = RV_NONE;
  g.pending_init[0] = g.pending_init[1] = g.pending_init[2] = RV_NONE;
  g.pending_tail[0] = RV_NONE;
  g.is_done = 1;      // until reset
  for (int s = 0; s < MAXP; s++) g.score[s] = game_mode >= 3 ? (s < 3 ? 35000 : 0) : 25000;   // state_3p/game_mode.rs:31-33
}

__global__ void refresh_kernel(Tables T, G* state) { refresh_caches(T, *state); }
// RiichiEnv::_reveal_kan_dora / _get_ura_markers (env.rs:624-631): the device routines the step path itself uses, run on one game
__global__ void debug_call_kernel(Tables T, G* state, uint32_t* log, uint32_t cap, int64_t game, int op, uint8_t* out) {
  G& g = *state;
  Ctx cx = make_ctx(T, log, cap, game);
  if (op == 0) {
    reveal_kan_dora(cx, g);
    out[7] = g.n_dora;
  } else if (op == 1) {
    out[7] = (uint8_t)ura_indicators(g, out);
  } else if (op == 2) {
    trigger_ryukyoku(cx, g, RV_RK_EXHAUSTIVE);
    out[7] = g.is_done;
  } else if (op == 6) {                          // replay: claim lists of every seat against the last discard
    if (g.last_discard_pid != RV_NONE) claims_after_tile(cx, g, g.last_discard_pid, g.last_discard_tile, false);
    out[7] = (uint8_t)__popc(g.active_mask);
  } else {
    next_round(cx, g, op == 4, op == 5);       // 3: (false, false)  4: (oya_won, false)  5: (false, is_draw)
    out[7] = g.is_done;
  }
}

__global__ void reseed_kernel(G* states, int64_t n, const uint64_t* seeds, uint64_t seed_base) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  states[i].seed = seeds ? seeds[i] : seed_base + (uint64_t)i;
  states[i].hand_index = 1;   // as a freshly constructed RiichiEnv(seed=...) (state/mod.rs:165)
}

__global__ void reset_kernel(Tables T, G* states, int64_t n, uint32_t* log, uint32_t cap, const uint8_t* oya, const uint8_t* rw,
                             const uint8_t* honba, const uint32_t* kyotaku, const int32_t* scores, const uint8_t* walls) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Ctx cx = make_ctx(T, log, cap, i);
  game_reset(cx, states[i], oya ? oya[i] : 0, rw ? rw[i] : 0, honba ? honba[i] : 0, kyotaku ? kyotaku[i] : 0,
             walls ? walls + (size_t)i * (states[i].game_mode >= 3 ? 108 : 136) : nullptr, scores ? scores + (size_t)i * MAXP : nullptr);
}

// Persistent rollout: each thread owns one game and advances it up to max_steps env steps.
// Divergence control: the one rare, expensive transition — shuffling and dealing the next round
// (~1 in 88 steps per game, ~10k instructions) — is not run inline.  A game whose round ended parks
// (g.pending_init) and the warp deals parked games together once INIT_BATCH lanes are waiting, or when
// no lane can make progress otherwise, so the deal runs convergent on many lanes instead of on one.
#ifndef RV_INIT_BATCH
#define RV_INIT_BATCH 8
#endif
// IDS (one-step launches of the observation pipeline): the legal lists' action-id sets go to idbits[game][seat][3].
template <bool IDS>
__global__ void __launch_bounds__(128) step_random_kernel(Tables T, G* states, int64_t n, uint32_t* log, uint32_t cap,
                                                          uint64_t agent_seed, uint32_t max_steps, unsigned long long* counters,
                                                          uint32_t* idbits) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long my_steps = 0, my_done = 0;
  bool alive = i < n;
  G& g = states[alive ? i : 0];
  if (IDS && alive) {
    // one-step launches of the observation pipeline: the rows streamed out since the last step have pushed the records out
    // of the L2; ask for all 14 lines of this game's record at once instead of meeting them one dependent miss at a time
    const char* rec = reinterpret_cast<const char*>(&g);
    #pragma unroll
    for (int k = 0; k < (int)((sizeof(G) + 127) / 128); k++) asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + 128 * k));
  }
  Ctx cx = make_ctx(T, log, cap, alive ? i : 0);
  cx.defer_init = true;
  uint32_t my_ids[3 * MAXP];          // built in (L1-cached) local memory, stored once: not a read-modify-write per id in HBM
  if (IDS) cx.idbits = my_ids;
  uint64_t gid = alive ? g.seed : 0;
  uint32_t taken = 0;
  while (true) {
    bool parked = alive && g.pending_init[0] != RV_NONE;
    bool can = alive && !parked && !g.is_done && taken < max_steps;
    if (can) {
      random_step<IDS>(cx, g, agent_seed, gid);
      if (IDS) {
        uint4* o = reinterpret_cast<uint4*>(idbits + (size_t)i * (3 * MAXP));
        o[0] = make_uint4(my_ids[0], my_ids[1], my_ids[2], my_ids[3]);
        o[1] = make_uint4(my_ids[4], my_ids[5], my_ids[6], my_ids[7]);
        o[2] = make_uint4(my_ids[8], my_ids[9], my_ids[10], my_ids[11]);
      }
      taken++;
      my_steps++;
      parked = g.pending_init[0] != RV_NONE;
      can = !parked && !g.is_done && taken < max_steps;
    }
    unsigned pm = __ballot_sync(0xFFFFFFFFu, parked), wm = __ballot_sync(0xFFFFFFFFu, can);
    if (pm == 0 && wm == 0) break;
    if (__popc(pm) >= RV_INIT_BATCH || wm == 0) {
      if (parked) run_pending_init(cx, g);
    }
  }
  if (alive && g.is_done && my_steps > 0) my_done = 1;
  // block reduction -> one atomic per block
  __shared__ unsigned long long sh[2];
  if (threadIdx.x == 0) sh[0] = sh[1] = 0;
  __syncthreads();
  for (int o = 16; o > 0; o >>= 1) {
    my_steps += __shfl_down_sync(0xFFFFFFFFu, my_steps, o);
    my_done += __shfl_down_sync(0xFFFFFFFFu, my_done, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&sh[0], my_steps);
    atomicAdd(&sh[1], my_done);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(&counters[0], sh[0]);
    atomicAdd(&counters[1], sh[1]);
  }
}

// One env step per game, lock-step, with the work regrouped inside the block so that warps run ONE kind of transition.
// (A thread-per-game warp executes the union of the code paths its 32 games need: ncu on the one-step launches of the
// observation pipeline shows 5.2 of 32 lanes active per instruction and 8,700 instructions per warp.)  A block owns 256 games:
//   1. every thread classifies its own game; the turns the compact act_fast routine can take are taken right here — all
//      lanes run the same routine — with the follow-up of a discard that somebody may claim parked (pending_tail);
//   2. the games that are left — claim windows (RESP), parked follow-ups (TAIL), turns act_fast declined (SLOW) — are compacted
//      into per-kind lists in shared memory and handed out again in warp-aligned ranges: a warp now runs one generic routine
//      on up to 32 games of that kind; a round that ends is dealt by the same thread right away.
// Id sets (IDS) are built in shared memory and stored once per game.
constexpr int SB = 256;
template <bool IDS>
__global__ void __launch_bounds__(SB) step_sorted_kernel(Tables T, G* states, int64_t n, uint32_t* log, uint32_t cap, uint64_t agent_seed,
                                                         unsigned long long* counters, uint32_t* idbits) {
  enum { K_RESP = 0, K_TAIL = 1, K_SLOW = 2, K_NONE = 3 };
  __shared__ uint16_t list[4][SB];          // [3]: games whose round ended in this launch (dealt below, a warp per game)
  __shared__ int cnt[4];
  __shared__ uint32_t s_ids[IDS ? SB : 1][3 * MAXP];
  __shared__ unsigned long long sh[2];
  __shared__ DealScratch deal_scratch[SB / 32];
  const int t = threadIdx.x, lane = t & 31;
  const int64_t base = (int64_t)blockIdx.x * SB, i = base + t;
  if (t < 4) cnt[t] = 0;
  if (t == 0) sh[0] = sh[1] = 0;
  __syncthreads();
  const bool alive = i < n;
  G& g = states[alive ? i : 0];
  if (IDS && alive) {
    const char* rec = reinterpret_cast<const char*>(&g);
    #pragma unroll
    for (int k = 0; k < (int)((sizeof(G) + 127) / 128); k++) asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + 128 * k));
  }
  const bool live = alive && !g.is_done;
  int kind = K_NONE;
  if (live) {
    if (g.phase == RV_WAIT_ACT) {
      Ctx cx = make_ctx(T, log, cap, i);
      cx.defer_init = true;
      cx.defer_tail = true;
      if (IDS) cx.idbits = s_ids[t];
      if (!act_fast<IDS>(cx, g, agent_seed, g.seed)) kind = K_SLOW;
      else if (g.pending_tail[0] != RV_NONE) kind = K_TAIL;
    } else {
      kind = K_RESP;
    }
  }
  #pragma unroll
  for (int k = 0; k < 3; k++) {
    const unsigned m = __ballot_sync(0xFFFFFFFFu, kind == k);
    if (m == 0) continue;
    int b = 0;
    if (lane == __ffs(m) - 1) b = atomicAdd(&cnt[k], __popc(m));
    b = __shfl_sync(0xFFFFFFFFu, b, __ffs(m) - 1);
    if (kind == k) list[k][b + __popc(m & ((1u << lane) - 1))] = (uint16_t)t;
  }
  __syncthreads();
  {
    const int n_resp = cnt[K_RESP], n_tail = cnt[K_TAIL], n_slow = cnt[K_SLOW];
    const int o_tail = (n_resp + 31) & ~31, o_slow = o_tail + ((n_tail + 31) & ~31), total = o_slow + n_slow;
    for (int slot = t; slot < ((total + 31) & ~31); slot += SB) {
      int k = K_NONE, li = 0;
      if (slot < n_resp) k = K_RESP, li = list[K_RESP][slot];
      else if (slot >= o_tail && slot < o_tail + n_tail) k = K_TAIL, li = list[K_TAIL][slot - o_tail];
      else if (slot >= o_slow && slot < total) k = K_SLOW, li = list[K_SLOW][slot - o_slow];
      if (k != K_NONE) {
        const int64_t gi = base + li;
        G& h = states[gi];
        Ctx cx = make_ctx(T, log, cap, gi);
        cx.defer_init = true;
        if (IDS) cx.idbits = s_ids[li];
        if (k == K_RESP) random_step_resp<IDS>(cx, h, agent_seed, h.seed);
        else if (k == K_TAIL) run_pending_tail(cx, h);
        else random_step_act<IDS>(cx, h, agent_seed, h.seed);
      }
      // a round that ended: the deal (~18 k instructions on one thread) is handed to a whole warp below
      const bool parked = k != K_NONE && states[base + li].pending_init[0] != RV_NONE;
      if (parked) list[3][atomicAdd(&cnt[3], 1)] = (uint16_t)li;
    }
  }
  __syncthreads();
  {
    const int n_deal = cnt[3];
    for (int d = t >> 5; d < n_deal; d += SB / 32) {
      const int64_t gi = base + list[3][d];
      Ctx cx = make_ctx(T, log, cap, gi);
      run_pending_init_coop(cx, states[gi], deal_scratch[t >> 5]);
    }
  }
  __syncthreads();
  if (IDS && live) {
    uint4* o = reinterpret_cast<uint4*>(idbits + (size_t)i * (3 * MAXP));
    const uint32_t* b = s_ids[t];
    o[0] = make_uint4(b[0], b[1], b[2], b[3]);
    o[1] = make_uint4(b[4], b[5], b[6], b[7]);
    o[2] = make_uint4(b[8], b[9], b[10], b[11]);
  }
  unsigned long long my_steps = live ? 1 : 0, my_done = (live && g.is_done) ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) {
    my_steps += __shfl_down_sync(0xFFFFFFFFu, my_steps, o);
    my_done += __shfl_down_sync(0xFFFFFFFFu, my_done, o);
  }
  if (lane == 0 && my_steps) {
    atomicAdd(&sh[0], my_steps);
    atomicAdd(&sh[1], my_done);
  }
  __syncthreads();
  if (t == 0 && sh[0]) {
    atomicAdd(&counters[0], sh[0]);
    if (sh[1]) atomicAdd(&counters[1], sh[1]);
  }
}

// ---- phase-sorted rollout -------------------------------------------------------------------
// Games are compacted by phase into index lists (warp ballot + prefix, one atomic per warp and phase),
// and each phase runs as its own kernel over its list, so a warp holds 32 games that all execute the
// same transition and a kernel only carries the code of its phase:
//   PH_ACT   the seat to move picks and applies a turn action (discard / riichi / kan / tsumo ...)
//   PH_RESP  the claim window: every active seat answers (pass / chi / pon / kan / ron)
//   PH_DEAL  shuffle + deal of the next round for games parked by next_round()
//   PH_SLOW  ACT games the fast path declined (see act_fast): the generic, large-footprint turn code.
//   PH_REACT / PH_REACT_D  games that return to the ACT phase out of a RESP/SLOW kernel / a DEAL kernel.
// The big-code kernels (RESP, SLOW, DEAL) run ASYNCHRONOUSLY on side streams, overlapping the following fast
// iterations; they are only joined right before the lists they append to are swapped.
// The ACT kernel only carries act_fast (a few KB of SASS, instruction-cache resident); a game it cannot
// handle is re-filed under PH_SLOW and takes its step one iteration later.
// After stepping its game a thread files it into the NEXT iteration's list (double buffering), so an
// iteration is {memset counts; ACT | RESP | DEAL concurrently on three streams}.
enum { PH_ACT = 0, PH_RESP = 1, PH_DEAL = 2, PH_SLOW = 3, PH_REACT = 4, PH_REACT_D = 5, N_LISTS = 6, PH_NONE = 7 };
// persistent scheduler only: a committed discard whose follow-up is parked (g.pending_tail).  Queue index 4 there.
enum { PH_TAIL = 4, N_QUEUES = 5 };

__device__ __forceinline__ int classify(const G& g, uint32_t budget) {
  if (g.pending_tail[0] != RV_NONE) return PH_TAIL;      // (persistent scheduler only; flushed even when the budget is spent)
  if (g.pending_init[0] != RV_NONE) return PH_DEAL;      // must be flushed even when the budget is spent
  if (g.is_done || budget == 0) return PH_NONE;
  if (g.phase == RV_WAIT_ACT && g.current_player == RV_NONE) return PH_NONE;   // event-driven record between turns: nothing to roll out
  return g.phase == RV_WAIT_ACT ? PH_ACT : PH_RESP;
}
// every lane of the warp must call this (cls = PH_NONE for lanes without a game).
// ACT games go to the per-iteration ACT list; RESP / DEAL / SLOW games go to the accumulating slow lists,
// which are only drained every few iterations (their kernels carry the large generic code and would
// otherwise put ~100 us of instruction-fetch latency on the critical path of every iteration).
struct Lists {
  int32_t* dst[N_LISTS];     // where games of each class are appended
  uint32_t* cnt[N_LISTS];
};
__device__ __forceinline__ void file_game(int cls, int32_t gi, const Lists& L, int64_t n) {
  int lane = threadIdx.x & 31;
  #pragma unroll
  for (int c = 0; c < N_LISTS; c++) {
    unsigned m = __ballot_sync(0xFFFFFFFFu, cls == c);
    if (m == 0) continue;
    int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(L.cnt[c], (uint32_t)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (cls == c) L.dst[c][base + __popc(m & ((1u << lane) - 1))] = gi;
  }
}
__global__ void sched_init_kernel(const G* states, int64_t n, uint32_t* budget, uint32_t max_steps, Lists out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int cls = PH_NONE;
  if (i < n) {
    budget[i] = max_steps;
    cls = classify(states[i], max_steps);
  }
  file_game(cls, (int32_t)i, out, n);
}
// ---- shared-memory staging of the hot prefix (per-lane bulk copies, TMA engine) ----
// A lane's game record is 13 cache lines of HBM and a step touches most of the hot ones through dependent, uncoalesced
// loads (every warp-level load waits for its slowest lane: ncu showed ~20 warps stalled on long_scoreboard per issue).
// So each lane pulls its record's hot prefix (RV_HOT_BYTES = 544 B) into shared memory with ONE cp.async.bulk, all 32
// copies of the warp in flight together, runs the step(s) there, and writes the prefix back with one bulk store.
// The cold arrays (wall, river, claims) stay in HBM; game code reaches them through cold(g) (game.cuh).
constexpr int PHB = 32;                              // threads (= games) per block: one warp, own barrier, own exit
constexpr int STG_STRIDE = RV_HOT_BYTES + 16;        // 560 B = 140 words (== 12 mod 32, gcd 4: same-field accesses are 4-way conflicts)
static_assert(offsetof(G, river) % 8 == 0, "river alignment");
static_assert(offsetof(G, wall) == RV_HOT_BYTES && RV_HOT_BYTES % 16 == 0 && sizeof(G) % 16 == 0, "hot prefix layout");
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void stage_in(unsigned char* slot, const G* src, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(slot)),
               "l"(src), "r"((uint32_t)RV_HOT_BYTES), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void stage_out(G* dst, const unsigned char* slot) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // this thread's generic-proxy writes -> visible to the copy engine
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(slot)),
               "r"((uint32_t)RV_HOT_BYTES)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// ---- one env step per game, records staged in shared memory -----------------------------------------------------------
// step_sorted_kernel (above) reads every record through dependent global loads; in the observation pipeline the tensor rows
// streamed out since the previous step have pushed the records out of the L2, so a one-step launch is a chain of ~20 HBM round
// trips per thread (113-190 us for 65,536 games, of which a few us are bandwidth).  Here every warp first pulls the hot prefixes
// of its 32 games into shared memory with one bulk copy per lane (all in flight together), the block then works exactly like
// step_sorted_kernel — act_fast on every game, the games that need generic code regrouped by kind across the block's warps,
// finished rounds dealt by a warp each — on the shared-memory copies, and each lane stores its game back with one bulk copy.
constexpr int SS = 128;
struct StepStagedSmem {
  alignas(128) unsigned char stage[SS * STG_STRIDE];
  alignas(8) uint64_t mbar[SS / 32];
  DealScratch deal[SS / 32];
  uint32_t ids[SS][3 * MAXP];
  uint16_t list[4][SS];
  int cnt[4];
  unsigned long long sh[2];
};
template <bool IDS>
__global__ void __launch_bounds__(SS) step_staged_kernel(Tables T, G* states, int64_t n, uint32_t* log, uint32_t cap, uint64_t agent_seed,
                                                         unsigned long long* counters, uint32_t* idbits) {
  enum { K_RESP = 0, K_TAIL = 1, K_SLOW = 2, K_NONE = 3 };
  extern __shared__ __align__(128) unsigned char smem_raw[];
  StepStagedSmem& S = *reinterpret_cast<StepStagedSmem*>(smem_raw);
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int64_t base = (int64_t)blockIdx.x * SS, i = base + t;
  const bool alive = i < n;
  unsigned char* slot = S.stage + t * STG_STRIDE;
  const uint32_t bar = smem_u32(&S.mbar[w]);
  if (t < 4) S.cnt[t] = 0;
  if (t == 0) S.sh[0] = S.sh[1] = 0;
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const unsigned alive_mask = __ballot_sync(0xFFFFFFFFu, alive);
  if (alive_mask) {
    if (lane == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)__popc(alive_mask) * (uint32_t)RV_HOT_BYTES) : "memory");
    __syncwarp();
    if (alive) {
      stage_in(slot, &states[i], bar);
      *reinterpret_cast<G**>(slot + RV_HOT_BYTES) = &states[i];
    }
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(ok)
                   : "r"(bar), "r"(0)
                   : "memory");
  }
  __syncthreads();
  G& g = *reinterpret_cast<G*>(slot);
  const bool live = alive && !g.is_done;
  int kind = K_NONE;
  if (live) {
    if (g.phase == RV_WAIT_ACT) {
      Ctx cx = make_ctx(T, log, cap, i);
      cx.defer_init = true;
      cx.defer_tail = true;
      if (IDS) cx.idbits = S.ids[t];
      if (!act_fast<IDS>(cx, g, agent_seed, g.seed)) kind = K_SLOW;
      else if (g.pending_tail[0] != RV_NONE) kind = K_TAIL;
    } else {
      kind = K_RESP;
    }
  }
  #pragma unroll
  for (int k = 0; k < 3; k++) {
    const unsigned m = __ballot_sync(0xFFFFFFFFu, kind == k);
    if (m == 0) continue;
    int b = 0;
    if (lane == __ffs(m) - 1) b = atomicAdd(&S.cnt[k], __popc(m));
    b = __shfl_sync(0xFFFFFFFFu, b, __ffs(m) - 1);
    if (kind == k) S.list[k][b + __popc(m & ((1u << lane) - 1))] = (uint16_t)t;
  }
  __syncthreads();
  {
    const int n_resp = S.cnt[K_RESP], n_tail = S.cnt[K_TAIL], n_slow = S.cnt[K_SLOW];
    const int o_tail = (n_resp + 31) & ~31, o_slow = o_tail + ((n_tail + 31) & ~31), total = o_slow + n_slow;
    for (int sl = t; sl < ((total + 31) & ~31); sl += SS) {
      int k = K_NONE, li = 0;
      if (sl < n_resp) k = K_RESP, li = S.list[K_RESP][sl];
      else if (sl >= o_tail && sl < o_tail + n_tail) k = K_TAIL, li = S.list[K_TAIL][sl - o_tail];
      else if (sl >= o_slow && sl < total) k = K_SLOW, li = S.list[K_SLOW][sl - o_slow];
      if (k != K_NONE) {
        G& h = *reinterpret_cast<G*>(S.stage + li * STG_STRIDE);
        Ctx cx = make_ctx(T, log, cap, base + li);
        cx.defer_init = true;
        if (IDS) cx.idbits = S.ids[li];
        if (k == K_RESP) random_step_resp<IDS>(cx, h, agent_seed, h.seed);
        else if (k == K_TAIL) run_pending_tail(cx, h);
        else random_step_act<IDS>(cx, h, agent_seed, h.seed);
        if (h.pending_init[0] != RV_NONE) S.list[3][atomicAdd(&S.cnt[3], 1)] = (uint16_t)li;
      }
    }
  }
  __syncthreads();
  for (int d = w; d < S.cnt[3]; d += SS / 32) {      // rounds that ended: a warp deals each (init_round_coop)
    const int li = S.list[3][d];
    Ctx cx = make_ctx(T, log, cap, base + li);
    run_pending_init_coop(cx, *reinterpret_cast<G*>(S.stage + li * STG_STRIDE), S.deal[w]);
  }
  __syncthreads();
  if (alive) stage_out(&states[i], slot);
  if (IDS && live) {
    uint4* o = reinterpret_cast<uint4*>(idbits + (size_t)i * (3 * MAXP));
    const uint32_t* b = S.ids[t];
    o[0] = make_uint4(b[0], b[1], b[2], b[3]);
    o[1] = make_uint4(b[4], b[5], b[6], b[7]);
    o[2] = make_uint4(b[8], b[9], b[10], b[11]);
  }
  unsigned long long my_steps = live ? 1 : 0, my_done = (live && g.is_done) ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) {
    my_steps += __shfl_down_sync(0xFFFFFFFFu, my_steps, o);
    my_done += __shfl_down_sync(0xFFFFFFFFu, my_done, o);
  }
  if (lane == 0 && my_steps) {
    atomicAdd(&S.sh[0], my_steps);
    atomicAdd(&S.sh[1], my_done);
  }
  __syncthreads();
  if (t == 0 && S.sh[0]) {
    atomicAdd(&counters[0], S.sh[0]);
    if (S.sh[1]) atomicAdd(&counters[1], S.sh[1]);
  }
}
template <bool IDS>
static cudaError_t launch_step_staged(rv_ctx* c, rv_vec* v, uint64_t agent_seed, uint32_t* idbits) {
  static bool configured[2] = {false, false};
  if (!configured[IDS]) {
    cudaError_t e = cudaFuncSetAttribute(step_staged_kernel<IDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StepStagedSmem));
    if (e != cudaSuccess) return e;
    configured[IDS] = true;
  }
  step_staged_kernel<IDS><<<(unsigned)((v->n + SS - 1) / SS), SS, sizeof(StepStagedSmem), c->stream>>>(c->T, v->d_states, v->n, v->d_log, v->log_cap,
                                                                                        agent_seed, v->d_steps, idbits);
  return cudaGetLastError();
}

template <int PH, int OUT_ACT>
__global__ void __launch_bounds__(PHB) phase_kernel(Tables T, G* states, int64_t n, uint32_t* log, uint32_t cap, uint64_t agent_seed,
                                                    uint32_t* budget, const int32_t* in_list, const uint32_t* in_count, Lists out,
                                                    unsigned long long* counters, int reps) {
  __shared__ __align__(128) unsigned char stage[PHB * STG_STRIDE];
  __shared__ __align__(8) uint64_t mbar;
  const int lane = threadIdx.x;
  int64_t i = (int64_t)blockIdx.x * PHB + lane;
  uint32_t count = *in_count;
  if ((int64_t)blockIdx.x * PHB >= count) return;     // whole block idle (block-uniform)
  const bool have = i < count;
  int cls = PH_NONE;
  int32_t gi = have ? in_list[i] : -1;
  unsigned char* slot = stage + lane * STG_STRIDE;
  const uint32_t bar = smem_u32(&mbar);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    uint32_t live = min((uint32_t)PHB, count - (uint32_t)blockIdx.x * PHB);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(live * (uint32_t)RV_HOT_BYTES) : "memory");
  }
  __syncwarp();
  if (have) {
    stage_in(slot, &states[gi], bar);
    *reinterpret_cast<G**>(slot + RV_HOT_BYTES) = &states[gi];    // cold(g) finds the HBM record here
  }
  {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(ok)
                   : "r"(bar), "r"(0)
                   : "memory");
  }
  unsigned long long stepped = 0, finished = 0;
  if (have) {
    G& g = *reinterpret_cast<G*>(slot);
    Ctx cx = make_ctx(T, log, cap, gi);
    cx.defer_init = true;
    uint32_t b = budget[gi];
    const uint32_t b0 = b;
    if (PH == PH_DEAL) {
      run_pending_init(cx, g);
      cls = classify(g, b);
    } else if (PH == PH_ACT) {
      // up to `reps` consecutive fast steps while the game stays in the ACT class (most discards draw no claim)
      for (int r = 0; r < reps; r++) {
        if (!act_fast(cx, g, agent_seed, g.seed)) {
          cls = PH_SLOW;
          break;
        }
        b--;
        cls = classify(g, b);
        if (cls != PH_ACT) break;
      }
    } else {
      if (PH == PH_SLOW) random_step_act(cx, g, agent_seed, g.seed);
      else random_step_resp(cx, g, agent_seed, g.seed);
      b--;
      cls = classify(g, b);
    }
    if (b != b0) {
      budget[gi] = b;
      stepped = b0 - b;
      finished = g.is_done ? 1 : 0;
    }
    if (cls == PH_ACT && OUT_ACT != PH_ACT) cls = OUT_ACT;   // async kernels return ACT games through their own list
    stage_out(&states[gi], slot);
  }
  file_game(cls, gi, out, n);
  if (PH != PH_DEAL) {   // (uniform per kernel)
    for (int o = 16; o > 0; o >>= 1) {
      stepped += __shfl_down_sync(0xFFFFFFFFu, stepped, o);
      finished += __shfl_down_sync(0xFFFFFFFFu, finished, o);
    }
    if (lane == 0 && stepped) {
      atomicAdd(&counters[0], stepped);
      if (finished) atomicAdd(&counters[1], finished);
    }
  }
}


// ---- persistent rollout: device-side class queues, no grid-wide barrier ------------------------------------------
// The list pipeline above advances every game of a class in one kernel and waits for the slowest warp before the next
// iteration starts (ncu: the SMs hold a resident warp for ~40 % of a kernel's duration).  Here a fixed crew of warps
// stays resident for the whole rollout; each warp repeatedly claims up to 32 games of ONE class from that class's ring
// queue, stages them, steps them, stores them and pushes each into the queue of its next class.  Nothing waits for
// anything but work.  Per-SM class preference (from %smid) keeps one class's code hot in an SM's instruction cache;
// a warp whose class has no full batch takes another class's, and only then a partial batch.
//   push:  state stores -> __threadfence() -> ticket = atomicAdd(tail) -> slot[ticket] = game
//   pop :  CAS on head (lane 0) -> lanes spin on their slots -> fence.acq_rel.gpu (invalidates L1: cold arrays and the
//          budget are read with plain loads) -> bulk copy in
enum { Q_HEAD = 0, Q_TAIL = 256, Q_LIVE = 512, Q_ERR = 513, Q_SMCLASS = 544, Q_CTL_WORDS = 544 + 256 };   // word offsets into d_q_ctl; heads/tails on own 128 B lines
struct Queues {
  int32_t* slots;      // [N_QUEUES][cap], -1 = empty
  uint32_t* ctl;
  uint32_t mask;       // cap - 1
};
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
// Slot states: -1 empty, -2 abandoned by its consumer, >= 0 a game.  Ticket k of the producers and ticket k of the
// consumers meet in slot k & mask.  Consumers take tickets with a plain fetch-add (no CAS loop: with ~1,500 warps on
// one head word a CAS retries ~40 times per success); a consumer whose ticket is not served within a few polls marks the
// slot abandoned and moves on, and the producer that later draws that ticket clears the mark and draws another ticket.
// every lane calls; cls in {PH_ACT, PH_RESP, PH_DEAL, PH_SLOW, PH_TAIL} or PH_NONE
__device__ __forceinline__ void q_push(const Queues& q, int cls, int32_t gi) {
  const int lane = threadIdx.x & 31;
  #pragma unroll
  for (int c = 0; c < N_QUEUES; c++) {
    unsigned m = __ballot_sync(0xFFFFFFFFu, cls == c);
    if (m == 0) continue;
    int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(&q.ctl[Q_TAIL + 32 * c], (uint32_t)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (cls == c) {
      uint32_t t = base + __popc(m & ((1u << lane) - 1));
      while (true) {
        int32_t* sl = q.slots + (size_t)c * (q.mask + 1) + (t & q.mask);
        int32_t old = atomicCAS(sl, -1, gi);
        if (old == -1) break;
        if (old == -2) {                                   // the consumer of this ticket gave up: clear, draw a new ticket
          *reinterpret_cast<volatile int32_t*>(sl) = -1;
          t = atomicAdd(&q.ctl[Q_TAIL + 32 * c], 1u);
        }
        // old >= 0: previous lap not consumed yet (cap >= 2n: practically never) -> retry the same slot
      }
    }
  }
}
// q_push in two halves, for callers that can put work between them: the ticket (one warp-aggregated fetch-add per class, nothing
// is published by it) and the publication (the slot store, after the game's record has reached memory).
__device__ __forceinline__ uint32_t q_ticket(const Queues& q, int cls) {
  const int lane = threadIdx.x & 31;
  uint32_t t = 0;
  #pragma unroll
  for (int c = 0; c < N_QUEUES; c++) {
    unsigned m = __ballot_sync(0xFFFFFFFFu, cls == c);
    if (m == 0) continue;
    int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(&q.ctl[Q_TAIL + 32 * c], (uint32_t)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (cls == c) t = base + __popc(m & ((1u << lane) - 1));
  }
  return t;
}
__device__ __forceinline__ void q_publish(const Queues& q, int cls, uint32_t t, int32_t gi) {
  if (cls < 0 || cls >= N_QUEUES) return;
  while (true) {
    int32_t* sl = q.slots + (size_t)cls * (q.mask + 1) + (t & q.mask);
    int32_t old = atomicCAS(sl, -1, gi);
    if (old == -1) break;
    if (old == -2) {                                     // the consumer of this ticket gave up: clear, draw a new ticket
      *reinterpret_cast<volatile int32_t*>(sl) = -1;
      t = atomicAdd(&q.ctl[Q_TAIL + 32 * cls], 1u);
    }
  }
}
__global__ void q_init_kernel(const G* states, int64_t n, uint32_t* budget, uint32_t max_steps, Queues q, int init_dist) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int cls = PH_NONE;
  if (i < n) {
    budget[i] = max_steps;
    cls = classify(states[i], max_steps);
  }
  if (i < 256) {
    // initial class of every SM; shares follow the measured per-class warp time
    const int r = (int)i & 15;
    if (init_dist == 0) q.ctl[Q_SMCLASS + i] = r < 7 ? PH_ACT : r < 10 ? PH_TAIL : r < 13 ? PH_RESP : r < 15 ? PH_SLOW : PH_DEAL;
    else q.ctl[Q_SMCLASS + i] = r < 6 ? PH_ACT : r < 13 ? PH_TAIL : r < 15 ? PH_SLOW : PH_DEAL;   // measured warp-time shares; RESP rides on TAIL
  }
  unsigned m = __ballot_sync(0xFFFFFFFFu, cls != PH_NONE);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(&q.ctl[Q_LIVE], (uint32_t)__popc(m));
  q_push(q, cls, (int32_t)i);
}
// Fault injection for the watchdog test (RV_FAULT_INJECT=lost_game): one more live game than the queues hold
__global__ void q_fault_kernel(Queues q) { atomicAdd(&q.ctl[Q_LIVE], 1u); }
// lane 0 only: take up to PHB consumer tickets of class c if the queue looks non-empty; returns the count and the first ticket
__device__ __forceinline__ int q_claim(const Queues& q, int c, uint32_t& h, unsigned long long* dbg = nullptr, int cap = PHB) {
  uint32_t hh = ld_volatile_u32(&q.ctl[Q_HEAD + 32 * c]), tt = ld_volatile_u32(&q.ctl[Q_TAIL + 32 * c]);
  int avail = (int)(tt - hh);
  if (avail <= 0) {
    if (dbg) dbg[1]++;
    return 0;
  }
  int want = avail < cap ? avail : cap;
  h = atomicAdd(&q.ctl[Q_HEAD + 32 * c], (uint32_t)want);
  return want;
}
// The per-warp scheduler loop (one warp = 32 staged games, its own mbarrier): runs until no game is live.  `stage` / `mbar` are
// the warp's shared memory (initialised by the caller), `parity` the phase its mbarrier is in.
template <int NPC>   // seat count of every game of the vector (4, or 3 for sanma)
__device__ __forceinline__ void persistent_warp_loop(const Tables& T, G* states, int64_t n, uint32_t* log, uint32_t cap,
                                                     uint64_t agent_seed, uint32_t* budget, const Queues& q,
                                                     unsigned long long* counters, int reps, uint32_t endgame_live,
                                                     int endgame_take, uint32_t switch_knobs, unsigned char* stage, uint64_t* mbar,
                                                     uint32_t parity) {
  const int lane = threadIdx.x & 31;
  unsigned char* slot = stage + lane * STG_STRIDE;
  const uint32_t bar = smem_u32(mbar);
  // All warps of an SM work on the SAME class (its code stays in the SM's instruction cache: the per-class hot code is
  // tens of KB, the L1.5 I-cache 32 KB).  The class is a per-SM word in global memory; a warp that finds its SM's class
  // empty twice in a row moves the whole SM to the class with the longest queue.
  unsigned smid;
  asm("mov.u32 %0, %%smid;" : "=r"(smid));
  uint32_t* const my_class = &q.ctl[Q_SMCLASS + (smid & 255)];
  unsigned long long stepped_total = 0, finished_total = 0;
  uint32_t idle = 0;
  // Lane refill.  A game whose next step is again a plain turn (ACT -> ACT, the most frequent transition) stays staged in its
  // lane between iterations instead of going through store / push / claim / load: the warp keeps such games, pushes the ones
  // that leave the class and refills only the FREE lanes from the ACT queue.  Every iteration then runs act_fast on a full
  // warp (the `reps` loop alone decays: ~16 of 32 lanes active), and an ACT -> ACT step costs no queue round trip.
  // when an SM changes class: after `switch_idle` empty looks in a row, and only to a queue of at least `switch_minlen` games
  const uint32_t switch_idle = switch_knobs & 0xFFFF;
  const int switch_minlen = (int)(switch_knobs >> 16);
  const bool hold_enabled = (reps >> 8) & 1;
  reps &= 0xFF;
  int32_t held_gi = -1;                 // the game kept in this lane's slot (class ACT only), -1 = the lane is free
  uint32_t held_b = 0, held_b0 = 0;     // its budget now / as last read from memory
#ifdef RV_QPROF
  // per-warp cycle accounting: [0] idle polls, [1] claim+slots+fence, [2] stage in, [3..7] compute per class, [8] stage out+fence,
  // [9] push; [10..14] batches per class, [15..19] games per class
  unsigned long long prof[20] = {0};
  unsigned long long dbg[4] = {0, 0, 0, 0};   // CAS retries, empty looks, claims given up, SM class switches
#define QDBG dbg
  long long t0 = clock64();
#define QP(i) { long long t1 = clock64(); prof[i] += (unsigned long long)(t1 - t0); t0 = t1; }
#else
#define QP(i)
#define QDBG nullptr
#endif
  while (true) {
    int cls = PH_NONE, take = 0;
    uint32_t h = 0;
    // Endgame.  A hanchan is a chain of ~1,000 dependent steps and a queue visit is tens of microseconds, so once fewer
    // games are live than the crew can keep busy the rollout is bound by the latency of the longest games, not by
    // throughput (per-class accounting: 39 % of all warp cycles were idle polls).  Below `endgame_live` live games a warp
    // takes a few games from ANY class and plays them to the end in place, every transition inline, no queue round trips.
    // ONE round trip for everything the decision needs: lane 0 reads the live-game counter, lane 1 the SM's class, lane 2 the
    // error flag, lanes 4.. the head and lanes 12.. the tail of every class queue; the warp then decides uniformly and lane 0
    // draws the tickets.  (Read one after the other by lane 0 these were four dependent L2 round trips per visit.)
    bool endgame = false;
    uint32_t ctl_v = 0;
    {
      const uint32_t* a = nullptr;
      if (lane == 0) a = &q.ctl[Q_LIVE];
      else if (lane == 1) a = my_class;
      else if (lane == 2) a = &q.ctl[Q_ERR];
      else if (lane >= 4 && lane < 4 + N_QUEUES) a = &q.ctl[Q_HEAD + 32 * (lane - 4)];
      else if (lane >= 12 && lane < 12 + N_QUEUES) a = &q.ctl[Q_TAIL + 32 * (lane - 12)];
      if (a) ctl_v = ld_volatile_u32(a);
    }
    const uint32_t live_now = __shfl_sync(0xFFFFFFFFu, ctl_v, 0);
    const uint32_t err_now = __shfl_sync(0xFFFFFFFFu, ctl_v, 2);
    auto q_len = [&](int c) { return (int)(__shfl_sync(0xFFFFFFFFu, ctl_v, 12 + c) - __shfl_sync(0xFFFFFFFFu, ctl_v, 4 + c)); };
    endgame = live_now < endgame_live;
    const unsigned held_mask = __ballot_sync(0xFFFFFFFFu, held_gi >= 0);
    const int n_held = __popc(held_mask);
    const int sm_cls = (int)__shfl_sync(0xFFFFFFFFu, ctl_v, 1);
    // games may be kept (and free lanes refilled) while the SM still serves ACT and the rollout is not in its endgame
    const bool hold_ok = hold_enabled && !endgame && sm_cls == PH_ACT;
    {
      int want_cap = PHB;
      if (n_held > 0) {
        cls = PH_ACT;
        want_cap = hold_ok ? PHB - n_held : 0;
      } else if (endgame) {
        cls = PH_NONE;
        want_cap = endgame_take;
        for (int c = N_QUEUES - 1; c >= 0; c--)
          if (q_len(c) > 0) cls = c;                       // first non-empty class
      } else {
        cls = sm_cls;
        // claim windows reached outside a TAIL visit are few (0.5 % of the visits) and their code is part of the TAIL code:
        // the TAIL SMs serve them too, no SM is ever switched to RESP for them
        if (cls == PH_TAIL && q_len(PH_TAIL) <= 0 && q_len(PH_RESP) > 0) cls = PH_RESP;
        const bool empty = q_len(cls) <= 0;
#ifdef RV_QPROF
        if (empty && lane == 0) dbg[1]++;
#endif
        if (empty && idle >= switch_idle) {
          // move the SM: longest queue wins
          int best = -1, best_len = switch_minlen - 1;
          for (int c = 0; c < N_QUEUES; c++) {
            const int len = q_len(c);
            if (len > best_len) best_len = len, best = c;
          }
          if (best >= 0) {
            if (lane == 0) *reinterpret_cast<volatile uint32_t*>(my_class) = (uint32_t)best;
            cls = best;
            idle = 0;
#ifdef RV_QPROF
            dbg[3]++;
#endif
          } else {
            cls = PH_NONE;
          }
        } else if (empty) {
          cls = PH_NONE;
        }
      }
      if (cls != PH_NONE && want_cap > 0) {
        const int avail = q_len(cls);
        take = avail < want_cap ? avail : want_cap;
        if (take < 0) take = 0;
        if (take > 0) {
          if (lane == 0) h = atomicAdd(&q.ctl[Q_HEAD + 32 * cls], (uint32_t)take);
          h = __shfl_sync(0xFFFFFFFFu, h, 0);
        }
      }
    }
    if (take == 0 && n_held == 0) {
      if (live_now != 0 && err_now == 0 && ++idle > (1u << 22)) {     // watchdog: seconds without work while games are live
        if (lane == 0) atomicExch(&q.ctl[Q_ERR], 1u);
        break;
      }
      if (live_now == 0 || err_now != 0) break;
      __nanosleep(300);
      QP(0);
      continue;
    }
    idle = 0;
    int32_t gi = -1;
    // the free lanes draw the tickets, in lane order
    const int my_ticket = held_gi >= 0 ? PHB : __popc(~held_mask & ((1u << lane) - 1));
    if (my_ticket < take) {
      int32_t* sl = q.slots + (size_t)cls * (q.mask + 1) + ((h + (uint32_t)my_ticket) & q.mask);
      for (int poll = 0; poll < 6 && gi < 0; poll++) gi = *reinterpret_cast<volatile int32_t*>(sl);
      if (gi < 0) {
        gi = atomicCAS(sl, -1, -2);          // racing consumers over-claimed, or the producer is slow: abandon the ticket
#ifdef RV_QPROF
        if (gi < 0) dbg[2]++;
#endif
      }
      if (gi >= 0) *reinterpret_cast<volatile int32_t*>(sl) = -1;
    }
    const bool fresh = gi >= 0;                                     // a game taken from the queue in this iteration
    const unsigned fresh_mask = __ballot_sync(0xFFFFFFFFu, fresh);
    if (take > 0) asm volatile("fence.acq_rel.gpu;" ::: "memory");
    QP(1);
    if (fresh_mask == 0 && n_held == 0) continue;
    uint32_t b_in = 0;
    if (fresh_mask != 0) {
      if (lane == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)__popc(fresh_mask) * (uint32_t)RV_HOT_BYTES) : "memory");
      __syncwarp();
      if (fresh) {
        stage_in(slot, &states[gi], bar);
        *reinterpret_cast<G**>(slot + RV_HOT_BYTES) = &states[gi];
        b_in = budget[gi];          // in flight together with the bulk copy
      }
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
      parity ^= 1;
    }
    QP(2);
    if (!fresh && held_gi >= 0) gi = held_gi;
    const bool have = gi >= 0;
    take = __popc(__ballot_sync(0xFFFFFFFFu, have));
    int next = PH_NONE;
    bool leave = true;                                              // the game goes back to HBM and into a queue after this visit
    if (have) {
      G& g = *reinterpret_cast<G*>(slot);
      Ctx cx = make_ctx(T, log, cap, gi);
      cx.defer_init = true;
      cx.defer_tail = true;
      uint32_t b = fresh ? b_in : held_b;
      const uint32_t b0 = fresh ? b_in : held_b0;
      held_gi = -1;
      if (endgame && n_held == 0) {
        cx.defer_init = false;
        cx.defer_tail = false;
        while (true) {
          if (g.pending_tail[0] != RV_NONE) run_pending_tail(cx, g);          // parked by an earlier visit
          else if (g.pending_init[0] != RV_NONE) run_pending_init(cx, g);
          else if (g.is_done || b == 0) break;
          else {
            if (g.phase != RV_WAIT_ACT) random_step_resp(cx, g, agent_seed, g.seed);
            else if (!act_fast<false, NPC>(cx, g, agent_seed, g.seed)) random_step_act(cx, g, agent_seed, g.seed);
            b--;
          }
        }
        next = PH_NONE;
      } else if (cls == PH_ACT) {
        for (int r = 0; r < reps; r++) {
          if (!act_fast<false, NPC>(cx, g, agent_seed, g.seed)) {
            next = PH_SLOW;
            break;
          }
          b--;
          next = classify(g, b);
          if (next != PH_ACT) break;
        }
        if (next == PH_ACT && hold_ok) {       // stays in this lane: no store, no push, no claim, no load
          leave = false;
          held_gi = gi;
          held_b = b;
          held_b0 = b0;
        }
      } else {
        if (cls == PH_TAIL) {
          // the parked follow-up of a discard; when it opens a claim window the answers are taken in the same visit
          run_pending_tail(cx, g);
          if (b > 0 && !g.is_done && g.pending_init[0] == RV_NONE && g.phase == RV_WAIT_RESPONSE) {
            random_step_resp(cx, g, agent_seed, g.seed);
            b--;
          }
        } else if (cls == PH_DEAL) run_pending_init(cx, g);
        else if (cls == PH_SLOW) random_step_act(cx, g, agent_seed, g.seed), b--;
        else random_step_resp(cx, g, agent_seed, g.seed), b--;
        next = classify(g, b);
        // (taking the plain turn most games leave these visits with right here, instead of through the ACT queue, measured
        // slower: 1.36 -> 1.30 / 1.26 G steps/s for one / two inline turns — the lanes of a generic batch diverge again)
      }
      if (leave && b != b0) {
        budget[gi] = b;
        stepped_total += b0 - b;
        finished_total += g.is_done ? 1 : 0;
      }
    }
#ifdef RV_QPROF
    __syncwarp();
    QP(3 + cls);
    prof[10 + cls]++;
    prof[15 + cls] += take;
#endif
    const bool out = have && leave;
    if (out) {
      stage_out(&states[gi], slot);
      __threadfence();
    }
    __syncwarp();
    QP(8);
    if (__any_sync(0xFFFFFFFFu, out)) {
      unsigned retired = __ballot_sync(0xFFFFFFFFu, out && next == PH_NONE);
      q_push(q, out ? next : (int)PH_NONE, gi);
      if (lane == 0 && retired) atomicSub(&q.ctl[Q_LIVE], (uint32_t)__popc(retired));
    }
    QP(9);
  }
#ifdef RV_QPROF
  if (lane == 0)
  {
    for (int i = 0; i < 20; i++) atomicAdd(&counters[8 + i], prof[i]);
    for (int i = 0; i < 4; i++) atomicAdd(&counters[28 + i], dbg[i]);
  }
#endif
  for (int o = 16; o > 0; o >>= 1) {
    stepped_total += __shfl_down_sync(0xFFFFFFFFu, stepped_total, o);
    finished_total += __shfl_down_sync(0xFFFFFFFFu, finished_total, o);
  }
  if (lane == 0 && stepped_total) {
    atomicAdd(&counters[0], stepped_total);
    if (finished_total) atomicAdd(&counters[1], finished_total);
  }
}

template <int NPC>
__global__ void __launch_bounds__(PHB) rollout_persistent_kernel(Tables T, G* states, int64_t n, uint32_t* log, uint32_t cap,
                                                                 uint64_t agent_seed, uint32_t* budget, Queues q,
                                                                 unsigned long long* counters, int reps, uint32_t endgame_live,
                                                                 int endgame_take, uint32_t switch_knobs) {
  __shared__ __align__(128) unsigned char stage[PHB * STG_STRIDE];
  __shared__ __align__(8) uint64_t mbar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  persistent_warp_loop<NPC>(T, states, n, log, cap, agent_seed, budget, q, counters, reps, endgame_live, endgame_take, switch_knobs,
                            stage, &mbar, 0u);
}

// ---- the crew kernel: ONE block per SM, its warps in step ---------------------------------------------------------------
// ncu on the per-warp scheduler (profiles/r02i_*): a batch visit executes ~7.5 k warp instructions = ~1,000 instruction-cache
// lines and ~420 of them MISS the SM's instruction cache (hit rate 60 %; the GPC-level cache behind it runs at 70 % of its
// request rate); at ~200 cycles a miss that IS the 55-130 k cycles of a visit.  The 12 one-warp blocks of an SM work on the same
// class but each at its own point of 40-60 KB of code, so they evict each other's lines.  Here the warps of an SM form one block:
// warp 0 reads the queue heads once and draws the tickets for the whole block, a block barrier starts the visit, and the warps
// then run the same class's code at the same time — a line is fetched once and used by all of them while it is resident.
// A short queue fills the first warps completely instead of every warp partially.  When the rollout reaches its endgame (or
// nothing is live any more) the warps leave the barrier loop and finish in persistent_warp_loop, each on its own.
struct CrewCtl {
  uint32_t mode, cls, take, h;      // mode: 0 = visit, 1 = nothing to do right now, 2 = leave the lock-step phase
};
template <int NPC>
__global__ void __launch_bounds__(384, 1) rollout_crew_kernel(Tables T, G* states, int64_t n, uint32_t* log, uint32_t cap,
                                                              uint64_t agent_seed, uint32_t* budget, Queues q,
                                                              unsigned long long* counters, int reps_knobs, uint32_t endgame_live,
                                                              int endgame_take, uint32_t switch_knobs) {
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  unsigned char* stage = dyn_smem + (size_t)warp * PHB * STG_STRIDE;
  uint64_t* mbars = reinterpret_cast<uint64_t*>(dyn_smem + (size_t)n_warps * PHB * STG_STRIDE);
  CrewCtl* ctl2 = reinterpret_cast<CrewCtl*>(mbars + n_warps);            // double-buffered by iteration parity
  uint64_t* mbar = mbars + warp;
  unsigned char* slot = stage + lane * STG_STRIDE;
  const uint32_t bar = smem_u32(mbar);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t parity = 0;
  unsigned smid;
  asm("mov.u32 %0, %%smid;" : "=r"(smid));
  uint32_t* const my_class = &q.ctl[Q_SMCLASS + (smid & 255)];
  const uint32_t switch_idle = switch_knobs & 0xFFFF;
  const int switch_minlen = (int)(switch_knobs >> 16);
  const int reps = reps_knobs & 0xFF;
  const int slot_polls = (reps_knobs >> 16) & 0xFF;      // looks at an empty slot before the ticket is abandoned
  const int follow_reps = (reps_knobs >> 24) & 0xF;      // plain turns taken at the end of a non-ACT visit
  unsigned long long stepped_total = 0, finished_total = 0;
  uint32_t idle = 0;
#ifdef RV_QPROF
  // per-warp cycles of the lock-step phase: [0] decision (warp 0), [1] barrier wait, [2] idle sleeps, [3] iterations without a ticket
  // (counts), [4] iterations (counts), [5] class changes (warp 0), [6] visit (claim .. push), [7] visits (counts)
  unsigned long long cp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long ct0 = clock64();
#define CQP(i) { long long t1 = clock64(); cp[i] += (unsigned long long)(t1 - ct0); ct0 = t1; }
#else
#define CQP(i)
#endif
  // warp 0's decision: (1) ONE load instruction for everything it needs (lane 0 the live-game counter, lane 1 the SM's class,
  // lane 2 the error flag, lanes 4.. / 12.. head and tail of every class queue), (2) the choice and the tickets.
  auto ctl_load = [&]() -> uint32_t {
    const uint32_t* a = nullptr;
    if (lane == 0) a = &q.ctl[Q_LIVE];
    else if (lane == 1) a = my_class;
    else if (lane == 2) a = &q.ctl[Q_ERR];
    else if (lane >= 4 && lane < 4 + N_QUEUES) a = &q.ctl[Q_HEAD + 32 * (lane - 4)];
    else if (lane >= 12 && lane < 12 + N_QUEUES) a = &q.ctl[Q_TAIL + 32 * (lane - 12)];
    return a ? ld_volatile_u32(a) : 0u;
  };
  auto decide = [&](uint32_t ctl_v, CrewCtl& out) {        // warp 0, all lanes; `out` is a register copy, published by lane 0 later
    const uint32_t live_now = __shfl_sync(0xFFFFFFFFu, ctl_v, 0);
    const uint32_t err_now = __shfl_sync(0xFFFFFFFFu, ctl_v, 2);
    auto q_len = [&](int c) { return (int)(__shfl_sync(0xFFFFFFFFu, ctl_v, 12 + c) - __shfl_sync(0xFFFFFFFFu, ctl_v, 4 + c)); };
    uint32_t mode = 0, take = 0, h = 0;
    int cls = (int)__shfl_sync(0xFFFFFFFFu, ctl_v, 1);
    if (live_now < endgame_live || live_now == 0 || err_now != 0) {
      mode = 2;
    } else {
      if (cls == PH_TAIL && q_len(PH_TAIL) <= 0 && q_len(PH_RESP) > 0) cls = PH_RESP;
      bool empty = q_len(cls) <= 0;
      // greedy variant (bit 9 of reps_knobs): a queue that cannot fill the block is left for one at least twice as long
      const bool short_q = ((reps_knobs >> 9) & 1) && q_len(cls) < (int)blockDim.x;
      if ((empty && idle >= switch_idle) || short_q) {
        int best = -1, best_len = short_q && !empty ? 2 * q_len(cls) : switch_minlen - 1;
        for (int c = 0; c < N_QUEUES; c++) {
          const int len = q_len(c);
          if (len > best_len) best_len = len, best = c;
        }
        if (best >= 0) {
          if (lane == 0) *reinterpret_cast<volatile uint32_t*>(my_class) = (uint32_t)best;
          cls = best;
          idle = 0;
          empty = false;
#ifdef RV_QPROF
          cp[5]++;
#endif
        }
      }
      if (empty) {
        mode = 1;
        if (++idle > (1u << 22)) {                      // watchdog, as in the per-warp loop
          if (lane == 0) atomicExch(&q.ctl[Q_ERR], 1u);
          mode = 2;
        }
      } else {
        idle = 0;
        // A blind fetch-add, not a compare-and-swap on exact tickets: deciders that race for a short queue over-claim, and the
        // tickets beyond the tail are RESERVATIONS — the next games pushed go straight into slots somebody is already polling.
        // Measured (profiles/r02p_ab_rollout.txt): exact tickets 2.05 G env steps/s, fetch-add 2.35 G.
        const int avail = q_len(cls), cap_all = (int)blockDim.x;
        take = (uint32_t)(avail < cap_all ? avail : cap_all);
        if (lane == 0) h = atomicAdd(&q.ctl[Q_HEAD + 32 * cls], take);
      }
    }
    out.mode = mode;
    out.cls = (uint32_t)cls;
    out.take = take;
    out.h = h;                                             // lane 0's copy holds the first ticket
  };
  // (Taking the decision one iteration ahead — loads issued before warp 0 polls its own slots, tickets drawn while its
  // records are in flight — was measured slower: 1.91 against 2.32 G env steps/s; the queue lengths it acts on are an
  // iteration old and the SMs over-claim by whole blocks.  profiles/r02o_*.)
  if (warp == 0) {                                         // the first decision
    CrewCtl d;
    decide(ctl_load(), d);
    if (lane == 0) ctl2[0] = d;
  }
  for (uint32_t iter = 0;; iter++) {
    CrewCtl* cc = ctl2 + (iter & 1);
    CrewCtl* cc_next = ctl2 + ((iter + 1) & 1);
    CQP(0);
    __syncthreads();
    CQP(1);
#ifdef RV_QPROF
    cp[4]++;
#endif
    const uint32_t mode = cc->mode;
    if (mode == 2) break;
    CrewCtl nd;
    if (mode == 1) {
      __nanosleep(300);
      if (warp == 0) {
        decide(ctl_load(), nd);
        if (lane == 0) *cc_next = nd;
      }
      CQP(2);
      continue;
    }
    const int cls = (int)cc->cls;
    const uint32_t take = cc->take, h = cc->h;
    if ((uint32_t)warp * 32u >= take) {                   // a short queue fills the first warps; the others wait at the barrier
#ifdef RV_QPROF
      cp[3]++;
#endif
      continue;                                            // (never warp 0: take > 0 here)
    }
    int32_t gi = -1;
    if (threadIdx.x < take) {
      int32_t* sl = q.slots + (size_t)cls * (q.mask + 1) + ((h + threadIdx.x) & q.mask);
      for (int poll = 0; poll < slot_polls && gi < 0; poll++) gi = *reinterpret_cast<volatile int32_t*>(sl);
      if (gi < 0) gi = atomicCAS(sl, -1, -2);
      if (gi >= 0) *reinterpret_cast<volatile int32_t*>(sl) = -1;
#ifdef RV_CREW_DEBUG
      if (gi >= n || cls < 0 || cls >= N_QUEUES) {
        printf("crew: bad ticket: block %d thread %d iter %u cls %d take %u h %u gi %d\n", blockIdx.x, threadIdx.x, iter, cls, take, h, gi);
        gi = -1;
      }
#endif
    }
    const bool have = gi >= 0;
    const unsigned have_mask = __ballot_sync(0xFFFFFFFFu, have);
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    if (have_mask == 0) {
      if (warp == 0) {
        decide(ctl_load(), nd);
        if (lane == 0) *cc_next = nd;
      }
      continue;
    }
    if (lane == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)__popc(have_mask) * (uint32_t)RV_HOT_BYTES) : "memory");
    __syncwarp();
    uint32_t b = 0;
    if (have) {
      stage_in(slot, &states[gi], bar);
      *reinterpret_cast<G**>(slot + RV_HOT_BYTES) = &states[gi];
      b = budget[gi];
    }
    {
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
      parity ^= 1;
    }
    int next = PH_NONE;
    if (have) {
      G& g = *reinterpret_cast<G*>(slot);
      Ctx cx = make_ctx(T, log, cap, gi);
      cx.defer_init = true;
      cx.defer_tail = true;
      const uint32_t b0 = b;
      if (cls == PH_ACT) {
        for (int r = 0; r < reps; r++) {
          if (!act_fast<false, NPC>(cx, g, agent_seed, g.seed)) {
            next = PH_SLOW;
            break;
          }
          b--;
          next = classify(g, b);
          if (next != PH_ACT) break;
        }
      } else {
        if (cls == PH_TAIL) {
          run_pending_tail(cx, g);
          if (b > 0 && !g.is_done && g.pending_init[0] == RV_NONE && g.phase == RV_WAIT_RESPONSE) {
            random_step_resp(cx, g, agent_seed, g.seed);
            b--;
          }
        } else if (cls == PH_DEAL) run_pending_init(cx, g);
        else if (cls == PH_SLOW) random_step_act(cx, g, agent_seed, g.seed), b--;
        else random_step_resp(cx, g, agent_seed, g.seed), b--;
        next = classify(g, b);
        // most games leave these visits with a plain turn next: take up to `follow_reps` of them here instead of through the
        // ACT queue (a queue round trip is a whole iteration of latency for the game)
        for (int r = 0; r < follow_reps && next == PH_ACT; r++) {
          if (!act_fast<false, NPC>(cx, g, agent_seed, g.seed)) {
            next = PH_SLOW;
            break;
          }
          b--;
          next = classify(g, b);
        }
      }
      if (b != b0) {
        budget[gi] = b;
        stepped_total += b0 - b;
        finished_total += g.is_done ? 1 : 0;
      }
    }
    __syncwarp();
    {
      // the ticket is drawn BEFORE the record is stored (its round trip runs under the bulk store), the game is published after
      const int dest = have ? next : (int)PH_NONE;
      const uint32_t ticket = q_ticket(q, dest);
      if (have) {
        stage_out(&states[gi], slot);
        __threadfence();
      }
      unsigned retired = __ballot_sync(0xFFFFFFFFu, have && next == PH_NONE);
      q_publish(q, dest, ticket, gi);
      if (lane == 0 && retired) atomicSub(&q.ctl[Q_LIVE], (uint32_t)__popc(retired));
    }
    if (warp == 0) {                                       // the decision for the next iteration, published by the barrier
      decide(ctl_load(), nd);
      if (lane == 0) *cc_next = nd;
    }
    CQP(6);
#ifdef RV_QPROF
    cp[7]++;
#endif
  }
#ifdef RV_QPROF
  if (lane == 0)
    for (int i = 0; i < 6; i++) atomicAdd(&counters[2 + i], i == 0 ? cp[1] : i == 1 ? cp[4] : i == 2 ? cp[3] : i == 3 ? cp[5] : i == 4 ? cp[2] : cp[6]);
#endif
  for (int o = 16; o > 0; o >>= 1) {
    stepped_total += __shfl_down_sync(0xFFFFFFFFu, stepped_total, o);
    finished_total += __shfl_down_sync(0xFFFFFFFFu, finished_total, o);
  }
  if (lane == 0 && stepped_total) {
    atomicAdd(&counters[0], stepped_total);
    if (finished_total) atomicAdd(&counters[1], finished_total);
  }
  // endgame / termination: every warp on its own
  persistent_warp_loop<NPC>(T, states, n, log, cap, agent_seed, budget, q, counters, reps_knobs & 0xFF, endgame_live, endgame_take,
                            switch_knobs, stage, mbar, parity);
}

__global__ void legal_kernel(Tables T, const G* states, int64_t n, rv_action* out, uint8_t* counts) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const G& g = states[i];
  Ctx cx = make_ctx(T, nullptr, 0, i);
  for (int p = 0; p < MAXP; p++) {
    bool owes = !g.is_done && ((g.phase == RV_WAIT_ACT && g.current_player == p) ||
                               (g.phase == RV_WAIT_RESPONSE && ((g.active_mask >> p) & 1)));
    int cnt = 0;
    if (owes) {
      uint32_t packed[RV_MAX_LEGAL];
      cnt = legal_actions(cx, g, p, packed, -1, nullptr);
      if (cnt > RV_MAX_LEGAL) cnt = RV_MAX_LEGAL;
      for (int k = 0; k < cnt; k++) out[((size_t)i * MAXP + p) * RV_MAX_LEGAL + k] = expand_act(g, p, packed[k]);
    }
    counts[(size_t)i * MAXP + p] = (uint8_t)cnt;
  }
}

__global__ void step_kernel(Tables T, G* states, int64_t n, uint32_t* log, uint32_t cap, const rv_action* actions,
                            unsigned long long* counters) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G& g = states[i];
  if (g.is_done) return;
  Ctx cx = make_ctx(T, log, cap, i);
  rv_action acts[MAXP];
  for (int p = 0; p < MAXP; p++) {
    acts[p] = actions[(size_t)i * MAXP + p];
    if (acts[p].type != RV_NO_ACTION) {   // Action::new sorts consume_tiles (action.rs:97-98)
      int nc = acts[p].n_consume > 4 ? 4 : acts[p].n_consume;
      for (int x = 1; x < nc; x++)
        for (int y = x; y > 0 && acts[p].consume[y - 1] > acts[p].consume[y]; y--) {
          uint8_t t = acts[p].consume[y]; acts[p].consume[y] = acts[p].consume[y - 1]; acts[p].consume[y - 1] = t;
        }
    }
  }
  g.step_count++;
  atomicAdd(&counters[0], 1ull);
  for (int p = 0; p < MAXP; p++) {
    if (acts[p].type == RV_NO_ACTION) continue;
    uint32_t packed[RV_MAX_LEGAL];
    int cnt = legal_actions(cx, g, p, packed, -1, nullptr);
    if (cnt > RV_MAX_LEGAL) cnt = RV_MAX_LEGAL;
    bool ok = false;
    for (int k = 0; k < cnt && !ok; k++) ok = action_matches(expand_act(g, p, packed[k]), acts[p]);
    if (!ok) {
      g.last_error = (uint8_t)p;
      trigger_ryukyoku(cx, g, RV_RK_ILLEGAL_BASE + p);
      if (g.is_done) atomicAdd(&counters[1], 1ull);
      return;
    }
  }
  step_apply(cx, g, acts);
  if (g.is_done) atomicAdd(&counters[1], 1ull);
}

// ---- observation encoding ------------------------------------------------------------------
__global__ void obs_count_kernel(const G* states, int64_t n, int32_t* counts) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  int c = 0;
  if (i < n && !states[i].is_done) c = __popc(states[i].active_mask & 0xF);
  counts[i] = c;   // counts[n] = 0 so that offsets[n] is the total
}

// Action-id sets (Observation::mask as a 96-bit set) of every seat that owes an action: one thread per game runs the
// generic legal-action enumeration once; obs_encode_kernel expands the bits into mask bytes.
// AVAIL (4P, rv_vec_encode_ext): bits 82..92 of the set carry the 11 action-availability flags of encode.rs:399-430.
constexpr int OBS_AVAIL_SHIFT = 18;   // bit 82 = word 2, bit 18
template <bool SANMA, bool AVAIL = false>
__global__ void __launch_bounds__(128) legal_ids_kernel(Tables T, const G* states, int64_t n, uint32_t* idbits) {
  int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= n) return;
  const G& g = states[gi];
  if (g.is_done) return;
  Ctx cx = make_ctx(T, nullptr, 0, gi);
  for (int pid = 0; pid < MAXP; pid++) {
    if (!((g.active_mask >> pid) & 1)) continue;
    uint32_t packed[RV_MAX_LEGAL];
    int cnt = legal_actions(cx, g, pid, packed, -1, nullptr);
    if (cnt > RV_MAX_LEGAL) cnt = RV_MAX_LEGAL;
    uint32_t bits[3] = {0, 0, 0};
    for (int k = 0; k < cnt; k++) {
      const rv_action a = expand_act(g, pid, packed[k]);
      obs_id_set<SANMA>(bits, a);
      if (AVAIL) bits[2] |= obs_avail_bit(a) << OBS_AVAIL_SHIFT;
    }
    uint32_t* o = idbits + ((size_t)gi * MAXP + pid) * 3;
    o[0] = bits[0], o[1] = bits[1], o[2] = bits[2];
  }
}
// One warp per game; for each seat that owes an action the warp writes the row with obs_encode_warp (obs.cuh) and
// expands the seat's action-id set into the 82 (sanma 60) mask bytes.
// The warp first copies the record's hot prefix and the rivers into shared memory with independent 16- and 8-byte loads (at
// most two per lane, all in flight together): read field by field the record costs a chain of ~8 dependent DRAM
// round trips per warp — the records are not L2-resident here, every step streams 0.67 GB of rows through the L2 — and that
// chain, not issue slots or HBM bandwidth, bounded the kernel (ncu: 2.9 TB/s of writes at 40 % DRAM utilisation, and a 29 %
// cut of the instruction count left the duration unchanged).
// (A persistent variant — 8 blocks per SM, each warp walking several games with the next record prefetched into registers —
// measured slower, 194 us against 157 us per 65,536 rows: with one short-lived warp per game the block scheduler keeps
// every SM topped up and other warps cover the one remaining round trip.)
constexpr int OBS_STAGE_BYTES = RV_HOT_BYTES + MAXP * RV_RIVER_CAP;   // 544 + 128
// Snapshot of what the encoder reads of every record (hot prefix + rivers, OBS_STAGE_BYTES per game, packed): taken before an
// env step so that the rows of the decision point can be streamed out WHILE the step kernel already advances the games
// (rv_vec_observe_step_random: the encoder is HBM-bound, the one-step kernel latency-bound — side by side they cost the longer
// of the two instead of the sum).
__global__ void __launch_bounds__(256) obs_snapshot_kernel(const G* states, int64_t n, unsigned char* snap) {
  const int64_t gi = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (gi >= n) return;
  const uint4* src = reinterpret_cast<const uint4*>(&states[gi]);
  const uint2* riv = reinterpret_cast<const uint2*>(&states[gi].river[0][0]);
  uint4* dst = reinterpret_cast<uint4*>(snap + (size_t)gi * OBS_STAGE_BYTES);
  __stcs(dst + lane, __ldcs(src + lane));
  if (lane < (RV_HOT_BYTES - 512) / 16) __stcs(dst + 32 + lane, __ldcs(src + 32 + lane));
  else if (lane >= 16) __stcs(reinterpret_cast<uint2*>(snap + (size_t)gi * OBS_STAGE_BYTES + RV_HOT_BYTES) + (lane - 16), __ldcs(riv + (lane - 16)));
}
template <bool SANMA>
__global__ void __launch_bounds__(128, 8) obs_encode_kernel(const G* states, int64_t n, const int32_t* offsets, const uint32_t* idbits,
                                                            float* obs, uint8_t* mask, int32_t* index, int64_t max_obs,
                                                            const unsigned char* snap = nullptr) {
  constexpr int W = SANMA ? OBS_W3 : OBS_W, IDS = SANMA ? OBS_IDS3 : OBS_IDS;
  __shared__ ObsScratch scratch[4];
  __shared__ __align__(16) unsigned char staged[4][OBS_STAGE_BYTES];
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t gi = (int64_t)blockIdx.x * 4 + w;
  if (gi >= n) return;
  {
    // source: the record itself, or its packed snapshot (hot prefix followed by the rivers)
    const uint4* src = snap ? reinterpret_cast<const uint4*>(snap + (size_t)gi * OBS_STAGE_BYTES) : reinterpret_cast<const uint4*>(&states[gi]);
    const uint2* riv = snap ? reinterpret_cast<const uint2*>(snap + (size_t)gi * OBS_STAGE_BYTES + RV_HOT_BYTES)
                            : reinterpret_cast<const uint2*>(&states[gi].river[0][0]);   // offset 776: 8-byte aligned only
    uint4* dst = reinterpret_cast<uint4*>(staged[w]);
    const uint4 a = __ldcs(src + lane);                                      // hot bytes 0..511
    uint4 b = make_uint4(0, 0, 0, 0);
    uint2 r = make_uint2(0, 0);
    if (lane < (RV_HOT_BYTES - 512) / 16) b = __ldcs(src + 32 + lane);      // hot bytes 512..RV_HOT_BYTES-1
    else if (lane >= 16) r = __ldcs(riv + (lane - 16));                      // rivers, 16 x 8 bytes
    dst[lane] = a;
    if (lane < (RV_HOT_BYTES - 512) / 16) dst[32 + lane] = b;
    else if (lane >= 16) reinterpret_cast<uint2*>(staged[w] + RV_HOT_BYTES)[lane - 16] = r;
  }
  int row = offsets[gi];
  __syncwarp();
  const G& g = *reinterpret_cast<const G*>(staged[w]);     // hot fields only; the rivers are passed separately
  const uint8_t* river = staged[w] + RV_HOT_BYTES;
  if (g.is_done) return;
  for (int pid = 0; pid < MAXP; pid++) {
    if (!((g.active_mask >> pid) & 1)) continue;
    if (row >= max_obs) break;
    if (obs) obs_encode_warp<SANMA>(g, river, pid, obs + (size_t)row * (OBS_CH * W), scratch[w], lane);
    if (mask) obs_mask_row_warp<SANMA>(idbits + ((size_t)gi * MAXP + pid) * 3, mask + (size_t)row * IDS, lane);
    if (index && lane == 0) index[row] = (int32_t)(gi * 4 + pid);
    row++;
  }
}

// Extended rows (Observation::encode_extended, 215 x 34): one warp per game, staged like obs_encode_kernel; obs_ext.cuh.
template <bool SANMA>
__global__ void __launch_bounds__(128, 6) obs_ext_kernel(Tables T, DecayTab D, const G* states, int64_t n, const int32_t* offsets,
                                                      const uint32_t* idbits, float* obs, uint8_t* mask, int32_t* index, int64_t max_obs) {
  __shared__ ObsScratch scratch[4];
  __shared__ ObsExtScratch xscratch[4];
  __shared__ __align__(16) unsigned char staged[4][OBS_STAGE_BYTES];
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t gi = (int64_t)blockIdx.x * 4 + w;
  if (gi >= n) return;
  {
    const uint4* src = reinterpret_cast<const uint4*>(&states[gi]);
    const uint2* riv = reinterpret_cast<const uint2*>(&states[gi].river[0][0]);
    uint4* dst = reinterpret_cast<uint4*>(staged[w]);
    const uint4 a = __ldcs(src + lane);
    uint4 b = make_uint4(0, 0, 0, 0);
    uint2 r = make_uint2(0, 0);
    if (lane < (RV_HOT_BYTES - 512) / 16) b = __ldcs(src + 32 + lane);
    else if (lane >= 16) r = __ldcs(riv + (lane - 16));
    dst[lane] = a;
    if (lane < (RV_HOT_BYTES - 512) / 16) dst[32 + lane] = b;
    else if (lane >= 16) reinterpret_cast<uint2*>(staged[w] + RV_HOT_BYTES)[lane - 16] = r;
  }
  int row = offsets[gi];
  __syncwarp();
  const G& g = *reinterpret_cast<const G*>(staged[w]);
  const uint8_t* river = staged[w] + RV_HOT_BYTES;
  if (g.is_done) return;
  for (int pid = 0; pid < MAXP; pid++) {
    if (!((g.active_mask >> pid) & 1)) continue;
    if (row >= max_obs) break;
    const uint32_t* bits = idbits + ((size_t)gi * MAXP + pid) * 3;
    if (obs) {
      if constexpr (SANMA)
        obs_ext3_encode_warp(T, D, g, states[gi], river, pid, (bits[2] >> OBS_AVAIL_SHIFT) & 0x7FFu, obs + (size_t)row * OBSX3_FLOATS,
                             scratch[w], xscratch[w], lane);
      else
        obs_ext_encode_warp(T, D, g, states[gi], river, pid, (bits[2] >> OBS_AVAIL_SHIFT) & 0x7FFu, obs + (size_t)row * (OBSX_CH * OBS_W),
                            scratch[w], xscratch[w], lane);
    }
    if (mask) obs_mask_row_warp<SANMA>(bits, mask + (size_t)row * (SANMA ? OBS_IDS3 : OBS_IDS), lane);
    if (index && lane == 0) index[row] = (int32_t)(gi * 4 + pid);
    row++;
  }
}

// Kawa overview rows (Observation::encode_kawa_overview, 4 x 7 x 34; sanma 3 x 7 x 27): one warp per game builds the block in
// shared memory (one lane per seat walks that seat's river) and streams it out once per acting seat — the block does not
// depend on the observer.
template <bool SANMA>
__global__ void __launch_bounds__(128) obs_kawa_kernel(const G* states, int64_t n, const int32_t* offsets, float* out, int32_t* index,
                                                       int64_t max_obs) {
  constexpr int FL = SANMA ? KAWA_FLOATS3 : KAWA_FLOATS, W = SANMA ? OBS_W3 : OBS_W, NPL = SANMA ? 3 : 4;
  __shared__ __align__(16) float blk[4][KAWA_FLOATS];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t gi = (int64_t)blockIdx.x * 4 + w;
  if (gi >= n) return;
  const G& g = states[gi];
  if (g.is_done || (g.active_mask & 0xF) == 0) return;
  for (int i = lane; i < FL; i += 32) blk[w][i] = 0.f;
  __syncwarp();
  if (lane < NPL) obs_kawa_seat<SANMA>(g, &g.river[0][0], lane, blk[w] + lane * 7 * W);
  __syncwarp();
  int row = offsets[gi];
  for (int pid = 0; pid < NPL; pid++) {
    if (!((g.active_mask >> pid) & 1)) continue;
    if (row >= max_obs) break;
    if (out) {
      float* o = out + (size_t)row * FL;               // a sanma row (2,268 B) is only 4-byte aligned
      if (SANMA) {
        for (int i = lane; i < FL; i += 32) __stcs(o + i, blk[w][i]);
      } else {
        const float4* b4 = reinterpret_cast<const float4*>(blk[w]);
        for (int i = lane; i < FL / 4; i += 32) __stcs(reinterpret_cast<float4*>(o) + i, b4[i]);
      }
    }
    if (index && lane == 0) index[row] = (int32_t)(gi * 4 + pid);
    row++;
  }
}

// Mask rows from the id sets a one-step rollout left in HBM: one warp per row, (game, seat) from the row index.
template <bool SANMA>
__global__ void __launch_bounds__(256) obs_mask_rows_kernel(const int32_t* index, const int32_t* total, const uint32_t* idbits, uint8_t* mask,
                                                            int64_t max_obs) {
  constexpr int IDS = SANMA ? OBS_IDS3 : OBS_IDS;
  const int64_t rows = min((int64_t)*total, max_obs);   // the row count is only known on the device: grid-stride over it
  for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * 8) {
    const int gs = index[row];                       // game * 4 + seat
    obs_mask_row_warp<SANMA>(idbits + (size_t)gs * 3, mask + (size_t)row * IDS, threadIdx.x & 31);
  }
}

// Observe + step, fused (BASELINE config 5: a rollout that emits FEATURE_ENCODING tensors and masks at every step).
// One warp owns 32 games.  Their hot prefixes are staged in shared memory with one bulk copy per game; then
//   1. the warp walks its games and writes the tensor row of every seat that owes an action (obs_encode_warp), all lanes
//      streaming one row at a time, straight from the staged state;
//   2. every lane takes its own game's env step with the on-device agent (random_step<true>): the routine that builds the
//      legal list to pick from leaves the list's action-id set in shared memory;
//   3. the warp expands those sets into the mask rows;
//   4. lanes whose round ended deal the next one together, and the records are stored back with one bulk copy each.
// The state is read once and written once per step; the observation rows are the only other HBM traffic.
template <bool SANMA>
__global__ void __launch_bounds__(32) observe_step_kernel(Tables T, G* states, int64_t n, uint32_t* log, uint32_t cap, uint64_t agent_seed,
                                                          const int32_t* offsets, float* obs, uint8_t* mask, int32_t* index,
                                                          int64_t max_obs, unsigned long long* counters) {
  constexpr int W = SANMA ? OBS_W3 : OBS_W, IDS = SANMA ? OBS_IDS3 : OBS_IDS;
  __shared__ __align__(128) unsigned char stage[32 * STG_STRIDE];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ ObsScratch scratch;
  __shared__ uint32_t s_bits[32][MAXP][3];
  const int lane = threadIdx.x;
  const int64_t gi = (int64_t)blockIdx.x * 32 + lane;
  const bool have = gi < n;
  unsigned char* slot = stage + lane * STG_STRIDE;
  const uint32_t bar = smem_u32(&mbar);
  const int take = (int)min((int64_t)32, n - (int64_t)blockIdx.x * 32);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)take * (uint32_t)RV_HOT_BYTES) : "memory");
  }
  __syncwarp();
  if (have) {
    stage_in(slot, &states[gi], bar);
    *reinterpret_cast<G**>(slot + RV_HOT_BYTES) = &states[gi];
  }
  {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(ok)
                   : "r"(bar), "r"(0u)
                   : "memory");
  }
  __syncwarp();
  G& g = *reinterpret_cast<G*>(slot);
  const bool live = have && !g.is_done;
  const uint32_t am = live ? (g.active_mask & 0xFu) : 0u;
  const int row0 = have ? offsets[gi] : 0;
  // 1. tensor rows, one at a time, all lanes
  const unsigned any_obs = __ballot_sync(0xFFFFFFFFu, am != 0);
  if (obs || index) {
    for (unsigned rest = any_obs; rest; rest &= rest - 1) {
      const int src = __ffs(rest) - 1;
      uint32_t am_s = __shfl_sync(0xFFFFFFFFu, am, src);
      int row = __shfl_sync(0xFFFFFFFFu, row0, src);
      const int64_t gi_s = (int64_t)blockIdx.x * 32 + src;
      const G& gs = *reinterpret_cast<const G*>(stage + src * STG_STRIDE);
      for (; am_s; am_s &= am_s - 1, row++) {
        const int pid = __ffs(am_s) - 1;
        if (row >= max_obs) break;
        if (obs) obs_encode_warp<SANMA>(gs, &states[gi_s].river[0][0], pid, obs + (size_t)row * (OBS_CH * W), scratch, lane);
        if (index && lane == 0) index[row] = (int32_t)(gi_s * 4 + pid);
      }
    }
  }
  // 2. the env step (leaves the id sets of the acting seats in s_bits)
  Ctx cx = make_ctx(T, log, cap, have ? gi : 0);
  cx.defer_init = true;
  cx.idbits = &s_bits[lane][0][0];
  unsigned long long stepped = 0, finished = 0;
  if (live) {
    random_step<true>(cx, g, agent_seed, g.seed);
    stepped = 1;
  }
  __syncwarp();
  // 3. mask rows
  if (mask) {
    for (unsigned rest = any_obs; rest; rest &= rest - 1) {
      const int src = __ffs(rest) - 1;
      uint32_t am_s = __shfl_sync(0xFFFFFFFFu, am, src);
      int row = __shfl_sync(0xFFFFFFFFu, row0, src);
      for (; am_s; am_s &= am_s - 1, row++) {
        const int pid = __ffs(am_s) - 1;
        if (row >= max_obs) break;
        obs_mask_row_warp<SANMA>(s_bits[src][pid], mask + (size_t)row * IDS, lane);
      }
    }
  }
  // 4. deal the rounds that ended in this step (convergent over the lanes that need it), store back
  if (live) {
    if (g.pending_init[0] != RV_NONE) run_pending_init(cx, g);
    finished = g.is_done ? 1 : 0;
    stage_out(&states[gi], slot);
  }
  for (int o = 16; o > 0; o >>= 1) {
    stepped += __shfl_down_sync(0xFFFFFFFFu, stepped, o);
    finished += __shfl_down_sync(0xFFFFFFFFu, finished, o);
  }
  if (lane == 0 && stepped) {
    atomicAdd(&counters[0], stepped);
    if (finished) atomicAdd(&counters[1], finished);
  }
}


// ---- sequence features: one thread per game, one row per seat that owes an action ----------------------------------
__global__ void seq_encode_kernel(Tables T, const G* states, int64_t n, const uint32_t* log, uint32_t cap, const int32_t* offsets,
                                  const uint32_t* start, uint32_t* cursor, int game_style, uint16_t* sparse, float* numeric,
                                  uint16_t* prog, int max_prog, uint16_t* cand, uint16_t* lens, int32_t* index, int64_t max_obs) {
  int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= n) return;
  const G& g = states[gi];
  if (g.is_done) return;
  Ctx cx = make_ctx(T, nullptr, 0, gi);
  const uint32_t* glog = log + (size_t)gi * cap;
  const uint32_t end = g.ev_words < cap ? g.ev_words : cap;
  int64_t row = offsets[gi];
  for (int pid = 0; pid < MAXP; pid++) {
    if (!((g.active_mask >> pid) & 1)) continue;
    if (row < max_obs) {
      uint32_t w0 = start ? start[gi * 4 + pid] : cursor[gi * 4 + pid];
      if (w0 > end) w0 = end;
      SeqOut o;
      o.sparse = sparse ? sparse + row * SEQ_MAX_SPARSE : nullptr;
      o.numeric = numeric ? numeric + row * SEQ_NUMERIC : nullptr;
      o.prog = prog ? prog + row * (int64_t)max_prog * 5 : nullptr;
      o.cand = cand ? cand + row * SEQ_MAX_CAND * 4 : nullptr;
      o.lens = lens ? lens + row * 3 : nullptr;
      o.max_prog = max_prog;
      seq_encode(cx, g, pid, glog, w0, end, game_style, o);
      if (index) index[row] = (int32_t)(gi * 4 + pid);
    }
    if (!start) cursor[gi * 4 + pid] = end;      // the delta advances on every observation (state/mod.rs:218)
    row++;
  }
}

__global__ void results_kernel(const G* states, int64_t n, uint8_t* done, int32_t* scores, uint8_t* ranks, uint32_t* step_count,
                               uint32_t* kyoku_count, uint32_t* ev_count, uint64_t* ev_hash) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const G& g = states[i];
  if (done) done[i] = g.is_done;
  if (scores)
    for (int p = 0; p < MAXP; p++) scores[i * MAXP + p] = g.score[p];
  if (ranks)   // env.rs:673-689: score desc, seat asc
    for (int p = 0; p < MAXP; p++) {
      int r = 1, np = num_players(g);
      for (int q = 0; q < np; q++)
        if (g.score[q] > g.score[p] || (g.score[q] == g.score[p] && q < p)) r++;
      ranks[i * MAXP + p] = (uint8_t)(p < np ? r : 0);
    }
  if (step_count) step_count[i] = g.step_count;
  if (kyoku_count) kyoku_count[i] = g.kyoku_count;
  if (ev_count) ev_count[i] = g.ev_count;
  if (ev_hash) ev_hash[i] = g.ev_hash;
}

// ------------------------------------------------------------------ host API
static inline int grid_for(int64_t n, int block) { return (int)((n + block - 1) / block); }

// Device scratch that is released on every return path (CK returns early on a CUDA error).
template <class T>
struct Scratch {
  T* p = nullptr;
  Scratch() = default;
  Scratch(const Scratch&) = delete;
  Scratch& operator=(const Scratch&) = delete;
  ~Scratch() {
    if (p) cudaFree(p);
  }
};
template <class T>
static int upload(rv_ctx* c, const T* h, size_t count, Scratch<T>& d) {
  if (!h) return RV_OK;
  CK(cudaMalloc(&d.p, sizeof(T) * count));
  CK(cudaMemcpyAsync(d.p, h, sizeof(T) * count, cudaMemcpyHostToDevice, c->stream));
  return RV_OK;
}
// Grow-only device buffer owned by a vector (staging of rv_vec_step / rv_vec_legal_actions: no allocation per call).
static int ensure(void** p, size_t* have, size_t need) {
  if (*have >= need) return RV_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *have = 0;
  CK(cudaMalloc(p, need));
  *have = need;
  return RV_OK;
}

struct rv_multi {
  std::vector<rv_ctx*> ctx;
  std::vector<rv_vec*> vec;
  std::vector<int64_t> first;     // global id of each shard's first game; first[G] = N
  bool own_ctx;
};
template <class F>
static int on_every_device(rv_multi* m, F&& f) {     // f(k) on its own host thread; first error wins
  const int G = (int)m->vec.size();
  std::vector<int> rc(G, RV_OK);
  std::vector<std::string> msg(G);
  std::vector<std::thread> th;
  for (int k = 0; k < G; k++)
    th.emplace_back([&, k] {
      rc[k] = f(k);
      if (rc[k] != RV_OK) msg[k] = g_err;            // g_err is thread-local: carry the message back
    });
  for (auto& t : th) t.join();
  for (int k = 0; k < G; k++)
    if (rc[k] != RV_OK) return fail(rc[k], "device " + std::to_string(m->ctx[k]->device) + ": " + msg[k]);
  return RV_OK;
}
extern "C" {

const char* rv_last_error(void) { return g_err.c_str(); }
int rv_version(void) { return 1; }

int rv_ctx_create(int device, rv_ctx** out) {
  if (!out) return fail(RV_ERR_INVALID, "out is null");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(RV_ERR_CUDA, std::string("no CUDA device available (this library has no CPU fallback): ") + cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail(RV_ERR_INVALID, "bad device index");
  CK(cudaSetDevice(device));
  rv_ctx* c = new rv_ctx();
  c->device = device;
  {
    // the context stream outranks the side streams: when the observation encoder (side stream, 16 k short blocks, HBM-bound)
    // runs beside the one-step kernel (context stream, a few hundred long, latency-bound blocks), freed SM resources go to
    // the step kernel first
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, hi));
  }
  for (int i = 0; i < 8; i++) CK(cudaEventCreate(&c->ev[i]));
  for (int i = 0; i < 3; i++) {
    CK(cudaStreamCreateWithFlags(&c->aux[i], cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->join_ev[i], cudaEventDisableTiming));
  }
  CK(cudaEventCreateWithFlags(&c->fork_ev, cudaEventDisableTiming));
  CK(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
  CK(cudaMalloc(&c->suit_info, sizeof(uint32_t) * SUIT_KEYS));
  CK(cudaMalloc(&c->honor_info, sizeof(uint32_t) * HONOR_KEYS));
  CK(cudaMalloc(&c->suit_cost, sizeof(uint64_t) * SUIT_KEYS));
  CK(cudaMalloc(&c->honor_cost, sizeof(uint64_t) * HONOR_KEYS));
  gen_cost_kernel<9, true><<<grid_for(SUIT_KEYS, 128), 128, 0, c->stream>>>(c->suit_cost, c->suit_info, SUIT_KEYS);
  gen_cost_kernel<7, false><<<grid_for(HONOR_KEYS, 128), 128, 0, c->stream>>>(c->honor_cost, c->honor_info, HONOR_KEYS);
  gen_wait_kernel<9><<<grid_for(SUIT_KEYS, 128), 128, 0, c->stream>>>(c->suit_info, SUIT_KEYS);
  gen_wait_kernel<7><<<grid_for(HONOR_KEYS, 128), 128, 0, c->stream>>>(c->honor_info, HONOR_KEYS);
  gen_discard_kernel<9><<<grid_for(SUIT_KEYS, 128), 128, 0, c->stream>>>(c->suit_info, SUIT_KEYS);
  gen_discard_kernel<7><<<grid_for(HONOR_KEYS, 128), 128, 0, c->stream>>>(c->honor_info, HONOR_KEYS);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  c->T.suit_info = c->suit_info;
  c->T.honor_info = c->honor_info;
  c->T.suit_cost = c->suit_cost;
  c->T.honor_cost = c->honor_cost;
  *out = c;
  return RV_OK;
}
int rv_ctx_destroy(rv_ctx* c) {
  if (!c) return RV_OK;
  cudaSetDevice(c->device);
  if (c->d_hand_list) cudaFree(c->d_hand_list);
  cudaFree(c->suit_info);
  cudaFree(c->honor_info);
  cudaFree(c->suit_cost);
  cudaFree(c->honor_cost);
  if (c->hands.d_q[0])
    for (int b = 0; b < 2; b++) {
      cudaFree(c->hands.d_q[b]);
      cudaFree(c->hands.d_r[b]);
      cudaFree(c->hands.d_list[b]);
      cudaFreeHost(c->hands.h_q[b]);
      cudaFreeHost(c->hands.h_r[b]);
      cudaStreamDestroy(c->hands.st[b]);
      cudaEventDestroy(c->hands.done[b]);
    }
  cudaStreamDestroy(c->stream);
  delete c;
  return RV_OK;
}
int rv_ctx_sync(rv_ctx* c) {
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  return RV_OK;
}
void* rv_ctx_stream(rv_ctx* c) { return (void*)c->stream; }
int rv_timer_mark(rv_ctx* c, int idx) {
  if (idx < 0 || idx >= 8) return fail(RV_ERR_INVALID, "timer index must be 0..7");
  CK(cudaSetDevice(c->device));
  CK(cudaEventRecord(c->ev[idx], c->stream));
  return RV_OK;
}
int rv_timer_elapsed(rv_ctx* c, int a, int b, float* ms) {
  if (a < 0 || a >= 8 || b < 0 || b >= 8) return fail(RV_ERR_INVALID, "timer index must be 0..7");
  CK(cudaSetDevice(c->device));
  CK(cudaEventSynchronize(c->ev[b]));
  CK(cudaEventElapsedTime(ms, c->ev[a], c->ev[b]));
  return RV_OK;
}

int rv_hand_eval_batch_device(rv_ctx* c, const rv_hand_query* d_q, rv_hand_result* d_out, int64_t n) {
  if (n <= 0) return RV_OK;
  CK(cudaSetDevice(c->device));
  if (n > 0x7FFFFFFF) return fail(RV_ERR_INVALID, "at most 2^31 - 1 hands per call");
  int rc = ensure((void**)&c->d_hand_list, &c->hand_list_bytes, sizeof(int32_t) * (size_t)n + 16);
  if (rc != RV_OK) return rc;
  // the counter sits behind the list (16-byte slack above)
  unsigned int* count = reinterpret_cast<unsigned int*>(c->d_hand_list + n);
  CK(launch_hand_eval(c->T, d_q, d_out, n, c->d_hand_list, count, c->sm_count, c->stream));
  return RV_OK;
}
// Host buffers: a two-deep pipeline over persistent pinned staging buffers and two streams, so that the host-side copy of
// chunk k+1 into pinned memory, the H2D/D2H transfers and the kernel of chunk k overlap.  Buffers the caller already
// pinned (cudaHostAlloc / cudaHostRegister) are transferred directly.
static constexpr int64_t HAND_CHUNK = 1 << 18;
static bool is_pinned(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}
int rv_hand_eval_batch(rv_ctx* c, const rv_hand_query* q, rv_hand_result* out, int64_t n) {
  if (n <= 0) return RV_OK;
  if (!q || !out) return fail(RV_ERR_INVALID, "null buffer");
  std::lock_guard<std::mutex> lock(c->hands_mu);   // callers on several host threads take turns
  CK(cudaSetDevice(c->device));
  HandStage& hs = c->hands;
  if (!hs.d_q[0]) {
    for (int b = 0; b < 2; b++) {
      CK(cudaMalloc(&hs.d_q[b], sizeof(rv_hand_query) * HAND_CHUNK));
      CK(cudaMalloc(&hs.d_r[b], sizeof(rv_hand_result) * HAND_CHUNK));
      CK(cudaMalloc(&hs.d_list[b], sizeof(int32_t) * HAND_CHUNK + 16));
      CK(cudaMallocHost(&hs.h_q[b], sizeof(rv_hand_query) * HAND_CHUNK));
      CK(cudaMallocHost(&hs.h_r[b], sizeof(rv_hand_result) * HAND_CHUNK));
      CK(cudaStreamCreateWithFlags(&hs.st[b], cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&hs.done[b], cudaEventDisableTiming));
    }
  }
  CK(cudaStreamSynchronize(c->stream));       // earlier work of this context (table generation) is complete
  const bool pin_in = is_pinned(q), pin_out = is_pinned(out);
  const int64_t chunks = (n + HAND_CHUNK - 1) / HAND_CHUNK;
  auto drain = [&](int64_t k) -> int {        // results of chunk k -> caller
    const int b = (int)(k & 1);
    CK(cudaEventSynchronize(hs.done[b]));
    const int64_t lo = k * HAND_CHUNK, m = std::min(HAND_CHUNK, n - lo);
    if (!pin_out) memcpy(out + lo, hs.h_r[b], sizeof(rv_hand_result) * m);
    return RV_OK;
  };
  for (int64_t k = 0; k < chunks; k++) {
    const int b = (int)(k & 1);
    if (k >= 2) {
      int rc = drain(k - 2);
      if (rc != RV_OK) return rc;
    }
    const int64_t lo = k * HAND_CHUNK, m = std::min(HAND_CHUNK, n - lo);
    const rv_hand_query* src = q + lo;
    if (!pin_in) {
      memcpy(hs.h_q[b], src, sizeof(rv_hand_query) * m);
      src = hs.h_q[b];
    }
    CK(cudaMemcpyAsync(hs.d_q[b], src, sizeof(rv_hand_query) * m, cudaMemcpyHostToDevice, hs.st[b]));
    CK(launch_hand_eval(c->T, hs.d_q[b], hs.d_r[b], m, hs.d_list[b], reinterpret_cast<unsigned int*>(hs.d_list[b] + HAND_CHUNK),
                        c->sm_count, hs.st[b]));
    CK(cudaMemcpyAsync(pin_out ? out + lo : hs.h_r[b], hs.d_r[b], sizeof(rv_hand_result) * m, cudaMemcpyDeviceToHost, hs.st[b]));
    CK(cudaEventRecord(hs.done[b], hs.st[b]));
  }
  for (int64_t k = std::max<int64_t>(0, chunks - 2); k < chunks; k++) {
    int rc = drain(k);
    if (rc != RV_OK) return rc;
  }
  return RV_OK;
}
int rv_hand_queries_seeded(rv_ctx* c, rv_hand_query* d_q, uint64_t first, int64_t n) {
  if (n <= 0) return RV_OK;
  if (!d_q) return fail(RV_ERR_INVALID, "null buffer");
  CK(cudaSetDevice(c->device));
  synth_hands_kernel<<<grid_for(n, 128), 128, 0, c->stream>>>(d_q, first, n);
  CK(cudaGetLastError());
  return RV_OK;
}
int rv_hand_query_seeded_host(uint64_t h, rv_hand_query* out) {
  if (!out) return fail(RV_ERR_INVALID, "null buffer");
  rv_synth_hand(h, out);
  return RV_OK;
}
int rv_calculate_score(int han, int fu, int is_oya, int is_tsumo, uint32_t honba, int num_players, uint32_t out[4]) {
  uint32_t total = 0;
  ScoreRes s = calc_score(han & 0xFF, fu & 0xFF, is_oya != 0, is_tsumo != 0, honba, (uint32_t)num_players, &total);
  out[0] = s.pay_ron;
  out[1] = is_tsumo ? s.pay_oya : 0;
  out[2] = is_tsumo ? s.pay_ko : 0;
  out[3] = total;
  return RV_OK;
}
int rv_wall_from_seed(uint64_t seed, uint64_t hand_index, int n_tiles, uint8_t* out) {
  if (n_tiles != 136 && n_tiles != 108) return fail(RV_ERR_INVALID, "n_tiles must be 136 or 108");
  wall_from_seed(seed, hand_index, n_tiles, out);
  return RV_OK;
}

int rv_vec_create(rv_ctx* c, int64_t n, int game_mode, uint32_t rule_bits, const uint64_t* seeds, uint64_t seed_base,
                  uint32_t log_cap_words, rv_vec** out) {
  if (!c || !out || n <= 0) return fail(RV_ERR_INVALID, "bad arguments");
  if (game_mode < 0 || game_mode > 5) return fail(RV_ERR_INVALID, "game_mode must be 0..5");
  CK(cudaSetDevice(c->device));
  rv_vec* v = new rv_vec();
  v->ctx = c;
  v->n = n;
  v->game_mode = game_mode;
  v->rule_bits = rule_bits;
  v->log_cap = log_cap_words;
  v->d_log = nullptr;
  v->d_seeds = nullptr;
  v->d_obs_counts = v->d_obs_offsets = nullptr;
  v->d_idbits = nullptr;
  v->h_obs_total = nullptr;
  v->d_row_index = nullptr;
  v->d_scan_tmp = nullptr;
  v->scan_tmp_bytes = 0;
  v->d_lists = nullptr;
  v->d_list_counts = nullptr;
  v->d_budget = nullptr;
  v->h_counts = nullptr;
  v->graph_exec = nullptr;
  v->d_seq_cursor = nullptr;
  v->d_gather = nullptr;
  v->h_gather = nullptr;
  v->d_seq_start = nullptr;
  v->d_q_slots = nullptr;
  v->d_q_ctl = nullptr;
  v->d_io_a = v->d_io_c = nullptr;
  v->d_obs_snap = nullptr;
  v->io_a_bytes = v->io_c_bytes = 0;
  v->d_rp_actions = nullptr;
  v->d_rp_first = nullptr;
  v->d_rp_cursor = v->d_rp_live = nullptr;
  v->q_cap = 0;
  v->graph_seed = 0;
  v->graph_period = 0;
  CK(cudaMalloc(&v->d_states, sizeof(G) * n));
  if (log_cap_words) CK(cudaMalloc(&v->d_log, sizeof(uint32_t) * (size_t)n * log_cap_words));
  CK(cudaMalloc(&v->d_steps, sizeof(unsigned long long) * 32));
  CK(cudaMemsetAsync(v->d_steps, 0, sizeof(unsigned long long) * 32, c->stream));
  uint64_t* d_seeds = nullptr;
  if (seeds) {
    CK(cudaMalloc(&d_seeds, sizeof(uint64_t) * n));
    CK(cudaMemcpyAsync(d_seeds, seeds, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, c->stream));
  }
  create_kernel<<<grid_for(n, 128), 128, 0, c->stream>>>(v->d_states, n, game_mode, rule_bits, d_seeds, seed_base);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  if (d_seeds) cudaFree(d_seeds);
  *out = v;
  return RV_OK;
}
int rv_vec_reseed(rv_vec* v, const uint64_t* seeds, uint64_t seed_base) {
  rv_ctx* c = v->ctx;
  CK(cudaSetDevice(c->device));
  if (seeds) {
    if (!v->d_seeds) CK(cudaMalloc(&v->d_seeds, sizeof(uint64_t) * v->n));
    CK(cudaMemcpyAsync(v->d_seeds, seeds, sizeof(uint64_t) * v->n, cudaMemcpyHostToDevice, c->stream));
  }
  reseed_kernel<<<grid_for(v->n, 256), 256, 0, c->stream>>>(v->d_states, v->n, seeds ? v->d_seeds : nullptr, seed_base);
  CK(cudaGetLastError());
  return RV_OK;
}
int rv_vec_destroy(rv_vec* v) {
  if (!v) return RV_OK;
  cudaSetDevice(v->ctx->device);
  cudaFree(v->d_states);
  if (v->d_log) cudaFree(v->d_log);
  if (v->d_seeds) cudaFree(v->d_seeds);
  if (v->d_obs_counts) cudaFree(v->d_obs_counts);
  if (v->d_idbits) cudaFree(v->d_idbits);
  if (v->h_obs_total) cudaFreeHost(v->h_obs_total);
  if (v->d_row_index) cudaFree(v->d_row_index);
  if (v->d_obs_offsets) cudaFree(v->d_obs_offsets);
  if (v->d_scan_tmp) cudaFree(v->d_scan_tmp);
  if (v->d_lists) cudaFree(v->d_lists);
  if (v->d_list_counts) cudaFree(v->d_list_counts);
  if (v->d_budget) cudaFree(v->d_budget);
  if (v->h_counts) cudaFreeHost(v->h_counts);
  if (v->graph_exec) cudaGraphExecDestroy(v->graph_exec);
  if (v->d_seq_cursor) cudaFree(v->d_seq_cursor);
  if (v->d_gather) cudaFree(v->d_gather);
  if (v->h_gather) cudaFreeHost(v->h_gather);
  if (v->d_seq_start) cudaFree(v->d_seq_start);
  if (v->d_q_slots) cudaFree(v->d_q_slots);
  if (v->d_q_ctl) cudaFree(v->d_q_ctl);
  if (v->d_obs_snap) cudaFree(v->d_obs_snap);
  if (v->d_io_a) cudaFree(v->d_io_a);
  if (v->d_io_c) cudaFree(v->d_io_c);
  if (v->d_rp_actions) cudaFree(v->d_rp_actions);
  if (v->d_rp_first) cudaFree(v->d_rp_first);
  if (v->d_rp_cursor) cudaFree(v->d_rp_cursor);
  if (v->d_rp_live) cudaFree(v->d_rp_live);
  cudaFree(v->d_steps);
  delete v;
  return RV_OK;
}
int64_t rv_vec_size(const rv_vec* v) { return v ? v->n : 0; }

int rv_vec_reset(rv_vec* v, const uint8_t* oya, const uint8_t* round_wind, const uint8_t* honba, const uint32_t* kyotaku,
                 const int32_t* scores, const uint8_t* walls) {
  rv_ctx* c = v->ctx;
  CK(cudaSetDevice(c->device));
  Scratch<uint8_t> d_oya, d_rw, d_honba, d_walls;    // freed on every return path
  Scratch<uint32_t> d_ky;
  Scratch<int32_t> d_sc;
  int rc;
  if ((rc = upload(c, oya, v->n, d_oya))) return rc;
  if ((rc = upload(c, round_wind, v->n, d_rw))) return rc;
  if ((rc = upload(c, honba, v->n, d_honba))) return rc;
  if ((rc = upload(c, kyotaku, v->n, d_ky))) return rc;
  if ((rc = upload(c, scores, v->n * MAXP, d_sc))) return rc;
  if ((rc = upload(c, walls, v->n * (v->game_mode >= 3 ? 108 : 136), d_walls))) return rc;
  if (v->d_seq_cursor) CK(cudaMemsetAsync(v->d_seq_cursor, 0, sizeof(uint32_t) * 4 * v->n, c->stream));   // player_event_counts = [0; NP]
  CK(cudaMemsetAsync(v->d_steps, 0, sizeof(unsigned long long) * 32, c->stream));
  reset_kernel<<<grid_for(v->n, 128), 128, 0, c->stream>>>(c->T, v->d_states, v->n, v->d_log, v->log_cap, d_oya.p, d_rw.p, d_honba.p,
                                                           d_ky.p, d_sc.p, d_walls.p);
  CK(cudaGetLastError());
  bool any = oya || round_wind || honba || kyotaku || scores || walls;
  if (any) CK(cudaStreamSynchronize(c->stream));   // the staging buffers go out of scope; a default reset stays asynchronous
  return RV_OK;
}

int rv_vec_legal_actions(rv_vec* v, rv_action* out_actions, uint8_t* out_counts) {
  rv_ctx* c = v->ctx;
  CK(cudaSetDevice(c->device));
  size_t na = (size_t)v->n * MAXP * RV_MAX_LEGAL;
  int rc = ensure(&v->d_io_a, &v->io_a_bytes, sizeof(rv_action) * na);
  if (rc == RV_OK) rc = ensure(&v->d_io_c, &v->io_c_bytes, (size_t)v->n * MAXP);
  if (rc != RV_OK) return rc;
  rv_action* d_a = (rv_action*)v->d_io_a;
  uint8_t* d_c = (uint8_t*)v->d_io_c;
  legal_kernel<<<grid_for(v->n, 64), 64, 0, c->stream>>>(c->T, v->d_states, v->n, d_a, d_c);
  CK(cudaGetLastError());
  if (out_actions) CK(cudaMemcpyAsync(out_actions, d_a, sizeof(rv_action) * na, cudaMemcpyDeviceToHost, c->stream));
  if (out_counts) CK(cudaMemcpyAsync(out_counts, d_c, (size_t)v->n * MAXP, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return RV_OK;
}

int rv_vec_step(rv_vec* v, const rv_action* actions) {
  rv_ctx* c = v->ctx;
  CK(cudaSetDevice(c->device));
  if (!actions) return fail(RV_ERR_INVALID, "actions is null");
  int rc = ensure(&v->d_io_a, &v->io_a_bytes, sizeof(rv_action) * v->n * MAXP);
  if (rc != RV_OK) return rc;
  rv_action* d_a = (rv_action*)v->d_io_a;
  CK(cudaMemcpyAsync(d_a, actions, sizeof(rv_action) * v->n * MAXP, cudaMemcpyHostToDevice, c->stream));
  step_kernel<<<grid_for(v->n, 64), 64, 0, c->stream>>>(c->T, v->d_states, v->n, v->d_log, v->log_cap, d_a, v->d_steps);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  return RV_OK;
}

static int env_int(const char* name, int dflt);
static int env_int0(const char* name, int dflt);
static int rollout_mono(rv_vec* v, uint64_t agent_seed, uint32_t max_steps) {
  rv_ctx* c = v->ctx;
  static const int sorted = env_int("RV_STEP_SORTED", 3) - 1;   // 1: thread per game; 2: regrouped in the block; 3: + staged records
  if (max_steps == 1 && sorted == 2) {
    CK(launch_step_staged<false>(c, v, agent_seed, nullptr));
    return RV_OK;
  }
  if (max_steps == 1 && sorted) {      // lock-step drivers: one env step per game, regrouped by kind inside the block
    step_sorted_kernel<false><<<grid_for(v->n, SB), SB, 0, c->stream>>>(c->T, v->d_states, v->n, v->d_log, v->log_cap, agent_seed, v->d_steps,
                                                                        nullptr);
    CK(cudaGetLastError());
    return RV_OK;
  }
  step_random_kernel<false><<<grid_for(v->n, 128), 128, 0, c->stream>>>(c->T, v->d_states, v->n, v->d_log, v->log_cap, agent_seed,
                                                                        max_steps, v->d_steps, nullptr);
  CK(cudaGetLastError());
  return RV_OK;
}
// Phase-sorted pipeline; returns after every game has used its budget (or finished) and no deal is pending.
// Every class has a double-buffered list.  The ACT list is swapped every iteration; RESP/SLOW lists are
// drained every `slow_every` iterations and the DEAL list every `deal_every` (their kernels are long
// single-warp latencies — a deal is ~270 us — that must not sit on the critical path of every iteration).
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  int v = e ? atoi(e) : dflt;
  return v < 1 ? 1 : v;
}
static int env_int0(const char* name, int dflt) {      // knobs for which 0 is a value (switches)
  const char* e = getenv(name);
  int v = e ? atoi(e) : dflt;
  return v < 0 ? 0 : v;
}
static int rollout_phased(rv_vec* v, uint64_t agent_seed, uint32_t max_steps) {
  rv_ctx* c = v->ctx;
  int64_t n = v->n;
  static int slow_every = env_int("RV_SLOW_EVERY", 1), deal_mult = env_int("RV_DEAL_MULT", 8);
  const int deal_every = slow_every * deal_mult;   // deal drains coincide with slow drains
  if (!v->d_lists) {
    CK(cudaMalloc(&v->d_lists, sizeof(int32_t) * 2 * N_LISTS * n));      // [class][buffer][n]
    CK(cudaMalloc(&v->d_list_counts, sizeof(uint32_t) * 16));            // [class][buffer]
    if (!v->d_budget) CK(cudaMalloc(&v->d_budget, sizeof(uint32_t) * n));
    CK(cudaMallocHost(&v->h_counts, sizeof(uint32_t) * 16));
  }
  auto list = [&](int ph, int b) { return v->d_lists + (size_t)(ph * 2 + b) * n; };
  auto count = [&](int ph, int b) { return v->d_list_counts + ph * 2 + b; };
  int wr[N_LISTS] = {0, 0, 0, 0, 0, 0};     // buffer currently being WRITTEN for each class
  auto mk = [&]() {
    Lists L;
    for (int ph = 0; ph < N_LISTS; ph++) {
      L.dst[ph] = list(ph, wr[ph]);
      L.cnt[ph] = count(ph, wr[ph]);
    }
    return L;
  };
  bool inflight[3] = {false, false, false};   // aux[0]=RESP, aux[1]=DEAL, aux[2]=SLOW
  auto join = [&](int a) -> int {
    if (inflight[a]) {
      CK(cudaStreamWaitEvent(c->stream, c->join_ev[a], 0));
      inflight[a] = false;
    }
    return RV_OK;
  };
  auto swap_class = [&](int ph) -> int {
    wr[ph] ^= 1;             // producers switch to the other buffer (drained earlier), cleared now
    CK(cudaMemsetAsync(count(ph, wr[ph]), 0, sizeof(uint32_t), c->stream));
    return RV_OK;
  };
#define LAUNCH(PH, OUT, STREAM, RD)                                                                                   \
  phase_kernel<PH, OUT><<<pgrid, PHB, 0, STREAM>>>(c->T, v->d_states, n, v->d_log, v->log_cap, agent_seed, v->d_budget, \
                                                  list(PH, RD), count(PH, RD), out, v->d_steps, act_reps)
  int grid = grid_for(n, 128);
  const int pgrid = grid_for(n, PHB);
  static int act_reps = env_int("RV_ACT_REPS", 4);
  int rc;
  // One period of the pipeline = lcm of the buffer-flip periods (ACT 2, RESP/SLOW/REACT 2*slow_every, DEAL 2*deal_every):
  // after it every write-buffer index is back where it started, so the whole period is captured once into a CUDA graph
  // (the rollout needs ~3,000 iterations x ~8 stream operations; issuing them one by one makes the HOST the bottleneck).
  const int period = 2 * deal_every;
  auto issue_period = [&]() -> int {
    for (int it = 0; it < period; it++) {
      bool slow_drain = (it % slow_every) == slow_every - 1;
      bool deal_drain = (it % deal_every) == deal_every - 1;
      int rd_act = wr[PH_ACT];
      if ((rc = swap_class(PH_ACT))) return rc;
      int rd_resp = 0, rd_slow = 0, rd_react = 0, rd_deal = 0, rd_react_d = 0;
      if (slow_drain) {
        // the RESP / SLOW kernels of the previous drain append to RESP, DEAL, REACT: finish them before swapping
        if ((rc = join(0)) || (rc = join(2))) return rc;
        rd_resp = wr[PH_RESP];
        rd_slow = wr[PH_SLOW];
        rd_react = wr[PH_REACT];
        if ((rc = swap_class(PH_RESP)) || (rc = swap_class(PH_SLOW)) || (rc = swap_class(PH_REACT))) return rc;
      }
      if (deal_drain) {
        if ((rc = join(1))) return rc;        // previous deal kernel appends to REACT_D and reads the other DEAL buffer
        rd_deal = wr[PH_DEAL];
        rd_react_d = wr[PH_REACT_D];
        if ((rc = swap_class(PH_DEAL)) || (rc = swap_class(PH_REACT_D))) return rc;
      }
      Lists out = mk();
      if (slow_drain || deal_drain) CK(cudaEventRecord(c->fork_ev, c->stream));
      LAUNCH(PH_ACT, PH_ACT, c->stream, rd_act);
      if (slow_drain) {
        // games returning from the previous drain's RESP/SLOW kernels re-enter through the fast kernel
        phase_kernel<PH_ACT, PH_ACT><<<pgrid, PHB, 0, c->stream>>>(c->T, v->d_states, n, v->d_log, v->log_cap, agent_seed, v->d_budget,
                                                                   list(PH_REACT, rd_react), count(PH_REACT, rd_react), out, v->d_steps, act_reps);
        CK(cudaStreamWaitEvent(c->aux[0], c->fork_ev, 0));
        CK(cudaStreamWaitEvent(c->aux[2], c->fork_ev, 0));
        LAUNCH(PH_RESP, PH_REACT, c->aux[0], rd_resp);
        LAUNCH(PH_SLOW, PH_REACT, c->aux[2], rd_slow);
        CK(cudaEventRecord(c->join_ev[0], c->aux[0]));
        CK(cudaEventRecord(c->join_ev[2], c->aux[2]));
        inflight[0] = inflight[2] = true;
      }
      if (deal_drain) {
        phase_kernel<PH_ACT, PH_ACT><<<pgrid, PHB, 0, c->stream>>>(c->T, v->d_states, n, v->d_log, v->log_cap, agent_seed, v->d_budget,
                                                                   list(PH_REACT_D, rd_react_d), count(PH_REACT_D, rd_react_d), out,
                                                                   v->d_steps, act_reps);
        CK(cudaStreamWaitEvent(c->aux[1], c->fork_ev, 0));
        LAUNCH(PH_DEAL, PH_REACT_D, c->aux[1], rd_deal);
        CK(cudaEventRecord(c->join_ev[1], c->aux[1]));
        inflight[1] = true;
      }
    }
    for (int a = 0; a < 3; a++)
      if ((rc = join(a))) return rc;          // the period is self-contained: side streams rejoin the main stream
    return RV_OK;
  };
  static int use_graph = getenv("RV_GRAPH") ? atoi(getenv("RV_GRAPH")) : 1;
  if (use_graph == 1 && (!v->graph_exec || v->graph_seed != agent_seed || v->graph_period != period * 64 + act_reps)) {
    if (v->graph_exec) {
      cudaGraphExecDestroy(v->graph_exec);
      v->graph_exec = nullptr;
    }
    cudaGraph_t graph;
    CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    rc = issue_period();
    cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
    if (rc) return rc;
    if (ce != cudaSuccess) return fail(RV_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
    CK(cudaGraphInstantiate(&v->graph_exec, graph, 0));
    cudaGraphDestroy(graph);
    v->graph_seed = agent_seed;
    v->graph_period = period * 64 + act_reps;
    for (int ph = 0; ph < N_LISTS; ph++) wr[ph] = 0;   // capture leaves every class back at buffer 0 (period property)
  }
  CK(cudaMemsetAsync(v->d_list_counts, 0, sizeof(uint32_t) * 16, c->stream));
  sched_init_kernel<<<grid, 128, 0, c->stream>>>(v->d_states, n, v->d_budget, max_steps, mk());
  uint64_t it_total = 0;
  while (true) {
    if (use_graph == 1) {
      CK(cudaGraphLaunch(v->graph_exec, c->stream));
    } else if ((rc = issue_period())) {
      return rc;
    }
    it_total += period;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(v->h_counts, v->d_list_counts, sizeof(uint32_t) * 16, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    bool pending = false;
    for (int ph = 0; ph < N_LISTS; ph++) pending |= v->h_counts[ph * 2 + wr[ph]] != 0;
    if (!pending) break;
  }
#undef LAUNCH
  if (getenv("RV_DEBUG")) fprintf(stderr, "[rollout_phased] iterations=%llu slow_every=%d deal_every=%d\n", (unsigned long long)it_total, slow_every, deal_every);
  return RV_OK;
}

// Persistent scheduler: one launch for the whole rollout (see rollout_persistent_kernel).
static int rollout_persistent(rv_vec* v, uint64_t agent_seed, uint32_t max_steps) {
  rv_ctx* c = v->ctx;
  int64_t n = v->n;
  if (!v->d_q_slots) {
    uint32_t capq = 64;
    while ((int64_t)capq < 2 * n) capq <<= 1;
    v->q_cap = capq;
    CK(cudaMalloc(&v->d_q_slots, sizeof(int32_t) * N_QUEUES * (size_t)capq));
    CK(cudaMalloc(&v->d_q_ctl, sizeof(uint32_t) * Q_CTL_WORDS));
    if (!v->d_budget) CK(cudaMalloc(&v->d_budget, sizeof(uint32_t) * n));
  }
  static int warps_per_sm = env_int("RV_WARPS_PER_SM", 12);
  // read per call (A/B in one process): reps of act_fast per iteration; RV_ACT_HOLD=0 turns the lane refill off
  const int act_reps = (env_int("RV_ACT_REPS", 4) & 0xFF) | (env_int0("RV_ACT_HOLD", 0) ? 0x100 : 0) | (env_int0("RV_CREW_GREEDY", 1) ? 0x200 : 0) | ((env_int("RV_SLOT_POLLS", 64) & 0xFF) << 16) | ((env_int0("RV_FOLLOW_REPS", 4) & 0xF) << 24);
  // endgame: below RV_ENDGAME_PER_WARP live games per crew warp (x1/4: the knob is in quarter games, default 8 = 2 games per
  // warp), warps own RV_ENDGAME_TAKE games each and play them out in place (see the kernel)
  static int eg_quarters = env_int("RV_ENDGAME_Q", 8), eg_take = env_int("RV_ENDGAME_TAKE", 1);
  const uint32_t switch_knobs = ((uint32_t)env_int("RV_SWITCH_IDLE", 3) & 0xFFFF) | ((uint32_t)env_int("RV_SWITCH_MINLEN", 1) << 16);
  Queues q{v->d_q_slots, v->d_q_ctl, v->q_cap - 1};
  CK(cudaMemsetAsync(v->d_q_ctl, 0, sizeof(uint32_t) * Q_CTL_WORDS, c->stream));
  CK(cudaMemsetAsync(v->d_q_slots, 0xFF, sizeof(int32_t) * N_QUEUES * (size_t)v->q_cap, c->stream));   // every slot empty (abandoned marks of the last call included)
  q_init_kernel<<<grid_for(n, 128), 128, 0, c->stream>>>(v->d_states, n, v->d_budget, max_steps, q, env_int0("RV_INIT_DIST", 1));
  {
    const char* fault = getenv("RV_FAULT_INJECT");     // read per call: the watchdog test arms it for one rollout
    if (fault && strcmp(fault, "lost_game") == 0) q_fault_kernel<<<1, 1, 0, c->stream>>>(q);
  }
  int64_t crew = (int64_t)c->sm_count * warps_per_sm, need = (n + PHB - 1) / PHB;
  const int grid = (int)(crew < need ? crew : need);
  const uint32_t eg_live = getenv("RV_ENDGAME_OFF") ? 0u : (uint32_t)((int64_t)grid * eg_quarters / 4);
  if (env_int0("RV_CREW", 1)) {
    // one block per SM, warps_per_sm warps each (see rollout_crew_kernel); the endgame threshold counts warps as before
    const int crew_warps = warps_per_sm > 12 ? 12 : warps_per_sm;   // 12 x (32 x 560 B) of staging fill the SM's shared memory; 168 registers x 384 threads its register file
    const int threads = 32 * crew_warps;
    const size_t smem = (size_t)crew_warps * PHB * STG_STRIDE + sizeof(uint64_t) * crew_warps + 2 * sizeof(CrewCtl);
    const int64_t need_b = (n + threads - 1) / threads;
    const int blocks = (int)(c->sm_count < need_b ? c->sm_count : need_b);
    const uint32_t eg_live_c = getenv("RV_ENDGAME_OFF") ? 0u : (uint32_t)((int64_t)blocks * crew_warps * eg_quarters / 4);
    // per launch: the attribute belongs to the device's context, and an rv_multi handle drives several devices from one process
    if (v->game_mode >= 3) CK(cudaFuncSetAttribute(rollout_crew_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else CK(cudaFuncSetAttribute(rollout_crew_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (v->game_mode >= 3)
      rollout_crew_kernel<3><<<blocks, threads, smem, c->stream>>>(c->T, v->d_states, n, v->d_log, v->log_cap, agent_seed, v->d_budget, q,
                                                                  v->d_steps, act_reps, eg_live_c, eg_take, switch_knobs);
    else
      rollout_crew_kernel<4><<<blocks, threads, smem, c->stream>>>(c->T, v->d_states, n, v->d_log, v->log_cap, agent_seed, v->d_budget, q,
                                                                  v->d_steps, act_reps, eg_live_c, eg_take, switch_knobs);
  } else if (v->game_mode >= 3)
    rollout_persistent_kernel<3><<<grid, PHB, 0, c->stream>>>(c->T, v->d_states, n, v->d_log, v->log_cap, agent_seed, v->d_budget, q,
                                                             v->d_steps, act_reps, eg_live, eg_take, switch_knobs);
  else
    rollout_persistent_kernel<4><<<grid, PHB, 0, c->stream>>>(c->T, v->d_states, n, v->d_log, v->log_cap, agent_seed, v->d_budget, q,
                                                             v->d_steps, act_reps, eg_live, eg_take, switch_knobs);
  CK(cudaGetLastError());
  return RV_OK;
}
static int rollout_mode() {   // RV_ROLLOUT = mono | phased | persistent (default)
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("RV_ROLLOUT");
    mode = (e && strcmp(e, "mono") == 0) ? 0 : (e && strcmp(e, "phased") == 0) ? 1 : 2;
  }
  return mode;
}
int rv_vec_step_random_async(rv_vec* v, uint64_t agent_seed, uint32_t max_steps) {
  rv_ctx* c = v->ctx;
  CK(cudaSetDevice(c->device));
  if (max_steps == 0) return RV_OK;
  // short calls (lock-step drivers, per-step observation loops) use the single persistent kernel: the phase pipeline
  // needs a few dozen iterations to drain its deferred lists and only pays off for long rollouts
  const int mode = rollout_mode();
  static const uint32_t mono_below = (uint32_t)env_int("RV_MONO_BELOW", 32);
  if (mode == 0 || max_steps < mono_below) return rollout_mono(v, agent_seed, max_steps);
  return mode == 1 ? rollout_phased(v, agent_seed, max_steps) : rollout_persistent(v, agent_seed, max_steps);
}
int rv_vec_steps_total(rv_vec* v, uint64_t* steps_total, int64_t* games_done) {
  rv_ctx* c = v->ctx;
  CK(cudaSetDevice(c->device));
  unsigned long long h[2];
  uint32_t sched_err = 0;
  CK(cudaMemcpyAsync(h, v->d_steps, sizeof h, cudaMemcpyDeviceToHost, c->stream));
  if (v->d_q_ctl) CK(cudaMemcpyAsync(&sched_err, v->d_q_ctl + Q_ERR, sizeof sched_err, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (sched_err) {   // reported once: the flag is cleared so that the vector stays usable
    CK(cudaMemsetAsync(v->d_q_ctl + Q_ERR, 0, sizeof(uint32_t), c->stream));
    return fail(RV_ERR_CUDA, "rollout scheduler watchdog: live games but no queued work");
  }
#ifdef RV_QPROF
  {
    unsigned long long p[32];
    CK(cudaMemcpy(p, v->d_steps, sizeof p, cudaMemcpyDeviceToHost));
    const char* nm[10] = {"idle", "claim", "stage_in", "ACT", "RESP", "DEAL", "SLOW", "TAIL", "stage_out", "push"};
    unsigned long long tot = 0;
    for (int i = 0; i < 10; i++) tot += p[8 + i];
    if (p[3])
      fprintf(stderr, "[qprof] crew phase: warp-iterations=%llu without ticket=%llu (%.1f%%) class changes=%llu | warp-cycles: barrier=%llu "
              "idle sleeps=%llu visits=%llu\n", p[3], p[4], 100.0 * p[4] / p[3], p[5], p[2], p[6], p[7]);
    if (tot) fprintf(stderr, "[qprof] steps=%llu total warp-cycles=%llu\n", h[0], tot);
    for (int i = 0; i < 10 && tot; i++) fprintf(stderr, "[qprof] %-9s %6.2f%%\n", nm[i], 100.0 * p[8 + i] / (tot ? tot : 1));
    if (tot) fprintf(stderr, "[qprof] (unused)=%llu empty_looks=%llu tickets_abandoned=%llu sm_switches=%llu\n", p[28], p[29], p[30], p[31]);
    for (int c2 = 0; c2 < 5 && tot; c2++)
      fprintf(stderr, "[qprof] class %s: batches=%llu games=%llu (%.1f/batch) cycles/batch=%.0f\n", nm[3 + c2], p[18 + c2], p[23 + c2],
              p[18 + c2] ? (double)p[23 + c2] / p[18 + c2] : 0.0, p[18 + c2] ? (double)p[11 + c2] / p[18 + c2] : 0.0);
  }
#endif
  if (steps_total) *steps_total = h[0];
  if (games_done) *games_done = (int64_t)h[1];
  return RV_OK;
}
int rv_vec_step_random(rv_vec* v, uint64_t agent_seed, uint32_t max_steps, uint64_t* steps_done) {
  uint64_t before = 0, after = 0;
  int rc = rv_vec_steps_total(v, &before, nullptr);
  if (rc) return rc;
  if ((rc = rv_vec_step_random_async(v, agent_seed, max_steps))) return rc;
  if ((rc = rv_vec_steps_total(v, &after, nullptr))) return rc;
  if (steps_done) *steps_done = after - before;
  return RV_OK;
}

// Results leave through a device gather buffer and a PINNED host staging buffer owned by the vector (allocated once):
// one kernel, one D2H copy, then plain memcpy into the caller's arrays (which may be pageable, freshly allocated numpy).
static int gather(rv_vec* v, uint8_t* done, int32_t* scores, uint8_t* ranks, uint32_t* sc, uint32_t* kc, uint32_t* ec, uint64_t* eh) {
  rv_ctx* c = v->ctx;
  CK(cudaSetDevice(c->device));
  const size_t n = (size_t)v->n;
  // layout (8-byte aligned sections): eh u64[n] | scores i32[4n] | sc, kc, ec u32[n] each | ranks u8[4n] | done u8[n]
  const size_t o_eh = 0, o_scores = o_eh + 8 * n, o_sc = o_scores + 16 * n, o_kc = o_sc + 4 * n, o_ec = o_kc + 4 * n,
               o_ranks = o_ec + 4 * n, o_done = o_ranks + 4 * n, total = o_done + n;
  if (!v->d_gather) {
    CK(cudaMalloc(&v->d_gather, total));
    CK(cudaMallocHost(&v->h_gather, total));
  }
  unsigned char *d = v->d_gather, *h = v->h_gather;
  const bool res = done || scores || ranks, cnt = sc || kc || ec || eh;
  results_kernel<<<grid_for((int64_t)n, 128), 128, 0, c->stream>>>(
      v->d_states, (int64_t)n, res ? d + o_done : nullptr, res ? (int32_t*)(d + o_scores) : nullptr, res ? d + o_ranks : nullptr,
      cnt ? (uint32_t*)(d + o_sc) : nullptr, cnt ? (uint32_t*)(d + o_kc) : nullptr, cnt ? (uint32_t*)(d + o_ec) : nullptr,
      cnt ? (uint64_t*)(d + o_eh) : nullptr);
  CK(cudaGetLastError());
  // one contiguous copy covering what was asked for
  const size_t from = cnt ? o_eh : o_scores, to = res ? total : o_ranks;
  CK(cudaMemcpyAsync(h + from, d + from, to - from, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (done) memcpy(done, h + o_done, n);
  if (scores) memcpy(scores, h + o_scores, 16 * n);
  if (ranks) memcpy(ranks, h + o_ranks, 4 * n);
  if (sc) memcpy(sc, h + o_sc, 4 * n);
  if (kc) memcpy(kc, h + o_kc, 4 * n);
  if (ec) memcpy(ec, h + o_ec, 4 * n);
  if (eh) memcpy(eh, h + o_eh, 8 * n);
  return RV_OK;
}
int rv_vec_results(rv_vec* v, uint8_t* done, int32_t* scores, uint8_t* ranks) {
  return gather(v, done, scores, ranks, nullptr, nullptr, nullptr, nullptr);
}
int rv_vec_counters(rv_vec* v, uint32_t* step_count, uint32_t* kyoku_count, uint32_t* ev_count, uint64_t* ev_hash) {
  return gather(v, nullptr, nullptr, nullptr, step_count, kyoku_count, ev_count, ev_hash);
}

int rv_vec_get_state(rv_vec* v, int64_t game, rv_game_state* out) {
  if (game < 0 || game >= v->n) return fail(RV_ERR_INVALID, "game index out of range");
  rv_ctx* c = v->ctx;
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpyAsync(out, v->d_states + game, sizeof(G), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return RV_OK;
}
int rv_vec_set_state(rv_vec* v, int64_t game, const rv_game_state* in) {
  if (game < 0 || game >= v->n) return fail(RV_ERR_INVALID, "game index out of range");
  if (const char* why = rv_state_defect(*in)) return fail(RV_ERR_INVALID, why);
  rv_ctx* c = v->ctx;
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpyAsync(v->d_states + game, in, sizeof(G), cudaMemcpyHostToDevice, c->stream));
  refresh_kernel<<<1, 1, 0, c->stream>>>(c->T, v->d_states + game);   // derived caches follow the canonical fields
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  return RV_OK;
}
int rv_vec_step_agent(rv_vec* v, int policy, uint64_t agent_seed, uint32_t max_steps, uint64_t* steps_done) {
  if (policy != RV_AGENT_RANDOM && policy != RV_AGENT_GREEDY) return fail(RV_ERR_INVALID, "unknown agent policy");
  if (policy == RV_AGENT_RANDOM) return rv_vec_step_random(v, agent_seed, max_steps, steps_done);
  rv_ctx* c = v->ctx;
  CK(cudaSetDevice(c->device));
  unsigned long long before = 0, after = 0;
  CK(cudaMemcpyAsync(&before, v->d_steps, sizeof before, cudaMemcpyDeviceToHost, c->stream));
  agent_kernel<<<grid_for(v->n, 128), 128, 0, c->stream>>>(c->T, v->d_states, v->n, v->d_log, v->log_cap, policy, agent_seed, max_steps,
                                                          v->d_steps);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(&after, v->d_steps, sizeof after, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (steps_done) *steps_done = after - before;
  return RV_OK;
}
int rv_vec_apply_events(rv_vec* v, const rv_mjai_event* events) {
  if (!events) return fail(RV_ERR_INVALID, "events is null");
  rv_ctx* c = v->ctx;
  CK(cudaSetDevice(c->device));
  int rc = ensure(&v->d_io_a, &v->io_a_bytes, sizeof(rv_mjai_event) * (size_t)v->n);
  if (rc != RV_OK) return rc;
  rv_mjai_event* d_ev = (rv_mjai_event*)v->d_io_a;
  CK(cudaMemcpyAsync(d_ev, events, sizeof(rv_mjai_event) * (size_t)v->n, cudaMemcpyHostToDevice, c->stream));
  apply_events_kernel<<<grid_for(v->n, 64), 64, 0, c->stream>>>(c->T, v->d_states, v->n, d_ev);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  return RV_OK;
}
int rv_vec_apply_log_actions(rv_vec* v, const rv_log_action* actions) {
  if (!actions) return fail(RV_ERR_INVALID, "actions is null");
  rv_ctx* c = v->ctx;
  CK(cudaSetDevice(c->device));
  int rc = ensure(&v->d_io_a, &v->io_a_bytes, sizeof(rv_log_action) * (size_t)v->n);
  if (rc != RV_OK) return rc;
  rv_log_action* d_a = (rv_log_action*)v->d_io_a;
  CK(cudaMemcpyAsync(d_a, actions, sizeof(rv_log_action) * (size_t)v->n, cudaMemcpyHostToDevice, c->stream));
  apply_log_actions_kernel<<<grid_for(v->n, 64), 64, 0, c->stream>>>(c->T, v->d_states, v->n, d_a);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  return RV_OK;
}
int rv_vec_replay_begin(rv_vec* v, const rv_log_kyoku* kyokus) {
  if (!kyokus) return fail(RV_ERR_INVALID, "kyokus is null");
  const int np = v->game_mode >= 3 ? 3 : 4;
  std::vector<uint8_t> oya(v->n), rw(v->n), honba(v->n);
  std::vector<uint32_t> ky(v->n);
  std::vector<int32_t> sc((size_t)v->n * MAXP, 0);
  for (int64_t i = 0; i < v->n; i++) {
    const rv_log_kyoku& k = kyokus[i];
    if (k.np != np) return fail(RV_ERR_INVALID, "kyoku " + std::to_string(i) + " has " + std::to_string(k.np) + " seats, the vector's game mode " + std::to_string(np));
    oya[i] = k.oya < np ? k.oya : 0;
    rw[i] = k.chang < 4 ? k.chang : 0;                       // LogKyoku::steps: chang -> Wind, anything else East
    honba[i] = k.ben;
    ky[i] = k.liqibang;
    for (int p = 0; p < np; p++) sc[(size_t)i * MAXP + p] = k.scores[p];
  }
  int rc = rv_vec_reset(v, oya.data(), rw.data(), honba.data(), ky.data(), sc.data(), nullptr);   // _initialize_round(.., None, scores)
  if (rc != RV_OK) return rc;
  rv_ctx* c = v->ctx;
  rc = ensure(&v->d_io_a, &v->io_a_bytes, sizeof(rv_log_kyoku) * (size_t)v->n);
  if (rc != RV_OK) return rc;
  rv_log_kyoku* d_k = (rv_log_kyoku*)v->d_io_a;
  CK(cudaMemcpyAsync(d_k, kyokus, sizeof(rv_log_kyoku) * (size_t)v->n, cudaMemcpyHostToDevice, c->stream));
  replay_begin_kernel<<<grid_for(v->n, 64), 64, 0, c->stream>>>(c->T, v->d_states, v->n, d_k);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  return RV_OK;
}
int rv_vec_replay_load(rv_vec* v, const rv_log_kyoku* kyokus, const rv_log_action* actions, const int64_t* first) {
  if (!kyokus || !first) return fail(RV_ERR_INVALID, "kyokus / first is null");
  const int64_t total = first[v->n];
  if (first[0] != 0 || total < 0) return fail(RV_ERR_INVALID, "first[] must start at 0 and be non-decreasing");
  for (int64_t i = 0; i < v->n; i++)
    if (first[i + 1] < first[i]) return fail(RV_ERR_INVALID, "first[] must start at 0 and be non-decreasing");
  if (total > 0 && !actions) return fail(RV_ERR_INVALID, "actions is null");
  int rc = rv_vec_replay_begin(v, kyokus);
  if (rc != RV_OK) return rc;
  rv_ctx* c = v->ctx;
  if (v->d_rp_actions) CK(cudaFree(v->d_rp_actions));
  v->d_rp_actions = nullptr;
  if (!v->d_rp_first) {
    CK(cudaMalloc(&v->d_rp_first, sizeof(int64_t) * (size_t)(v->n + 1)));
    CK(cudaMalloc(&v->d_rp_cursor, sizeof(int32_t) * (size_t)v->n));
    CK(cudaMalloc(&v->d_rp_live, sizeof(int32_t)));
  }
  if (total > 0) {
    CK(cudaMalloc(&v->d_rp_actions, sizeof(rv_log_action) * (size_t)total));
    CK(cudaMemcpyAsync(v->d_rp_actions, actions, sizeof(rv_log_action) * (size_t)total, cudaMemcpyHostToDevice, c->stream));
  }
  CK(cudaMemcpyAsync(v->d_rp_first, first, sizeof(int64_t) * (size_t)(v->n + 1), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemsetAsync(v->d_rp_cursor, 0, sizeof(int32_t) * (size_t)v->n, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return RV_OK;
}
int rv_vec_replay_advance(rv_vec* v, int64_t* n_applied) {
  if (!v->d_rp_first) return fail(RV_ERR_INVALID, "no log loaded (rv_vec_replay_load)");
  rv_ctx* c = v->ctx;
  CK(cudaSetDevice(c->device));
  CK(cudaMemsetAsync(v->d_rp_live, 0, sizeof(int32_t), c->stream));
  replay_advance_kernel<<<grid_for(v->n, 64), 64, 0, c->stream>>>(c->T, v->d_states, v->n, v->d_rp_actions, v->d_rp_first, v->d_rp_cursor,
                                                                 v->d_rp_live);
  CK(cudaGetLastError());
  if (n_applied) {                                   // NULL: asynchronous on the context's stream
    int32_t live = 0;
    CK(cudaMemcpyAsync(&live, v->d_rp_live, sizeof live, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    *n_applied = live;
  }
  return RV_OK;
}
int rv_vec_debug_call(rv_vec* v, int64_t game, int op, uint8_t out_tiles[5], int* n_out) {
  if (game < 0 || game >= v->n) return fail(RV_ERR_INVALID, "game index out of range");
  if (op < 0 || op > 6) return fail(RV_ERR_INVALID, "op must be 0..6");
  rv_ctx* c = v->ctx;
  CK(cudaSetDevice(c->device));
  uint8_t* d_out = nullptr;
  CK(cudaMalloc(&d_out, 8));
  debug_call_kernel<<<1, 1, 0, c->stream>>>(c->T, v->d_states + game, v->d_log, v->log_cap, game, op, d_out);
  uint8_t h[8] = {0};
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(h, d_out, 8, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(d_out);
  CK(e);
  if (n_out) *n_out = h[7];
  if (out_tiles && op == 1)
    for (int k = 0; k < 5; k++) out_tiles[k] = k < h[7] ? h[k] : RV_NONE;
  return RV_OK;
}
int rv_vec_clone(rv_vec* v, rv_vec** out) {
  if (!v || !out) return fail(RV_ERR_INVALID, "bad arguments");
  rv_ctx* c = v->ctx;
  rv_vec* w = nullptr;
  int rc = rv_vec_create(c, v->n, v->game_mode, v->rule_bits, nullptr, 0, v->log_cap, &w);
  if (rc != RV_OK) return rc;
  cudaError_t e = cudaMemcpyAsync(w->d_states, v->d_states, sizeof(G) * v->n, cudaMemcpyDeviceToDevice, c->stream);
  if (e == cudaSuccess && v->d_log)
    e = cudaMemcpyAsync(w->d_log, v->d_log, sizeof(uint32_t) * (size_t)v->n * v->log_cap, cudaMemcpyDeviceToDevice, c->stream);
  if (e == cudaSuccess && v->d_seq_cursor) {
    e = cudaMalloc(&w->d_seq_cursor, sizeof(uint32_t) * v->n * MAXP);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(w->d_seq_cursor, v->d_seq_cursor, sizeof(uint32_t) * v->n * MAXP, cudaMemcpyDeviceToDevice, c->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) {
    rv_vec_destroy(w);
    CK(e);
  }
  *out = w;
  return RV_OK;
}
int rv_vec_state_device_ptr(rv_vec* v, void** d_states) {
  *d_states = v->d_states;
  return RV_OK;
}
int rv_vec_events(rv_vec* v, int64_t game, uint32_t* out_words, uint32_t cap, uint32_t* n_words) {
  if (game < 0 || game >= v->n) return fail(RV_ERR_INVALID, "game index out of range");
  rv_ctx* c = v->ctx;
  CK(cudaSetDevice(c->device));
  G g;
  CK(cudaMemcpyAsync(&g, v->d_states + game, sizeof(G), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (!v->d_log) {
    if (n_words) *n_words = 0;
    return fail(RV_ERR_INVALID, "event logging is off for this VecEnv (log_cap_words == 0)");
  }
  // readable length: what the log holds, cut back to the last whole event when the capacity was exceeded (ev_push
  // sets overflow bit 2 and drops the words that do not fit, possibly in the middle of an event)
  uint32_t have = g.ev_words < v->log_cap ? g.ev_words : v->log_cap;
  if (g.ev_words > v->log_cap) {
    std::vector<uint32_t> all(have);
    if (have) {
      CK(cudaMemcpyAsync(all.data(), v->d_log + (size_t)game * v->log_cap, sizeof(uint32_t) * have, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
    }
    uint32_t i = 0;
    while (i < have) {
      uint32_t nw = (all[i] >> 8) & 0xFF;
      if (nw == 0 || i + nw > have) break;
      i += nw;
    }
    have = i;
  }
  if (n_words) *n_words = have;
  uint32_t ncopy = have < cap ? have : cap;
  if (out_words && ncopy) {
    CK(cudaMemcpyAsync(out_words, v->d_log + (size_t)game * v->log_cap, sizeof(uint32_t) * ncopy, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  return RV_OK;
}
// rows of the observation buffers: active seats per game and their exclusive scan (offsets[n] = total)
static int obs_offsets(rv_vec* v) {
  rv_ctx* c = v->ctx;
  int64_t n = v->n;
  if (!v->d_obs_counts) {
    CK(cudaMalloc(&v->d_obs_counts, sizeof(int32_t) * (n + 1)));
    CK(cudaMalloc(&v->d_obs_offsets, sizeof(int32_t) * (n + 1)));
    CK(cub::DeviceScan::ExclusiveSum(nullptr, v->scan_tmp_bytes, v->d_obs_counts, v->d_obs_offsets, (int)(n + 1), c->stream));
    CK(cudaMalloc(&v->d_scan_tmp, v->scan_tmp_bytes));
  }
  obs_count_kernel<<<grid_for(n + 1, 256), 256, 0, c->stream>>>(v->d_states, n, v->d_obs_counts);
  CK(cub::DeviceScan::ExclusiveSum(v->d_scan_tmp, v->scan_tmp_bytes, v->d_obs_counts, v->d_obs_offsets, (int)(n + 1), c->stream));
  return RV_OK;
}
static int obs_row_count(rv_vec* v, int64_t* n_obs) {
  if (!n_obs) return RV_OK;
  rv_ctx* c = v->ctx;
  if (!v->h_obs_total) CK(cudaMallocHost(&v->h_obs_total, sizeof(int32_t)));
  CK(cudaMemcpyAsync(v->h_obs_total, v->d_obs_offsets + v->n, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  *n_obs = *v->h_obs_total;
  return RV_OK;
}
int rv_vec_encode(rv_vec* v, float* d_obs, uint8_t* d_mask, int32_t* d_index, int64_t max_obs, int64_t* n_obs) {
  rv_ctx* c = v->ctx;
  CK(cudaSetDevice(c->device));
  int64_t n = v->n;
  int rc = obs_offsets(v);
  if (rc != RV_OK) return rc;
  const bool sanma = v->game_mode >= 3;
  if (d_mask) {
    if (!v->d_idbits) CK(cudaMalloc(&v->d_idbits, sizeof(uint32_t) * 3 * MAXP * n));
    if (sanma) legal_ids_kernel<true><<<grid_for(n, 128), 128, 0, c->stream>>>(c->T, v->d_states, n, v->d_idbits);
    else legal_ids_kernel<false><<<grid_for(n, 128), 128, 0, c->stream>>>(c->T, v->d_states, n, v->d_idbits);
  }
  if (d_obs || d_mask || d_index) {
    if (sanma)
      obs_encode_kernel<true><<<grid_for(n, 4), 128, 0, c->stream>>>(v->d_states, n, v->d_obs_offsets, v->d_idbits, d_obs, d_mask, d_index, max_obs);
    else
      obs_encode_kernel<false><<<grid_for(n, 4), 128, 0, c->stream>>>(v->d_states, n, v->d_obs_offsets, v->d_idbits, d_obs, d_mask, d_index, max_obs);
  }
  CK(cudaGetLastError());
  return obs_row_count(v, n_obs);
}
int rv_vec_encode_ext(rv_vec* v, float* d_obs, uint8_t* d_mask, int32_t* d_index, int64_t max_obs, int64_t* n_obs) {
  rv_ctx* c = v->ctx;
  const bool sanma = v->game_mode >= 3;                // sanma: Observation3P::encode_extended, 215 x 27 (obs_ext3.cuh)
  CK(cudaSetDevice(c->device));
  int64_t n = v->n;
  int rc = obs_offsets(v);
  if (rc != RV_OK) return rc;
  static DecayTab D = [] {
    DecayTab d;
    for (int age = 0; age < RV_RIVER_CAP; age++) d.w[age] = expf(-0.2f * (float)age);   // encode.rs:296-308, the C library's expf
    return d;
  }();
  if (!v->d_idbits) CK(cudaMalloc(&v->d_idbits, sizeof(uint32_t) * 3 * MAXP * n));
  if (sanma) {
    legal_ids_kernel<true, true><<<grid_for(n, 128), 128, 0, c->stream>>>(c->T, v->d_states, n, v->d_idbits);
    obs_ext_kernel<true><<<grid_for(n, 4), 128, 0, c->stream>>>(c->T, D, v->d_states, n, v->d_obs_offsets, v->d_idbits, d_obs, d_mask, d_index, max_obs);
  } else {
    legal_ids_kernel<false, true><<<grid_for(n, 128), 128, 0, c->stream>>>(c->T, v->d_states, n, v->d_idbits);
    obs_ext_kernel<false><<<grid_for(n, 4), 128, 0, c->stream>>>(c->T, D, v->d_states, n, v->d_obs_offsets, v->d_idbits, d_obs, d_mask, d_index, max_obs);
  }
  CK(cudaGetLastError());
  return obs_row_count(v, n_obs);
}
int rv_vec_encode_kawa(rv_vec* v, float* d_out, int32_t* d_index, int64_t max_obs, int64_t* n_obs) {
  rv_ctx* c = v->ctx;
  CK(cudaSetDevice(c->device));
  int rc = obs_offsets(v);
  if (rc != RV_OK) return rc;
  if (v->game_mode >= 3) obs_kawa_kernel<true><<<grid_for(v->n, 4), 128, 0, c->stream>>>(v->d_states, v->n, v->d_obs_offsets, d_out, d_index, max_obs);
  else obs_kawa_kernel<false><<<grid_for(v->n, 4), 128, 0, c->stream>>>(v->d_states, v->n, v->d_obs_offsets, d_out, d_index, max_obs);
  CK(cudaGetLastError());
  return obs_row_count(v, n_obs);
}
int rv_vec_observe_step_random(rv_vec* v, uint64_t agent_seed, float* d_obs, uint8_t* d_mask, int32_t* d_index, int64_t max_obs,
                               int64_t* n_obs) {
  rv_ctx* c = v->ctx;
  CK(cudaSetDevice(c->device));
  int64_t n = v->n;
  int rc = obs_offsets(v);
  if (rc != RV_OK) return rc;
  const char* impl = getenv("RV_OBS_STEP");                // "fused" selects the single kernel; read per call (the parity
  const bool use_queue = !(impl && strcmp(impl, "fused") == 0);   // tests exercise both implementations in one process)
  if (use_queue) {
    // tensor rows from the state as it is (many warps per SM: the encoder is a streaming kernel); then ONE env step per game
    // (thread per game — few, fat warps), which leaves the legal lists' id sets in HBM; then the mask rows.  Measured on a
    // B200 (65,536 hanchan): 203 + 270 + 10 us per step against 814 us for the single fused kernel, whose 2,048 one-warp
    // CTAs (168 registers, 24 KB of staging) leave the encoder 9 warps per SM.  A one-step launch of the class-queue
    // scheduler is no alternative: 194 us at best, milliseconds when the last games of a class keep SMs switching.
    const bool sanma = v->game_mode >= 3;
    if (!v->d_idbits) CK(cudaMalloc(&v->d_idbits, sizeof(uint32_t) * 3 * MAXP * n));
    if (!d_index) {
      if (!v->d_row_index) CK(cudaMalloc(&v->d_row_index, sizeof(int32_t) * MAXP * n));
      d_index = v->d_row_index;
      if (max_obs > (int64_t)MAXP * n) max_obs = (int64_t)MAXP * n;
    }
    // RV_OBS_OVERLAP (default on): snapshot what the encoder reads, then stream the rows out on a side stream while the step
    // kernel advances the games on the context stream; the streams join before the mask rows.
    static const int overlap = env_int("RV_OBS_OVERLAP", 2) - 1;
    const unsigned char* snap = nullptr;
    cudaStream_t enc_stream = c->stream;
    if (overlap && d_obs) {
      if (!v->d_obs_snap) CK(cudaMalloc(&v->d_obs_snap, (size_t)n * OBS_STAGE_BYTES));
      obs_snapshot_kernel<<<grid_for(n, 8), 256, 0, c->stream>>>(v->d_states, n, v->d_obs_snap);
      CK(cudaEventRecord(c->fork_ev, c->stream));
      CK(cudaStreamWaitEvent(c->aux[0], c->fork_ev, 0));
      snap = v->d_obs_snap;
      enc_stream = c->aux[0];
    }
    // overlapped: the step kernel is issued FIRST (its few long blocks take their SM share), the encoder fills what is left
    if (!snap) {
      if (sanma) obs_encode_kernel<true><<<grid_for(n, 4), 128, 0, enc_stream>>>(v->d_states, n, v->d_obs_offsets, nullptr, d_obs, nullptr, d_index, max_obs, snap);
      else obs_encode_kernel<false><<<grid_for(n, 4), 128, 0, enc_stream>>>(v->d_states, n, v->d_obs_offsets, nullptr, d_obs, nullptr, d_index, max_obs, snap);
      if (snap) CK(cudaEventRecord(c->join_ev[0], c->aux[0]));
    }
    static const int sorted = env_int("RV_STEP_SORTED", 3) - 1;    // RV_STEP_SORTED=1: thread per game, 2: regrouped, 3: staged (A/B)
    if (sorted == 2)
      CK(launch_step_staged<true>(c, v, agent_seed, v->d_idbits));
    else if (sorted)
      step_sorted_kernel<true><<<grid_for(n, SB), SB, 0, c->stream>>>(c->T, v->d_states, n, v->d_log, v->log_cap, agent_seed, v->d_steps, v->d_idbits);
    else
      step_random_kernel<true><<<grid_for(n, 128), 128, 0, c->stream>>>(c->T, v->d_states, n, v->d_log, v->log_cap, agent_seed, 1, v->d_steps,
                                                                        v->d_idbits);
    if (snap) {
      if (sanma) obs_encode_kernel<true><<<grid_for(n, 4), 128, 0, enc_stream>>>(v->d_states, n, v->d_obs_offsets, nullptr, d_obs, nullptr, d_index, max_obs, snap);
      else obs_encode_kernel<false><<<grid_for(n, 4), 128, 0, enc_stream>>>(v->d_states, n, v->d_obs_offsets, nullptr, d_obs, nullptr, d_index, max_obs, snap);
      if (snap) CK(cudaEventRecord(c->join_ev[0], c->aux[0]));
    }
    if (snap) CK(cudaStreamWaitEvent(c->stream, c->join_ev[0], 0));   // rows (and the row index) are complete from here on
    if (d_mask) {
      const int grid = (int)std::min<int64_t>(grid_for(std::min<int64_t>(max_obs, (int64_t)MAXP * n), 8), (int64_t)c->sm_count * 8);
      if (sanma) obs_mask_rows_kernel<true><<<grid, 256, 0, c->stream>>>(d_index, v->d_obs_offsets + n, v->d_idbits, d_mask, max_obs);
      else obs_mask_rows_kernel<false><<<grid, 256, 0, c->stream>>>(d_index, v->d_obs_offsets + n, v->d_idbits, d_mask, max_obs);
    }
    CK(cudaGetLastError());
    return obs_row_count(v, n_obs);
  }
  const unsigned grid = (unsigned)((n + 31) / 32);
  if (v->game_mode >= 3)
    observe_step_kernel<true><<<grid, 32, 0, c->stream>>>(c->T, v->d_states, n, v->d_log, v->log_cap, agent_seed, v->d_obs_offsets, d_obs,
                                                          d_mask, d_index, max_obs, v->d_steps);
  else
    observe_step_kernel<false><<<grid, 32, 0, c->stream>>>(c->T, v->d_states, n, v->d_log, v->log_cap, agent_seed, v->d_obs_offsets, d_obs,
                                                           d_mask, d_index, max_obs, v->d_steps);
  CK(cudaGetLastError());
  return obs_row_count(v, n_obs);
}
int rv_vec_encode_seq(rv_vec* v, int game_style, const uint32_t* start_words, uint16_t* d_sparse, float* d_numeric, uint16_t* d_prog,
                      int max_prog, uint16_t* d_cand, uint16_t* d_lens, int32_t* d_index, int64_t max_obs, int64_t* n_obs) {
  if (v->game_mode >= 3) return fail(RV_ERR_UNSUPPORTED, "sequence features are 4P only (observation/sequence_features.rs:811)");
  if (!v->d_log || v->log_cap == 0) return fail(RV_ERR_INVALID, "sequence features read the event log: create the vector with log_cap_words > 0");
  if (max_prog < 0 || (d_prog && max_prog == 0)) return fail(RV_ERR_INVALID, "max_prog must be positive");
  rv_ctx* c = v->ctx;
  CK(cudaSetDevice(c->device));
  int64_t n = v->n;
  if (!v->d_obs_counts) {
    CK(cudaMalloc(&v->d_obs_counts, sizeof(int32_t) * (n + 1)));
    CK(cudaMalloc(&v->d_obs_offsets, sizeof(int32_t) * (n + 1)));
    CK(cub::DeviceScan::ExclusiveSum(nullptr, v->scan_tmp_bytes, v->d_obs_counts, v->d_obs_offsets, (int)(n + 1), c->stream));
    CK(cudaMalloc(&v->d_scan_tmp, v->scan_tmp_bytes));
  }
  if (!v->d_seq_cursor) {
    CK(cudaMalloc(&v->d_seq_cursor, sizeof(uint32_t) * 4 * n));
    CK(cudaMemsetAsync(v->d_seq_cursor, 0, sizeof(uint32_t) * 4 * n, c->stream));
  }
  const uint32_t* d_start = nullptr;
  if (start_words) {
    if (!v->d_seq_start) CK(cudaMalloc(&v->d_seq_start, sizeof(uint32_t) * 4 * n));
    CK(cudaMemcpyAsync(v->d_seq_start, start_words, sizeof(uint32_t) * 4 * n, cudaMemcpyHostToDevice, c->stream));
    d_start = v->d_seq_start;
  }
  obs_count_kernel<<<grid_for(n + 1, 256), 256, 0, c->stream>>>(v->d_states, n, v->d_obs_counts);
  CK(cub::DeviceScan::ExclusiveSum(v->d_scan_tmp, v->scan_tmp_bytes, v->d_obs_counts, v->d_obs_offsets, (int)(n + 1), c->stream));
  seq_encode_kernel<<<grid_for(n, 64), 64, 0, c->stream>>>(c->T, v->d_states, n, v->d_log, v->log_cap, v->d_obs_offsets, d_start,
                                                          v->d_seq_cursor, game_style, d_sparse, d_numeric, d_prog, max_prog, d_cand,
                                                          d_lens, d_index, max_obs);
  CK(cudaGetLastError());
  if (n_obs || start_words) {      // (the caller's start_words must stay valid until the copy has run)
    int32_t total = 0;
    CK(cudaMemcpyAsync(&total, v->d_obs_offsets + n, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (n_obs) *n_obs = total;
  }
  return RV_OK;
}

// ---- several GPUs behind one handle -----------------------------------------------------------------------------------
// Games are independent: device k of G owns the contiguous global ids [k*N/G, (k+1)*N/G) and game g is seeded seed_base + g
// whatever G is, so results do not depend on the number of devices.  One host thread per device drives that device's
// vector on its own context and stream (the rollout call blocks until its device is done); nothing crosses devices on the
// step path.  The only cross-device work is the host-side sum of the per-device episode statistics at the end of a run.
int rv_multi_create(const int* devices, int n_devices, int64_t n_games, int game_mode, uint32_t rule_bits, uint64_t seed_base,
                    uint32_t log_cap_words, rv_multi** out) {
  if (!devices || n_devices <= 0 || !out || n_games < n_devices) return fail(RV_ERR_INVALID, "bad arguments");
  rv_multi* m = new rv_multi();
  m->own_ctx = true;
  m->ctx.assign(n_devices, nullptr);
  m->vec.assign(n_devices, nullptr);
  m->first.resize(n_devices + 1);
  for (int k = 0; k <= n_devices; k++) m->first[k] = n_games * k / n_devices;
  int rc = on_every_device(m, [&](int k) -> int {
    int r = rv_ctx_create(devices[k], &m->ctx[k]);
    if (r != RV_OK) return r;
    return rv_vec_create(m->ctx[k], m->first[k + 1] - m->first[k], game_mode, rule_bits, nullptr, seed_base + (uint64_t)m->first[k],
                         log_cap_words, &m->vec[k]);
  });
  if (rc != RV_OK) {
    std::string why = g_err;
    for (int k = 0; k < n_devices; k++) {
      if (m->vec[k]) rv_vec_destroy(m->vec[k]);
      if (m->ctx[k]) rv_ctx_destroy(m->ctx[k]);
    }
    delete m;
    return fail(rc, why);
  }
  *out = m;
  return RV_OK;
}
int rv_multi_destroy(rv_multi* m) {
  if (!m) return RV_OK;
  for (size_t k = 0; k < m->vec.size(); k++) {
    rv_vec_destroy(m->vec[k]);
    if (m->own_ctx) rv_ctx_destroy(m->ctx[k]);
  }
  delete m;
  return RV_OK;
}
int rv_multi_devices(const rv_multi* m) { return m ? (int)m->vec.size() : 0; }
int64_t rv_multi_size(const rv_multi* m) { return m ? m->first.back() : 0; }
int rv_multi_shard(rv_multi* m, int k, rv_vec** vec, int64_t* first_game, int64_t* n_games) {
  if (!m || k < 0 || k >= (int)m->vec.size()) return fail(RV_ERR_INVALID, "shard index out of range");
  if (vec) *vec = m->vec[k];
  if (first_game) *first_game = m->first[k];
  if (n_games) *n_games = m->first[k + 1] - m->first[k];
  return RV_OK;
}
int rv_multi_reset(rv_multi* m) {
  return on_every_device(m, [&](int k) -> int {
    int r = rv_vec_reset(m->vec[k], nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    return r != RV_OK ? r : rv_ctx_sync(m->ctx[k]);
  });
}
int rv_multi_reseed(rv_multi* m, uint64_t seed_base) {
  return on_every_device(m, [&](int k) -> int { return rv_vec_reseed(m->vec[k], nullptr, seed_base + (uint64_t)m->first[k]); });
}
int rv_multi_step_random(rv_multi* m, uint64_t agent_seed, uint32_t max_steps, uint64_t* steps_done) {
  std::vector<uint64_t> done(m->vec.size(), 0);
  int rc = on_every_device(m, [&](int k) -> int { return rv_vec_step_random(m->vec[k], agent_seed, max_steps, &done[k]); });
  if (steps_done) {
    *steps_done = 0;
    for (uint64_t d : done) *steps_done += d;
  }
  return rc;
}
int rv_multi_results(rv_multi* m, uint8_t* done, int32_t* scores, uint8_t* ranks) {
  return on_every_device(m, [&](int k) -> int {
    const int64_t f = m->first[k];
    return rv_vec_results(m->vec[k], done ? done + f : nullptr, scores ? scores + f * MAXP : nullptr, ranks ? ranks + f * MAXP : nullptr);
  });
}
int rv_multi_counters(rv_multi* m, uint32_t* step_count, uint32_t* kyoku_count, uint32_t* ev_count, uint64_t* ev_hash) {
  return on_every_device(m, [&](int k) -> int {
    const int64_t f = m->first[k];
    return rv_vec_counters(m->vec[k], step_count ? step_count + f : nullptr, kyoku_count ? kyoku_count + f : nullptr,
                           ev_count ? ev_count + f : nullptr, ev_hash ? ev_hash + f : nullptr);
  });
}
// End-of-run episode statistics (SURVEY §8 e): per-device partial sums (one gather per device, in parallel), summed on the
// host — a few hundred bytes cross the host, no device-to-device traffic.
int rv_multi_stats(rv_multi* m, rv_run_stats* out) {
  if (!m || !out) return fail(RV_ERR_INVALID, "bad arguments");
  const int G = (int)m->vec.size();
  std::vector<rv_run_stats> part(G);
  int rc = on_every_device(m, [&](int k) -> int {
    const int64_t n = m->first[k + 1] - m->first[k];
    std::vector<uint8_t> done(n), ranks(n * MAXP);
    std::vector<int32_t> scores(n * MAXP);
    std::vector<uint32_t> steps(n), kyoku(n);
    int r = rv_vec_results(m->vec[k], done.data(), scores.data(), ranks.data());
    if (r != RV_OK) return r;
    r = rv_vec_counters(m->vec[k], steps.data(), kyoku.data(), nullptr, nullptr);
    if (r != RV_OK) return r;
    rv_run_stats& s = part[k];
    memset(&s, 0, sizeof s);
    const int np = m->vec[k]->game_mode >= 3 ? 3 : 4;
    for (int64_t g = 0; g < n; g++) {
      s.games += 1;
      s.games_done += done[g] ? 1 : 0;
      s.env_steps += steps[g];
      s.rounds += kyoku[g];
      for (int p = 0; p < np; p++) {
        s.score_sum[p] += scores[g * MAXP + p];
        if (done[g]) s.rank_hist[p][ranks[g * MAXP + p] - 1] += 1;
      }
    }
    return RV_OK;
  });
  if (rc != RV_OK) return rc;
  memset(out, 0, sizeof *out);
  for (const rv_run_stats& s : part) {
    out->games += s.games;
    out->games_done += s.games_done;
    out->env_steps += s.env_steps;
    out->rounds += s.rounds;
    for (int p = 0; p < MAXP; p++) {
      out->score_sum[p] += s.score_sum[p];
      for (int r = 0; r < MAXP; r++) out->rank_hist[p][r] += s.rank_hist[p][r];
    }
  }
  return RV_OK;
}
int rv_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(rv_game_state);
    case 1: return (int)sizeof(rv_hand_query);
    case 2: return (int)sizeof(rv_hand_result);
    case 3: return (int)sizeof(rv_action);
    case 4: return (int)sizeof(rv_mjai_event);
    case 5: return (int)sizeof(rv_run_stats);
    case 6: return (int)sizeof(rv_log_action);
    case 7: return (int)sizeof(rv_log_kyoku);
  }
  return -1;
}
}  // extern "C"
