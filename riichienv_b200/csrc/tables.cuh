// Suit-pattern lookup tables, generated ON THE DEVICE at context creation.
//
// The reference decides agari / waits by recursive backtracking over a 34-histogram
// (agari.rs:63-245, hand_evaluator.rs:178-213) and shanten by Cryolite's nyanten
// perfect-hash tables (shanten.rs:163-196, 536 KB of data files).  Here both come
// from one direct-indexed table per tile class, keyed by the base-5 number of the
// nine (seven) tile counts of a suit, so a whole-hand query is four L2-resident
// loads plus bit operations:
//
//   info[key]  (u32)  bit0  M : the suit splits into mentsu only
//                     bit1  P : the suit splits into mentsu + exactly one pair
//                     bits 2..10   waitM[i] : count[i] < 4 and (suit + tile i) is M
//                     bits 11..19  waitP[i] : count[i] < 4 and (suit + tile i) is P
//                     bits 20..23  DM, DP, DWM, DWP : some tile i of the suit can be removed so that
//                                  (suit - tile i) is M / is P / has a non-empty waitM / waitP
//                                  (answers "does any discard leave the hand tenpai" in 4 loads)
//   cost[key]  (u64)  ten nibbles: min #tiles missing to hold k mentsu (k=0..4),
//                     without (nibble 2k) / with (nibble 2k+1) a pair, when no tile
//                     may be used more than four times — the quantity nyanten encodes.
//
// Number suits: 5^9 = 1,953,125 entries; honors: 5^7 = 78,125 entries.
// 12 B/entry -> 24.4 MB, resident in the 126 MB L2 after first touch.
#pragma once
#include <cstdint>
#ifdef __CUDACC__
#include <cuda_runtime.h>
#endif

namespace rv {

constexpr int SUIT_KEYS = 1953125;   // 5^9
constexpr int HONOR_KEYS = 78125;    // 5^7

struct Tables {
  const uint32_t* suit_info;
  const uint32_t* honor_info;
  const uint64_t* suit_cost;
  const uint64_t* honor_cost;
};

__host__ __device__ inline int pow5(int i) {
  // computed, not tabulated: keeps the value in registers (no local/constant array indexing)
  int v = 1;
  if (i & 1) v *= 5;
  if (i & 2) v *= 25;
  if (i & 4) v *= 625;
  if (i & 8) v *= 390625;
  return v;
}

// D-bits contributed by the entry `o` of (suit - one tile)
__host__ __device__ inline uint32_t discard_bits(uint32_t o) {
  uint32_t e = 0;
  if (o & 1u) e |= 1u << 20;
  if (o & 2u) e |= 1u << 21;
  if ((o >> 2) & 0x1FFu) e |= 1u << 22;
  if ((o >> 11) & 0x1FFu) e |= 1u << 23;
  return e;
}

// DP over tile positions.  State: (s1 = sequences started at i-1, s2 = sequences
// started at i-2, k = mentsu so far, h = pair used).  best[k][h] = min missing tiles.
template <int N, bool SEQ>
__device__ inline void suit_cost_dp(const uint8_t* c, uint8_t best[5][2]) {
  // dp[s1][s2][k][h]
  uint8_t cur[5][5][5][2], nxt[5][5][5][2];
  #pragma unroll 1
  for (int a = 0; a < 5; a++)
    for (int b = 0; b < 5; b++)
      for (int k = 0; k < 5; k++) cur[a][b][k][0] = cur[a][b][k][1] = 99;
  cur[0][0][0][0] = 0;
  #pragma unroll 1
  for (int i = 0; i < N; i++) {
    for (int a = 0; a < 5; a++)
      for (int b = 0; b < 5; b++)
        for (int k = 0; k < 5; k++) nxt[a][b][k][0] = nxt[a][b][k][1] = 99;
    for (int s1 = 0; s1 < 5; s1++)
      for (int s2 = 0; s1 + s2 < 5; s2++)
        for (int k = 0; k < 5; k++)
          for (int h = 0; h < 2; h++) {
            int base = cur[s1][s2][k][h];
            if (base >= 99) continue;
            int max_s0 = (SEQ && i <= N - 3) ? 4 : 0;
            for (int s0 = 0; s0 <= max_s0; s0++)
              for (int q = 0; q < 2; q++)
                for (int p = 0; p + h < 2; p++) {
                  int use = s2 + s1 + s0 + 3 * q + 2 * p;
                  if (use > 4) continue;
                  int nk = k + s0 + q;
                  if (nk > 4) continue;
                  int cost = base + (use > c[i] ? use - c[i] : 0);
                  uint8_t& dst = nxt[s0][s1][nk][h + p];
                  if (cost < dst) dst = (uint8_t)cost;
                }
          }
    for (int a = 0; a < 5; a++)
      for (int b = 0; b < 5; b++)
        for (int k = 0; k < 5; k++) {
          cur[a][b][k][0] = nxt[a][b][k][0];
          cur[a][b][k][1] = nxt[a][b][k][1];
        }
  }
  for (int k = 0; k < 5; k++) {
    best[k][0] = cur[0][0][k][0];
    best[k][1] = cur[0][0][k][1];
  }
}

#ifdef __CUDACC__
template <int N, bool SEQ>
__global__ void gen_cost_kernel(uint64_t* cost, uint32_t* info, int n_keys) {
  int key = blockIdx.x * blockDim.x + threadIdx.x;
  if (key >= n_keys) return;
  uint8_t c[9];
  int k = key, sum = 0;
  for (int i = 0; i < N; i++) {
    c[i] = (uint8_t)(k % 5);
    k /= 5;
    sum += c[i];
  }
  if (sum > 14) {
    cost[key] = 0xFFFFFFFFFFull;
    info[key] = 0;
    return;
  }
  uint8_t best[5][2];
  suit_cost_dp<N, SEQ>(c, best);
  uint64_t packed = 0;
  for (int m = 0; m < 5; m++) {
    uint64_t a = best[m][0] > 15 ? 15 : best[m][0], b = best[m][1] > 15 ? 15 : best[m][1];
    packed |= a << (8 * m);
    packed |= b << (8 * m + 4);
  }
  cost[key] = packed;
  uint32_t e = 0;
  if (sum % 3 == 0 && best[sum / 3][0] == 0) e |= 1u;
  if (sum % 3 == 2 && best[sum / 3][1] == 0) e |= 2u;
  info[key] = e;
}

template <int N>
__global__ void gen_wait_kernel(uint32_t* info, int n_keys) {
  int key = blockIdx.x * blockDim.x + threadIdx.x;
  if (key >= n_keys) return;
  int k = key, sum = 0;
  uint8_t c[9];
  for (int i = 0; i < N; i++) {
    c[i] = (uint8_t)(k % 5);
    k /= 5;
    sum += c[i];
  }
  if (sum > 13) return;
  uint32_t e = info[key] & 3u;
  for (int i = 0; i < N; i++) {
    if (c[i] >= 4) continue;
    uint32_t o = info[key + pow5(i)] & 3u;   // written by gen_cost_kernel (previous launch)
    if (o & 1u) e |= 1u << (2 + i);
    if (o & 2u) e |= 1u << (11 + i);
  }
  info[key] = e;
}

template <int N>
__global__ void gen_discard_kernel(uint32_t* info, int n_keys) {
  int key = blockIdx.x * blockDim.x + threadIdx.x;
  if (key >= n_keys) return;
  int k = key, sum = 0;
  uint8_t c[9];
  for (int i = 0; i < N; i++) {
    c[i] = (uint8_t)(k % 5);
    k /= 5;
    sum += c[i];
  }
  if (sum > 14 || sum == 0) return;
  uint32_t e = info[key] & 0xFFFFFu;
  for (int i = 0; i < N; i++) {
    if (c[i] == 0) continue;
    uint32_t o = info[key - pow5(i)];   // bits 0..19 were final after gen_wait_kernel (previous launch)
    e |= discard_bits(o);
  }
  info[key] = e;
}
#endif  // __CUDACC__

}  // namespace rv
