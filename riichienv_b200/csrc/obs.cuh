// FEATURE_ENCODING observation tensor (74 x 34 f32, channel-major) and 82-id action mask.
//
// Replaces Observation::encode (observation/python.rs:457-806), Observation::mask (98-111) and
// Action::encode (action.rs:158-227) for the seat snapshots GameState::get_observation builds
// (state/mod.rs:189-263).  Every channel is one of three shapes, so a channel is described by
// (kind, 34-bit mask, broadcast value) and the 10 KB tensor is then streamed out by the whole warp
// with 16-byte stores:
//   MASK  : 1.0 at the set tile kinds, else 0.0        (hand, melds, dora, discards, waits, winds)
//   BCAST : one value in all 34 columns                (counts / scores / flags, each = small int / const,
//                                                       IEEE f32 division as the reference does)
//   SEEN  : per-kind visible count / 4                 (channel 63)
//
// Sanma (Observation3P::encode, observation_3p/python.rs:402-708; mask 102-114; ActionEncoder::encode_3p,
// action.rs:262-346): the same 74 channels over 27 compact columns (1m, 9m, 1-9p, 1-9s, honors —
// observation_3p/helpers.rs:8-15), three relative seats per group (the fourth channel of each group stays zero),
// 108 tiles, the 1m<->9m dora wrap, a 60-id action space.  `SANMA` selects it at compile time.
#pragma once
#include "game.cuh"

namespace rv {

constexpr int OBS_CH = 74;
constexpr int OBS_W = 34;
constexpr int OBS_W3 = 27;     // TILE_DIM_3P
constexpr int OBS_IDS = 82, OBS_IDS3 = 60;
// compact column -> tile kind (inverse of tile34_to_compact, observation_3p/helpers.rs:8-15)
__device__ __forceinline__ int obs_col_kind3(int col) { return col == 0 ? 0 : col + 7; }
enum { OBS_MASK = 0, OBS_BCAST = 1, OBS_SEEN = 2 };

__device__ __forceinline__ uint64_t river_tail_mask(const G& g, int p, int from_last) {
  // kind of the (from_last)-th most recent discard, as a one-hot mask (0 if absent)
  int n = min((int)g.n_river[p], RV_RIVER_CAP);
  int i = n - 1 - from_last;
  return i >= 0 ? 1ull << (cold(g).river[p][i] >> 2) : 0;
}
__device__ __forceinline__ int obs_next_kind(int k) {  // observation/helpers.rs:25-50, k in 0..33
  // k + 1, wrapping 9 -> 1 inside a suit, North -> East, Red -> White (no integer modulo: this is inlined many times)
  constexpr uint64_t NINES = (1ull << 8) | (1ull << 17) | (1ull << 26);
  return k + 1 - (int)((NINES >> k) & 1) * 9 - (k == 30 ? 4 : 0) - (k == 33 ? 3 : 0);
}
__device__ __forceinline__ int obs_next_kind_sanma(int k) {  // observation_3p/helpers.rs:41-50
  if (k == 0) return 8;
  if (k == 8) return 0;
  if (k <= 7) return k;
  return obs_next_kind(k);
}
// visible dora count of seat q as seen by pid (observation/python.rs:684-725, observation_3p/python.rs:595-630)
template <bool SANMA>
__device__ inline int obs_dora_count(const G& g, int pid, int q) {
  int cnt = 0;
  for (int d = 0; d < g.n_dora; d++) {
    int dk = SANMA ? obs_next_kind_sanma(g.dora_ind[d] >> 2) : obs_next_kind(g.dora_ind[d] >> 2);
    for (int m = 0; m < g.n_melds[q]; m++)
      for (int k = 0; k < 4; k++) {
        int t = g.meld_tiles[q][m][k];
        if (t != RV_NONE && (t >> 2) == dk) cnt++;
      }
    int n = min((int)g.n_river[q], RV_RIVER_CAP);
    for (int i = 0; i < n; i++)
      if ((cold(g).river[q][i] >> 2) == dk) cnt++;
    if (q == pid) cnt += (int)((g.c_cnt[pid][dk / 9] >> (4 * (dk % 9))) & 15);
  }
  return cnt & 0xFF;
}

template <bool SANMA>
__device__ inline void obs_channel(const G& g, int pid, int ch, int& kind, uint64_t& mask, float& val) {
  kind = OBS_MASK;
  mask = 0;
  val = 0.0f;
  constexpr int NPV = SANMA ? 3 : 4;
  auto rel = [&](int i) { return SANMA ? (pid + i) % 3 : (pid + i) & 3; };
  // sanma: each per-seat group has three channels; the fourth (22-25, 29, 34, 42, 46, 52, 58, 62) is never written
  if (SANMA && ((ch >= 22 && ch <= 25) || ch == 29 || ch == 34 || ch == 42 || ch == 46 || ch == 52 || ch == 58 || ch == 62)) return;
  auto bc = [&](float v) { kind = OBS_BCAST; val = v; };
  if (ch <= 3) {           // hand count >= ch+1
    for (int su = 0; su < 4; su++) {
      uint64_t x = g.c_cnt[pid][su];
      for (int i = 0; i < 9; i++)
        if ((int)((x >> (4 * i)) & 15) >= ch + 1) mask |= 1ull << (9 * su + i);
    }
  } else if (ch == 4) {    // red fives in hand
    for (int k = 0; k < g.hand_len[pid]; k++) {
      int t = g.hand[pid][k];
      if (t == 16 || t == 52 || t == 88) mask |= 1ull << (t >> 2);
    }
  } else if (ch <= 8) {    // own melds 1..4
    int m = ch - 5;
    if (m < g.n_melds[pid])
      for (int k = 0; k < 4; k++) {
        int t = g.meld_tiles[pid][m][k];
        if (t != RV_NONE) mask |= 1ull << (t >> 2);
      }
  } else if (ch == 9) {
    for (int d = 0; d < g.n_dora; d++) mask |= 1ull << (g.dora_ind[d] >> 2);
  } else if (ch <= 13) {
    mask = river_tail_mask(g, pid, ch - 10);
  } else if (ch <= 25) {
    int o = (ch - 14) / 4, j = (ch - 14) % 4;
    mask = river_tail_mask(g, rel(o + 1), j);
  } else if (ch <= 29) {
    bc((float)g.n_river[rel(ch - 26)] / 24.0f);
  } else if (ch == 30) {
    int used = g.hand_len[pid] + g.n_dora;
    for (int p = 0; p < NPV; p++) {
      used += g.n_river[p];
      for (int m = 0; m < g.n_melds[p]; m++) used += (g.meld_tiles[p][m][3] != RV_NONE) ? 4 : 3;
    }
    int left = (SANMA ? 108 : 136) - used;
    bc((float)(left < 0 ? 0 : left) / 70.0f);
  } else if (ch <= 34) {
    bc((g.flags[rel(ch - 31)] & RV_F_RIICHI_DECLARED) ? 1.0f : 0.0f);
  } else if (ch == 35) {
    if (27 + g.round_wind < 34) mask = 1ull << (27 + g.round_wind);
  } else if (ch == 36) {
    mask = 1ull << (27 + (SANMA ? (pid + 3 - g.oya) % 3 : (pid + 4 - g.oya) & 3));
  } else if (ch == 37) {
    bc((float)g.honba / 10.0f);
  } else if (ch == 38) {
    bc((float)g.riichi_sticks / 5.0f);
  } else if (ch <= 42) {
    int s = g.score[rel(ch - 39)];
    s = s < 0 ? 0 : (s > 100000 ? 100000 : s);
    bc((float)s / 100000.0f);
  } else if (ch <= 46) {
    int s = g.score[rel(ch - 43)];
    s = s < 0 ? 0 : (s > 30000 ? 30000 : s);
    bc((float)s / 30000.0f);
  } else if (ch == 47) {
    mask = g.c_waits[pid];
  } else if (ch == 48) {
    bc(g.c_waits[pid] != 0 ? 1.0f : 0.0f);
  } else if (ch <= 52) {
    int rank = 0;
    for (int p = 0; p < NPV; p++)
      if (g.score[p] > g.score[pid]) rank++;
    bc(rank == ch - 49 ? 1.0f : 0.0f);
  } else if (ch == 53) {
    bc((float)g.kyoku_idx / 8.0f);
  } else if (ch == 54) {
    bc(((float)g.round_wind * 4.0f + (float)g.kyoku_idx) / 7.0f);
  } else if (ch <= 58) {
    bc((float)obs_dora_count<SANMA>(g, pid, rel(ch - 55)) / 12.0f);
  } else if (ch <= 62) {
    bc((float)g.n_melds[rel(ch - 59)] / 4.0f);
  } else if (ch == 63) {
    kind = OBS_SEEN;
  } else if (ch <= 67) {
    mask = river_tail_mask(g, pid, 4 + (ch - 64));
  } else if (ch <= 69) {
    mask = river_tail_mask(g, rel(1), 4 + (ch - 68));
  } else {
    // 70-73: tsumogiri flags are always empty in the live env (observation/mod.rs:105, observation_3p/mod.rs:100) -> zeros
  }
}
// channel 63: own hand + all melds + all rivers + dora indicators, per kind
__device__ inline int obs_seen(const G& g, int pid, int kind) {
  int c = (int)((g.c_cnt[pid][kind / 9] >> (4 * (kind % 9))) & 15);
  for (int p = 0, np = num_players(g); p < np; p++) {
    for (int m = 0; m < g.n_melds[p]; m++)
      for (int k = 0; k < 4; k++) {
        int t = g.meld_tiles[p][m][k];
        if (t != RV_NONE && (t >> 2) == kind) c++;
      }
    int n = min((int)g.n_river[p], RV_RIVER_CAP);
    for (int i = 0; i < n; i++)
      if ((cold(g).river[p][i] >> 2) == kind) c++;
  }
  for (int d = 0; d < g.n_dora; d++)
    if ((g.dora_ind[d] >> 2) == kind) c++;
  return c;
}
__device__ __forceinline__ float obs_value(int kind, uint64_t mask, float val, int seen, int col) {
  if (kind == OBS_BCAST) return val;
  if (kind == OBS_SEEN) return (float)seen / 4.0f;
  return ((mask >> col) & 1) ? 1.0f : 0.0f;
}

// ActionEncoder::encode_3p (action.rs:262-346); -1 if not encodable
__device__ inline int action_id_3p(const rv_action& a) {
  auto compact = [](int t34) { return t34 == 0 ? 0 : (t34 >= 8 && t34 < 34 ? t34 - 7 : -1); };
  switch (a.type) {
    case RV_DISCARD: return a.tile == RV_NONE ? -1 : compact(a.tile >> 2);
    case RV_RIICHI: return 27;
    case RV_CHI: return -1;
    case RV_PON: return 28;
    case RV_DAIMINKAN: {
      int c = a.tile == RV_NONE ? -1 : compact(a.tile >> 2);
      return c < 0 ? -1 : 29 + c;
    }
    case RV_ANKAN:
    case RV_KAKAN: {
      int c = a.n_consume == 0 ? -1 : compact(a.consume[0] >> 2);
      return c < 0 ? -1 : 29 + c;
    }
    case RV_RON:
    case RV_TSUMO: return 56;
    case RV_KYUSHU_KYUHAI: return 57;
    case RV_PASS: return 58;
    case RV_KITA: return 59;
  }
  return -1;
}
// Action::encode (action.rs:158-227); -1 if not encodable
__device__ inline int action_id(const rv_action& a) {
  switch (a.type) {
    case RV_DISCARD: return a.tile == RV_NONE ? -1 : (a.tile >> 2);
    case RV_RIICHI: return 37;
    case RV_CHI: {
      if (a.tile == RV_NONE || a.n_consume < 2) return -1;
      int t = a.tile >> 2, x = a.consume[0] >> 2, y = a.consume[1] >> 2;
      int lo = min(t, min(x, y)), hi = max(t, max(x, y));
      if (hi - lo != 2 || x == y || x == t || y == t) return -1;
      return t == lo ? 38 : (t == hi ? 40 : 39);
    }
    case RV_PON: return 41;
    case RV_DAIMINKAN: return a.tile == RV_NONE ? -1 : 42 + (a.tile >> 2);
    case RV_ANKAN:
    case RV_KAKAN: return a.n_consume == 0 ? -1 : 42 + (a.consume[0] >> 2);
    case RV_RON:
    case RV_TSUMO: return 79;
    case RV_KYUSHU_KYUHAI: return 80;
    case RV_PASS: return 81;
  }
  return -1;
}

}  // namespace rv

#ifdef __CUDACC__
namespace rv {
// ---------------------------------------------------------------------------------------------------------------------
// Warp-cooperative encoder (the production path; obs_channel/obs_value above are the per-element statement of the same
// tensor, kept for single observations and for the host-compiled parity tests).
//
// One warp writes one observation row.  The work is arranged so that lanes run the same instructions:
//   gather      every lane owns one 4-byte word of the seat-major river array (32 words = the four 32-tile rivers), lanes
//               0-15 one meld each, lanes 0-4 one dora indicator, and lane k the hand count of tile kind k; per-kind "seen"
//               counts, per-seat dora counts and the last eight discards of every seat land in shared memory through
//               shared-memory atomics, the hand>=k masks come out of warp ballots;
//   describe    a channel is (34-bit column mask, value): lane L describes channels L, L+32, L+64 from those gathered
//               pieces (a handful of instructions per branch — no loops over tiles);
//   stream      2,516 floats (sanma 1,998) leave as 16-byte (8-byte) streaming stores, each built from two table reads
//               and a funnel of the masks of the (at most two) channels it spans.  Channel 63 (seen/4) is patched in.
struct __align__(16) ObsDesc {   // one channel: out[col] = (mask >> col) & 1 ? val : 0
  uint64_t mask;
  float val;
  uint32_t pad;
};
struct ObsScratch {
  ObsDesc d[OBS_CH + 2];
  int seen[36];
  int dora[4];
  uint8_t rtail[4][8];   // kind of the j-th most recent discard of seat q, 0xFF = none
  uint8_t dm[36];        // per tile kind: how many dora indicators point at it
};
__device__ __forceinline__ uint64_t obs_compact3(uint64_t m) { return (m & 1) | ((m >> 7) & ~1ull); }   // 34 kinds -> 27 columns

// gather + describe: leaves the channel descriptors S.d[0..73] and the per-kind seen counts S.seen.  `ch30_adj` is added to
// the "tiles left" count of channel 30 (0 for Observation::encode; encode_base_into, which encode_extended uses, counts every
// called meld one tile short — observation/encode.rs:94-110).
template <bool SANMA>
__device__ __forceinline__ void obs_describe_warp(const G& g, const uint8_t* river, int pid, ObsScratch& S, int lane, int ch30_adj = 0) {
  constexpr int NPV = SANMA ? 3 : 4, W = SANMA ? OBS_W3 : OBS_W;
  constexpr uint64_t ALL = (1ull << W) - 1;
  auto rel = [&](int i) { return SANMA ? (pid + i) % 3 : (pid + i) & 3; };
  // ---- gather
  S.seen[lane] = 0;
  if (lane < 4) S.seen[32 + lane] = 0, S.dora[lane] = 0;
  S.rtail[lane >> 3][lane & 7] = 0xFF;
  const int nd = g.n_dora;
  uint64_t dkp = ~0ull;                                     // dora kinds, one per byte
  for (int d = 0; d < nd; d++) {
    int k = g.dora_ind[d] >> 2;
    k = SANMA ? obs_next_kind_sanma(k) : obs_next_kind(k);
    dkp = (dkp & ~(0xFFull << (8 * d))) | ((uint64_t)k << (8 * d));
  }
  {
    int c = 0, c2 = 0;
    for (int d = 0; d < nd; d++) {
      const int k = (int)((dkp >> (8 * d)) & 0xFF);
      c += k == lane ? 1 : 0;
      c2 += k == 32 + lane ? 1 : 0;
    }
    S.dm[lane] = (uint8_t)c;
    if (lane < 2) S.dm[32 + lane] = (uint8_t)c2;
  }
  __syncwarp();
  auto dmatch = [&](int kind) { return (int)S.dm[kind]; };
  {
    const int q = lane >> 3, w = lane & 7;
    const int n = min((int)g.n_river[q], RV_RIVER_CAP);
    if (4 * w < n) {
      const uint32_t rw = reinterpret_cast<const uint32_t*>(river)[lane];
      #pragma unroll
      for (int b = 0; b < 4; b++) {
        const int i = 4 * w + b;
        if (i < n) {
          const int kind = (int)((rw >> (8 * b)) & 0xFF) >> 2;
          atomicAdd(&S.seen[kind], 1);
          const int dm = dmatch(kind);
          if (dm) atomicAdd(&S.dora[q], dm);
          const int fl = n - 1 - i;
          if (fl < 8) S.rtail[q][fl] = (uint8_t)kind;
        }
      }
    }
  }
  if (lane < 16) {
    const int q = lane >> 2, m = lane & 3;
    if (m < g.n_melds[q]) {
      const uint32_t mw = *reinterpret_cast<const uint32_t*>(&g.meld_tiles[q][m][0]);
      #pragma unroll
      for (int b = 0; b < 4; b++) {
        const int t = (int)((mw >> (8 * b)) & 0xFF);
        if (t != RV_NONE) {
          atomicAdd(&S.seen[t >> 2], 1);
          const int dm = dmatch(t >> 2);
          if (dm) atomicAdd(&S.dora[q], dm);
        }
      }
    }
  }
  if (lane < nd) atomicAdd(&S.seen[g.dora_ind[lane] >> 2], 1);
  uint64_t hge[4];                                          // hand count >= 1..4, per kind
  {
    const int c = (int)((g.c_cnt[pid][lane / 9] >> (4 * (lane % 9))) & 15);                 // kinds 0..31
    const int c2 = lane < 2 ? (int)((g.c_cnt[pid][3] >> (4 * (lane + 5))) & 15) : 0;        // kinds 32, 33
    if (c) {
      atomicAdd(&S.seen[lane], c);
      const int dm = dmatch(lane);
      if (dm) atomicAdd(&S.dora[pid], c * dm);
    }
    if (c2) {
      atomicAdd(&S.seen[32 + lane], c2);
      const int dm = dmatch(32 + lane);
      if (dm) atomicAdd(&S.dora[pid], c2 * dm);
    }
    #pragma unroll
    for (int k = 0; k < 4; k++)
      hge[k] = (uint64_t)__ballot_sync(0xFFFFFFFFu, c > k) | ((uint64_t)(__ballot_sync(0xFFFFFFFFu, c2 > k) & 3u) << 32);
  }
  uint64_t red;
  {
    const int t = lane < g.hand_len[pid] ? g.hand[pid][lane < RV_HAND_CAP ? lane : 0] : RV_NONE;
    red = (__any_sync(0xFFFFFFFFu, t == 16) ? 1ull << 4 : 0) | (__any_sync(0xFFFFFFFFu, t == 52) ? 1ull << 13 : 0) |
          (__any_sync(0xFFFFFFFFu, t == 88) ? 1ull << 22 : 0);
  }
  __syncwarp();
  int used = S.seen[lane] + (lane < 2 ? S.seen[32 + lane] : 0);      // tiles visible to the seat = rivers + melds + hand + indicators
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) used += __shfl_xor_sync(0xFFFFFFFFu, used, o);
  // ---- describe
  for (int ch = lane; ch < OBS_CH + 2; ch += 32) {
    uint64_t m = 0;
    float v = 1.0f;
    bool bcast = false;
    auto bc = [&](float x) { bcast = true; v = x; };
    auto tail = [&](int q, int j) { int k = S.rtail[q][j]; return k != 0xFF ? 1ull << k : 0ull; };
    const bool dead = ch >= OBS_CH ||
                      (SANMA && ((ch >= 22 && ch <= 25) || ch == 29 || ch == 34 || ch == 42 || ch == 46 || ch == 52 || ch == 58 || ch == 62));
    if (dead) {
    } else if (ch <= 3) {
      m = hge[ch];
    } else if (ch == 4) {
      m = red;
    } else if (ch <= 8) {
      const int mi = ch - 5;
      if (mi < g.n_melds[pid]) {
        const uint32_t mw = *reinterpret_cast<const uint32_t*>(&g.meld_tiles[pid][mi][0]);
        #pragma unroll
        for (int b = 0; b < 4; b++) {
          const int t = (int)((mw >> (8 * b)) & 0xFF);
          if (t != RV_NONE) m |= 1ull << (t >> 2);
        }
      }
    } else if (ch == 9) {
      for (int d = 0; d < nd; d++) m |= 1ull << (g.dora_ind[d] >> 2);
    } else if (ch <= 13) {
      m = tail(pid, ch - 10);
    } else if (ch <= 25) {
      m = tail(rel(((ch - 14) >> 2) + 1), (ch - 14) & 3);
    } else if (ch <= 29) {
      bc((float)g.n_river[rel(ch - 26)] / 24.0f);
    } else if (ch == 30) {
      const int left = (SANMA ? 108 : 136) - used + ch30_adj;
      bc((float)(left < 0 ? 0 : left) / 70.0f);
    } else if (ch <= 34) {
      bc((g.flags[rel(ch - 31)] & RV_F_RIICHI_DECLARED) ? 1.0f : 0.0f);
    } else if (ch == 35) {
      if (27 + g.round_wind < 34) m = 1ull << (27 + g.round_wind);
    } else if (ch == 36) {
      m = 1ull << (27 + (SANMA ? (pid + 3 - g.oya) % 3 : (pid + 4 - g.oya) & 3));
    } else if (ch == 37) {
      bc((float)g.honba / 10.0f);
    } else if (ch == 38) {
      bc((float)g.riichi_sticks / 5.0f);
    } else if (ch <= 46) {
      const bool wide = ch <= 42;
      int s = g.score[rel(wide ? ch - 39 : ch - 43)];
      const int cap = wide ? 100000 : 30000;
      s = s < 0 ? 0 : (s > cap ? cap : s);
      bc(wide ? (float)s / 100000.0f : (float)s / 30000.0f);
    } else if (ch == 47) {
      m = g.c_waits[pid];
    } else if (ch == 48) {
      bc(g.c_waits[pid] != 0 ? 1.0f : 0.0f);
    } else if (ch <= 52) {
      int rank = 0;
      #pragma unroll
      for (int p = 0; p < NPV; p++) rank += g.score[p] > g.score[pid] ? 1 : 0;
      bc(rank == ch - 49 ? 1.0f : 0.0f);
    } else if (ch == 53) {
      bc((float)g.kyoku_idx / 8.0f);
    } else if (ch == 54) {
      bc(((float)g.round_wind * 4.0f + (float)g.kyoku_idx) / 7.0f);
    } else if (ch <= 58) {
      bc((float)(S.dora[rel(ch - 55)] & 0xFF) / 12.0f);
    } else if (ch <= 62) {
      bc((float)g.n_melds[rel(ch - 59)] / 4.0f);
    } else if (ch == 63) {
      // seen/4: patched in by the store loop
    } else if (ch <= 67) {
      m = tail(pid, 4 + (ch - 64));
    } else if (ch <= 69) {
      m = tail(rel(1), 4 + (ch - 68));
    }
    // 70-73: tsumogiri flags are always empty in the live env -> zeros
    if (bcast) m = ALL;
    else if (SANMA) m = obs_compact3(m);
    S.d[ch].mask = m & ALL;
    S.d[ch].val = v;
  }
  __syncwarp();
}
// stream: the descriptors of one row -> 74 x 34 (sanma 74 x 27) floats
template <bool SANMA>
__device__ __forceinline__ void obs_stream_warp(float* dst, ObsScratch& S, int lane) {
  constexpr int W = SANMA ? OBS_W3 : OBS_W;
  if constexpr (!SANMA) {
    // Two channels are 68 floats = 17 sixteen-byte stores: 8 inside the even channel, one straddling both (columns 32,33 |
    // 0,1), 8 inside the odd channel.  The 16 one-channel stores of a pair are item (p, u): u = lane & 15 never changes for
    // a lane, so column and half are loop invariants and a store costs one 16-byte table read, one funnel shift and four
    // bit-field-extract + and pairs.  The 37 straddling stores are a second, short pass.
    auto bit_and = [](uint32_t bits, int q, uint32_t v) {
      int m;
      asm("bfe.s32 %0, %1, %2, 1;" : "=r"(m) : "r"(bits), "r"(q));   // 0 or -1
      return __uint_as_float(v & (uint32_t)m);
    };
    const int u = lane & 15, odd = u >> 3, col = odd ? 4 * u - 30 : 4 * u;
    float4* const out4 = reinterpret_cast<float4*>(dst);
    for (int pr = lane >> 4; pr < OBS_CH / 2; pr += 2) {
      const int ch = 2 * pr + odd;
      const uint4 dsc = *reinterpret_cast<const uint4*>(&S.d[ch]);
      const uint32_t bits = __funnelshift_r(dsc.x, dsc.y, col);
      float4 o = make_float4(bit_and(bits, 0, dsc.z), bit_and(bits, 1, dsc.z), bit_and(bits, 2, dsc.z), bit_and(bits, 3, dsc.z));
      if (ch == 63) o = make_float4((float)S.seen[col] / 4.0f, (float)S.seen[col + 1] / 4.0f, (float)S.seen[col + 2] / 4.0f,
                                    (float)S.seen[col + 3] / 4.0f);
      __stcs(out4 + 17 * pr + u + odd, o);
    }
    for (int pr = lane; pr < OBS_CH / 2; pr += 32) {
      const uint4 a = *reinterpret_cast<const uint4*>(&S.d[2 * pr]), b = *reinterpret_cast<const uint4*>(&S.d[2 * pr + 1]);
      float4 o = make_float4(bit_and(a.y, 0, a.z), bit_and(a.y, 1, a.z), bit_and(b.x, 0, b.z), bit_and(b.x, 1, b.z));   // a: columns 32, 33
      if (2 * pr + 1 == 63) o.z = (float)S.seen[0] / 4.0f, o.w = (float)S.seen[1] / 4.0f;
      __stcs(out4 + 17 * pr + 8, o);
    }
  } else {
    constexpr int VEC = 2;
    for (int j = lane; j < OBS_CH * W / VEC; j += 32) {
      const int e0 = VEC * j, ch0 = e0 / W, col0 = e0 - ch0 * W, rem = W - col0;   // elements q < rem belong to ch0
      const uint64_t bits = (S.d[ch0].mask >> col0) | (S.d[ch0 + 1].mask << rem);
      const float v0 = S.d[ch0].val, v1 = S.d[ch0 + 1].val;
      float o[VEC];
      #pragma unroll
      for (int q = 0; q < VEC; q++) o[q] = ((bits >> q) & 1) ? (q < rem ? v0 : v1) : 0.0f;
      if (ch0 == 63 || (ch0 == 62 && rem < VEC)) {
        #pragma unroll
        for (int q = 0; q < VEC; q++) {
          const int ch = q < rem ? ch0 : ch0 + 1, col = q < rem ? col0 + q : q - rem;
          if (ch == 63) o[q] = (float)S.seen[obs_col_kind3(col)] / 4.0f;
        }
      }
      __stcs(reinterpret_cast<float2*>(dst) + j, make_float2(o[0], o[1]));
    }
  }
  __syncwarp();
}
template <bool SANMA>
__device__ __forceinline__ void obs_encode_warp(const G& g, const uint8_t* river, int pid, float* dst, ObsScratch& S, int lane) {
  obs_describe_warp<SANMA>(g, river, pid, S, lane);
  obs_stream_warp<SANMA>(dst, S, lane);
}

// 82 (sanma 60) mask bytes of one row from the id bitset (3 words) — lanes write 4 bytes each when the row is 4-byte aligned
template <bool SANMA>
__device__ __forceinline__ void obs_mask_row_warp(const uint32_t bits[3], uint8_t* mrow, int lane) {
  constexpr int IDS = SANMA ? OBS_IDS3 : OBS_IDS;
  for (int k = lane; k < IDS; k += 32) mrow[k] = (uint8_t)((bits[k >> 5] >> (k & 31)) & 1);
}
// id bitset of a legal-action list
template <bool SANMA>
__device__ __forceinline__ void obs_id_set(uint32_t bits[3], const rv_action& a) {
  const int id = SANMA ? action_id_3p(a) : action_id(a);
  if (id >= 0 && id < (SANMA ? OBS_IDS3 : OBS_IDS)) bits[id >> 5] |= 1u << (id & 31);
}
}  // namespace rv
#endif
