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
__device__ __forceinline__ int obs_next_kind(int k) {  // observation/helpers.rs:25-50
  if (k < 27) return (k % 9 == 8) ? k - 8 : k + 1;
  if (k < 31) return 27 + (k - 27 + 1) % 4;
  return 31 + (k - 31 + 1) % 3;
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
