// Sanma rows of the extended observation tensor: Observation3P::encode_extended (observation_3p/python.rs:1117-1140; channel
// blocks observation_3p/encode.rs:22-620), 215 x 27 f32.  Same block offsets as the 4P layout (obs_ext.cuh); differences:
//   * 27 compact columns (1m, 9m, 1-9p, 1-9s, honors — observation_3p/helpers.rs:8-15);
//   * three relative seats per group (the fourth channel of the decay / shanten / ankan groups and the fourth 20-channel fuuro
//     group stay zero), two opponents in the last-tedashi / riichi-sutehai blocks (absolute order);
//   * shanten through calculate_shanten_3p (1m / 9m koutsu only), draws over the 27 kinds that exist, effective tiles / 27
//     (shanten.rs:470-615);
//   * tile context: compact column / 26, the sanma dora successor (1m <-> 9m wrap);
//   * channel 30 counts 108 tiles.
// The quirks of the 4P encoder are the same here: a called meld counts one tile short in channel 30 (encode.rs:101-114 vs
// python.rs), "is dora" compares a tile ID with the successor's id, the pass context is fed the discarder's seat
// (state_3p/mod.rs:220).
#pragma once
#include "obs_ext.cuh"

namespace rv {

constexpr int OBSX3_FLOATS = OBSX_CH * OBS_W3;   // 5,805 floats per row (a row is only 4-byte aligned)
__device__ __forceinline__ bool obs_sanma_kind(int k) { return k == 0 || k >= 8; }
__device__ __forceinline__ int obs_compact_col3(int k) { return k == 0 ? 0 : (k >= 8 ? k - 7 : -1); }   // tile34_to_compact

__device__ __forceinline__ uint64_t obs_ext3_dora_kinds(const G& g) {
  uint64_t m = 0;
  for (int d = 0; d < g.n_dora; d++) m |= 1ull << obs_next_kind_sanma(g.dora_ind[d] >> 2);
  return m;
}

// shanten.rs:470-615 for one seat, one thread (the definition; vis[k] = discards + meld tiles + dora indicators of kind k)
__device__ inline void obs_ext3_shanten_scalar(const Tables& T, const G& g, int pid, const int* vis, ObsExtInfo& I) {
  const Cnt c = obs_hand_cnt(g, pid);
  const int n = g.hand_len[pid];
  const int cur = shanten_counts_3p(T, c, n / 3);
  int keep = 0, inc = 0, best_uke = 0, best_eff = 0;
  for (int d = 0; d < 34; d++) {
    const int cd = cnt_get(c, d);
    if (!cd) continue;
    Cnt sub = c;
    cnt_sub(sub, d);
    const int ss = shanten_counts_3p(T, sub, (n - 1) / 3);
    if (ss == cur) keep += cd;
    else if (ss > cur) inc += cd;
    if (ss > cur) continue;
    int uke = 0, eff = 0;
    for (int k = 0; k < 34; k++) {
      if (!obs_sanma_kind(k)) continue;
      const int sc = cnt_get(sub, k);
      if (sc >= 4) continue;
      Cnt t = sub;
      cnt_add(t, k);
      if (shanten_counts_3p(T, t, n / 3) < ss) {
        int rem = 4 - vis[k];
        rem = rem < 0 ? 0 : rem;
        rem -= sc;
        uke += rem < 0 ? 0 : rem;
        eff++;
      }
    }
    best_uke = max(best_uke, uke);
    best_eff = max(best_eff, eff);
  }
  if (n % 3 == 1) {
    best_eff = 0;
    for (int k = 0; k < 34; k++) {
      if (!obs_sanma_kind(k) || cnt_get(c, k) >= 4) continue;
      Cnt t = c;
      cnt_add(t, k);
      if (shanten_counts_3p(T, t, (n + 1) / 3) < cur) best_eff++;
    }
  }
  I.shanten = cur, I.eff = best_eff, I.uke = best_uke, I.keep = keep, I.inc = inc;
}

// discard decay row (27 columns) of seat q (observation_3p/encode.rs:315-333), oldest discard first
__device__ inline void obs_ext3_decay_row(const G& g, const uint8_t* river, int q, const DecayTab& D, float* row) {
  const int n = min((int)g.n_river[q], RV_RIVER_CAP);
  #pragma unroll 1
  for (int i = 0; i < n; i++) {
    const int col = obs_compact_col3(river[q * RV_RIVER_CAP + i] >> 2);
    if (col >= 0) row[col] += D.w[n - 1 - i];
  }
}

// channels 78..214 as (27-bit column mask, value)
__device__ inline void obs_ext3_channel(const G& g, const G& rec, int pid, int ch, const ObsExtInfo& I, uint64_t& mask, float& val) {
  constexpr uint64_t ALL = (1ull << OBS_W3) - 1;
  uint64_t m34 = 0;        // mask over the 34 tile kinds; compacted at the end
  bool bcast = false;
  val = 1.0f;
  auto bc = [&](float x) { bcast = true, val = x; };
  auto rel = [&](int i) { return (pid + i) % 3; };
  auto tile_ctx = [&](int f, int tile) {
    if (f == 0) {
      const int col = obs_compact_col3(tile >> 2);
      if (col >= 0) bc((float)col / 26.0f);
    } else if (f == 1) {
      bc((tile == 16 || tile == 52 || tile == 88) ? 1.0f : 0.0f);
    } else {
      bc(((tile & 3) == 0 && ((I.dora_kinds >> (tile >> 2)) & 1)) ? 1.0f : 0.0f);
    }
  };
  if (ch < 94) {                                   // shanten efficiency: three seats
    const int r = (ch - 78) >> 2, f = (ch - 78) & 3;
    if (r < 3) {
      if (f == 3) {
        const float v = (float)g.n_river[rel(r)] / 18.0f;
        bc(v < 1.0f ? v : 1.0f);
      } else if (r != 0) {
        bc(0.5f);
      } else if (f == 0) {
        const float s = (float)I.shanten;
        bc((s > 0.0f ? s : 0.0f) / 8.0f);
      } else if (f == 1) {
        bc((float)I.eff / 27.0f);
      } else {
        bc((float)I.uke / 80.0f);
      }
    }
  } else if (ch < 98) {                            // ankan overview
    if (ch - 94 < 3) {
      const int q = rel(ch - 94);
      for (int m = 0; m < g.n_melds[q]; m++)
        if (g.meld_type[q][m] == RV_MELD_ANKAN && g.meld_tiles[q][m][0] != RV_NONE) m34 |= 1ull << (g.meld_tiles[q][m][0] >> 2);
    }
  } else if (ch < 178) {                           // fuuro overview
    const int x = ch - 98, r = x / 20, m = (x % 20) / 5, slot = x % 5;
    if (r < 3) {
      const int q = rel(r);
      if (m < g.n_melds[q]) {
        if (slot < 4) {
          const int t = g.meld_tiles[q][m][slot];
          if (t != RV_NONE) m34 = 1ull << (t >> 2);
        } else {
          for (int k = 0; k < 4; k++) {
            const int t = g.meld_tiles[q][m][k];
            if (t == 16 || t == 52 || t == 88) m34 |= 1ull << (t >> 2);
          }
        }
      }
    }
  } else if (ch < 189) {                           // action availability
    if ((I.avail >> (ch - 178)) & 1) bc(1.0f);
  } else if (ch < 194) {                           // discard candidates
    const int n = g.hand_len[pid];
    if (ch == 189) bc((float)n / 34.0f);
    else if (ch == 190) { if (n) bc((float)I.keep / (float)n); }
    else if (ch == 191) { if (n) bc((float)I.inc / (float)n); }
    else if (ch == 192) bc(I.shanten == -1 ? 1.0f : 0.0f);
    else bc((g.flags[pid] & RV_F_RIICHI_DECLARED) ? 1.0f : 0.0f);
  } else if (ch < 197) {                           // pass context: the "tile" is the discarder's seat (state_3p/mod.rs:220)
    if (g.last_discard_pid != RV_NONE) tile_ctx(ch - 194, g.last_discard_pid);
  } else {                                         // last tedashi / riichi sutehai: the two opponents in absolute seat order
    const bool ted = ch < 206;
    const int x = ch - (ted ? 197 : 206), o = x / 3;
    if (o < 2) {
      const int p = o < pid ? o : o + 1;
      const int t = ted ? rec.last_tedashi[p] : rec.riichi_sutehai[p];
      if (t != RV_NONE) tile_ctx(x % 3, t);
    }
  }
  mask = bcast ? ALL : (((m34 & 1) | ((m34 >> 7) & ~1ull)) & ALL);   // 34 kinds -> 27 columns
}

}  // namespace rv
