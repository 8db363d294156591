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

#ifdef __CUDACC__
// shanten.rs:470-615, one warp: lane l = the l-th distinct kind in hand (shanten after discarding it); the discards that do
// not raise shanten are compacted (+ a "no discard" entry for a 3n+1 hand) and their (discard, draw) pairs — 27 draws each —
// are dealt out flat over the lanes, one full shanten_counts_3p per pair (the incremental evaluation of obs_ext.cuh does not
// carry over: relocating 1m / 9m into honor slots couples the manzu and honor suits).
__device__ __forceinline__ void obs_ext3_shanten_warp(const Tables& T, const G& g, int pid, const int* seen, int lane, ObsExtScratch& X,
                                                     ObsExtInfo& I) {
  const Cnt c = obs_hand_cnt(g, pid);
  const int n = g.hand_len[pid];
  const int cur = shanten_counts_3p(T, c, n / 3);
  uint64_t present = cnt_present(c);
  const int D = __popcll(present);
  int kind = -1;
  for (int i = 0; i < RV_HAND_CAP; i++) {
    if (i == lane && present) kind = __ffsll((long long)present) - 1;
    present &= present - 1;
  }
  int ss = 99, cd = 0;
  if (lane < D) {
    cd = cnt_get(c, kind);
    Cnt sub = c;
    cnt_sub(sub, kind);
    ss = shanten_counts_3p(T, sub, (n - 1) / 3);
  }
  int keep = ss == cur ? cd : 0, inc = (lane < D && ss > cur) ? cd : 0;
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    keep += __shfl_xor_sync(0xFFFFFFFFu, keep, o);
    inc += __shfl_xor_sync(0xFFFFFFFFu, inc, o);
  }
  const bool valid = lane < D && ss <= cur;
  const uint32_t vmask = __ballot_sync(0xFFFFFFFFu, valid);
  const int V = __popc(vmask);
  if (valid) X.vk[__popc(vmask & ((1u << lane) - 1))] = (uint16_t)(kind | ((ss + 2) << 8));
  const bool self13 = n % 3 == 1;
  if (lane == 0 && self13) X.vk[V] = (uint16_t)(0xFF | ((cur + 2) << 8));
  if (lane < 16) X.uke[lane] = 0, X.eff[lane] = 0;
  __syncwarp();
  const int total = (V + (self13 ? 1 : 0)) * OBS_W3;
  for (int j = lane; j < total; j += 32) {
    const int di = j / OBS_W3, k = obs_col_kind3(j - di * OBS_W3);
    const int e = X.vk[di], dk = e & 0xFF, base_s = (e >> 8) - 2;
    Cnt t = c;
    if (dk != 0xFF) cnt_sub(t, dk);
    const int sc = cnt_get(t, k);
    if (sc < 4) {
      cnt_add(t, k);
      if (shanten_counts_3p(T, t, dk != 0xFF ? n / 3 : (n + 1) / 3) < base_s) {
        const int vis = seen[k] - cnt_get(c, k);           // seen = hand + rivers + melds + indicators
        atomicAdd(&X.uke[di], max(max(4 - vis, 0) - sc, 0));
        atomicAdd(&X.eff[di], 1);
      }
    }
  }
  __syncwarp();
  int uke = lane < V ? X.uke[lane] : 0;
  int eff = self13 ? (lane == V ? X.eff[lane] : 0) : (lane < V ? X.eff[lane] : 0);
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    uke = max(uke, __shfl_xor_sync(0xFFFFFFFFu, uke, o));
    eff = max(eff, __shfl_xor_sync(0xFFFFFFFFu, eff, o));
  }
  I.shanten = cur, I.eff = eff, I.uke = uke, I.keep = keep, I.inc = inc;
}

// One warp, one sanma row: base gather + describe (obs.cuh, SANMA), the extended descriptors, 5,805 four-byte streaming stores
// (a 23,220-byte row is only 4-byte aligned).
__device__ __forceinline__ void obs_ext3_encode_warp(const Tables& T, const DecayTab& D, const G& g, const G& rec, const uint8_t* river,
                                                     int pid, uint32_t avail, float* dst, ObsScratch& S, ObsExtScratch& X, int lane) {
  int called = 0;
  if (lane < 12 && (lane & 3) < g.n_melds[lane >> 2] && rec.meld_called[lane >> 2][lane & 3] != RV_NONE) called = 1;
  called = __popc(__ballot_sync(0xFFFFFFFFu, called));
  obs_describe_warp<true>(g, river, pid, S, lane, called);
  for (int i = lane; i < 4 * 36; i += 32) (&X.decay[0][0])[i] = 0.0f;
  __syncwarp();
  if (lane < 3) obs_ext3_decay_row(g, river, (pid + lane) % 3, D, X.decay[lane]);
  ObsExtInfo I;
  obs_ext3_shanten_warp(T, g, pid, S.seen, lane, X, I);
  I.avail = avail;
  I.dora_kinds = obs_ext3_dora_kinds(g);
  for (int ch = lane; ch < 78; ch += 32) {
    uint4 dd = make_uint4(0, 0, 0, 0);
    if (ch < OBS_CH && ch != 63) dd = *reinterpret_cast<const uint4*>(&S.d[ch]);
    *reinterpret_cast<uint4*>(&X.d[ch]) = dd;
  }
  for (int ch = 78 + lane; ch < OBSX_CH; ch += 32) {
    uint64_t m;
    float v;
    obs_ext3_channel(g, rec, pid, ch, I, m, v);
    X.d[ch].mask = m;
    X.d[ch].val = v;
  }
  __syncwarp();
  #pragma unroll 1
  for (int e = lane; e < OBSX3_FLOATS; e += 32) {
    const int ch = e / OBS_W3, col = e - ch * OBS_W3;
    float o;
    if (ch == 63) o = (float)S.seen[obs_col_kind3(col)] * 0.25f;
    else if (ch >= 74 && ch < 78) o = X.decay[ch - 74][col];
    else o = ((X.d[ch].mask >> col) & 1) ? X.d[ch].val : 0.0f;
    __stcs(dst + e, o);
  }
  __syncwarp();
}
#endif

}  // namespace rv
