// Extended observation tensor: Observation::encode_extended (observation/python.rs:1272-1294), 215 x 34 f32.
//
//   ch   0- 73  base encoder          observation/encode.rs:12-291   (== Observation::encode except channel 30, which here
//                                                                      counts a called meld one tile short, encode.rs:94-110)
//   ch  74- 77  discard decay         encode.rs:295-313   sum over a seat's discards of exp(-0.2 * age), relative seats
//   ch  78- 93  shanten efficiency    encode.rs:317-350   self: shanten/8, effective tiles/34, best ukeire/80; others 0.5;
//                                                         + min(discards/18, 1) per relative seat   (shanten.rs:250-393)
//   ch  94- 97  ankan overview        encode.rs:354-368
//   ch  98-177  fuuro overview        encode.rs:372-396   seat(4) x meld(4) x {tile slot 0-3, red five}
//   ch 178-188  action availability   encode.rs:399-430   over the seat's legal list
//   ch 189-193  discard candidates    encode.rs:433-476   hand size/34, share of discards keeping / raising shanten, ...
//   ch 194-196  pass context          encode.rs:479-510   (fed the discarder's SEAT, see obs_ext_channel)
//   ch 197-205  last tedashi          encode.rs:513-547   opponents in absolute seat order
//   ch 206-214  riichi sutehai        encode.rs:550-584
//
// One warp writes one row (29,240 B).  Channels 78+ are (34-bit mask, value) descriptors like the base encoder's, lane =
// channel; the decay rows are accumulated in the reference's order (oldest discard first) by one lane per seat from a table of
// exp(-0.2 * age) computed on the host with the C library's expf (the reference calls the same libm through f32::exp), so the
// row is bit-identical to the oracle's.  The shanten channels cost up to 14 x 34 + 15 shanten evaluations
// (shanten.rs:331-393 removes each hand tile and tries each of the 34 kinds): the (discard, draw) pairs are dealt out flat
// over the lanes (obs_ext_shanten_warp).
#pragma once
#include "obs.cuh"

namespace rv {

constexpr int OBSX_CH = 215;
struct DecayTab {
  float w[RV_RIVER_CAP];   // w[age] = expf(-0.2f * age)
};
struct ObsExtInfo {
  int shanten, eff, uke, keep, inc;
  uint32_t avail;   // bit i = channel 178 + i
};

// encode.rs:399-430: which availability channel a legal action raises (0 = none)
__device__ inline uint32_t obs_avail_bit(const rv_action& a) {
  switch (a.type) {
    case RV_RIICHI: return 1u << 0;
    case RV_CHI: {
      if (a.n_consume != 2) return 0;
      const int t0 = a.consume[0] >> 2, t1 = a.consume[1] >> 2, diff = t1 > t0 ? t1 - t0 : t0 - t1;
      if (diff == 1) return t0 < t1 ? 1u << 1 : 1u << 3;
      return diff == 2 ? 1u << 2 : 0;
    }
    case RV_PON: return 1u << 4;
    case RV_DAIMINKAN: return 1u << 5;
    case RV_ANKAN: return 1u << 6;
    case RV_KAKAN: return 1u << 7;
    case RV_TSUMO:
    case RV_RON: return 1u << 8;
    case RV_KYUSHU_KYUHAI: return 1u << 9;
    case RV_PASS: return 1u << 10;
    default: return 0;
  }
}

// hand histogram of seat pid as a Cnt (shanten.rs:250-261 counts tile / 4 of the concealed tiles)
__device__ __forceinline__ Cnt obs_hand_cnt(const G& g, int pid) {
  Cnt c;
  c.s[0] = g.c_cnt[pid][0], c.s[1] = g.c_cnt[pid][1], c.s[2] = g.c_cnt[pid][2], c.s[3] = g.c_cnt[pid][3];
  return c;
}

// shanten.rs:250-393 for one seat, one thread.  vis[k] = discards + meld tiles + dora indicators of kind k.
// (the warp version below computes the same numbers; this one is the definition the host compile of the sources tests)
__device__ inline void obs_ext_shanten_scalar(const Tables& T, const G& g, int pid, const int* vis, ObsExtInfo& I) {
  const Cnt c = obs_hand_cnt(g, pid);
  const int n = g.hand_len[pid];
  const int cur = shanten_counts(T, c, n / 3);
  int keep = 0, inc = 0, best_uke = 0, best_eff = 0;
  for (int d = 0; d < 34; d++) {
    const int cd = cnt_get(c, d);
    if (!cd) continue;
    Cnt sub = c;
    cnt_sub(sub, d);
    const int ss = shanten_counts(T, sub, (n - 1) / 3);
    if (ss == cur) keep += cd;
    else if (ss > cur) inc += cd;
    if (ss > cur) continue;
    int uke = 0, eff = 0;
    for (int k = 0; k < 34; k++) {
      const int sc = cnt_get(sub, k);
      if (sc >= 4) continue;
      Cnt t = sub;
      cnt_add(t, k);
      if (shanten_counts(T, t, n / 3) < ss) {
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
  if (n % 3 == 1) {   // shanten.rs:265-296 on the hand itself
    best_eff = 0;
    for (int k = 0; k < 34; k++) {
      if (cnt_get(c, k) >= 4) continue;
      Cnt t = c;
      cnt_add(t, k);
      if (shanten_counts(T, t, (n + 1) / 3) < cur) best_eff++;
    }
  }
  I.shanten = cur, I.eff = best_eff, I.uke = best_uke, I.keep = keep, I.inc = inc;
}

// discard decay row of seat q (encode.rs:295-313): row[kind] += exp(-0.2 * age), oldest discard first
__device__ inline void obs_ext_decay_row(const G& g, const uint8_t* river, int q, const DecayTab& D, float* row) {
  const int n = min((int)g.n_river[q], RV_RIVER_CAP);
  for (int i = 0; i < n; i++) row[river[q * RV_RIVER_CAP + i] >> 2] += D.w[n - 1 - i];
}

// channels 78..214 as (mask, value): out[col] = (mask >> col) & 1 ? val : 0
__device__ inline void obs_ext_channel(const G& g, int pid, int ch, const ObsExtInfo& I, uint64_t& mask, float& val) {
  constexpr uint64_t ALL = (1ull << OBS_W) - 1;
  mask = 0;
  val = 1.0f;
  auto bc = [&](float x) { mask = ALL, val = x; };
  auto rel = [&](int i) { return (pid + i) & 3; };
  // tile context triple (encode.rs:479-584): kind/33, is red five, "is dora" — the reference compares the TILE ID with
  // get_next_tile(indicator), which is the id of copy 0 of the dora kind (helpers.rs:25-50), so only copy 0 matches
  auto tile_ctx = [&](int f, int tile) {
    if (f == 0) bc((float)(tile >> 2) / 33.0f);
    else if (f == 1) bc((tile == 16 || tile == 52 || tile == 88) ? 1.0f : 0.0f);
    else {
      bool dora = false;
      for (int d = 0; d < g.n_dora; d++) dora |= obs_next_kind(g.dora_ind[d] >> 2) * 4 == tile;
      bc(dora ? 1.0f : 0.0f);
    }
  };
  if (ch < 94) {                                   // shanten efficiency
    const int r = (ch - 78) >> 2, f = (ch - 78) & 3;
    if (f == 3) {
      const float v = (float)g.n_river[rel(r)] / 18.0f;
      bc(v < 1.0f ? v : 1.0f);
    } else if (r != 0) {
      bc(0.5f);
    } else if (f == 0) {
      const float s = (float)I.shanten;
      bc((s > 0.0f ? s : 0.0f) / 8.0f);
    } else if (f == 1) {
      bc((float)I.eff / 34.0f);
    } else {
      bc((float)I.uke / 80.0f);
    }
  } else if (ch < 98) {                            // ankan overview
    const int q = rel(ch - 94);
    for (int m = 0; m < g.n_melds[q]; m++)
      if (g.meld_type[q][m] == RV_MELD_ANKAN && g.meld_tiles[q][m][0] != RV_NONE) mask |= 1ull << (g.meld_tiles[q][m][0] >> 2);
  } else if (ch < 178) {                           // fuuro overview
    const int x = ch - 98, q = rel(x / 20), m = (x % 20) / 5, slot = x % 5;
    if (m < g.n_melds[q]) {
      if (slot < 4) {
        const int t = g.meld_tiles[q][m][slot];
        if (t != RV_NONE) mask = 1ull << (t >> 2);
      } else {
        for (int k = 0; k < 4; k++) {
          const int t = g.meld_tiles[q][m][k];
          if (t == 16 || t == 52 || t == 88) mask |= 1ull << (t >> 2);
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
  } else if (ch < 197) {                           // pass context
    // GameState::get_observation hands the encoder `last_discard.map(|(tile, _pid)| tile)` of a tuple stored as
    // (pid, tile) (state/mod.rs:252 vs 1329): the "tile" these channels see is the DISCARDER'S SEAT.  Kept.
    if (g.last_discard_pid != RV_NONE) tile_ctx(ch - 194, g.last_discard_pid);
  } else {                                         // last tedashi / riichi sutehai, opponents in absolute seat order
    const bool ted = ch < 206;
    const int x = ch - (ted ? 197 : 206), o = x / 3, p = o < pid ? o : o + 1;
    const int t = ted ? g.last_tedashi[p] : g.riichi_sutehai[p];
    if (t != RV_NONE) tile_ctx(x % 3, t);
  }
}

#ifdef __CUDACC__
struct ObsExtScratch {
  ObsDesc d[OBSX_CH - 78 + 1];     // channels 78..214
  float decay[4][36];
  int uke[16], eff[16];          // per compacted discard (obs_ext_shanten_warp)
  uint16_t vk[16];               // compacted discards: kind | (shanten after the discard + 2) << 8; kind 0xFF = no discard
};

// shanten.rs:250-393, one warp.  Three passes, each one shanten evaluation per lane:
//   1. lane l = the l-th distinct tile kind of the hand: shanten after discarding it (keep / increase counts come from here);
//      the discards that do not raise shanten are compacted into a list (+ one pseudo entry "no discard" for a 3n+1 hand,
//      whose effective-tile count is taken on the hand itself, shanten.rs:265-296);
//   2. the (discard, draw kind) pairs of that list, 34 per discard, are dealt out flat over the lanes: a draw that lowers the
//      shanten adds its unseen copies to the discard's ukeire and 1 to its effective-tile count (shared-memory atomics);
//   3. maxima over the list.
// A hand has at most 14 distinct kinds, so pass 2 is at most 15 rounds (the per-discard loop it replaces took 3 per discard).
__device__ __forceinline__ void obs_ext_shanten_warp(const Tables& T, const G& g, int pid, const int* seen, int lane, ObsExtScratch& X,
                                                    ObsExtInfo& I) {
  const Cnt c = obs_hand_cnt(g, pid);
  const int n = g.hand_len[pid];
  const int cur = shanten_counts(T, c, n / 3);
  // pass 1
  uint64_t present = cnt_present(c);
  const int D = __popcll(present);
  int kind = -1;
  for (int i = 0; i < 14; i++) {
    if (i == lane && present) kind = __ffsll((long long)present) - 1;
    present &= present - 1;
  }
  int ss = 99, cd = 0;
  if (lane < D) {
    cd = cnt_get(c, kind);
    Cnt sub = c;
    cnt_sub(sub, kind);
    ss = shanten_counts(T, sub, (n - 1) / 3);
  }
  int keep = ss == cur ? cd : 0, inc = (lane < D && ss > cur) ? cd : 0;
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    keep += __shfl_xor_sync(0xFFFFFFFFu, keep, o);
    inc += __shfl_xor_sync(0xFFFFFFFFu, inc, o);
  }
  const bool valid = lane < D && ss <= cur;
  const uint32_t vmask = __ballot_sync(0xFFFFFFFFu, valid);
  int V = __popc(vmask);
  if (valid) X.vk[__popc(vmask & ((1u << lane) - 1))] = (uint16_t)(kind | ((ss + 2) << 8));
  const bool self13 = n % 3 == 1;
  if (lane == 0 && self13) X.vk[V] = (uint16_t)(0xFF | ((cur + 2) << 8));
  if (lane < 16) X.uke[lane] = 0, X.eff[lane] = 0;
  __syncwarp();
  // pass 2
  const int total = (V + (self13 ? 1 : 0)) * 34;
  for (int j = lane; j < total; j += 32) {
    const int di = j / 34, k = j - di * 34;
    const int e = X.vk[di], dk = e & 0xFF, base_s = (e >> 8) - 2;
    Cnt t = c;
    if (dk != 0xFF) cnt_sub(t, dk);
    const int sc = cnt_get(t, k);
    if (sc < 4) {
      cnt_add(t, k);
      if (shanten_counts(T, t, dk != 0xFF ? n / 3 : (n + 1) / 3) < base_s) {
        const int vis = seen[k] - cnt_get(c, k);           // seen = hand + rivers + melds + indicators
        atomicAdd(&X.uke[di], max(max(4 - vis, 0) - sc, 0));
        atomicAdd(&X.eff[di], 1);
      }
    }
  }
  __syncwarp();
  // pass 3
  int uke = lane < V ? X.uke[lane] : 0;
  int eff = self13 ? (lane == V ? X.eff[lane] : 0) : (lane < V ? X.eff[lane] : 0);
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    uke = max(uke, __shfl_xor_sync(0xFFFFFFFFu, uke, o));
    eff = max(eff, __shfl_xor_sync(0xFFFFFFFFu, eff, o));
  }
  I.shanten = cur, I.eff = eff, I.uke = uke, I.keep = keep, I.inc = inc;
}

// One warp, one row: base gather + describe (obs.cuh), extended describe, then 3,655 eight-byte streaming stores
// (a 29,240-byte row is 8-byte aligned; 34 columns = 17 pairs, so a pair never straddles two channels).
__device__ __forceinline__ void obs_ext_encode_warp(const Tables& T, const DecayTab& D, const G& g, const uint8_t* river, int pid,
                                                    uint32_t avail, float* dst, ObsScratch& S, ObsExtScratch& X, int lane) {
  int called = 0;                                 // channel 30 of encode_base_into: called melds count one tile short
  if (lane < 16 && (lane & 3) < g.n_melds[lane >> 2] && g.meld_called[lane >> 2][lane & 3] != RV_NONE) called = 1;
  called = __popc(__ballot_sync(0xFFFFFFFFu, called));
  obs_describe_warp<false>(g, river, pid, S, lane, called);
  for (int i = lane; i < 4 * 36; i += 32) (&X.decay[0][0])[i] = 0.0f;
  __syncwarp();
  if (lane < 4) obs_ext_decay_row(g, river, (pid + lane) & 3, D, X.decay[lane]);
  ObsExtInfo I;
  obs_ext_shanten_warp(T, g, pid, S.seen, lane, X, I);
  I.avail = avail;
  for (int ch = 78 + lane; ch < OBSX_CH; ch += 32) {
    uint64_t m;
    float v;
    obs_ext_channel(g, pid, ch, I, m, v);
    X.d[ch - 78].mask = m;
    X.d[ch - 78].val = v;
  }
  __syncwarp();
  float2* const out2 = reinterpret_cast<float2*>(dst);
  for (int j = lane; j < OBSX_CH * 17; j += 32) {
    const int ch = j / 17, col = 2 * (j - ch * 17);
    float2 o;
    if (ch == 63) {
      o = make_float2((float)S.seen[col] / 4.0f, (float)S.seen[col + 1] / 4.0f);
    } else if (ch >= 74 && ch < 78) {
      o = make_float2(X.decay[ch - 74][col], X.decay[ch - 74][col + 1]);
    } else {
      const ObsDesc& dd = ch < OBS_CH ? S.d[ch] : X.d[ch - 78];
      const uint32_t bits = (uint32_t)(dd.mask >> col);
      o = make_float2((bits & 1) ? dd.val : 0.0f, (bits & 2) ? dd.val : 0.0f);
    }
    __stcs(out2 + j, o);
  }
  __syncwarp();
}
#endif

}  // namespace rv
