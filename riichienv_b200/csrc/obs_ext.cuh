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
  uint32_t avail;        // bit i = channel 178 + i
  uint64_t dora_kinds;   // bit k = some indicator points at kind k (obs_ext_dora_kinds)
};
__device__ __forceinline__ uint64_t obs_ext_dora_kinds(const G& g) {
  uint64_t m = 0;
  for (int d = 0; d < g.n_dora; d++) m |= 1ull << obs_next_kind(g.dora_ind[d] >> 2);
  return m;
}

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

// ---- shanten of "hand - one tile + one tile" without the full four-suit combination --------------------------------------
// shanten_counts (hand.cuh) folds the four suits' cost vectors (tiles missing for k mentsu, with / without the pair) with three
// (min,+) convolutions.  Discarding a tile and drawing one changes at most two suits, so per hand the convolutions of the
// UNCHANGED suits are taken once (R1[A]: all suits but A; R2[A,B]: the two suits other than A and B), per discard one
// convolution per other suit (Y), and a (discard, draw) pair costs one table load and the single entry [m mentsu, with pair]
// of the last convolution.  Chiitoitsu / kokushi (shanten.rs:198-239, hands without melds) follow from four counters.
// Vectors are packed like the cost table: byte k = c0[k] | c1[k] << 4 (every entry <= 14).
__device__ __noinline__ uint64_t sh_conv(uint64_t a, uint64_t b) {
  uint64_t out = 0;
  #pragma unroll
  for (int k = 0; k < 5; k++) {
    int n0 = 99, n1 = 99;
    #pragma unroll
    for (int i = 0; i <= k; i++) {
      const int a0 = (int)(a >> (8 * i)) & 15, a1 = (int)(a >> (8 * i + 4)) & 15;
      const int b0 = (int)(b >> (8 * (k - i))) & 15, b1 = (int)(b >> (8 * (k - i) + 4)) & 15;
      n0 = min(n0, a0 + b0);
      n1 = min(n1, min(a1 + b0, a0 + b1));
    }
    out |= (uint64_t)(n0 | (n1 << 4)) << (8 * k);
  }
  return out;
}
// entry [m][with pair] of sh_conv(a, b)
__device__ __forceinline__ int sh_single(uint64_t a, uint64_t b, int m) {
  int best = 99;
  #pragma unroll
  for (int i = 0; i < 5; i++) {
    if (i > m) break;
    const int a0 = (int)(a >> (8 * i)) & 15, a1 = (int)(a >> (8 * i + 4)) & 15;
    const int b0 = (int)(b >> (8 * (m - i))) & 15, b1 = (int)(b >> (8 * (m - i) + 4)) & 15;
    best = min(best, min(a1 + b0, a0 + b1));
  }
  return best;
}
__device__ __forceinline__ uint64_t sh_cost(const Tables& T, int suit, int key) {
  return suit == 3 ? __ldg(&T.honor_cost[clamp_key7((uint32_t)key)]) : __ldg(&T.suit_cost[clamp_key9((uint32_t)key)]);
}
__device__ __forceinline__ int sh_pair_index(int a, int b) {   // unordered pair of distinct suits -> 0..5
  const int lo = a < b ? a : b, hi = a < b ? b : a;
  return lo == 0 ? hi - 1 : lo + hi;                            // 01 02 03 12 13 23 -> 0 1 2 3 4 5
}
struct ShStats {   // what calc_chitoi / calc_kokushi need
  int kinds, pairs, tk, tp2;   // kinds present, kinds with >= 2, terminal/honor kinds present, terminal/honor kinds with >= 2
};
__device__ __forceinline__ bool sh_is_terminal(int k) { return (MASK_TERMINAL_HONOR >> k) & 1; }
__device__ __forceinline__ void sh_remove(ShStats& s, int k, int cnt_before) {
  s.kinds -= cnt_before == 1, s.pairs -= cnt_before == 2;
  if (sh_is_terminal(k)) s.tk -= cnt_before == 1, s.tp2 -= cnt_before == 2;
}
__device__ __forceinline__ void sh_add(ShStats& s, int k, int cnt_before) {
  s.kinds += cnt_before == 0, s.pairs += cnt_before == 1;
  if (sh_is_terminal(k)) s.tk += cnt_before == 0, s.tp2 += cnt_before == 1;
}
// shanten.rs:228-239 from the normal-form replacement number and the counters
__device__ __forceinline__ int sh_finish(int normal_repl, int m, const ShStats& st) {
  int sh = normal_repl - 1;
  if (sh <= 0 || m < 4) return sh;
  sh = min(sh, 7 - st.pairs + (st.kinds < 7 ? 7 - st.kinds : 0) - 1);
  if (sh > 0) sh = min(sh, 14 - st.tk - (st.tp2 > 0 ? 1 : 0) - 1);
  return sh;
}
struct ShHand {    // per hand
  uint64_t cs[4];  // cost vectors of the four suits
  uint64_t R2[6];  // sh_pair_index(A, B) -> convolution of the two suits other than A, B
  uint64_t R1[4];  // A -> convolution of the three suits other than A
  int key[4];
  ShStats st;
};
__device__ inline void sh_hand_init(const Tables& T, const G& g, int pid, ShHand& H) {
  const Cnt c = obs_hand_cnt(g, pid);
  for (int s = 0; s < 4; s++) {
    H.key[s] = (int)g.c_key[pid][s];
    H.cs[s] = sh_cost(T, s, H.key[s]);
  }
  for (int a = 0; a < 4; a++)
    for (int b = a + 1; b < 4; b++) {
      int o[2], no = 0;
      for (int s = 0; s < 4; s++)
        if (s != a && s != b) o[no++] = s;
      H.R2[sh_pair_index(a, b)] = sh_conv(H.cs[o[0]], H.cs[o[1]]);
    }
  for (int a = 0; a < 4; a++) {
    const int b = a == 0 ? 1 : 0;                                   // R1[a] = cs[b] (x) (the two suits other than a, b)
    H.R1[a] = sh_conv(H.cs[b], H.R2[sh_pair_index(a, b)]);
  }
  const uint64_t present = cnt_present(c);
  H.st.kinds = __popcll(present);
  H.st.tk = __popcll(present & MASK_TERMINAL_HONOR);
  H.st.pairs = H.st.tp2 = 0;
  for (int k = 0; k < 34; k++)
    if (cnt_get(c, k) >= 2) H.st.pairs++, H.st.tp2 += sh_is_terminal(k);
}
// shanten of the hand itself (m = len / 3)
__device__ __forceinline__ int sh_hand_shanten(const ShHand& H, int m) { return sh_finish(sh_single(H.R1[0], H.cs[0], m), m, H.st); }
// hand minus one tile of kind d (cd copies before): shanten, the suit's new vector, the counters
__device__ __forceinline__ int sh_discard(const Tables& T, const ShHand& H, int d, int cd, int m, uint64_t& xa, ShStats& st) {
  const int A = d / 9;
  xa = sh_cost(T, A, H.key[A] - pow5(d - 9 * A));
  st = H.st;
  sh_remove(st, d, cd);
  return sh_finish(sh_single(H.R1[A], xa, m), m, st);
}
// (hand minus d) plus one tile of kind k (ck copies after the discard).  d < 0: no discard.  y = sh_conv(xa, R2[A, B]) for
// B != A (precomputed per discard by the caller), ignored otherwise.
__device__ __forceinline__ int sh_draw(const Tables& T, const ShHand& H, int d, int k, int ck, int m, uint64_t y, ShStats st) {
  const int B = k / 9, b = k - 9 * B;
  int repl;
  if (d < 0) {
    repl = sh_single(H.R1[B], sh_cost(T, B, H.key[B] + pow5(b)), m);
  } else {
    const int A = d / 9;
    if (A == B) repl = sh_single(H.R1[A], sh_cost(T, A, H.key[A] - pow5(d - 9 * A) + pow5(b)), m);
    else repl = sh_single(y, sh_cost(T, B, H.key[B] + pow5(b)), m);
  }
  sh_add(st, k, ck);
  return sh_finish(repl, m, st);
}
// the definition above, evaluated through the fast path (one thread): must equal obs_ext_shanten_scalar
__device__ inline void obs_ext_shanten_fast_scalar(const Tables& T, const G& g, int pid, const int* vis, ObsExtInfo& I) {
  ShHand H;
  sh_hand_init(T, g, pid, H);
  const Cnt c = obs_hand_cnt(g, pid);
  const int n = g.hand_len[pid];
  const int cur = sh_hand_shanten(H, n / 3);
  int keep = 0, inc = 0, best_uke = 0, best_eff = 0;
  for (int d = 0; d < 34; d++) {
    const int cd = cnt_get(c, d);
    if (!cd) continue;
    uint64_t xa;
    ShStats st;
    const int ss = sh_discard(T, H, d, cd, (n - 1) / 3, xa, st);
    if (ss == cur) keep += cd;
    else if (ss > cur) inc += cd;
    if (ss > cur) continue;
    int uke = 0, eff = 0;
    for (int k = 0; k < 34; k++) {
      const int sc = cnt_get(c, k) - (k == d ? 1 : 0);
      if (sc >= 4) continue;
      const int A = d / 9, B = k / 9;
      const uint64_t y = A != B ? sh_conv(xa, H.R2[sh_pair_index(A, B)]) : 0;
      if (sh_draw(T, H, d, k, sc, n / 3, y, st) < ss) {
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
      const int sc = cnt_get(c, k);
      if (sc >= 4) continue;
      if (sh_draw(T, H, -1, k, sc, (n + 1) / 3, 0, H.st) < cur) best_eff++;
    }
  }
  I.shanten = cur, I.eff = best_eff, I.uke = best_uke, I.keep = keep, I.inc = inc;
}

// discard decay row of seat q (encode.rs:295-313): row[kind] += exp(-0.2 * age), oldest discard first
__device__ inline void obs_ext_decay_row(const G& g, const uint8_t* river, int q, const DecayTab& D, float* row) {
  const int n = min((int)g.n_river[q], RV_RIVER_CAP);
  #pragma unroll 1
  for (int i = 0; i < n; i++) row[river[q * RV_RIVER_CAP + i] >> 2] += D.w[n - 1 - i];
}

// channels 78..214 as (mask, value): out[col] = (mask >> col) & 1 ? val : 0
// `rec`: the record in HBM, for the cold fields (last_tedashi, riichi_sutehai); `g` may be a staged copy of the hot prefix
__device__ inline void obs_ext_channel(const G& g, const G& rec, int pid, int ch, const ObsExtInfo& I, uint64_t& mask, float& val) {
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
    else bc(((tile & 3) == 0 && ((I.dora_kinds >> (tile >> 2)) & 1)) ? 1.0f : 0.0f);
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
    const int t = ted ? rec.last_tedashi[p] : rec.riichi_sutehai[p];
    if (t != RV_NONE) tile_ctx(x % 3, t);
  }
}

// Observation::encode_kawa_overview (observation/python.rs:881-930) of seat p into a ZEROED (7, 34) block: the n-th copy of a
// kind the seat discarded (n capped at 4) raises channel n-1; channels 4-6 are the reference's "aka" flags, raised for tile
// ids 20 / 24 / 28 at columns 5 / 14 / 23.  `river` = the [4][32] discard bytes.
// Sanma (observation_3p/python.rs:760-808): 3 seats over the 27 compact columns; id 20 has no column, ids 24 / 28 raise channel
// 5 at column 6 and channel 6 at column 15.
constexpr int KAWA_FLOATS = 4 * 7 * OBS_W;
constexpr int KAWA_FLOATS3 = 3 * 7 * OBS_W3;
template <bool SANMA = false>
__device__ inline void obs_kawa_seat(const G& g, const uint8_t* river, int p, float* out) {
  constexpr int W = SANMA ? OBS_W3 : OBS_W;
  const int n = min((int)g.n_river[p], RV_RIVER_CAP);
  uint64_t c0 = 0, c1 = 0;            // 3-bit counters per kind: kinds 0..20 in c0, 21..33 in c1
  #pragma unroll 1
  for (int i = 0; i < n; i++) {
    const int t = river[p * RV_RIVER_CAP + i], k = t >> 2;
    const int col = SANMA ? (k == 0 ? 0 : (k >= 8 ? k - 7 : -1)) : k;
    if (col >= 0) {
      uint64_t& c = k < 21 ? c0 : c1;
      const int sh = 3 * (k < 21 ? k : k - 21), have = (int)((c >> sh) & 7);
      out[(have < 3 ? have : 3) * W + col] = 1.0f;
      if (have < 7) c += 1ull << sh;
    }
    if (SANMA) {
      if (t == 24) out[5 * W + 6] = 1.0f;
      else if (t == 28) out[6 * W + 15] = 1.0f;
    } else {
      if (t == 20) out[4 * W + 5] = 1.0f;
      else if (t == 24) out[5 * W + 14] = 1.0f;
      else if (t == 28) out[6 * W + 23] = 1.0f;
    }
  }
}

#ifdef __CUDACC__
struct ObsExtScratch {
  ObsDesc d[OBSX_CH + 1];          // all channels: 0..73 copied from the base describe, 63 and 74..77 empty (streamed apart)
  float decay[4][36];
  int uke[16], eff[16];          // per compacted discard (obs_ext_shanten_warp)
  uint16_t vk[16];               // compacted discards: kind | (shanten after the discard + 2) << 8; kind 0xFF = no discard
  uint32_t vst[16];              // their chiitoi / kokushi counters (ShStats, a byte each)
  uint64_t Y[16][4];             // their cross-suit vectors: Y[B] = (suit of the discard, after it) (x) R2[A, B]
  ShHand H;
};

// shanten.rs:250-393, one warp, on the incremental evaluation above (sh_*).  Passes:
//   0. lanes 0-5 take the six two-suit convolutions, lanes 0-3 then the four three-suit ones (ShHand in shared memory);
//   1. lane l = the l-th distinct tile kind of the hand: shanten after discarding it (keep / increase counts come from here);
//      the discards that do not raise shanten are compacted into a list together with their three cross-suit vectors Y
//      (+ one pseudo entry "no discard" for a 3n+1 hand, whose effective-tile count is taken on the hand itself,
//      shanten.rs:265-296);
//   2. the (discard, draw kind) pairs of that list, 34 per discard, are dealt out flat over the lanes — one table load and
//      one five-term minimum each; a draw that lowers the shanten adds its unseen copies to the discard's ukeire and 1 to
//      its effective-tile count (shared-memory atomics);
//   3. maxima over the list.
__device__ __forceinline__ void obs_ext_shanten_warp(const Tables& T, const G& g, int pid, const int* seen, int lane, ObsExtScratch& X,
                                                    ObsExtInfo& I) {
  const Cnt c = obs_hand_cnt(g, pid);
  const int n = g.hand_len[pid];
  ShHand& H = X.H;
  // pass 0
  if (lane < 4) {
    H.key[lane] = (int)g.c_key[pid][lane];
    H.cs[lane] = sh_cost(T, lane, H.key[lane]);
  }
  __syncwarp();
  if (lane < 6) {
    const int a = lane < 3 ? 0 : lane < 5 ? 1 : 2, b = lane < 3 ? lane + 1 : lane < 5 ? lane - 1 : 3;   // 01 02 03 12 13 23
    int o0 = -1, o1 = -1;
    #pragma unroll
    for (int s = 0; s < 4; s++)
      if (s != a && s != b) (o0 < 0 ? o0 : o1) = s;
    H.R2[lane] = sh_conv(H.cs[o0], H.cs[o1]);
  }
  __syncwarp();
  if (lane < 4) {
    const int b = lane == 0 ? 1 : 0;
    H.R1[lane] = sh_conv(H.cs[b], H.R2[sh_pair_index(lane, b)]);
  }
  uint64_t present = cnt_present(c);
  ShStats st0;
  st0.kinds = __popcll(present);
  st0.tk = __popcll(present & MASK_TERMINAL_HONOR);
  {
    const int cA = cnt_get(c, lane), cB = lane < 2 ? cnt_get(c, 32 + lane) : 0;
    const uint32_t p2 = __ballot_sync(0xFFFFFFFFu, cA >= 2), p2b = __ballot_sync(0xFFFFFFFFu, cB >= 2) & 3u;
    const uint64_t ge2 = (uint64_t)p2 | ((uint64_t)p2b << 32);
    st0.pairs = __popcll(ge2);
    st0.tp2 = __popcll(ge2 & MASK_TERMINAL_HONOR);
  }
  if (lane == 0) H.st = st0;
  __syncwarp();
  const int cur = sh_finish(sh_single(H.R1[0], H.cs[0], n / 3), n / 3, st0);
  // pass 1
  const int D = st0.kinds;
  int kind = -1;
  for (int i = 0; i < RV_HAND_CAP; i++) {
    if (i == lane && present) kind = __ffsll((long long)present) - 1;
    present &= present - 1;
  }
  int ss = 99, cd = 0;
  uint64_t xa = 0;
  ShStats st1 = st0;
  if (lane < D) {
    cd = cnt_get(c, kind);
    ss = sh_discard(T, H, kind, cd, (n - 1) / 3, xa, st1);
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
  if (valid) {
    const int slot = __popc(vmask & ((1u << lane) - 1)), A = kind / 9;
    X.vk[slot] = (uint16_t)(kind | ((ss + 2) << 8));
    X.vst[slot] = (uint32_t)st1.kinds | ((uint32_t)st1.pairs << 8) | ((uint32_t)st1.tk << 16) | ((uint32_t)st1.tp2 << 24);
    #pragma unroll
    for (int B = 0; B < 4; B++)
      if (B != A) X.Y[slot][B] = sh_conv(xa, H.R2[sh_pair_index(A, B)]);
  }
  const bool self13 = n % 3 == 1;
  if (lane == 0 && self13) {
    X.vk[V] = (uint16_t)(0xFF | ((cur + 2) << 8));
    X.vst[V] = (uint32_t)st0.kinds | ((uint32_t)st0.pairs << 8) | ((uint32_t)st0.tk << 16) | ((uint32_t)st0.tp2 << 24);
  }
  if (lane < 16) X.uke[lane] = 0, X.eff[lane] = 0;
  __syncwarp();
  // pass 2
  const int total = (V + (self13 ? 1 : 0)) * 34;
  for (int j = lane; j < total; j += 32) {
    const int di = j / 34, k = j - di * 34;
    const int e = X.vk[di], dk = e & 0xFF, base_s = (e >> 8) - 2;
    const bool none = dk == 0xFF;
    const int ck = cnt_get(c, k), sc = ck - ((!none && dk == k) ? 1 : 0);
    if (sc < 4) {
      const uint32_t pv = X.vst[di];
      ShStats st;
      st.kinds = pv & 0xFF, st.pairs = (pv >> 8) & 0xFF, st.tk = (pv >> 16) & 0xFF, st.tp2 = pv >> 24;
      const int B = k / 9;
      const uint64_t y = (!none && dk / 9 != B) ? X.Y[di][B] : 0;
      if (sh_draw(T, H, none ? -1 : dk, k, sc, none ? (n + 1) / 3 : n / 3, y, st) < base_s) {
        const int vis = seen[k] - ck;                       // seen = hand + rivers + melds + indicators
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
__device__ __forceinline__ void obs_ext_encode_warp(const Tables& T, const DecayTab& D, const G& g, const G& rec, const uint8_t* river, int pid,
                                                    uint32_t avail, float* dst, ObsScratch& S, ObsExtScratch& X, int lane) {
  int called = 0;                                 // channel 30 of encode_base_into: called melds count one tile short
  if (lane < 16 && (lane & 3) < g.n_melds[lane >> 2] && rec.meld_called[lane >> 2][lane & 3] != RV_NONE) called = 1;   // `rec`: the HBM record (cold field)
  called = __popc(__ballot_sync(0xFFFFFFFFu, called));
  obs_describe_warp<false>(g, river, pid, S, lane, called);
  for (int i = lane; i < 4 * 36; i += 32) (&X.decay[0][0])[i] = 0.0f;
  __syncwarp();
  if (lane < 4) obs_ext_decay_row(g, river, (pid + lane) & 3, D, X.decay[lane]);
  ObsExtInfo I;
  obs_ext_shanten_warp(T, g, pid, S.seen, lane, X, I);
  I.avail = avail;
  I.dora_kinds = obs_ext_dora_kinds(g);
  for (int ch = lane; ch < 78; ch += 32) {          // base descriptors next to the extended ones: one array for the stream
    uint4 dd = make_uint4(0, 0, 0, 0);
    if (ch < OBS_CH && ch != 63) dd = *reinterpret_cast<const uint4*>(&S.d[ch]);
    *reinterpret_cast<uint4*>(&X.d[ch]) = dd;
  }
  for (int ch = 78 + lane; ch < OBSX_CH; ch += 32) {
    uint64_t m;
    float v;
    obs_ext_channel(g, rec, pid, ch, I, m, v);
    X.d[ch].mask = m;
    X.d[ch].val = v;
  }
  __syncwarp();
  float2* const out2 = reinterpret_cast<float2*>(dst);
  // (mask, value) channels: pair j = lane + 32 t covers columns 2p, 2p+1 of channel ch, (ch, p) = divmod(j, 17) kept
  // incrementally (32 = 17 + 15).  Channels 63 and 74-77 have empty descriptors here and are written below.
  {
    int ch = lane >= 17 ? 1 : 0, p = lane >= 17 ? lane - 17 : lane;
    float2* o = out2 + lane;
    #pragma unroll 1
    for (int j = lane; j < OBSX_CH * 17; j += 32) {
      const uint4 dd = *reinterpret_cast<const uint4*>(&X.d[ch]);
      const uint32_t bits = __funnelshift_rc(dd.x, dd.y, 2 * p);            // clamped: p == 16 takes the high word
      __stcs(o, make_float2(__uint_as_float((bits & 1) ? dd.z : 0u), __uint_as_float((bits & 2) ? dd.z : 0u)));
      o += 32;
      p += 15, ch += 1;
      if (p >= 17) p -= 17, ch += 1;
    }
  }
  __syncwarp();                                        // the per-column channels overwrite zeros: order the two passes
  // the five channels with per-column values (63: seen / 4; 74-77: decay): 85 pairs
  for (int j = lane; j < 5 * 17; j += 32) {
    const int r = j / 17, col = 2 * (j - r * 17);
    const float2 o = r == 0 ? make_float2((float)S.seen[col] * 0.25f, (float)S.seen[col + 1] * 0.25f)   // == / 4.0f, exactly
                            : make_float2(X.decay[r - 1][col], X.decay[r - 1][col + 1]);
    __stcs(out2 + (r == 0 ? 63 : 73 + r) * 17 + (j - r * 17), o);
  }
  __syncwarp();
}
#endif

}  // namespace rv
