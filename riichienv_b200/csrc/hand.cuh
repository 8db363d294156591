// Device-side hand evaluation: agari test, wait set, shanten, and full
// yaku / han / fu / score calculation.
//
// Replaces (reference paths relative to riichienv-core/src):
//   agari.rs:63-245            -> table lookups (agari14 / waits13)
//   hand_evaluator.rs:24-213   -> hand_calc(), waits13(), tenpai checks
//   yaku.rs:232-1280           -> eval_division() over an iterative enumeration of
//                                 (head, body) divisions in the reference's DFS order
//   score.rs:13-99             -> calc_score()
//   shanten.rs:186-239         -> shanten_counts() from the cost tables
// A hand is held as four 64-bit words of 4-bit tile counts (m, p, s, z), so every
// predicate is register arithmetic; the only memory traffic is four table loads.
#pragma once
#include "../../include/riichienv_b200.h"
#include "tables.cuh"

#ifdef RV_HOSTSIM_STATS
extern unsigned long long g_rv_stats[48];
#define RV_STAT(i) (g_rv_stats[i]++)
#else
#define RV_STAT(i)
#endif

namespace rv {

// ------------------------------------------------------------------ counts
struct Cnt {
  uint64_t s[4];
};
__device__ __forceinline__ void cnt_zero(Cnt& c) { c.s[0] = c.s[1] = c.s[2] = c.s[3] = 0; }
__device__ __forceinline__ void cnt_add(Cnt& c, int t34, int n = 1) {
  int k = t34 / 9;
  uint64_t inc = (uint64_t)n << (4 * (t34 - 9 * k));
  c.s[0] += (k == 0) ? inc : 0;
  c.s[1] += (k == 1) ? inc : 0;
  c.s[2] += (k == 2) ? inc : 0;
  c.s[3] += (k == 3) ? inc : 0;
}
__device__ __forceinline__ void cnt_sub(Cnt& c, int t34, int n = 1) {
  int k = t34 / 9;
  uint64_t inc = (uint64_t)n << (4 * (t34 - 9 * k));
  c.s[0] -= (k == 0) ? inc : 0;
  c.s[1] -= (k == 1) ? inc : 0;
  c.s[2] -= (k == 2) ? inc : 0;
  c.s[3] -= (k == 3) ? inc : 0;
}
__device__ __forceinline__ uint64_t cnt_suit(const Cnt& c, int k) {
  return k == 0 ? c.s[0] : k == 1 ? c.s[1] : k == 2 ? c.s[2] : c.s[3];
}
__device__ __forceinline__ int cnt_get(const Cnt& c, int t34) {
  int k = t34 / 9;
  return (int)((cnt_suit(c, k) >> (4 * (t34 - 9 * k))) & 15);
}
__device__ __forceinline__ int cnt_total(const Cnt& c) {
  // sum of nibbles
  int tot = 0;
  #pragma unroll
  for (int k = 0; k < 4; k++) {
    uint64_t x = c.s[k];
    x = (x & 0x0F0F0F0F0F0F0F0Full) + ((x >> 4) & 0x0F0F0F0F0F0F0F0Full);
    tot += (int)((x * 0x0101010101010101ull) >> 56);
  }
  return tot;
}
// 34-bit mask of tile kinds present
__device__ __forceinline__ uint64_t cnt_present(const Cnt& c) {
  uint64_t m = 0;
  #pragma unroll
  for (int k = 0; k < 4; k++) {
    uint64_t x = c.s[k];
    x |= x >> 2;
    x |= x >> 1;
    x &= 0x1111111111111111ull;  // bit 4i set if nibble i nonzero
    // compress every 4th bit
    uint32_t r = 0;
    #pragma unroll
    for (int i = 0; i < 9; i++) r |= (uint32_t)((x >> (4 * i)) & 1) << i;
    m |= (uint64_t)r << (9 * k);
  }
  return m;
}
// Base-5 table index of a suit histogram.  Physical hands hold at most four copies of a kind; a record injected through
// the API may not (the reference's own tests park thirteen copies of one tile in seats they do not care about), so the
// index is clamped: such a seat gets meaningless answers, never an out-of-bounds table read.
// 34-bit mask of tile kinds held at least twice
__device__ __forceinline__ uint64_t cnt_ge2(const Cnt& c) {
  uint64_t m = 0;
  #pragma unroll
  for (int k = 0; k < 4; k++) {
    uint64_t x = c.s[k];
    x = ((x >> 1) | (x >> 2)) & 0x1111111111111111ull;   // bit 4i set if nibble i >= 2
    uint32_t r = 0;
    #pragma unroll
    for (int i = 0; i < 9; i++) r |= (uint32_t)((x >> (4 * i)) & 1) << i;
    m |= (uint64_t)r << (9 * k);
  }
  return m;
}
// lowest tile kind >= i with a nonzero count, 34 if there is none (one find-first-set per suit word instead of a scan)
__device__ __forceinline__ int cnt_next(const Cnt& c, int i) {
  if (i >= 34) return 34;
  const int su = i / 9;
  #pragma unroll
  for (int k = 0; k < 4; k++) {
    if (k < su) continue;
    uint64_t x = c.s[k];
    if (k == su) x &= ~0ull << (4 * (i - 9 * su));
    if (x) return 9 * k + ((__ffsll((long long)x) - 1) >> 2);
  }
  return 34;
}
template <int N>
__device__ __forceinline__ int suit_key(uint64_t x) {
  int k = 0;
  #pragma unroll
  for (int i = N - 1; i >= 0; i--) k = k * 5 + (int)((x >> (4 * i)) & 15);
  constexpr int LAST = (N == 9 ? 1953125 : 78125) - 1;
  return min(max(k, 0), LAST);
}
__device__ __forceinline__ uint32_t clamp_key9(uint32_t k) { return min(k, 1953124u); }   // cached keys (c_key), same reason
__device__ __forceinline__ uint32_t clamp_key7(uint32_t k) { return min(k, 78124u); }

constexpr uint64_t MASK_TERMINAL_HONOR =
    (1ull << 0) | (1ull << 8) | (1ull << 9) | (1ull << 17) | (1ull << 18) | (1ull << 26) | (0x7Full << 27);
constexpr uint64_t MASK_HONOR = 0x7Full << 27;
constexpr uint64_t MASK_NUM_TERMINAL = MASK_TERMINAL_HONOR & ~MASK_HONOR;
constexpr uint64_t MASK_GREEN = (1ull << 19) | (1ull << 20) | (1ull << 21) | (1ull << 23) | (1ull << 25) | (1ull << 32);

struct SuitInfo {
  uint32_t e[4];
};
__device__ __forceinline__ void load_info(const Tables& T, const Cnt& c, SuitInfo& si) {
  si.e[0] = __ldg(&T.suit_info[suit_key<9>(c.s[0])]);
  si.e[1] = __ldg(&T.suit_info[suit_key<9>(c.s[1])]);
  si.e[2] = __ldg(&T.suit_info[suit_key<9>(c.s[2])]);
  si.e[3] = __ldg(&T.honor_info[suit_key<7>(c.s[3])]);
}
__device__ __forceinline__ uint32_t load_info_suit(const Tables& T, uint64_t x, int k) {
  return k == 3 ? __ldg(&T.honor_info[suit_key<7>(x)]) : __ldg(&T.suit_info[suit_key<9>(x)]);
}

// agari.rs:166-177 — seven distinct pairs
__device__ __forceinline__ bool chiitoi14(const Cnt& c) {
  uint64_t bad = 0;
  int pairs = 0;
  #pragma unroll
  for (int k = 0; k < 4; k++) {
    bad |= c.s[k] & ~0x2222222222222222ull;
    pairs += __popcll(c.s[k]);
  }
  return bad == 0 && pairs == 7;
}
// agari.rs:141-164
__device__ __forceinline__ bool kokushi14(const Cnt& c) {
  // all 13 terminal kinds present once, exactly one twice
  const uint64_t TM = 0xF0000000Full;  // nibbles 0 and 8
  uint64_t x0 = c.s[0] & TM, x1 = c.s[1] & TM, x2 = c.s[2] & TM, x3 = c.s[3] & 0xFFFFFFFull;
  uint64_t hi = 0;
  int n1 = 0, n2 = 0;
  uint64_t xs[4] = {x0, x1, x2, x3};
  #pragma unroll
  for (int k = 0; k < 4; k++) {
    uint64_t o = xs[k] & 0x1111111111111111ull, t = xs[k] & 0x2222222222222222ull;
    hi |= xs[k] & 0xCCCCCCCCCCCCCCCCull;
    uint64_t three = (o << 1) & t;
    hi |= three;
    n1 += __popcll(o);
    n2 += __popcll(t);
  }
  return hi == 0 && n1 == 12 && n2 == 1;
}

// standard-form agari from the four suit entries
__device__ __forceinline__ bool standard_agari(const SuitInfo& si) {
  int nM = (si.e[0] & 1) + (si.e[1] & 1) + (si.e[2] & 1) + (si.e[3] & 1);
  int nP = ((si.e[0] >> 1) & 1) + ((si.e[1] >> 1) & 1) + ((si.e[2] >> 1) & 1) + ((si.e[3] >> 1) & 1);
  return nM == 3 && nP == 1;
}
// agari::is_agari on a 3n+2 concealed hand (agari.rs:63-71)
__device__ __forceinline__ bool agari14(const Tables& T, const Cnt& c) {
  SuitInfo si;
  load_info(T, c, si);
  if (standard_agari(si)) return true;
  return chiitoi14(c) || kokushi14(c);
}

// get_waits_u8 (hand_evaluator.rs:196-213) of a 3n+1 concealed hand: 34-bit mask.
// `si` holds the suit entries of `c`.
// the standard-form part of waits13 (4 mentsu + pair), from the four suit entries alone
__device__ __forceinline__ uint64_t waits13_std(const SuitInfo& si) {
  int nM = (si.e[0] & 1) + (si.e[1] & 1) + (si.e[2] & 1) + (si.e[3] & 1);
  int nP = ((si.e[0] >> 1) & 1) + ((si.e[1] >> 1) & 1) + ((si.e[2] >> 1) & 1) + ((si.e[3] >> 1) & 1);
  uint64_t w = 0;
  #pragma unroll
  for (int k = 0; k < 4; k++) {
    int m = si.e[k] & 1, p = (si.e[k] >> 1) & 1;
    uint64_t wm = (si.e[k] >> 2) & 0x1FF, wp = (si.e[k] >> 11) & 0x1FF;
    if (nM - m == 3 && nP - p == 0) w |= wp << (9 * k);   // others all M: this suit supplies the pair
    if (nM - m == 2 && nP - p == 1) w |= wm << (9 * k);   // one other suit holds the pair
  }
  return w;
}
__device__ __forceinline__ uint64_t waits13_inl(const Cnt& c, const SuitInfo& si) {
  int nM = (si.e[0] & 1) + (si.e[1] & 1) + (si.e[2] & 1) + (si.e[3] & 1);
  int nP = ((si.e[0] >> 1) & 1) + ((si.e[1] >> 1) & 1) + ((si.e[2] >> 1) & 1) + ((si.e[3] >> 1) & 1);
  uint64_t w = 0;
  #pragma unroll
  for (int k = 0; k < 4; k++) {
    int m = si.e[k] & 1, p = (si.e[k] >> 1) & 1;
    uint64_t wm = (si.e[k] >> 2) & 0x1FF, wp = (si.e[k] >> 11) & 0x1FF;
    if (nM - m == 3 && nP - p == 0) w |= wp << (9 * k);   // others all M: this suit supplies the pair
    if (nM - m == 2 && nP - p == 1) w |= wm << (9 * k);   // one other suit holds the pair
  }
  // chiitoitsu: six pairs + one single, nothing else
  {
    uint64_t bad = 0;
    int n1 = 0, n2 = 0;
    #pragma unroll
    for (int k = 0; k < 4; k++) {
      uint64_t o = c.s[k] & 0x1111111111111111ull, t = c.s[k] & 0x2222222222222222ull;
      bad |= (c.s[k] & 0xCCCCCCCCCCCCCCCCull) | ((o << 1) & t);
      n1 += __popcll(o);
      n2 += __popcll(t);
    }
    if (bad == 0 && n1 == 1 && n2 == 6) {
      #pragma unroll
      for (int k = 0; k < 4; k++) {
        uint64_t o = c.s[k] & 0x1111111111111111ull;
        if (o) w |= 1ull << (9 * k + (__ffsll((long long)o) - 1) / 4);
      }
    }
  }
  // kokushi: 13 tiles, all terminal/honor kinds
  {
    const uint64_t TM = 0xF0000000Full;
    uint64_t other = (c.s[0] & ~TM) | (c.s[1] & ~TM) | (c.s[2] & ~TM);
    if (other == 0) {
      uint64_t xs[4] = {c.s[0] & TM, c.s[1] & TM, c.s[2] & TM, c.s[3] & 0xFFFFFFFull};
      uint64_t hi = 0;
      int n1 = 0, n2 = 0;
      #pragma unroll
      for (int k = 0; k < 4; k++) {
        uint64_t o = xs[k] & 0x1111111111111111ull, t = xs[k] & 0x2222222222222222ull;
        hi |= (xs[k] & 0xCCCCCCCCCCCCCCCCull) | ((o << 1) & t);
        n1 += __popcll(o);
        n2 += __popcll(t);
      }
      if (hi == 0) {
        uint64_t present = cnt_present(c);
        if (n1 == 13 && n2 == 0) w |= MASK_TERMINAL_HONOR;            // 13-sided
        if (n1 == 11 && n2 == 1) w |= MASK_TERMINAL_HONOR & ~present;  // the missing kind
      }
    }
  }
  return w;
}
__device__ __noinline__ uint64_t waits13(const Cnt& c, const SuitInfo& si) { return waits13_inl(c, si); }
__device__ __forceinline__ uint64_t waits13(const Tables& T, const Cnt& c) {
  SuitInfo si;
  load_info(T, c, si);
  return waits13(c, si);
}

// ------------------------------------------------------------------ shanten
// normal-form replacement number minus one (shanten.rs:186-196): three (min,+) convolutions of the four suits' cost vectors
__device__ __noinline__ int shanten_normal(const Tables& T, const Cnt& c, int m) {
  uint64_t cs[4] = {__ldg(&T.suit_cost[suit_key<9>(c.s[0])]), __ldg(&T.suit_cost[suit_key<9>(c.s[1])]),
                    __ldg(&T.suit_cost[suit_key<9>(c.s[2])]), __ldg(&T.honor_cost[suit_key<7>(c.s[3])])};
  int b0[5], b1[5];
  #pragma unroll
  for (int k = 0; k < 5; k++) {
    b0[k] = (int)((cs[0] >> (8 * k)) & 15);
    b1[k] = (int)((cs[0] >> (8 * k + 4)) & 15);
  }
  #pragma unroll
  for (int s = 1; s < 4; s++) {
    int n0[5], n1[5];
    #pragma unroll
    for (int k = 0; k < 5; k++) n0[k] = n1[k] = 99;
    #pragma unroll
    for (int a = 0; a < 5; a++)
      #pragma unroll
      for (int b = 0; b < 5; b++) {
        if (a + b >= 5) continue;
        int c0 = (int)((cs[s] >> (8 * b)) & 15), c1 = (int)((cs[s] >> (8 * b + 4)) & 15);
        n0[a + b] = min(n0[a + b], b0[a] + c0);
        n1[a + b] = min(n1[a + b], min(b1[a] + c0, b0[a] + c1));
      }
    #pragma unroll
    for (int k = 0; k < 5; k++) {
      b0[k] = n0[k];
      b1[k] = n1[k];
    }
  }
  int sh = 99;
  #pragma unroll
  for (int k = 0; k < 5; k++)
    if (k == m) sh = b1[k] - 1;
  return sh;
}
// chiitoitsu / kokushi closing of shanten.rs:198-239; `skip` = tile kinds chiitoitsu ignores (3P: 2m-8m, shanten.rs:437-453)
__device__ __forceinline__ int shanten_close(int sh, int m, const Cnt& c, uint64_t skip) {
  if (sh <= 0 || m < 4) return sh;
  const uint64_t present = cnt_present(c);
  int kinds = __popcll(present & ~skip), pairs = 0;
  int tk = __popcll(present & MASK_TERMINAL_HONOR);
  bool tpair = false;
  #pragma unroll
  for (int k = 0; k < 4; k++) {
    uint64_t x = c.s[k];
    uint64_t ge2 = ((x >> 1) | (x >> 2)) & 0x1111111111111111ull;   // bit 4i set if nibble i >= 2
    if (k == 0) {
      uint64_t sk = 0;                                                // nibble mask of the skipped manzu kinds
      #pragma unroll
      for (int i = 0; i < 9; i++) sk |= ((skip >> i) & 1) << (4 * i);
      pairs += __popcll(ge2 & ~sk);
    } else {
      pairs += __popcll(ge2);
    }
    uint64_t tm = (k == 3) ? 0x1111111ull : 0x100000001ull;
    if (ge2 & tm) tpair = true;
  }
  int chi = 7 - pairs + (kinds < 7 ? 7 - kinds : 0) - 1;
  sh = min(sh, chi);
  if (sh > 0) sh = min(sh, 14 - tk - (tpair ? 1 : 0) - 1);
  return sh;
}
__device__ __noinline__ int shanten_counts(const Tables& T, const Cnt& c, int len_div3) {
  const int m = len_div3 > 4 ? 4 : len_div3;
  return shanten_close(shanten_normal(T, c, m), m, c, 0);
}
// calc_shanten_from_counts_3p (shanten.rs:407-468): 1m / 9m are relocated into EMPTY honor slots before the normal-form lookup
// (they cannot form sequences; the honor tables are position independent), staying in manzu when no slot is left — the
// reference's overflow fallback; chiitoitsu skips 2m-8m; kokushi as in 4P.
__device__ __noinline__ int shanten_counts_3p(const Tables& T, const Cnt& c, int len_div3) {
  const int m = len_div3 > 4 ? 4 : len_div3;
  Cnt t = c;
  const int mc[2] = {(int)(c.s[0] & 15), (int)((c.s[0] >> 32) & 15)};
  t.s[0] &= ~(0xFull | (0xFull << 32));
  int slot = 0;
  #pragma unroll
  for (int i = 0; i < 2; i++) {
    if (mc[i] == 0) continue;
    while (slot < 7 && ((t.s[3] >> (4 * slot)) & 15) != 0) slot++;
    if (slot < 7) {
      t.s[3] |= (uint64_t)mc[i] << (4 * slot);
      slot++;
    } else {
      t.s[0] |= (uint64_t)mc[i] << (i == 0 ? 0 : 32);
    }
  }
  return shanten_close(shanten_normal(T, t, m), m, c, 0xFEull);
}

// ------------------------------------------------------------------ score.rs
struct ScoreRes {
  uint32_t pay_ron, pay_oya, pay_ko;
};
__host__ __device__ inline uint32_t ceil100(uint32_t v) { return (v + 99) / 100 * 100; }
__host__ __device__ inline ScoreRes calc_score(int han, int fu, bool is_oya, bool is_tsumo, uint32_t honba, uint32_t np,
                                               uint32_t* total = nullptr) {
  uint32_t base;
  if (han >= 5) {
    base = han == 5 ? 2000 : han <= 7 ? 3000 : han <= 10 ? 4000 : han <= 12 ? 6000 : 8000u * ((uint32_t)han / 13u);
  } else {
    uint32_t f = (fu == 25) ? 25u : (uint32_t)((fu + 9) / 10 * 10);
    base = f << (2 + han);
    if (base > 2000) base = 2000;
  }
  ScoreRes r{0, 0, 0};
  uint32_t tot;
  if (is_tsumo) {
    r.pay_oya = is_oya ? 0 : ceil100(base * 2);
    r.pay_ko = is_oya ? ceil100(base * 2) : ceil100(base);
    tot = is_oya ? r.pay_ko * (np - 1) : r.pay_oya + r.pay_ko * (np - 2);
    r.pay_oya += honba * 100;
    r.pay_ko += honba * 100;
    tot += honba * 100 * (np - 1);
  } else {
    r.pay_ron = is_oya ? ceil100(base * 6) : ceil100(base * 4);
    r.pay_ron += honba * 100 * (np - 1);
    tot = r.pay_ron;
  }
  if (total) *total = tot;
  return r;
}

// ------------------------------------------------------------------ yaku
struct MeldView {           // melds in 34-space, as HandEvaluator::new normalises them
  uint8_t n;
  uint8_t type[4];          // rv_meld_type
  uint8_t t0[4];            // tiles[0] after normalisation (chi: lowest tile kind)
};
struct WinCtx {
  bool tsumo, riichi, double_riichi, ippatsu, haitei, houtei, rinshan, chankan, first_turn, menzen;
  uint8_t dora, aka, ura, nuki;   // nuki: nukidora (kita) count, 3P only (yaku_3p.rs:706-709)
  uint8_t round_wind, seat_wind;  // 27..30
};
struct YakuRes {
  int han, fu, yakuman;
  uint64_t mask;
};
struct WinRes {
  bool is_win, yakuman, has_shape;
  int han, fu;
  uint64_t yaku_mask;
  uint32_t ron, oya, ko;
};

__device__ __forceinline__ bool t_terminal(int t) { return t >= 27 || t % 9 == 0 || t % 9 == 8; }
__device__ __forceinline__ bool t_numterm(int t) { return t < 27 && (t % 9 == 0 || t % 9 == 8); }

__device__ __forceinline__ void static_yaku(YakuRes& r, const WinCtx& x) {  // yaku.rs:843-890
  if (x.riichi && !x.double_riichi) { r.han += 1; r.mask |= 1ull << 2; }
  if (x.double_riichi) { r.han += 2; r.mask |= 1ull << 18; }
  if (x.ippatsu) { r.han += 1; r.mask |= 1ull << 30; }
  if (x.menzen && x.tsumo) { r.han += 1; r.mask |= 1ull << 1; }
  if (x.haitei && x.tsumo) { r.han += 1; r.mask |= 1ull << 5; }
  if (x.houtei && !x.tsumo) { r.han += 1; r.mask |= 1ull << 6; }
  if (x.rinshan && x.tsumo) { r.han += 1; r.mask |= 1ull << 4; }
  if (x.chankan && !x.tsumo) { r.han += 1; r.mask |= 1ull << 3; }
  if (x.dora > 0) { r.han += x.dora; r.mask |= 1ull << 31; }
  if (x.aka > 0) { r.han += x.aka; r.mask |= 1ull << 32; }
  if (x.ura > 0) { r.han += x.ura; r.mask |= 1ull << 33; }
  if (x.nuki > 0) { r.han += x.nuki; r.mask |= 1ull << 34; }
}

struct Division {
  int8_t head;     // tile kind, or -1 for the chiitoitsu pseudo-division
  int8_t nb;
  uint8_t bt[4];   // body tile (start tile for shuntsu)
  uint8_t bk[4];   // 1 = koutsu, 0 = shuntsu
};
constexpr int YAKU_CAND_CAP = 12;   // (division, winning group) candidates kept before a flush; more are flushed in batches

// yaku.rs:892-1055.  `all` = kinds present in hand_14 or any meld; wg = -1 for head.
__device__ __noinline__ void yakuman_eval(YakuRes& r, const Cnt& hand, uint64_t all, const MeldView& mv, const WinCtx& x,
                                    const Division& d, int wg, int win) {
  int yc = 0;
  if ((all & ~MASK_HONOR) == 0) { yc += 1; r.mask |= 1ull << 39; }
  if ((all & ~MASK_NUM_TERMINAL) == 0) { yc += 1; r.mask |= 1ull << 41; }
  if ((all & ~MASK_GREEN) == 0) { yc += 1; r.mask |= 1ull << 40; }
  int kans = 0, ankans = 0;
  uint64_t meld_kou = 0;   // non-chi melds by tiles[0]
  for (int m = 0; m < mv.n; m++) {
    if (mv.type[m] >= RV_MELD_DAIMINKAN) kans++;
    if (mv.type[m] == RV_MELD_ANKAN) ankans++;
    if (mv.type[m] != RV_MELD_CHI) meld_kou |= 1ull << mv.t0[m];
  }
  if (kans == 4) { yc += 1; r.mask |= 1ull << 44; }
  if (x.menzen && d.nb + mv.n == 4) {
    // is_chuuren_poutou (yaku.rs:1101-1132) on the concealed hand
    uint64_t pres = cnt_present(hand);
    int suit = -1;
    bool one = false;
    if ((pres & MASK_HONOR) == 0) {
      int ns = ((pres & 0x1FF) != 0) + (((pres >> 9) & 0x1FF) != 0) + (((pres >> 18) & 0x1FF) != 0);
      if (ns == 1) {
        suit = (pres & 0x1FF) ? 0 : ((pres >> 9) & 0x1FF) ? 1 : 2;
        one = true;
      }
    }
    if (one) {
      uint64_t s = cnt_suit(hand, suit);
      bool ok = (s & 15) >= 3 && ((s >> 32) & 15) >= 3 && ((pres >> (9 * suit)) & 0x1FF) == 0x1FF;
      if (ok) {
        bool nine = false;
        if (win < 27) {
          int v = win % 9, cw = cnt_get(hand, win);
          nine = (v == 0 || v == 8) ? cw == 4 : cw == 2;
        }
        if (nine) { yc += 2; r.mask |= 1ull << 47; }
        else { yc += 1; r.mask |= 1ull << 45; }
      }
    }
  }
  if (x.first_turn && x.menzen && x.tsumo) {
    yc += 1;
    r.mask |= 1ull << (x.seat_wind == 27 ? 35 : 36);
  }
  int closed = ankans;
  uint64_t body_kou = 0;
  for (int i = 0; i < d.nb; i++)
    if (d.bk[i]) {
      body_kou |= 1ull << d.bt[i];
      if (!x.tsumo && i == wg) continue;
      closed++;
    }
  if (closed == 4) {
    if (wg < 0) { yc += 2; r.mask |= 1ull << 48; }
    else { yc += 1; r.mask |= 1ull << 38; }
  }
  uint64_t kou = body_kou | meld_kou;   // chi melds never contain honors, so `tiles.contains` == non-chi t0 here
  if (((kou >> 31) & 7) == 7) { yc += 1; r.mask |= 1ull << 37; }
  int wk = __popcll((kou >> 27) & 15);
  int wp = (d.head >= 27 && d.head <= 30 && !((kou >> d.head) & 1)) ? 1 : 0;
  if (wk == 4) { yc += 2; r.mask |= 1ull << 50; }
  else if (wk == 3 && wp == 1) { yc += 1; r.mask |= 1ull << 43; }
  if (yc > 0) {
    r.han = 13 * yc;
    r.yakuman = yc;
  }
}

// One (division, winning group) candidate: yaku.rs:319-555
__device__ __noinline__ YakuRes eval_division(const Cnt& hand, uint64_t all, const MeldView& mv, const WinCtx& x,
                                        const Division& d, int wg, int win) {
  YakuRes r{0, 0, 0, 0};
  yakuman_eval(r, hand, all, mv, x, d, wg, win);
  if (r.han >= 13) return r;
  static_yaku(r, x);
  bool tanyao = (all & MASK_TERMINAL_HONOR) == 0;
  if (tanyao) { r.han += 1; r.mask |= 1ull << 12; }
  // gather sets
  uint64_t body_kou = 0, meld_kou = 0;
  uint32_t seq_starts = 0;     // shuntsu starts (body + chi melds), 27 bits
  int n_body_kou = 0, n_meld_nonchi = 0, ankans = 0, kans = 0;
  bool any_body_kou = false;
  for (int i = 0; i < d.nb; i++) {
    if (d.bk[i]) { body_kou |= 1ull << d.bt[i]; n_body_kou++; any_body_kou = true; }
    else seq_starts |= 1u << d.bt[i];
  }
  for (int m = 0; m < mv.n; m++) {
    if (mv.type[m] == RV_MELD_CHI) seq_starts |= 1u << mv.t0[m];
    else { meld_kou |= 1ull << mv.t0[m]; n_meld_nonchi++; }
    if (mv.type[m] == RV_MELD_ANKAN) ankans++;
    if (mv.type[m] >= RV_MELD_DAIMINKAN) kans++;
  }
  uint64_t kou = body_kou | meld_kou;
  bool head_yakuhai = d.head >= 31 || d.head == x.round_wind || d.head == x.seat_wind;
  // pinfu (yaku.rs:644-686)
  bool pinfu = false;
  if (x.menzen && mv.n == 0 && !any_body_kou && !head_yakuhai && wg >= 0 && !d.bk[wg]) {
    int t = d.bt[wg];
    if (win == t) pinfu = (t % 9 != 6);
    else if (win == t + 2) pinfu = (t % 9 != 0);
  }
  if (pinfu) {
    r.han += 1;
    r.mask |= 1ull << 14;
    r.fu = x.tsumo ? 20 : 30;
  } else {
    // yaku.rs:561-642
    int fu = 20;
    if (x.tsumo) fu += 2;
    else if (x.menzen) fu += 10;
    if (d.head == x.round_wind) fu += 2;
    if (d.head == x.seat_wind) fu += 2;
    if (d.head >= 31) fu += 2;
    if (wg < 0) fu += 2;
    else if (!d.bk[wg]) {
      int t = d.bt[wg];
      if (win == t + 1 || (win == t + 2 && t % 9 == 0) || (win == t && t % 9 == 6)) fu += 2;
    }
    for (int i = 0; i < d.nb; i++)
      if (d.bk[i]) {
        int f = (!x.tsumo && i == wg) ? 2 : 4;
        if (t_terminal(d.bt[i])) f *= 2;
        fu += f;
      }
    for (int m = 0; m < mv.n; m++)
      if (mv.type[m] != RV_MELD_CHI) {
        int f = mv.type[m] == RV_MELD_ANKAN ? 4 : 2;
        if (t_terminal(mv.t0[m])) f *= 2;
        if (mv.type[m] >= RV_MELD_DAIMINKAN) f *= 4;
        fu += f;
      }
    if (fu == 20 && !x.tsumo) fu = 30;
    r.fu = (fu + 9) / 10 * 10;
  }
  // yakuhai (yaku.rs:353-386): body koutsu + non-chi melds of that kind (0/1 each)
  {
    int yt[5] = {31, 32, 33, x.round_wind, x.seat_wind};
    int yid[5] = {7, 8, 9, 11, 10};
    #pragma unroll
    for (int i = 0; i < 5; i++) {
      int cnt = (int)((body_kou >> yt[i]) & 1) + (int)((meld_kou >> yt[i]) & 1);
      if (cnt > 0) {
        r.han += cnt;
        int id = yt[i] == 31 ? 7 : yt[i] == 32 ? 8 : yt[i] == 33 ? 9 : yid[i];
        r.mask |= 1ull << id;
      }
    }
  }
  // shousangen
  {
    int dk = __popcll((kou >> 31) & 7);
    if (dk == 2 && d.head >= 31 && d.head <= 33) { r.han += 2; r.mask |= 1ull << 23; }
  }
  if (n_body_kou + n_meld_nonchi == 4) { r.han += 2; r.mask |= 1ull << 21; }
  {
    int closed = ankans;
    for (int i = 0; i < d.nb; i++)
      if (d.bk[i] && !(!x.tsumo && i == wg)) closed++;
    if (closed == 3) { r.han += 2; r.mask |= 1ull << 22; }
  }
  if (kans == 3) { r.han += 2; r.mask |= 1ull << 20; }
  if (x.menzen) {
    // identical shuntsu pairs among body shuntsu (body order is ascending, equal starts adjacent)
    int pairs = 0, i = 0;
    uint8_t st[4];
    int ns = 0;
    for (int k = 0; k < d.nb; k++)
      if (!d.bk[k]) st[ns++] = d.bt[k];
    // body shuntsu appear in ascending start order in the DFS, no sort needed
    while (i + 1 < ns) {
      if (st[i] == st[i + 1]) { pairs++; i += 2; }
      else i += 1;
    }
    if (pairs == 2) { r.han += 3; r.mask |= 1ull << 28; }
    else if (pairs == 1) { r.han += 1; r.mask |= 1ull << 13; }
  }
  {
    bool ittsu = false, sanshoku = false, doukou = false;
    #pragma unroll
    for (int off = 0; off < 27; off += 9)
      if (((seq_starts >> off) & 0x49) == 0x49) ittsu = true;
    #pragma unroll
    for (int i = 0; i < 7; i++)
      if (((seq_starts >> i) & 0x40201) == 0x40201) sanshoku = true;
    #pragma unroll
    for (int i = 0; i < 9; i++)
      if (((kou >> i) & 0x40201) == 0x40201) doukou = true;
    if (ittsu) { r.han += x.menzen ? 2 : 1; r.mask |= 1ull << 16; }
    if (sanshoku) { r.han += x.menzen ? 2 : 1; r.mask |= 1ull << 17; }
    if (doukou) { r.han += 2; r.mask |= 1ull << 19; }
  }
  {
    int ns = ((all & 0x1FF) != 0) + (((all >> 9) & 0x1FF) != 0) + (((all >> 18) & 0x1FF) != 0);
    bool honor = (all & MASK_HONOR) != 0;
    if (ns == 1 && !honor) { r.han += x.menzen ? 6 : 5; r.mask |= 1ull << 29; }
    else if (ns == 1 && honor) { r.han += x.menzen ? 3 : 2; r.mask |= 1ull << 27; }
  }
  {
    bool honroutou = (all & ~MASK_TERMINAL_HONOR) == 0;
    if (honroutou) { r.han += 2; r.mask |= 1ull << 24; }
    else {
      // junchan / chanta (yaku.rs:707-761)
      bool jun = d.head >= 0 && t_numterm(d.head), chan = d.head >= 0 && t_terminal(d.head);
      bool has_honor = d.head >= 27;
      for (int i = 0; i < d.nb; i++) {
        int t = d.bt[i];
        if (d.bk[i]) {
          if (!t_numterm(t)) jun = false;
          if (!t_terminal(t)) chan = false;
          if (t >= 27) has_honor = true;
        } else {
          bool edge = (t % 9 == 0) || (t % 9 == 6);
          if (!edge) { jun = false; chan = false; }
        }
      }
      for (int m = 0; m < mv.n; m++) {
        int t = mv.t0[m];
        if (mv.type[m] == RV_MELD_CHI) {
          bool edge = (t % 9 == 0) || (t % 9 == 6);
          if (!edge) { jun = false; chan = false; }
        } else {
          if (!t_numterm(t)) jun = false;
          if (!t_terminal(t)) chan = false;
          if (t >= 27) has_honor = true;
        }
      }
      if (jun) { r.han += x.menzen ? 3 : 2; r.mask |= 1ull << 26; }
      else if (chan && has_honor) { r.han += x.menzen ? 2 : 1; r.mask |= 1ull << 15; }
    }
  }
  return r;
}

// yaku::calculate_yaku (yaku.rs:232-559) on the concealed 3n+2 hand.
// A soft block-wide rendezvous for kernels whose threads each run a long, data-dependent routine (hand_yaku_kernel): every
// thread ARRIVES once per generation — as early as it knows that it will not take the expensive path — and the threads that do
// take it WAIT (bounded spin, so a thread that never arrives costs time, never a hang) until the whole block has arrived.
// The point: 28 warps of an SM at 28 different places of ~80 KB of yaku code miss the instruction cache on four lines out of
// ten; after the rendezvous they run the candidate evaluation — 90 % of the instructions — at the same time.
struct BlockPace {
  unsigned* cnt;        // shared memory, monotonic
  unsigned target;      // arrivals that complete the current generation
  bool arrived;
};
__device__ __forceinline__ void pace_arrive(BlockPace* p) {
#ifdef __CUDA_ARCH__
  if (p && !p->arrived) {
    atomicAdd(p->cnt, 1u);
    p->arrived = true;
  }
#else
  (void)p;
#endif
}
__device__ __forceinline__ void pace_wait(BlockPace* p) {
#ifdef __CUDA_ARCH__
  if (!p) return;
  pace_arrive(p);
  for (int spins = 0; spins < 4096; spins++)
    if ((int)(*reinterpret_cast<volatile unsigned*>(p->cnt) - p->target) >= 0) break;
#else
  (void)p;
#endif
}
__device__ __noinline__ YakuRes calculate_yaku(const Tables& T, const Cnt& hand, uint64_t all, const MeldView& mv,
                                         const WinCtx& x, int win, bool std_shape, BlockPace* pace = nullptr) {
  YakuRes best{0, 0, 0, 0};
  if (!std_shape) {
    pace_arrive(pace);                       // no candidate evaluation on this path
    if (kokushi14(hand)) {
      if (cnt_get(hand, win) == 2) { best.han = 26; best.yakuman = 2; best.mask = 1ull << 49; }
      else { best.han = 13; best.yakuman = 1; best.mask = 1ull << 42; }
      return best;
    }
    if (chiitoi14(hand)) {
      best.han = 2;
      best.fu = 25;
      best.mask = 1ull << 25;
      if ((all & MASK_TERMINAL_HONOR) == 0) { best.han += 1; best.mask |= 1ull << 12; }
      int ns = ((all & 0x1FF) != 0) + (((all >> 9) & 0x1FF) != 0) + (((all >> 18) & 0x1FF) != 0);
      bool honor = (all & MASK_HONOR) != 0;
      if (ns == 1 && !honor) { best.han += 6; best.mask |= 1ull << 29; }
      else if (ns == 1 && honor) { best.han += 3; best.mask |= 1ull << 27; }
      if ((all & ~MASK_TERMINAL_HONOR) == 0) { best.han += 2; best.mask |= 1ull << 24; }
      Division d0;
      d0.head = 0;   // the reference passes Division{head:0, body:[]} (yaku.rs:281-290)
      d0.nb = 0;
      YakuRes y{best.han, best.fu, 0, best.mask};
      yakuman_eval(y, hand, all, mv, x, d0, -1, win);
      static_yaku(y, x);
      return y;
    }
    return best;
  }
  // enumerate divisions in the reference's order: head ascending, then DFS over the
  // remaining tiles (lowest tile first; koutsu before shuntsu) — agari.rs:73-139
  Division cand[YAKU_CAND_CAP];
  int8_t cand_wg[YAKU_CAND_CAP];
  int ncand = 0;
  auto flush = [&]() {            // candidates in discovery order: the first maximum wins, as in the reference's loop
    for (int k = 0; k < ncand; k++) {
      YakuRes r = eval_division(hand, all, mv, x, cand[k], cand_wg[k], win);
      if (r.han >= 13 && r.yakuman > 0) { if (r.han > best.han) best = r; }
      else if (r.han > best.han || (r.han == best.han && r.fu > best.fu)) best = r;
    }
    ncand = 0;
  };
  Cnt work = hand;
  // heads in ascending order, but each lane walks only ITS candidate heads (kinds held at least twice): in a loop over all
  // 34 kinds a warp spent every iteration with the few lanes that happen to hold that pair (ncu: 3 of 32 lanes active)
  uint64_t heads = cnt_ge2(hand);
  {
    // first the quick reject for every candidate (cheap, all lanes busy): without the pair every suit must be mentsu-only;
    // only the suit of the pair changes, the other three entries are those of the full hand
    SuitInfo full;
    load_info(T, hand, full);
    uint64_t keep = 0, cand_heads = heads;
    while (cand_heads) {
      const int head = __ffsll((long long)cand_heads) - 1;
      cand_heads &= cand_heads - 1;
      const int su = head / 9;
      cnt_sub(work, head, 2);
      const uint32_t e = load_info_suit(T, cnt_suit(work, su), su);
      cnt_add(work, head, 2);
      const uint32_t e0 = su == 0 ? e : full.e[0], e1 = su == 1 ? e : full.e[1], e2 = su == 2 ? e : full.e[2], e3 = su == 3 ? e : full.e[3];
      if ((e0 & e1 & e2 & e3 & 1) != 0) keep |= 1ull << head;
    }
    heads = keep;
  }
  while (heads) {
    const int head = __ffsll((long long)heads) - 1;
    heads &= heads - 1;
    cnt_sub(work, head, 2);
    {
      Division d;
      d.head = (int8_t)head;
      d.nb = 0;
      // explicit DFS stack: stage[l] = 0 try koutsu, 1 try shuntsu, 2 exhausted
      int pos[5], stage[5];
      int depth = 0;
      pos[0] = 0;
      stage[0] = 0;
      while (depth >= 0) {
        // advance to first nonzero tile
        int i = pos[depth];
        if (stage[depth] == 0) {
          i = cnt_next(work, i);            // lowest kind >= i still in the hand (34: none)
          pos[depth] = i;
        }
        if (i >= 34) {
          // complete division: every winning group is a candidate.  Candidates are only RECORDED here and evaluated in one
          // loop after the search (flush): eval_division is ~90 % of the instructions of a winning hand, and called from
          // inside this data-dependent search the lanes of a warp reached it at different iterations — ncu counted 2 of 32
          // lanes active per instruction.  In the flush loop the lanes run it together.
          if (d.head == win) {
            if (ncand == YAKU_CAND_CAP) flush();
            cand[ncand] = d;
            cand_wg[ncand++] = -1;
          }
          for (int g = 0; g < d.nb; g++) {
            bool hit = d.bk[g] ? (d.bt[g] == win) : (win >= d.bt[g] && win <= d.bt[g] + 2);
            if (!hit) continue;
            if (ncand == YAKU_CAND_CAP) flush();
            cand[ncand] = d;
            cand_wg[ncand++] = (int8_t)g;
          }
          depth--;
          if (depth >= 0) {
            // undo the mentsu taken at this level
            int g = d.nb - 1;
            if (d.bk[g]) cnt_add(work, d.bt[g], 3);
            else { cnt_add(work, d.bt[g]); cnt_add(work, d.bt[g] + 1); cnt_add(work, d.bt[g] + 2); }
            d.nb--;
          }
          continue;
        }
        bool advanced = false;
        if (stage[depth] == 0) {
          stage[depth] = 1;
          if (cnt_get(work, i) >= 3) {
            cnt_sub(work, i, 3);
            d.bt[d.nb] = (uint8_t)i;
            d.bk[d.nb] = 1;
            d.nb++;
            advanced = true;
          }
        }
        if (!advanced && stage[depth] == 1) {
          stage[depth] = 2;
          if (i < 27 && (i % 9) <= 6 && cnt_get(work, i + 1) > 0 && cnt_get(work, i + 2) > 0) {
            cnt_sub(work, i);
            cnt_sub(work, i + 1);
            cnt_sub(work, i + 2);
            d.bt[d.nb] = (uint8_t)i;
            d.bk[d.nb] = 0;
            d.nb++;
            advanced = true;
          }
        }
        if (advanced) {
          depth++;
          pos[depth] = i;
          stage[depth] = 0;
        } else {
          depth--;
          if (depth >= 0) {
            int g = d.nb - 1;
            if (d.bk[g]) cnt_add(work, d.bt[g], 3);
            else { cnt_add(work, d.bt[g]); cnt_add(work, d.bt[g] + 1); cnt_add(work, d.bt[g] + 2); }
            d.nb--;
          }
        }
      }
    }
    cnt_add(work, head, 2);
  }
  pace_wait(pace);                           // the block evaluates its candidates together
  flush();
  return best;
}

__device__ __forceinline__ int next_dora_tile(int t) {  // hand_evaluator.rs:286-300
  if (t < 27) return (t % 9 == 8) ? t - 8 : t + 1;
  if (t < 31) return t == 30 ? 27 : t + 1;
  return t == 33 ? 31 : t + 1;
}
__device__ __forceinline__ int next_dora_tile_sanma(int t) {  // hand_evaluator_3p.rs:300-311
  if (t == 0) return 8;
  if (t == 8) return 0;
  if (t >= 1 && t <= 7) return t;
  return next_dora_tile(t);
}
__device__ __forceinline__ bool tid_is_aka(int tid) { return tid == 16 || tid == 52 || tid == 88; }

// HandEvaluator::new + calc (hand_evaluator.rs:24-176).
//  tiles/n: concealed tids; if n + 3*n_melds == 13 the win tile is added (as the
//  reference does), otherwise it is assumed to be among `tiles`.
//  meld_tiles: tids (RV_NONE pad).  cond: RV_C_* bits.
__device__ __noinline__ WinRes hand_calc(const Tables& T, const uint8_t* tiles, int n, int n_melds, const uint8_t* meld_type,
                                   const uint8_t (*meld_tiles)[4], int win_tid, const uint8_t* dora, int n_dora,
                                   const uint8_t* ura, int n_ura, uint32_t cond, int player_wind, int round_wind,
                                   uint32_t honba, bool sanma = false, int kita_count = 0, BlockPace* pace = nullptr) {
  WinRes out{false, false, false, 0, 0, 0, 0, 0, 0};
  RV_STAT(0);
  Cnt hand, full;
  cnt_zero(hand);
  int aka = 0;
  for (int i = 0; i < n; i++) {
    int t = tiles[i];
    cnt_add(hand, t >> 2);
    aka += tid_is_aka(t);
  }
  full = hand;
  MeldView mv;
  mv.n = (uint8_t)n_melds;
  uint64_t meld_present = 0;
  bool menzen = true;
  for (int m = 0; m < n_melds; m++) {
    int ty = meld_type[m];
    mv.type[m] = (uint8_t)ty;
    int lo = 99;
    for (int k = 0; k < 4; k++) {
      int t = meld_tiles[m][k];
      if (t == RV_NONE) continue;
      aka += tid_is_aka(t);
      cnt_add(full, t >> 2);
      meld_present |= 1ull << (t >> 2);
      lo = min(lo, t >> 2);
    }
    // chi melds are sorted by the reference; pon/kan tiles share one kind.  tiles[0]/4 == lowest kind in all cases
    // that occur (a pon/kan has a single kind).
    mv.t0[m] = (uint8_t)lo;
    if (ty != RV_MELD_ANKAN) menzen = false;
    // hand_evaluator.rs:44-51: a kan whose 4 tiles are also in `tiles` counts 3 in the agari hand
    if (ty >= RV_MELD_DAIMINKAN && cnt_get(hand, lo) == 4) cnt_sub(hand, lo, 1);
  }
  int total = cnt_total(hand) + 3 * n_melds;
  int win34 = win_tid >> 2;
  if (total == 13) {
    cnt_add(hand, win34);
    cnt_add(full, win34);
    aka += tid_is_aka(win_tid);
  }
  SuitInfo si;
  load_info(T, hand, si);
  bool std_shape = standard_agari(si);
  if (!std_shape && !chiitoi14(hand) && !kokushi14(hand)) {
    pace_arrive(pace);
    return out;
  }
  out.has_shape = true;
  RV_STAT(1);
  WinCtx x;
  x.tsumo = cond & RV_C_TSUMO;
  x.riichi = cond & RV_C_RIICHI;
  x.double_riichi = cond & RV_C_DOUBLE_RIICHI;
  x.ippatsu = cond & RV_C_IPPATSU;
  x.haitei = cond & RV_C_HAITEI;
  x.houtei = cond & RV_C_HOUTEI;
  x.rinshan = cond & RV_C_RINSHAN;
  x.chankan = cond & RV_C_CHANKAN;
  x.first_turn = cond & RV_C_TSUMO_FIRST_TURN;
  x.menzen = menzen;
  int dc = 0, uc = 0;
  for (int i = 0; i < n_dora; i++) {
    int nt = sanma ? next_dora_tile_sanma(dora[i] >> 2) : next_dora_tile(dora[i] >> 2);
    dc += cnt_get(full, nt) + ((sanma && nt == 30) ? kita_count : 0);   // hand_evaluator_3p.rs:108-115
  }
  for (int i = 0; i < n_ura; i++) {
    int nt = sanma ? next_dora_tile_sanma(ura[i] >> 2) : next_dora_tile(ura[i] >> 2);
    uc += cnt_get(full, nt) + ((sanma && nt == 30) ? kita_count : 0);
  }
  x.dora = (uint8_t)dc;
  x.ura = (uint8_t)uc;
  x.aka = (uint8_t)aka;
  x.nuki = (uint8_t)(sanma ? kita_count : 0);
  x.round_wind = (uint8_t)(27 + round_wind);
  x.seat_wind = (uint8_t)(27 + player_wind);
  uint64_t all = cnt_present(hand) | meld_present;
  YakuRes y = calculate_yaku(T, hand, all, mv, x, win34, std_shape, pace);
  bool is_oya = player_wind == 0;
  int scoring_han = (y.yakuman == 0 && y.han >= 13) ? 13 : y.han;
  ScoreRes sc = calc_score(scoring_han & 0xFF, y.fu, is_oya, x.tsumo, honba, sanma ? 3 : 4);
  bool has_yaku = (y.mask & ~((1ull << 31) | (1ull << 32) | (1ull << 33) | (sanma ? (1ull << 34) : 0ull))) != 0;
  out.is_win = (has_yaku || y.yakuman > 0) && y.han >= 1;
  out.yakuman = y.yakuman > 0;
  out.han = y.han;
  out.fu = y.fu;
  out.yaku_mask = y.mask;
  out.ron = sc.pay_ron;
  out.oya = sc.pay_oya;
  out.ko = sc.pay_ko;
  return out;
}

// Split in two so that a batch can run the uniform part on every hand and the yaku evaluation — long, divergent, and needed by
// the few hands that have a winning shape — on a compacted list (hand_shape_kernel / hand_yaku_kernel):
//   hand_eval_shape  validation, histograms, wait set, both shanten numbers; returns true when hand_calc has work to do
//                    (the 14-tile-equivalent hand is a standard, seven-pairs or thirteen-orphans shape);
//   hand_eval_win    hand_calc and the win fields of the result.
__device__ __forceinline__ bool hand_eval_shape(const Tables& T, const rv_hand_query& h, rv_hand_result& o) {
  memset(&o, 0, sizeof o);
  bool valid = h.n_tiles <= 14 && h.n_melds <= 4 && h.win_tile < 136 && h.n_dora <= 5 && h.n_ura <= 5 && h.player_wind < 4 &&
               h.round_wind < 4;
  Cnt c;
  cnt_zero(c);
  for (int k = 0; valid && k < h.n_tiles; k++) {
    if (h.tiles[k] >= 136 || cnt_get(c, h.tiles[k] >> 2) >= 4) valid = false;
    else cnt_add(c, h.tiles[k] >> 2);
  }
  for (int m = 0; valid && m < h.n_melds; m++) {
    if (h.meld_type[m] > RV_MELD_KAKAN) valid = false;
    for (int k = 0; k < 3; k++)
      if (h.meld_tiles[m][k] >= 136) valid = false;
    if (h.meld_tiles[m][3] >= 136 && h.meld_tiles[m][3] != RV_NONE) valid = false;
  }
  for (int k = 0; valid && k < h.n_dora; k++) valid = h.dora_ind[k] < 136;
  for (int k = 0; valid && k < h.n_ura; k++) valid = h.ura_ind[k] < 136;
  if (!valid) {
    o.shanten = o.shanten13 = 127;
    return false;
  }
  // concealed histogram (kan melds whose tiles are also listed count 3, as HandEvaluator::new)
  Cnt raw = c;
  for (int m = 0; m < h.n_melds; m++)
    if (h.meld_type[m] >= RV_MELD_DAIMINKAN) {
      int kind = h.meld_tiles[m][0] >> 2;
      if (cnt_get(c, kind) == 4) cnt_sub(c, kind, 1);
    }
  int total = cnt_total(c) + 3 * h.n_melds;
  int win34 = h.win_tile >> 2;
  Cnt c13 = c, r13 = raw, c14 = c;
  bool ok13 = total == 13;
  if (total == 14 && cnt_get(c, win34) > 0) {
    cnt_sub(c13, win34);
    cnt_sub(r13, win34);
    ok13 = true;
  }
  if (total == 13) cnt_add(c14, win34);            // HandEvaluator::calc adds the win tile to a 13-tile hand
  o.wait_mask = ok13 ? waits13(T, c13) : 0;
  Cnt r14 = raw;
  int n14 = h.n_tiles;
  if (total == 13 && cnt_get(r14, win34) < 4) {
    cnt_add(r14, win34);
    n14++;
  }
  const bool sanma = h.sanma & 1;   // sanma queries: calculate_shanten_3p (shanten.rs:470-484)
  o.shanten = (int8_t)(sanma ? shanten_counts_3p(T, r14, n14 / 3) : shanten_counts(T, r14, n14 / 3));
  o.shanten13 = ok13 ? (int8_t)(sanma ? shanten_counts_3p(T, r13, cnt_total(r13) / 3) : shanten_counts(T, r13, cnt_total(r13) / 3))
                     : (int8_t)127;
  return agari14(T, c14);           // the shape test hand_calc starts with (a five-of-a-kind c14 is clamped, hence no shape)
}
__device__ __noinline__ void hand_eval_win(const Tables& T, const rv_hand_query& h, rv_hand_result& o, BlockPace* pace = nullptr) {
  WinRes r = hand_calc(T, h.tiles, h.n_tiles, h.n_melds, h.meld_type, h.meld_tiles, h.win_tile, h.dora_ind, h.n_dora,
                       h.ura_ind, h.n_ura, h.cond, h.player_wind, h.round_wind, h.honba, h.sanma & 1, h.kita_count, pace);
  o.is_win = r.is_win;
  o.yakuman = r.yakuman;
  o.has_win_shape = r.has_shape;
  o.han = (uint8_t)r.han;
  o.fu = (uint8_t)r.fu;
  o.ron_agari = r.ron;
  o.tsumo_agari_oya = r.oya;
  o.tsumo_agari_ko = r.ko;
  o.yaku_mask = r.yaku_mask;
  o.n_yaku = (uint8_t)__popcll(r.yaku_mask);
}
// One query of rv_hand_eval_batch: HandEvaluator::new + calc + get_waits_u8 (hand_evaluator.rs:24-213) and
// calculate_shanten[_3p] (shanten.rs:250-261, 470-484).  A query that is not a hand (more than 14 concealed tiles or 4
// melds, a tile id outside 0..135, more than four copies of a kind) gets a zeroed result with shanten = shanten13 = 127.
__device__ __noinline__ void hand_eval_one(const Tables& T, const rv_hand_query& h, rv_hand_result& o) {
  if (hand_eval_shape(T, h, o)) hand_eval_win(T, h, o);
}

}  // namespace rv
