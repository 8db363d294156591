// Device-side game state machine for 4-player Riichi (one rv_game_state record per game).
//
// Replaces (reference paths relative to riichienv-core/src):
//   state/mod.rs:330-1315  GameState::step            -> step_apply()
//   state/mod.rs:1317-1413 _resolve_discard           -> resolve_discard()
//   state/mod.rs:1415-1547 _resolve_kan               -> resolve_kan()
//   state/mod.rs:1549-1593 _accept_riichi/_deal_next  -> accept_riichi()/deal_next()
//   state/mod.rs:1595-1844 round / game flow          -> next_round()/init_round()
//   state/mod.rs:1846-2081 ryukyoku, abortive draws, kan dora, ura, end game
//   state/legal_actions.rs:11-252  turn actions       -> TurnInfo + enum_turn_actions()
//   state/legal_actions.rs:254-508 claim actions      -> gen_claims()
//   state/wall.rs:36-88    seeded wall                -> wall_shuffle()
// The reference walks Vec/HashMap structures and re-runs a backtracking agari test
// up to 14x34 times per turn; here hands are fixed byte arrays in the HBM record and
// every agari / tenpai / wait question is answered from the suit tables (hand.cuh).
#pragma once
#include "hand.cuh"

namespace rv {

typedef rv_game_state G;
constexpr int MAXP = 4;   // seat capacity of the record; the seat count is num_players(g): 4, or 3 for sanma
__device__ __forceinline__ bool is_sanma(const rv_game_state& g) { return g.game_mode >= 3; }   // game_variant.rs:12-36
__device__ __forceinline__ int num_players(const rv_game_state& g) { return g.game_mode >= 3 ? 3 : 4; }

struct Ctx {
  Tables T;
  uint32_t* log;      // this game's event log region or nullptr
  uint32_t log_cap;   // words
  bool defer_init;    // rollout kernels: next_round() parks the deal in g.pending_init so warps can batch it
  bool defer_tail;    // rollout kernels: act_fast() parks the follow-up of a discard it cannot finish inline in g.pending_tail
  uint32_t* idbits;   // observe+step kernel: [MAXP][3] words; the random_step<true> routines leave here, per acting seat, the
                      // set of action ids (Action::encode) of the legal list they picked from — Observation::mask for free
};

// The rollout kernels run on a shared-memory copy of the record's hot prefix (RV_HOT_BYTES); the word after that prefix
// holds the address of the HBM record, where the cold arrays (wall, river, claims) stay.  Everywhere else &g IS the record.
#ifndef RV_COLD_HOOK
__device__ __forceinline__ G& cold(G& g) {
#ifdef __CUDA_ARCH__
  if (__isShared(&g)) return **reinterpret_cast<G**>(reinterpret_cast<char*>(&g) + RV_HOT_BYTES);
#endif
  return g;
}
#else
__device__ __forceinline__ G& cold(G& g) { return RV_COLD_HOOK(g); }
#endif
__device__ __forceinline__ const G& cold(const G& g) { return cold(const_cast<G&>(g)); }

__device__ __forceinline__ bool rule(const G& g, uint32_t bit) { return (g.rule_bits & bit) != 0; }

// ------------------------------------------------------------------ events
__device__ __noinline__ void ev_push(const Ctx& cx, G& g, const uint32_t* w, int n) {
  uint64_t h = g.ev_hash;
  uint32_t base = g.ev_words;
  for (int i = 0; i < n; i++) {
    h = (h ^ w[i]) * 0x100000001b3ull;
    if (cx.log && base + i < cx.log_cap) cx.log[base + i] = w[i];
  }
  if (cx.log && base + n > cx.log_cap) g.overflow |= 4;   // the log is full: words dropped (hash and counters go on)
  g.ev_hash = h;
  g.ev_words = base + n;
  g.ev_count++;
}
__device__ __forceinline__ uint32_t ev_w0(int type, int n, int a, int b) {
  return (uint32_t)type | ((uint32_t)n << 8) | ((uint32_t)(a & 0xFF) << 16) | ((uint32_t)(b & 0xFF) << 24);
}
__device__ __noinline__ void ev_simple(const Ctx& cx, G& g, int type, int a = 0, int b = 0) {
  uint32_t w = ev_w0(type, 1, a, b);
  ev_push(cx, g, &w, 1);
}
__device__ __noinline__ void ev_meld(const Ctx& cx, G& g, int type, int actor, int tile, int b0, int b1, int b2, int b3) {
  uint32_t w[2] = {ev_w0(type, 2, actor, tile),
                   (uint32_t)(b0 & 0xFF) | ((uint32_t)(b1 & 0xFF) << 8) | ((uint32_t)(b2 & 0xFF) << 16) | ((uint32_t)(b3 & 0xFF) << 24)};
  ev_push(cx, g, w, 2);
}

// ------------------------------------------------------------------ hand helpers
__device__ __forceinline__ void cache_add(G& g, int p, int kind) {
  int su = kind / 9, pos = kind - 9 * su;
  g.c_cnt[p][su] += 1ull << (4 * pos);
  g.c_key[p][su] += (uint32_t)pow5(pos);
}
__device__ __forceinline__ void cache_sub(G& g, int p, int kind) {
  int su = kind / 9, pos = kind - 9 * su;
  g.c_cnt[p][su] -= 1ull << (4 * pos);
  g.c_key[p][su] -= (uint32_t)pow5(pos);
}
__device__ __noinline__ bool hand_remove_first(G& g, int p, int tile) {
  int n = g.hand_len[p];
  #pragma unroll 1
  for (int i = 0; i < n; i++)
    if (g.hand[p][i] == tile) {
      #pragma unroll 1
      for (int j = i; j + 1 < n; j++) g.hand[p][j] = g.hand[p][j + 1];
      g.hand[p][n - 1] = RV_NONE;
      g.hand_len[p] = (uint8_t)(n - 1);
      cache_sub(g, p, tile >> 2);
      return true;
    }
  return false;
}
__device__ __noinline__ void hand_push(G& g, int p, int tile) {
  int n = g.hand_len[p];
  if (n < RV_HAND_CAP) {
    g.hand[p][n] = (uint8_t)tile;
    g.hand_len[p] = (uint8_t)(n + 1);
    cache_add(g, p, tile >> 2);
  } else {
    g.overflow = 1;
  }
}
__device__ __noinline__ void hand_sort(G& g, int p) {
  int n = g.hand_len[p];
  for (int i = 1; i < n; i++) {
    uint8_t v = g.hand[p][i];
    int j = i - 1;
    while (j >= 0 && g.hand[p][j] > v) {
      g.hand[p][j + 1] = g.hand[p][j];
      j--;
    }
    g.hand[p][j + 1] = v;
  }
}
__device__ __forceinline__ Cnt hand_cnt(const G& g, int p) {
  Cnt c;
  c.s[0] = g.c_cnt[p][0];
  c.s[1] = g.c_cnt[p][1];
  c.s[2] = g.c_cnt[p][2];
  c.s[3] = g.c_cnt[p][3];
  return c;
}
__device__ __forceinline__ void hand_info(const Tables& T, const G& g, int p, SuitInfo& si) {
  si.e[0] = __ldg(&T.suit_info[clamp_key9(g.c_key[p][0])]);
  si.e[1] = __ldg(&T.suit_info[clamp_key9(g.c_key[p][1])]);
  si.e[2] = __ldg(&T.suit_info[clamp_key9(g.c_key[p][2])]);
  si.e[3] = __ldg(&T.honor_info[clamp_key7(g.c_key[p][3])]);
}
__device__ __forceinline__ uint64_t river_kinds(const G& g, int p) { return g.c_river_kinds[p]; }
// c_waits[p] = get_waits_u8 when the hand is 13-tile-equivalent, else 0 (hand_evaluator.rs:196-201)
__device__ __noinline__ void waits_update(const Tables& T, G& g, int p) {
  uint64_t w = 0;
  if (g.hand_len[p] + 3 * g.n_melds[p] == 13) {
    SuitInfo si;
    hand_info(T, g, p, si);
    w = waits13(hand_cnt(g, p), si);
  }
  g.c_waits[p] = w;
}
// Rebuild every derived cache from the canonical fields (after rv_vec_set_state)
__device__ __noinline__ void refresh_caches(const Tables& T, G& g) {
  for (int p = 0; p < 4; p++) {
    for (int k = 0; k < 4; k++) g.c_cnt[p][k] = 0, g.c_key[p][k] = 0;
    for (int i = 0; i < g.hand_len[p]; i++) {
      int kind = g.hand[p][i] >> 2, su = kind / 9, pos = kind - 9 * su;
      g.c_cnt[p][su] += 1ull << (4 * pos);
      g.c_key[p][su] += (uint32_t)pow5(pos);
    }
    uint64_t m = 0;
    int n = min((int)g.n_river[p], RV_RIVER_CAP);
    for (int i = 0; i < n; i++) m |= 1ull << (cold(g).river[p][i] >> 2);
    g.c_river_kinds[p] = m;
    waits_update(T, g, p);
  }
}
__device__ __forceinline__ bool tid_terminal(int t) {  // types.rs:364-369
  int k = t >> 2;
  return k >= 27 || k % 9 == 0 || k % 9 == 8;
}
__device__ __forceinline__ bool any_open_meld(const G& g, int p) {
  for (int m = 0; m < g.n_melds[p]; m++)
    if (g.meld_type[p][m] != RV_MELD_ANKAN) return true;
  return false;
}
__device__ __forceinline__ bool all_meldless(const G& g) {
  return (g.n_melds[0] | g.n_melds[1] | g.n_melds[2] | g.n_melds[3]) == 0;
}

// Conditions common to every calc call site (riichi flags, winds, honba)
__device__ __forceinline__ uint32_t base_cond(const G& g, int p) {
  uint32_t f = g.flags[p], c = 0;
  if (f & RV_F_RIICHI_DECLARED) c |= RV_C_RIICHI;
  if (f & RV_F_DOUBLE_RIICHI) c |= RV_C_DOUBLE_RIICHI;
  if (f & RV_F_IPPATSU_CYCLE) c |= RV_C_IPPATSU;
  return c;
}
// HandEvaluator::new(hand, melds).calc(win_tile, dora, ura, cond) for a seat
// ura indicators of the revealed doras.  4P: wall[5+2i] while still in the Vec (state/mod.rs:2048-2058; absolute index:
// the Vec lost rinshan_draw_count front tiles).  3P: pre-extracted tiles[9+2i] (state_3p/wall.rs:104-112).
__device__ __forceinline__ int ura_indicators(const G& g, uint8_t* ura) {
  int n = 0;
  bool sanma = is_sanma(g);
  for (int i = 0; i < g.n_dora; i++) {
    int idx = (sanma ? 9 : 5) + 2 * i;
    if (sanma || idx < g.wall_top) ura[n++] = cold(g).wall[idx];
  }
  return n;
}
__device__ inline WinRes seat_calc(const Ctx& cx, const G& g, int p, int win_tile, uint32_t cond, bool with_ura, uint32_t honba,
                                   bool with_kita = false) {
  const int np = num_players(g);
  uint8_t ura[5];
  int n_ura = with_ura ? ura_indicators(g, ura) : 0;
  return hand_calc(cx.T, g.hand[p], g.hand_len[p], g.n_melds[p], g.meld_type[p], g.meld_tiles[p], win_tile, g.dora_ind,
                   g.n_dora, ura, n_ura, cond, (p + np - g.oya) % np, g.round_wind % 4, honba, is_sanma(g),
                   with_kita ? cold(g).n_kita[p] : 0);
}

// ------------------------------------------------------------------ wall (state/wall.rs:36-88)
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  uint64_t z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
struct ChaCha12 {
  uint32_t key[8];
  uint32_t buf[64];     // output words of blocks 0..3, in order.  A 136-tile shuffle draws at most 28 chunks x 2 words.
  uint64_t counter;
  int idx, filled;
  __host__ __device__ static inline uint32_t rotl(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }
  __host__ __device__ inline void init(uint64_t state) {
    // rand_core SeedableRng::seed_from_u64: PCG32 (XSH-RR) expansion
    for (int i = 0; i < 8; i++) {
      state = state * 6364136223846793005ull + 11634580027462260723ull;
      uint32_t xs = (uint32_t)(((state >> 18) ^ state) >> 27);
      uint32_t rot = (uint32_t)(state >> 59);
      key[i] = (xs >> rot) | (xs << ((32 - rot) & 31));
    }
    counter = 0;
    idx = 0;
    filled = 0;
    // Three blocks up front, the same instructions in every lane of a warp: drawn on demand, the lanes of a warp cross a block
    // boundary at different iterations (the bias-correction draw of random_below is data dependent), so each refill ran
    // with a handful of lanes active — ~10 block computations per warp and shuffle instead of 3.
    for (int b = 0; b < 3; b++) refill();
  }
  __host__ __device__ inline void refill() {   // appends block `counter`
    uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3],
                       key[4], key[5], key[6], key[7], (uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u};
    uint32_t s[16];
    for (int i = 0; i < 16; i++) s[i] = in[i];
#define RV_QR(a, b, c, d)                                                  \
  s[a] += s[b]; s[d] ^= s[a]; s[d] = rotl(s[d], 16); s[c] += s[d]; s[b] ^= s[c]; s[b] = rotl(s[b], 12); \
  s[a] += s[b]; s[d] ^= s[a]; s[d] = rotl(s[d], 8);  s[c] += s[d]; s[b] ^= s[c]; s[b] = rotl(s[b], 7);
    for (int r = 0; r < 6; r++) {
      RV_QR(0, 4, 8, 12) RV_QR(1, 5, 9, 13) RV_QR(2, 6, 10, 14) RV_QR(3, 7, 11, 15)
      RV_QR(0, 5, 10, 15) RV_QR(1, 6, 11, 12) RV_QR(2, 7, 8, 13) RV_QR(3, 4, 9, 14)
    }
#undef RV_QR
    const int o = filled & 63;            // (a fifth block cannot be needed; the mask only keeps the store in bounds)
    for (int i = 0; i < 16; i++) buf[o + i] = s[i] + in[i];
    counter++;
    filled += 16;
  }
  __host__ __device__ inline uint32_t next_u32() {
    if (idx >= filled) refill();
    return buf[(idx++) & 63];
  }
};
// rand UniformInt<u32>::sample_single_inclusive(0, range-1): widening multiply + one bias-correction draw
__host__ __device__ inline uint32_t random_below(ChaCha12& rng, uint32_t range) {
  uint64_t m = (uint64_t)rng.next_u32() * range;
  uint32_t hi = (uint32_t)(m >> 32), lo = (uint32_t)m;
  if (lo > (uint32_t)(0u - range)) {
    uint64_t m2 = (uint64_t)rng.next_u32() * range;
    if ((uint64_t)lo + (uint32_t)(m2 >> 32) > 0xFFFFFFFFull) hi += 1;
  }
  return hi;
}
// SliceRandom::shuffle via IncreasingUniform (one u32 draw feeds several indices), then reverse.
// `out` receives the reference's `wall.tiles` order.  n = 136 (4P) or 108 (3P tile set).
__host__ __device__ __forceinline__ uint32_t rv_umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
__host__ __device__ __noinline__ void wall_from_seed(uint64_t seed, uint64_t hand_index, int n, uint8_t* out) {
  static const uint32_t RECIP[137] = {   // floor(2^32 / d) (d = 1: 2^32 - 1; the fix-ups below absorb the difference)
      0u, 0xFFFFFFFFu, 2147483648u, 1431655765u, 1073741824u, 858993459u, 715827882u, 613566756u, 536870912u, 477218588u, 429496729u, 390451572u, 357913941u, 330382099u, 306783378u, 286331153u, 268435456u, 252645135u, 238609294u, 226050910u, 214748364u, 204522252u, 195225786u, 186737708u, 178956970u, 171798691u, 165191049u, 159072862u, 153391689u, 148102320u, 143165576u, 138547332u, 134217728u, 130150524u, 126322567u, 122713351u, 119304647u, 116080197u, 113025455u, 110127366u, 107374182u, 104755299u, 102261126u, 99882960u, 97612893u, 95443717u, 93368854u, 91382282u, 89478485u, 87652393u, 85899345u, 84215045u, 82595524u, 81037118u, 79536431u, 78090314u, 76695844u, 75350303u, 74051160u, 72796055u, 71582788u, 70409299u, 69273666u, 68174084u, 67108864u, 66076419u, 65075262u, 64103989u, 63161283u, 62245902u, 61356675u, 60492497u, 59652323u, 58835168u, 58040098u, 57266230u, 56512727u, 55778796u, 55063683u, 54366674u, 53687091u, 53024287u, 52377649u, 51746593u, 51130563u, 50529027u, 49941480u, 49367440u, 48806446u, 48258059u, 47721858u, 47197442u, 46684427u, 46182444u, 45691141u, 45210182u, 44739242u, 44278013u, 43826196u, 43383508u, 42949672u, 42524428u, 42107522u, 41698711u, 41297762u, 40904450u, 40518559u, 40139881u, 39768215u, 39403369u, 39045157u, 38693399u, 38347922u, 38008560u, 37675151u, 37347541u, 37025580u, 36709122u, 36398027u, 36092162u, 35791394u, 35495597u, 35204649u, 34918433u, 34636833u, 34359738u, 34087042u, 33818640u, 33554432u, 33294320u, 33038209u, 32786009u, 32537631u, 32292987u, 32051994u, 31814572u, 31580641u};
  uint8_t w[136];
  if (n == 136) {
    for (int i = 0; i < 136; i++) w[i] = (uint8_t)i;
  } else {
    int k = 0;
    for (int i = 0; i < 136; i++) {
      int t = i >> 2;
      if (t >= 1 && t <= 7) continue;
      w[k++] = (uint8_t)i;
    }
  }
  ChaCha12 rng;
  rng.init(splitmix64(seed + hand_index));
  uint32_t cur_n = 0, chunk = 0, remaining = 1;
  for (int i = 0; i < n; i++) {
    uint32_t next_n = cur_n + 1, rem;
    if (remaining == 0) {
      uint32_t product = next_n, current = next_n + 1;
      while (true) {
        uint64_t p = (uint64_t)product * current;
        if (p > 0xFFFFFFFFull) break;
        product = (uint32_t)p;
        current++;
      }
      chunk = random_below(rng, product);
      rem = (current - next_n) - 1;
    } else {
      rem = remaining - 1;
    }
    uint32_t j;
    if (rem == 0) {
      j = chunk;
    } else {
      // chunk / next_n and chunk % next_n without an integer division (~25 instructions each on the GPU, 108 of them per
      // shuffle): q = hi32(chunk * floor(2^32 / d)) is q or short of it by at most 2
      uint32_t q = rv_umulhi(chunk, RECIP[next_n]);
      uint32_t r = chunk - q * next_n;
      if (r >= next_n) r -= next_n, q++;
      if (r >= next_n) r -= next_n, q++;
      j = r;
      chunk = q;
    }
    remaining = rem;
    cur_n = next_n;
    uint8_t t = w[i];
    w[i] = w[j];
    w[j] = t;
  }
  for (int i = 0; i < n; i++) out[i] = w[n - 1 - i];
}

// ------------------------------------------------------------------ forward decls
__device__ void trigger_ryukyoku(const Ctx& cx, G& g, int reason);
__device__ void next_round(const Ctx& cx, G& g, bool oya_won, bool is_draw);

// state/mod.rs:2021-2046
__device__ __noinline__ void reveal_kan_dora(const Ctx& cx, G& g) {
  int count = g.n_dora;
  if (count < 5) {
    bool sanma = is_sanma(g);
    int idx = (sanma ? 8 : 4) + 2 * count;   // 3P: pre-extracted dora_indicator_tiles (state_3p/mod.rs:1893-1915)
    if (sanma || idx < g.wall_top) {
      g.dora_ind[count] = cold(g).wall[idx];
      g.n_dora = (uint8_t)(count + 1);
      ev_simple(cx, g, RV_EV_DORA, 0, cold(g).wall[idx]);
    }
  }
}
__device__ inline void flush_pending_kan_dora(const Ctx& cx, G& g) {
  while (g.pending_kan_dora_count > 0) {
    g.pending_kan_dora_count--;
    reveal_kan_dora(cx, g);
  }
}
// state/mod.rs:1549-1567
__device__ __noinline__ void accept_riichi(const Ctx& cx, G& g) {
  int p = g.riichi_pending_acceptance;
  if (p != RV_NONE) {
    g.score[p] -= 1000;
    cold(g).score_delta[p] -= 1000;
    g.riichi_sticks += 1;
    g.flags[p] |= RV_F_RIICHI_DECLARED | RV_F_IPPATSU_CYCLE;
    ev_simple(cx, g, RV_EV_REACH_ACCEPTED, p);
    g.riichi_pending_acceptance = RV_NONE;
  }
}
// state/mod.rs:1569-1593
__device__ __noinline__ void deal_next(const Ctx& cx, G& g) {
  g.is_rinshan_flag = 0;
  if (g.drawable_count == 0) {
    trigger_ryukyoku(cx, g, RV_RK_EXHAUSTIVE);
    return;
  }
  if (g.wall_top > g.rinshan_draw_count) {
    int t = cold(g).wall[g.wall_top - 1];
    g.wall_top--;
    g.drawable_count--;
    int pid = g.current_player;
    hand_push(g, pid, t);
    g.c_waits[pid] = 0;   // 14-tile-equivalent now
    g.drawn_tile = (uint8_t)t;
    g.needs_tsumo = 0;
    g.phase = RV_WAIT_ACT;
    g.active_mask = (uint8_t)(1u << pid);
    ev_simple(cx, g, RV_EV_TSUMO, pid, t);
    g.forbidden[pid][0] = g.forbidden[pid][1] = RV_NONE;
  }
}

// ------------------------------------------------------------------ round setup (state/mod.rs:1695-1844)
// `custom_wall`: nullptr -> seeded shuffle; else 136 tids in the order passed to reset(wall=) (load_wall reverses).
__device__ __noinline__ void init_round(const Ctx& cx, G& g, int oya, int round_wind, int honba, uint32_t kyotaku,
                                  const uint8_t* custom_wall, const int32_t* scores) {
  const int np = num_players(g);
  RV_STAT(2);
  g.oya = g.kyoku_idx = g.current_player = (uint8_t)oya;
  g.honba = (uint8_t)honba;
  g.riichi_sticks = kyotaku;
  g.round_wind = (uint8_t)round_wind;
  for (int p = 0; p < MAXP; p++) {  // PlayerState::reset_round (state/player.rs:66-86); the unused 3P slot is reset too
    for (int i = 0; i < RV_HAND_CAP; i++) g.hand[p][i] = RV_NONE;
    g.hand_len[p] = 0;
    for (int m = 0; m < 4; m++) {
      for (int k = 0; k < 4; k++) g.meld_tiles[p][m][k] = RV_NONE;
      g.meld_type[p][m] = cold(g).meld_from[p][m] = cold(g).meld_called[p][m] = RV_NONE;
    }
    g.n_melds[p] = 0;
    for (int i = 0; i < RV_RIVER_CAP; i++) cold(g).river[p][i] = RV_NONE;
    g.n_river[p] = 0;
    g.river_tedashi[p] = 0;
    cold(g).river_riichi[p] = 0;
    cold(g).riichi_decl_idx[p] = RV_NONE;
    g.flags[p] = RV_F_NAGASHI_ELIGIBLE;
    cold(g).pao[p][0] = cold(g).pao[p][1] = RV_NONE;
    g.forbidden[p][0] = g.forbidden[p][1] = RV_NONE;
    cold(g).score_delta[p] = 0;
    g.n_claims[p] = 0;
    cold(g).riichi_sutehai[p] = cold(g).last_tedashi[p] = RV_NONE;
    cold(g).n_kita[p] = 0;
    for (int k = 0; k < 4; k++) g.c_cnt[p][k] = 0, g.c_key[p][k] = 0;
    g.c_river_kinds[p] = 0;
    g.c_waits[p] = 0;
    if (scores && p < np) g.score[p] = scores[p];
  }
  if (np == 3) {   // the record has four seat slots; the unused one stays inert
    g.flags[3] = 0;
    g.score[3] = 0;
  }
  g.is_done = 0;
  g.pending_kan_pid = g.pending_kan_type = g.pending_kan_tile = RV_NONE;
  g.is_rinshan_flag = 0;
  g.rinshan_draw_count = 0;
  g.pending_kan_dora_count = 0;
  g.is_first_turn = 1;
  g.riichi_pending_acceptance = RV_NONE;
  g.turn_count = 0;
  g.needs_tsumo = 1;
  g.last_discard_pid = g.last_discard_tile = RV_NONE;
  g.pending_init[0] = g.pending_init[1] = g.pending_init[2] = RV_NONE;
  g.pending_tail[0] = RV_NONE;
  g.pending_tail[1] = 0;
  const int wl = np == 3 ? 108 : 136;
  // The wall is built in a local array, written to the record with 8-byte stores and DEALT FROM THE LOCAL COPY: reading the
  // 53 dealt tiles back from the record one byte at a time was a chain of 53 dependent L2 round trips (the cold part of the
  // record lives in global memory) — most of the 17 k cycles a deal took in the rollout's DEAL class.
  alignas(8) uint8_t w[136];
  if (custom_wall) {
    for (int i = 0; i < wl; i++) w[i] = custom_wall[wl - 1 - i];  // wall.rs:69-72 / state_3p/wall.rs:148-149
  } else {
    wall_from_seed(g.seed, cold(g).hand_index, wl, w);
    cold(g).hand_index++;
  }
  for (int i = wl; i < 136; i++) w[i] = RV_NONE;
  {
    uint64_t* dst = reinterpret_cast<uint64_t*>(cold(g).wall);          // offset RV_HOT_BYTES of a 16-byte aligned record
    const uint64_t* src = reinterpret_cast<const uint64_t*>(w);
    for (int i = 0; i < 136 / 8; i++) dst[i] = src[i];
  }
  g.wall_len = (uint8_t)wl;
  g.wall_top = (uint8_t)wl;
  g.n_dora = 1;
  g.dora_ind[0] = w[np == 3 ? 8 : 4];   // state_3p/wall.rs:104-112
  for (int i = 1; i < 5; i++) g.dora_ind[i] = RV_NONE;
  cold(g).kyoku_count++;
  // deal: 3 x (4 tiles per seat from oya), then 1 each; tiles pop from the back
  for (int r = 0; r < 3; r++)
    for (int idx = 0; idx < np; idx++) {
      int p = (idx + oya) % np;
      for (int k = 0; k < 4; k++) hand_push(g, p, w[--g.wall_top]);
    }
  for (int idx = 0; idx < np; idx++) {
    int p = (idx + oya) % np;
    hand_push(g, p, w[--g.wall_top]);
  }
  for (int p = 0; p < np; p++) {
    hand_sort(g, p);
    if (p != oya) waits_update(cx.T, g, p);   // the dealer draws a 14th tile right below
  }
  g.drawable_count = (uint8_t)(g.wall_top - 14);
  {  // start_kyoku: 4P 19 words, 3P 15 words (np scores, 13*np tehai bytes)
    uint32_t ew[19];
    const int nb = 13 * np, ntw = (nb + 3) / 4, nwords = 2 + np + ntw;
    ew[0] = ev_w0(RV_EV_START_KYOKU, nwords, round_wind % 4, oya);
    ew[1] = (uint32_t)honba | ((uint32_t)g.dora_ind[0] << 8) | ((kyotaku & 0xFFFF) << 16);
    for (int i = 0; i < np; i++) ew[2 + i] = (uint32_t)g.score[i];
    for (int k = 0; k < ntw; k++) {
      uint32_t v = 0;
      for (int b = 0; b < 4; b++) {
        int flat = k * 4 + b;
        int t = flat < nb ? g.hand[flat / 13][flat % 13] : RV_NONE;
        v |= (uint32_t)t << (8 * b);
      }
      ew[2 + np + k] = v;
    }
    ev_push(cx, g, ew, nwords);
  }
  g.phase = RV_WAIT_ACT;
  g.active_mask = (uint8_t)(1u << oya);
  {
    int t = w[--g.wall_top];
    g.drawable_count--;
    hand_push(g, oya, t);
    g.drawn_tile = (uint8_t)t;
    g.needs_tsumo = 0;
    ev_simple(cx, g, RV_EV_TSUMO, oya, t);
  }
}

// RiichiEnv::new + reset (env.rs:82-118, 799-851): fresh game, logs cleared
__device__ __noinline__ void game_reset(const Ctx& cx, G& g, int oya, int round_wind, int honba, uint32_t kyotaku,
                                  const uint8_t* custom_wall, const int32_t* scores) {
  const int np = num_players(g);
  g.ev_hash = 0xcbf29ce484222325ull;
  g.ev_count = g.ev_words = g.step_count = cold(g).kyoku_count = 0;
  g.last_error = RV_NONE;   // NOTE: the reference never clears last_error on reset (state/mod.rs:171-187); a fresh
                            // VecEnv has none, and rv_vec_reset is documented to clear it.
  g.overflow = 0;
  ev_simple(cx, g, RV_EV_START_GAME);
  const int32_t st = np == 3 ? 35000 : 25000;   // state_3p/game_mode.rs:31-33
  int32_t def[MAXP] = {st, st, st, st};
  init_round(cx, g, oya, round_wind, honba, kyotaku, custom_wall, scores ? scores : def);
}

// state/mod.rs:2071-2081
__device__ __noinline__ void end_game(const Ctx& cx, G& g) {
  g.is_done = 1;
  ev_simple(cx, g, RV_EV_END_KYOKU);
  ev_simple(cx, g, RV_EV_END_GAME);
}

// state/mod.rs:1595-1688
__device__ __noinline__ void next_round(const Ctx& cx, G& g, bool oya_won, bool is_draw) {
  const int np = num_players(g);
  if (g.is_done) return;
  int32_t mx = g.score[0];
  bool tobi = false;
  for (int p = 0; p < np; p++) {
    if (g.score[p] < 0) tobi = true;
    mx = max(mx, g.score[p]);
  }
  if (tobi) {
    end_game(cx, g);
    return;
  }
  int oya = g.oya;
  int32_t ds = g.score[oya];
  bool top = true;
  for (int s = 0; s < np; s++)
    if (!(s == oya || ds > g.score[s] || (ds == g.score[s] && oya <= s))) top = false;
  int gm = g.game_mode;
  bool last = false;
  if (gm == 1 || gm == 4) last = g.round_wind == 0 && oya == np - 1;
  if (gm == 2 || gm == 5) last = g.round_wind == 1 && oya == np - 1;
  const int32_t target = np == 3 ? 40000 : 30000;   // state_3p/mod.rs:1524,1553,1561
  if (oya_won && last && top && ds >= target) {
    end_game(cx, g);
    return;
  }
  int nh = g.honba, no = oya, nw = g.round_wind;
  if (oya_won) {
    nh = nh == 255 ? 255 : nh + 1;
  } else {
    nh = is_draw ? (nh == 255 ? 255 : nh + 1) : 0;
    no = (no + 1) % np;
    if (no == 0) nw += 1;
  }
  bool fin;
  if (gm == 1 || gm == 4) fin = nw >= 1 && (mx >= target || nw > 1);
  else if (gm == 2 || gm == 5) fin = nw >= 2 && (mx >= target || nw > 2);
  else if (gm == 0 || gm == 3) fin = true;
  else fin = nw >= 1;
  if (fin) {
    end_game(cx, g);
    return;
  }
  ev_simple(cx, g, RV_EV_END_KYOKU);
  if (cx.defer_init) {   // the (expensive, rare) shuffle + deal is batched across the warp by the rollout kernel
    g.pending_init[0] = (uint8_t)no;
    g.pending_init[1] = (uint8_t)nw;
    g.pending_init[2] = (uint8_t)nh;
    return;
  }
  init_round(cx, g, no, nw, nh, g.riichi_sticks, nullptr, nullptr);
}
__device__ __noinline__ void run_pending_init(const Ctx& cx, G& g) {
  int no = g.pending_init[0], nw = g.pending_init[1], nh = g.pending_init[2];
  init_round(cx, g, no, nw, nh, g.riichi_sticks, nullptr, nullptr);
}

// ---- the deal, one warp per game -------------------------------------------------------------------------------------
// init_round above is ~18 k serial instructions (ChaCha blocks, the chunked shuffle, dealing, sorting, wait sets, the
// start_kyoku event).  A lock-step launch (one env step for every game, the observation pipeline of BASELINE configs[4])
// waits for whichever thread happens to deal a round — ~190 us, longer than the rest of the step.  Here a whole warp deals
// ONE round: lanes 0-3 compute the four ChaCha blocks, the chunk draws and the swaps stay serial (they are a dependency
// chain) but run out of shared memory, and everything else — clearing the record, writing the wall, dealing the 53 tiles,
// rank-sorting the hands, histograms, wait sets, the event words — is spread over the lanes.  Same record, bit for bit.
// Written as lane loops between warp barriers so that the host compile of these sources (tests/hostsim) runs the same code
// with the lanes in sequence.
#ifdef __CUDA_ARCH__
#define RV_FOR_LANES(l) for (int l = (int)(threadIdx.x & 31), rv_once_ = 1; rv_once_; rv_once_ = 0)
#define RV_WARP_BARRIER() __syncwarp()
#else
#define RV_FOR_LANES(l) for (int l = 0; l < 32; l++)
#define RV_WARP_BARRIER()
#endif
struct DealScratch {            // per warp, in shared memory (on the host: on the stack)
  uint32_t words[64];           // ChaCha12 output of blocks 0..3
  alignas(8) uint8_t w[136];    // the wall being shuffled, then reversed in place
  alignas(8) uint8_t rev[136];
  uint8_t idx[136];             // swap partner of position i
  uint32_t recip[137];          // floor((2^32 - 1) / d): the chunk / n, chunk % n of the shuffle without integer divisions
  uint8_t hand[MAXP][16];       // dealt tiles per seat (13), unsorted
};
// One ChaCha12 block (key, 64-bit counter `ctr`, nonce 0) -> 16 words.
__host__ __device__ inline void chacha12_block(const uint32_t* key, uint64_t ctr, uint32_t* out) {
  uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3],
                     key[4], key[5], key[6], key[7], (uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
  uint32_t s[16];
  for (int i = 0; i < 16; i++) s[i] = in[i];
#define RV_ROTL(v, n) (((v) << (n)) | ((v) >> (32 - (n))))
#define RV_QR2(a, b, c, d)                                                      \
  s[a] += s[b]; s[d] ^= s[a]; s[d] = RV_ROTL(s[d], 16); s[c] += s[d]; s[b] ^= s[c]; s[b] = RV_ROTL(s[b], 12); \
  s[a] += s[b]; s[d] ^= s[a]; s[d] = RV_ROTL(s[d], 8);  s[c] += s[d]; s[b] ^= s[c]; s[b] = RV_ROTL(s[b], 7);
  for (int r = 0; r < 6; r++) {
    RV_QR2(0, 4, 8, 12) RV_QR2(1, 5, 9, 13) RV_QR2(2, 6, 10, 14) RV_QR2(3, 7, 11, 15)
    RV_QR2(0, 5, 10, 15) RV_QR2(1, 6, 11, 12) RV_QR2(2, 7, 8, 13) RV_QR2(3, 4, 9, 14)
  }
#undef RV_QR2
#undef RV_ROTL
  for (int i = 0; i < 16; i++) out[i] = s[i] + in[i];
}
// Every lane of the warp calls this with the same arguments; `S` is the warp's scratch.  Seeded walls only (the rollout
// never passes a custom wall or scores).
__device__ __noinline__ void init_round_coop(const Ctx& cx, G& g, DealScratch& S, int oya, int round_wind, int honba, uint32_t kyotaku) {
  const int np = num_players(g);
  const int wl = np == 3 ? 108 : 136;
  G& cg = cold(g);
  // ---- phase 1: ChaCha blocks (lanes 0-3), identity wall (all lanes), per-seat reset (lanes 4-7), scalars (lane 8)
  uint32_t key[8];
  {
    uint64_t state = splitmix64(g.seed + cg.hand_index);   // rand_core SeedableRng::seed_from_u64: PCG32 (XSH-RR) expansion
    for (int i = 0; i < 8; i++) {
      state = state * 6364136223846793005ull + 11634580027462260723ull;
      uint32_t xs = (uint32_t)(((state >> 18) ^ state) >> 27);
      uint32_t rot = (uint32_t)(state >> 59);
      key[i] = (xs >> rot) | (xs << ((32 - rot) & 31));
    }
  }
  RV_WARP_BARRIER();   // every lane has read hand_index before lane 8 advances it
  RV_FOR_LANES(l) {
    if (l < 4) chacha12_block(key, (uint64_t)l, S.words + 16 * l);
    for (int i = l + 1; i <= 136; i += 32) S.recip[i] = 0xFFFFFFFFu / (uint32_t)i;
    for (int i = l; i < 136; i += 32) {
      int t = i;
      if (np == 3) {                       // the 108-tile set: kinds 1..7 (2m-8m) removed (state_3p/wall.rs:77-79)
        t = i < 4 ? i : i + 28;
        if (i >= 108) t = RV_NONE;
      }
      S.w[i] = (uint8_t)t;
    }
    if (l >= 4 && l < 8) {                 // PlayerState::reset_round (state/player.rs:66-86)
      const int p = l - 4;
      for (int i = 0; i < RV_HAND_CAP; i++) g.hand[p][i] = RV_NONE;
      g.hand_len[p] = 0;
      for (int m = 0; m < 4; m++) {
        for (int k = 0; k < 4; k++) g.meld_tiles[p][m][k] = RV_NONE;
        g.meld_type[p][m] = cg.meld_from[p][m] = cg.meld_called[p][m] = RV_NONE;
      }
      g.n_melds[p] = 0;
      for (int i = 0; i < RV_RIVER_CAP; i++) cg.river[p][i] = RV_NONE;
      g.n_river[p] = 0;
      g.river_tedashi[p] = 0;
      cg.river_riichi[p] = 0;
      cg.riichi_decl_idx[p] = RV_NONE;
      g.flags[p] = (np == 3 && p == 3) ? 0 : RV_F_NAGASHI_ELIGIBLE;
      cg.pao[p][0] = cg.pao[p][1] = RV_NONE;
      g.forbidden[p][0] = g.forbidden[p][1] = RV_NONE;
      cg.score_delta[p] = 0;
      g.n_claims[p] = 0;
      cg.riichi_sutehai[p] = cg.last_tedashi[p] = RV_NONE;
      cg.n_kita[p] = 0;
      g.c_river_kinds[p] = 0;
      g.c_waits[p] = 0;
      if (np == 3 && p == 3) g.score[3] = 0;
    }
    if (l == 8) {
      g.oya = g.kyoku_idx = g.current_player = (uint8_t)oya;
      g.honba = (uint8_t)honba;
      g.riichi_sticks = kyotaku;
      g.round_wind = (uint8_t)round_wind;
      g.is_done = 0;
      g.pending_kan_pid = g.pending_kan_type = g.pending_kan_tile = RV_NONE;
      g.is_rinshan_flag = 0;
      g.rinshan_draw_count = 0;
      g.pending_kan_dora_count = 0;
      g.is_first_turn = 1;
      g.riichi_pending_acceptance = RV_NONE;
      g.turn_count = 0;
      g.last_discard_pid = g.last_discard_tile = RV_NONE;
      g.pending_init[0] = g.pending_init[1] = g.pending_init[2] = RV_NONE;
      g.pending_tail[0] = RV_NONE;
      g.pending_tail[1] = 0;
      g.wall_len = (uint8_t)wl;
      g.n_dora = 1;
      for (int i = 1; i < 5; i++) g.dora_ind[i] = RV_NONE;
      cg.hand_index++;
      cg.kyoku_count++;
    }
  }
  RV_WARP_BARRIER();
  // ---- phase 2 (lane 0): the chunk draws and the swap partners (IncreasingUniform; as wall_from_seed)
  RV_FOR_LANES(l) {
    if (l == 0) {
      int widx = 0;
      uint32_t cur_n = 0, chunk = 0, remaining = 1;
      for (int i = 0; i < wl; i++) {
        uint32_t next_n = cur_n + 1, rem;
        if (remaining == 0) {
          uint32_t product = next_n, current = next_n + 1;
          while (true) {
            uint64_t pr = (uint64_t)product * current;
            if (pr > 0xFFFFFFFFull) break;
            product = (uint32_t)pr;
            current++;
          }
          uint64_t m = (uint64_t)S.words[widx++ & 63] * product;      // random_below: widening multiply + one bias draw
          uint32_t hi = (uint32_t)(m >> 32), lo = (uint32_t)m;
          if (lo > (uint32_t)(0u - product)) {
            uint64_t m2 = (uint64_t)S.words[widx++ & 63] * product;
            if ((uint64_t)lo + (uint32_t)(m2 >> 32) > 0xFFFFFFFFull) hi += 1;
          }
          chunk = hi;
          rem = (current - next_n) - 1;
        } else {
          rem = remaining - 1;
        }
        uint32_t j;
        if (rem == 0) {
          j = chunk;
        } else {
          uint32_t q = rv_umulhi(chunk, S.recip[next_n]);   // q or short of it by at most 2
          uint32_t r = chunk - q * next_n;
          if (r >= next_n) r -= next_n, q++;
          if (r >= next_n) r -= next_n, q++;
          j = r;
          chunk = q;
        }
        remaining = rem;
        cur_n = next_n;
        S.idx[i] = (uint8_t)j;
      }
      for (int i = 0; i < wl; i++) {       // the swaps: a chain through memory, kept in shared memory
        const int j = S.idx[i];
        uint8_t t = S.w[i];
        S.w[i] = S.w[j];
        S.w[j] = t;
      }
    }
  }
  RV_WARP_BARRIER();
  // ---- phase 3: reverse (wall.rs:57-58), dead-wall pad, dealt tiles per seat
  RV_FOR_LANES(l) {
    for (int i = l; i < 136; i += 32) S.rev[i] = i < wl ? S.w[wl - 1 - i] : (uint8_t)RV_NONE;
  }
  RV_WARP_BARRIER();
  RV_FOR_LANES(l) {
    if (l < 17) reinterpret_cast<uint64_t*>(cg.wall)[l] = reinterpret_cast<const uint64_t*>(S.rev)[l];
    // deal (state/mod.rs:1750-1765): 3 rounds of 4 tiles per seat from the dealer, then one each; pop from the back
    for (int d = l; d < 13 * np; d += 32) {
      int idx, k;                           // d-th tile dealt: seat slot idx (from the dealer), position k in that hand
      if (d < 12 * np) {
        const int r = d / (4 * np), within = d % (4 * np);
        idx = within / 4;
        k = 4 * r + (within & 3);
      } else {
        idx = d - 12 * np;
        k = 12;
      }
      S.hand[(idx + oya) % np][k] = S.rev[wl - 1 - d];
    }
    if (l == 31) g.dora_ind[0] = S.rev[np == 3 ? 8 : 4];   // state_3p/wall.rs:104-112
  }
  RV_WARP_BARRIER();
  // ---- phase 4: rank sort (tile ids are distinct), 13 lanes at a time per seat; the dealer's 14th tile goes last
  const int drawn = S.rev[wl - 1 - 13 * np];
  RV_FOR_LANES(l) {
    for (int e = l; e < 13 * np; e += 32) {
      const int p = e / 13, k = e - 13 * p, t = S.hand[p][k];
      int rank = 0;
      for (int j = 0; j < 13; j++) rank += S.hand[p][j] < t ? 1 : 0;
      g.hand[p][rank] = (uint8_t)t;
    }
  }
  RV_WARP_BARRIER();
  // ---- phase 5: histograms + wait sets (one lane per seat), event (lane 4 builds it once the hands are in place)
  RV_FOR_LANES(l) {
    if (l < np) {
      const int p = l;
      uint64_t c[4] = {0, 0, 0, 0};
      uint32_t key5[4] = {0, 0, 0, 0};
      for (int k = 0; k < 13; k++) {
        const int kind = g.hand[p][k] >> 2, su = kind / 9, pos = kind - 9 * su;
        c[su] += 1ull << (4 * pos);
        key5[su] += (uint32_t)pow5(pos);
      }
      for (int k = 0; k < 4; k++) g.c_cnt[p][k] = c[k], g.c_key[p][k] = key5[k];
      g.hand_len[p] = 13;
      if (p != oya) {
        waits_update(cx.T, g, p);
      } else {
        g.hand[p][13] = (uint8_t)drawn;      // the dealer draws right away
        g.hand_len[p] = 14;
        cache_add(g, p, drawn >> 2);
      }
    } else if (l < 4) {                      // sanma: the unused fourth seat keeps empty caches
      for (int k = 0; k < 4; k++) g.c_cnt[l][k] = 0, g.c_key[l][k] = 0;
    }
  }
  RV_WARP_BARRIER();
  RV_FOR_LANES(l) {
    if (l == 0) {
      g.wall_top = (uint8_t)(wl - 13 * np - 1);
      g.drawable_count = (uint8_t)(wl - 13 * np - 14 - 1);
      uint32_t ew[19];
      const int nb = 13 * np, ntw = (nb + 3) / 4, nwords = 2 + np + ntw;
      ew[0] = ev_w0(RV_EV_START_KYOKU, nwords, round_wind % 4, oya);
      ew[1] = (uint32_t)honba | ((uint32_t)g.dora_ind[0] << 8) | ((kyotaku & 0xFFFF) << 16);
      for (int i = 0; i < np; i++) ew[2 + i] = (uint32_t)g.score[i];
      for (int k = 0; k < ntw; k++) {
        uint32_t v = 0;
        for (int b = 0; b < 4; b++) {
          int flat = k * 4 + b;
          int t = flat < nb ? g.hand[flat / 13][flat % 13] : RV_NONE;
          v |= (uint32_t)t << (8 * b);
        }
        ew[2 + np + k] = v;
      }
      ev_push(cx, g, ew, nwords);
      g.phase = RV_WAIT_ACT;
      g.active_mask = (uint8_t)(1u << oya);
      g.drawn_tile = (uint8_t)drawn;
      g.needs_tsumo = 0;
      ev_simple(cx, g, RV_EV_TSUMO, oya, drawn);
    }
  }
  RV_WARP_BARRIER();
}
// the parked deal of `g`, by the whole warp (every lane calls with the same game)
__device__ __forceinline__ void run_pending_init_coop(const Ctx& cx, G& g, DealScratch& S) {
  const int no = g.pending_init[0], nw = g.pending_init[1], nh = g.pending_init[2];
  const uint32_t ky = g.riichi_sticks;
  RV_WARP_BARRIER();              // every lane has read the parked parameters before lane 8 clears them
  init_round_coop(cx, g, S, no, nw, nh, ky);
}

// is_tenpai of a seat's 13-tile-equivalent hand (hand_evaluator.rs:178-194)
__device__ __forceinline__ bool seat_tenpai(const Ctx& cx, const G& g, int p) { return g.c_waits[p] != 0; }

// state/mod.rs:1846-1968
__device__ __noinline__ void trigger_ryukyoku(const Ctx& cx, G& g, int reason) {
  const int np = num_players(g);
  accept_riichi(cx, g);
  bool tenpai[MAXP] = {false, false, false, false};
  int final_reason = reason;
  int nagashi_mask = 0;
  int oya = g.oya;
  if (reason == RV_RK_EXHAUSTIVE) {
    for (int p = 0; p < np; p++) {
      tenpai[p] = seat_tenpai(cx, g, p);
      if (g.flags[p] & RV_F_NAGASHI_ELIGIBLE) nagashi_mask |= 1 << p;
    }
    if (nagashi_mask) {
      final_reason = RV_RK_NAGASHI;
      for (int w = 0; w < np; w++) {
        if (!(nagashi_mask & (1 << w))) continue;
        bool is_oya = w == oya;   // mangan tsumo: calculate_score(5,30,is_oya,true,0,4) -> oya 4000 / ko 2000
        for (int i = 0; i < np; i++) {
          if (i == w) continue;
          int32_t pay = is_oya ? 4000 : (i == oya ? 4000 : 2000);
          g.score[i] -= pay;
          cold(g).score_delta[i] -= pay;
          g.score[w] += pay;
          cold(g).score_delta[w] += pay;
        }
      }
    } else {
      int ntp = 0;
      for (int i = 0; i < np; i++) ntp += tenpai[i];
      if (ntp > 0 && ntp < np) {
        const int32_t pool = np == 3 ? 2000 : 3000;   // state_3p/game_mode.rs:39-41
        int32_t pk = pool / ntp, pn = pool / (np - ntp);
        for (int i = 0; i < np; i++) {
          int32_t d = tenpai[i] ? pk : -pn;
          g.score[i] += d;
          cold(g).score_delta[i] = d;
        }
      }
    }
  } else if (reason >= RV_RK_ILLEGAL_BASE && reason - RV_RK_ILLEGAL_BASE < np) {
    int pid = reason - RV_RK_ILLEGAL_BASE;
    for (int i = 0; i < np; i++) {
      int32_t d;
      if (pid == oya) d = (i == pid) ? -4000 * (np - 1) : 4000;
      else d = (i == pid) ? -(4000 + 2000 * (np - 2)) : (i == oya ? 4000 : 2000);
      g.score[i] += d;
      cold(g).score_delta[i] = d;
    }
  }
  bool renchan = final_reason == RV_RK_EXHAUSTIVE ? tenpai[oya]
               : final_reason == RV_RK_NAGASHI ? ((nagashi_mask >> oya) & 1) != 0 : true;
  uint32_t w[5] = {ev_w0(RV_EV_RYUKYOKU, 1 + np, final_reason, 0), (uint32_t)cold(g).score_delta[0], (uint32_t)cold(g).score_delta[1],
                   (uint32_t)cold(g).score_delta[2], (uint32_t)cold(g).score_delta[3]};
  ev_push(cx, g, w, 1 + np);
  next_round(cx, g, renchan, true);
}

// state/mod.rs:1970-2019
__device__ __noinline__ bool check_abortive_draw(const Ctx& cx, G& g) {
  const int np = num_players(g);
  bool turns_ok = np == 4;   // sufuurenta and suucha riichi do not exist in 3P (state_3p/mod.rs:1860-1888)
  for (int p = 0; p < np; p++)
    if (g.n_river[p] != 1) turns_ok = false;
  if (turns_ok && all_meldless(g)) {
    int first = cold(g).river[0][0] >> 2;
    if (first >= 27 && first <= 30 && (cold(g).river[1][0] >> 2) == first && (cold(g).river[2][0] >> 2) == first &&
        (cold(g).river[3][0] >> 2) == first) {
      trigger_ryukyoku(cx, g, RV_RK_SUFUURENTA);
      return true;
    }
  }
  int kans = 0, first_owner = -1;
  bool same = true;
  for (int p = 0; p < np; p++)
    for (int m = 0; m < g.n_melds[p]; m++)
      if (g.meld_type[p][m] >= RV_MELD_DAIMINKAN) {
        kans++;
        if (first_owner < 0) first_owner = p;
        else if (p != first_owner) same = false;
      }
  if (kans == 4 && !same) {
    trigger_ryukyoku(cx, g, RV_RK_SUUKANSANSEN);
    return true;
  }
  if (np == 4 && ((g.flags[0] & g.flags[1] & g.flags[2] & g.flags[3]) & RV_F_RIICHI_DECLARED)) {
    trigger_ryukyoku(cx, g, RV_RK_SUUCHA_RIICHI);
    return true;
  }
  return false;
}

// ------------------------------------------------------------------ claims (state/legal_actions.rs:254-508)
__device__ __forceinline__ uint32_t pack_act(int type, int tile, int c0, int c1) {
  return (uint32_t)type | ((uint32_t)(tile & 0xFF) << 8) | ((uint32_t)(c0 & 0xFF) << 16) | ((uint32_t)(c1 & 0xFF) << 24);
}
__device__ __noinline__ void claim_push(G& g, int i, uint32_t a) {
  int n = g.n_claims[i];
  if (n < RV_MAX_CLAIMS) {
    cold(g).claims[i][n] = a;
    g.n_claims[i] = (uint8_t)(n + 1);
  } else {
    g.overflow = 1;
  }
}
// Ron part of gen_claims (legal_actions.rs:262-300): furiten tests, then the yaku evaluation.  Returns `missed_agari`.
__device__ __noinline__ bool claim_ron(const Ctx& cx, G& g, int i, int tile) {
  const int kind = tile >> 2;
  const uint32_t f = g.flags[i];
  const bool riichi = f & RV_F_RIICHI_DECLARED;
  const uint64_t rk = g.c_river_kinds[i], waits = g.c_waits[i];
  bool in_discards = (rk >> kind) & 1;
  bool in_missed = (f & RV_F_MISSED_AGARI_DOUJUN) || (riichi && (f & RV_F_MISSED_AGARI_RIICHI));
  if (in_discards || in_missed) return false;
  bool furiten = (waits & rk) != 0 || (f & (RV_F_MISSED_AGARI_RIICHI | RV_F_MISSED_AGARI_DOUJUN));
  if (furiten) return false;
  uint32_t cond = base_cond(g, i);
  if (g.drawable_count == 0 && !g.is_rinshan_flag) cond |= RV_C_HOUTEI;
  WinRes r = seat_calc(cx, g, i, tile, cond, false, g.honba);
  if (r.is_win) {
    claim_push(g, i, pack_act(RV_RON, tile, RV_NONE, RV_NONE));
    return false;
  }
  return r.has_shape;
}
// Fills cold(g).claims[i]; returns the reference's `missed_agari` flag.
// Fast path: the cached wait mask / histogram answer "nothing to claim" with a few bit tests;
// the hand is only scanned for tile ids when a pon / chi pattern actually exists.
__device__ __noinline__ bool gen_claims(const Ctx& cx, G& g, int i, int pid, int tile) {
  const int np = num_players(g);
  bool missed = false;
  g.n_claims[i] = 0;
  int kind = tile >> 2;
  int hl = g.hand_len[i];
  uint32_t f = g.flags[i];
  bool riichi = f & RV_F_RIICHI_DECLARED;
  // 1. Ron
  uint64_t waits = g.c_waits[i];
  if ((waits >> kind) & 1) missed = claim_ron(cx, g, i, tile);   // hand + tile has a winning shape (rare: kept out of line)
  if (riichi || g.drawable_count == 0 || hl < 3) return missed;
  RV_STAT(6);
  int su = kind / 9, r9 = kind - 9 * su;
  uint64_t sc = g.c_cnt[i][su];
  int n_same = (int)((sc >> (4 * r9)) & 15);
  bool kuikae = rule(g, RV_RULE_KUIKAE_FORBIDDEN);
  // Which neighbours a chi needs (shimocha only, number suits; no chi in 3P: state_3p/legal_actions.rs:386)
  const bool chi_seat = np == 4 && i == (pid + 1) % np && su < 3;
  auto at = [&](int r) -> int { return (r < 0 || r > 8) ? 0 : (int)((sc >> (4 * r)) & 15); };
  int m2 = 0, m1 = 0, p1 = 0, p2 = 0;
  if (chi_seat) m2 = at(r9 - 2), m1 = at(r9 - 1), p1 = at(r9 + 1), p2 = at(r9 + 2);
  const bool chi_any = (m2 && m1) || (m1 && p1) || (p1 && p2);
  if (n_same < 2 && !chi_any) return missed;
  // ONE pass over the hand collects, in hand order, the tile ids of the five kinds involved (kind-2 .. kind+2): a byte
  // queue per kind, packed in a 32-bit word (at most four copies).  The claim lists below are cross products of these.
  uint32_t q[5] = {0, 0, 0, 0, 0};
  int qn[5] = {0, 0, 0, 0, 0};
  #pragma unroll 1
  for (int k = 0; k < hl; k++) {
    const int t = g.hand[i][k];
    const unsigned d = (unsigned)((t >> 2) - kind + 2);
    if (d < 5u) {   // (a neighbour kind across a suit boundary is collected but never used: its pattern is off)
      #pragma unroll
      for (int e = 0; e < 5; e++)
        if (d == (unsigned)e && qn[e] < 4) q[e] |= (uint32_t)t << (8 * qn[e]), qn[e]++;
    }
  }
  // 2. Pon / Daiminkan
  if (n_same >= 2) {
    RV_STAT(7);
    const int cnt = qn[2];
    // kuikae: some tile other than the consumed pair must be discardable (legal_actions.rs:316-338)
    bool ok = kuikae ? (hl - cnt) > 0 : true;
    if (ok)
      #pragma unroll 1
      for (int a = 0; a < cnt; a++)
        #pragma unroll 1
        for (int b = a + 1; b < cnt; b++) claim_push(g, i, pack_act(RV_PON, tile, (q[2] >> (8 * a)) & 0xFF, (q[2] >> (8 * b)) & 0xFF));
    if (cnt >= 3) claim_push(g, i, pack_act(RV_DAIMINKAN, tile, q[2] & 0xFF, (q[2] >> 8) & 0xFF));
  }
  // 3. Chi
  if (chi_any) {
    RV_STAT(8);
    #pragma unroll 1
    for (int pat = 0; pat < 3; pat++) {
      int forb2 = -1;
      if (pat == 0) { if (!(m2 && m1)) continue; if (r9 >= 3) forb2 = kind - 3; }
      else if (pat == 1) { if (!(m1 && p1)) continue; }
      else { if (!(p1 && p2)) continue; if (r9 <= 5) forb2 = kind + 3; }
      // leftover tiles (hand minus c1,c2) need one tile that is neither `kind` nor forb2 (legal_actions.rs:394-431)
      int free_tiles = hl;
      if (kuikae) free_tiles -= n_same + (forb2 >= 0 ? at(forb2 - 9 * su) : 0);
      if (free_tiles - 2 <= 0) continue;
      const uint32_t qa = pat == 0 ? q[0] : pat == 1 ? q[1] : q[3], qb = pat == 0 ? q[1] : pat == 1 ? q[3] : q[4];
      const int na = pat == 0 ? qn[0] : pat == 1 ? qn[1] : qn[3], nb = pat == 0 ? qn[1] : pat == 1 ? qn[3] : qn[4];
      #pragma unroll 1
      for (int a = 0; a < na; a++)
        #pragma unroll 1
        for (int b = 0; b < nb; b++) claim_push(g, i, pack_act(RV_CHI, tile, (qa >> (8 * a)) & 0xFF, (qb >> (8 * b)) & 0xFF));
    }
  }
  return missed;
}

// Expand a packed claim / turn action into the public rv_action (consume lists as the reference builds them)
__device__ __noinline__ rv_action expand_act(const G& g, int seat, uint32_t a) {
  rv_action r;
  r.type = (uint8_t)(a & 0xFF);
  r.tile = (uint8_t)((a >> 8) & 0xFF);
  r.actor = (uint8_t)seat;
  r.n_consume = 0;
  for (int k = 0; k < 4; k++) r.consume[k] = RV_NONE;
  int c0 = (a >> 16) & 0xFF, c1 = (a >> 24) & 0xFF;
  switch (r.type) {
    case RV_PON:
    case RV_CHI:
      r.consume[0] = (uint8_t)min(c0, c1);
      r.consume[1] = (uint8_t)max(c0, c1);
      r.n_consume = 2;
      break;
    case RV_DAIMINKAN: {  // first three matching tiles in hand order, sorted
      int n = 0, kind = r.tile >> 2;
      for (int k = 0; k < g.hand_len[seat] && n < 3; k++)
        if ((g.hand[seat][k] >> 2) == kind) r.consume[n++] = g.hand[seat][k];
      r.n_consume = (uint8_t)n;
      for (int x = 1; x < n; x++)
        for (int y = x; y > 0 && r.consume[y - 1] > r.consume[y]; y--) {
          uint8_t t = r.consume[y]; r.consume[y] = r.consume[y - 1]; r.consume[y - 1] = t;
        }
      break;
    }
    case RV_ANKAN: {
      int lo = (r.tile >> 2) << 2;
      for (int k = 0; k < 4; k++) r.consume[k] = (uint8_t)(lo + k);
      r.n_consume = 4;
      break;
    }
    case RV_KAKAN: {  // consume = the pon's tiles (legal_actions.rs:158-172)
      int kind = r.tile >> 2;
      for (int m = 0; m < g.n_melds[seat]; m++)
        if (g.meld_type[seat][m] == RV_MELD_PON && (g.meld_tiles[seat][m][0] >> 2) == kind) {
          for (int k = 0; k < 3; k++) r.consume[k] = g.meld_tiles[seat][m][k];
          r.n_consume = 3;
          break;
        }
      break;
    }
    default:
      break;
  }
  return r;
}

// ------------------------------------------------------------------ turn actions (state/legal_actions.rs:11-252)
struct TurnInfo {
  bool can_tsumo, can_riichi, riichi_ankan, kyushu;
  uint16_t tenpai_keep;   // bit k: hand minus hand[k] is tenpai (valid when riichi_stage or can_riichi was evaluated)
  uint64_t quads;         // kinds with four tiles in hand
  uint64_t present;       // kinds present in hand
};

// bit k set iff removing hand[k] leaves a tenpai hand (exact per-tile answer; used in riichi_stage)
__device__ __noinline__ uint16_t tenpai_discard_mask(const Ctx& cx, const G& g, int p, bool stop_at_first) {
  RV_STAT(5);
  int hl = g.hand_len[p];
  if (hl - 1 + 3 * g.n_melds[p] != 13) return 0;
  Cnt c = hand_cnt(g, p);
  SuitInfo si;
  hand_info(cx.T, g, p, si);
  uint16_t mask = 0;
  uint64_t done = 0, good = 0;
  for (int k = 0; k < hl; k++) {
    int kind = g.hand[p][k] >> 2;
    if (!((done >> kind) & 1)) {
      done |= 1ull << kind;
      Cnt c2 = c;
      cnt_sub(c2, kind);
      SuitInfo s2 = si;
      int su = kind / 9;
      uint32_t key = g.c_key[p][su] - (uint32_t)pow5(kind - 9 * su);
      uint32_t e = su == 3 ? __ldg(&cx.T.honor_info[clamp_key7(key)]) : __ldg(&cx.T.suit_info[clamp_key9(key)]);
      s2.e[0] = su == 0 ? e : si.e[0];
      s2.e[1] = su == 1 ? e : si.e[1];
      s2.e[2] = su == 2 ? e : si.e[2];
      s2.e[3] = su == 3 ? e : si.e[3];
      if (waits13(c2, s2) != 0) good |= 1ull << kind;
    }
    if ((good >> kind) & 1) {
      mask |= (uint16_t)(1u << k);
      if (stop_at_first) return mask;
    }
  }
  return mask;
}

// "Does some discard leave the hand tenpai?" (legal_actions.rs:112-132) from the four suit entries of the
// 14-tile-equivalent hand: the D-bits say whether a tile can leave a suit so that it becomes M / P / waitM / waitP.
__device__ __noinline__ bool any_tenpai_discard(const Cnt& c, const SuitInfo& si, int hl, int n_melds) {
  if (hl + 3 * n_melds != 14) return false;
  uint32_t M = 0, P = 0, WM = 0, WP = 0, DM = 0, DP = 0, DWM = 0, DWP = 0;
  #pragma unroll
  for (int k = 0; k < 4; k++) {
    uint32_t e = si.e[k];
    M |= (e & 1u) << k;
    P |= ((e >> 1) & 1u) << k;
    WM |= (((e >> 2) & 0x1FFu) != 0 ? 1u : 0u) << k;
    WP |= (((e >> 11) & 0x1FFu) != 0 ? 1u : 0u) << k;
    DM |= ((e >> 20) & 1u) << k;
    DP |= ((e >> 21) & 1u) << k;
    DWM |= ((e >> 22) & 1u) << k;
    DWP |= ((e >> 23) & 1u) << k;
  }
  bool ok = false;
  #pragma unroll
  for (int s = 0; s < 4; s++) {
    uint32_t others = 0xFu & ~(1u << s);
    uint32_t mo = M & others, po = P & others;
    int nM = __popc(mo), nP = __popc(po);
    // the discard suit also holds the wait
    if (((DWP >> s) & 1) && nM == 3) ok = true;
    if (((DWM >> s) & 1) && nP == 1 && nM == 2) ok = true;
    // the discard suit ends complete (M or P); another suit u holds the wait
    #pragma unroll
    for (int u = 0; u < 4; u++) {
      if (u == s) continue;
      uint32_t rest = others & ~(1u << u);
      int rM = __popc(M & rest), rP = __popc(P & rest);
      if ((DM >> s) & 1) {
        if (((WP >> u) & 1) && rM == 2) ok = true;
        if (((WM >> u) & 1) && rM == 1 && rP == 1) ok = true;
      }
      if ((DP >> s) & 1) {
        if (((WM >> u) & 1) && rM == 2) ok = true;
      }
    }
  }
  if (ok || n_melds != 0) return ok;
  // chiitoitsu: 7 pairs | 6 pairs + 2 singles | 5 pairs + triplet + single (all kinds distinct)
  {
    int n1 = 0, n2 = 0, n3 = 0;
    uint64_t bad = 0;
    #pragma unroll
    for (int k = 0; k < 4; k++) {
      uint64_t x = c.s[k];
      uint64_t b0 = x & 0x1111111111111111ull, b1 = (x >> 1) & 0x1111111111111111ull;
      bad |= x & 0xCCCCCCCCCCCCCCCCull;
      n3 += __popcll(b0 & b1);
      n1 += __popcll(b0 & ~b1);
      n2 += __popcll(b1 & ~b0);
    }
    if (bad == 0 && ((n2 == 7 && n1 == 0 && n3 == 0) || (n2 == 6 && n1 == 2 && n3 == 0) || (n2 == 5 && n3 == 1 && n1 == 1)))
      return true;
  }
  // kokushi: at least 13 terminal/honor tiles covering at least 12 kinds
  {
    const uint64_t TM = 0xF0000000Full;
    uint64_t xs[4] = {c.s[0] & TM, c.s[1] & TM, c.s[2] & TM, c.s[3] & 0xFFFFFFFull};
    int tiles = 0, kinds = 0;
    #pragma unroll
    for (int k = 0; k < 4; k++) {
      uint64_t x = xs[k];
      uint64_t nz = (x | (x >> 1) | (x >> 2)) & 0x1111111111111111ull;
      kinds += __popcll(nz);
      uint64_t y = (x & 0x0F0F0F0F0F0F0F0Full) + ((x >> 4) & 0x0F0F0F0F0F0F0F0Full);
      tiles += (int)((y * 0x0101010101010101ull) >> 56);
    }
    if (tiles >= 13 && kinds >= 12) return true;
  }
  return false;
}

__device__ __noinline__ void turn_info(const Ctx& cx, const G& g, int pid, TurnInfo& ti) {
  ti.can_tsumo = ti.can_riichi = ti.riichi_ankan = ti.kyushu = false;
  ti.tenpai_keep = 0;
  ti.quads = 0;
  ti.present = 0;
  uint32_t f = g.flags[pid];
  bool riichi = f & RV_F_RIICHI_DECLARED, stage = f & RV_F_RIICHI_STAGE;
  int drawn = g.drawn_tile;
  int hl = g.hand_len[pid], nm = g.n_melds[pid];
  Cnt c = hand_cnt(g, pid);
  SuitInfo si;
  hand_info(cx.T, g, pid, si);
  if (drawn != RV_NONE && !stage) {
    // HandEvaluator::calc adds the win tile to a 13-tile hand (hand_evaluator.rs:93-96): a record injected through the
    // setters may carry `drawn_tile` without the tile in the hand — the cached wait set answers for that shape
    const bool shape = hl + 3 * nm == 14 ? (standard_agari(si) || chiitoi14(c) || kokushi14(c))
                                         : (hl + 3 * nm == 13 && ((g.c_waits[pid] >> (drawn >> 2)) & 1));
    if (shape) {
      uint32_t cond = base_cond(g, pid) | RV_C_TSUMO;
      if (g.drawable_count == 0 && !g.is_rinshan_flag) cond |= RV_C_HAITEI;
      if (g.is_rinshan_flag) cond |= RV_C_RINSHAN;
      if (g.is_first_turn && g.n_river[pid] == 0) cond |= RV_C_TSUMO_FIRST_TURN;
      WinRes r = seat_calc(cx, g, pid, drawn, cond, false, g.honba);
      ti.can_tsumo = r.is_win && (r.yakuman || r.han >= 1);
    }
  }
  if (stage) {
    ti.tenpai_keep = tenpai_discard_mask(cx, g, pid, false);
  } else if (!riichi) {
    // 3P needs only one drawable tile (state_3p/legal_actions.rs:116)
    if (g.score[pid] >= 1000 && (is_sanma(g) ? g.drawable_count > 0 : g.drawable_count >= 4) && !any_open_meld(g, pid))
      ti.can_riichi = any_tenpai_discard(c, si, hl, nm);
  }
  if (g.drawable_count > 0 && drawn != RV_NONE) {
    if (!riichi && !stage) {
      #pragma unroll
      for (int k = 0; k < 4; k++) {
        uint64_t x = (c.s[k] >> 2) & 0x1111111111111111ull;   // nibble == 4
        while (x) {
          int b = __ffsll((long long)x) - 1;
          x &= x - 1;
          ti.quads |= 1ull << (9 * k + (b >> 2));
        }
      }
    } else if (riichi) {
      int kind = drawn >> 2;
      if (cnt_get(c, kind) == 4 && hl + 3 * nm == 14) {
        Cnt pre = c;
        cnt_sub(pre, kind);
        uint64_t wpre = waits13(cx.T, pre);
        Cnt post = c;
        cnt_sub(post, kind, 4);
        uint64_t wpost = waits13(cx.T, post);
        ti.riichi_ankan = wpre == wpost && wpre != 0;
      }
    }
  }
  ti.present = cnt_present(c);
  if (g.is_first_turn && all_meldless(g) && !stage)
    ti.kyushu = __popcll(ti.present & MASK_TERMINAL_HONOR) >= 9;
}

__device__ __forceinline__ bool discard_forbidden(const G& g, int p, int tile) {
  int k = tile >> 2;
  return (g.forbidden[p][0] != RV_NONE && (g.forbidden[p][0] >> 2) == k) ||
         (g.forbidden[p][1] != RV_NONE && (g.forbidden[p][1] >> 2) == k);
}

// Calls f(packed_action) for every legal turn action in the reference's order; returns the count.
template <class F>
__device__ inline int enum_turn_actions(const G& g, int pid, const TurnInfo& ti, F&& f) {
  int n = 0;
  uint32_t fl = g.flags[pid];
  bool riichi = fl & RV_F_RIICHI_DECLARED, stage = fl & RV_F_RIICHI_STAGE;
  int drawn = g.drawn_tile, hl = g.hand_len[pid];
  if (ti.can_tsumo) { f(pack_act(RV_TSUMO, drawn, RV_NONE, RV_NONE)); n++; }
  if (riichi) {
    if (drawn != RV_NONE) { f(pack_act(RV_DISCARD, drawn, RV_NONE, RV_NONE)); n++; }
  } else if (stage) {
    for (int k = 0; k < hl; k++) {
      int t = g.hand[pid][k];
      if (discard_forbidden(g, pid, t)) continue;
      if ((ti.tenpai_keep >> k) & 1) { f(pack_act(RV_DISCARD, t, RV_NONE, RV_NONE)); n++; }
    }
  } else {
    for (int k = 0; k < hl; k++) {
      int t = g.hand[pid][k];
      if (!discard_forbidden(g, pid, t)) { f(pack_act(RV_DISCARD, t, RV_NONE, RV_NONE)); n++; }
    }
    if (ti.can_riichi) { f(pack_act(RV_RIICHI, RV_NONE, RV_NONE, RV_NONE)); n++; }
  }
  if (g.drawable_count > 0 && drawn != RV_NONE) {
    if (!riichi && !stage) {
      uint64_t q = ti.quads;
      while (q) {
        int kind = __ffsll((long long)q) - 1;
        q &= q - 1;
        f(pack_act(RV_ANKAN, kind * 4, RV_NONE, RV_NONE));
        n++;
      }
      for (int m = 0; m < g.n_melds[pid]; m++)
        if (g.meld_type[pid][m] == RV_MELD_PON) {
          int target = g.meld_tiles[pid][m][0] >> 2;
          if (!((ti.present >> target) & 1)) continue;   // no 4th tile in hand: skip the scan
          for (int k = 0; k < hl; k++)
            if ((g.hand[pid][k] >> 2) == target) { f(pack_act(RV_KAKAN, g.hand[pid][k], RV_NONE, RV_NONE)); n++; }
        }
    } else if (riichi && ti.riichi_ankan) {
      f(pack_act(RV_ANKAN, (drawn >> 2) * 4, RV_NONE, RV_NONE));
      n++;
    }
  }
  if (ti.kyushu) { f(pack_act(RV_KYUSHU_KYUHAI, RV_NONE, RV_NONE, RV_NONE)); n++; }
  // Kita: one action per North tile in hand (state_3p/sanma.rs:146-169)
  if (is_sanma(g) && drawn != RV_NONE && g.drawable_count > 0)
    for (int k = 0; k < hl; k++)
      if ((g.hand[pid][k] >> 2) == 30) { f(pack_act(RV_KITA, g.hand[pid][k], RV_NONE, RV_NONE)); n++; }
  return n;
}

// Legal actions of `pid` as the reference's _get_legal_actions_internal would list them (packed).
// Returns count; out may be nullptr (count only).  `pick` >= 0: stores only that index into *picked.
__device__ __noinline__ int legal_actions(const Ctx& cx, const G& g, int pid, uint32_t* out, int pick, uint32_t* picked) {
  if (g.is_done) return 0;
  if (g.phase == RV_WAIT_ACT) {
    if (pid != g.current_player) return 0;
    TurnInfo ti;
    turn_info(cx, g, pid, ti);
    int idx = 0;
    return enum_turn_actions(g, pid, ti, [&](uint32_t a) {
      if (out && idx < RV_MAX_LEGAL) out[idx] = a;
      if (idx == pick && picked) *picked = a;
      idx++;
    });
  }
  int n = g.n_claims[pid];
  for (int k = 0; k < n; k++) {
    if (out) out[k] = cold(g).claims[pid][k];
    if (k == pick && picked) *picked = cold(g).claims[pid][k];
  }
  uint32_t pass = pack_act(RV_PASS, RV_NONE, RV_NONE, RV_NONE);
  if (out && n < RV_MAX_LEGAL) out[n] = pass;
  if (n == pick && picked) *picked = pass;
  return n + 1;
}

// ------------------------------------------------------------------ settlement helpers
__device__ __noinline__ void register_pao(G& g, int claimer, int tile, int discarder) {  // state/mod.rs:1229-1259
  int tv = tile >> 2;
  int dragons = 0, winds = 0;
  for (int m = 0; m < g.n_melds[claimer]; m++) {
    if (g.meld_type[claimer][m] == RV_MELD_CHI) continue;
    int t = g.meld_tiles[claimer][m][0] >> 2;
    if (t >= 31 && t <= 33) dragons++;
    if (t >= 27 && t <= 30) winds++;
  }
  if (tv >= 31 && tv <= 33) {
    if (dragons == 3) cold(g).pao[claimer][0] = (uint8_t)discarder;
  } else if (tv >= 27 && tv <= 30) {
    if (winds == 4) cold(g).pao[claimer][1] = (uint8_t)discarder;
  }
}
__device__ __forceinline__ int yakuman_val(const G& g, int yid) {
  if (yid == 47 && rule(g, RV_RULE_JUNSEI_CHUUREN_DOUBLE)) return 2;
  if (yid == 48 && rule(g, RV_RULE_SUUANKOU_TANKI_DOUBLE)) return 2;
  if (yid == 49 && rule(g, RV_RULE_KOKUSHI13_DOUBLE)) return 2;
  if (yid == 50 && rule(g, RV_RULE_DAISUUSHII_DOUBLE)) return 2;
  return 1;
}
// state/mod.rs:720-745 / 1009-1034
__device__ __noinline__ void cap_double_yakuman(const G& g, WinRes& r, bool is_oya, bool tsumo, uint32_t honba) {
  const int np = num_players(g);
  if (r.yakuman && r.han > 13) {
    int cap = 0;
    if (((r.yaku_mask >> 47) & 1) && !rule(g, RV_RULE_JUNSEI_CHUUREN_DOUBLE)) cap += 13;
    if (((r.yaku_mask >> 48) & 1) && !rule(g, RV_RULE_SUUANKOU_TANKI_DOUBLE)) cap += 13;
    if (((r.yaku_mask >> 49) & 1) && !rule(g, RV_RULE_KOKUSHI13_DOUBLE)) cap += 13;
    if (((r.yaku_mask >> 50) & 1) && !rule(g, RV_RULE_DAISUUSHII_DOUBLE)) cap += 13;
    if (cap > 0) {
      r.han = max(r.han - cap, 13);
      ScoreRes s = calc_score(r.han, 0, is_oya, tsumo, honba, np);
      r.ron = s.pay_ron;
      r.oya = s.pay_oya;
      r.ko = s.pay_ko;
    }
  }
}
__device__ __noinline__ void ev_hora(const Ctx& cx, G& g, int actor, int target, bool tsumo, const WinRes& r, const int32_t* d,
                               bool with_ura) {
  const int np = num_players(g);
  uint8_t ub[8] = {RV_NONE, RV_NONE, RV_NONE, RV_NONE, RV_NONE, 0, 0, 0};
  int n_ura = with_ura ? ura_indicators(g, ub) : 0;
  for (int i = n_ura; i < 5; i++) ub[i] = RV_NONE;
  ub[5] = r.yakuman ? 1 : 0;
  uint32_t w[10];
  const int nw = 6 + np;   // 4P: 10 words, 3P: 9 words
  w[0] = ev_w0(RV_EV_HORA, nw, actor, target);
  w[1] = (tsumo ? 1u : 0u) | ((uint32_t)n_ura << 8) | ((uint32_t)(r.han & 0xFF) << 16) | ((uint32_t)(r.fu & 0xFF) << 24);
  w[2] = ub[0] | (ub[1] << 8) | (ub[2] << 16) | ((uint32_t)ub[3] << 24);
  w[3] = ub[4] | (ub[5] << 8);
  for (int i = 0; i < np; i++) w[4 + i] = (uint32_t)d[i];
  w[4 + np] = (uint32_t)r.yaku_mask;
  w[5 + np] = (uint32_t)(r.yaku_mask >> 32);
  ev_push(cx, g, w, nw);
}

// ------------------------------------------------------------------ kan (state/mod.rs:1415-1547)
__device__ __noinline__ void resolve_kan(const Ctx& cx, G& g, int pid, const rv_action& act) {
  const int np = num_players(g);
  int c_ev[4] = {RV_NONE, RV_NONE, RV_NONE, RV_NONE};
  for (int k = 0; k < act.n_consume && k < 4; k++) c_ev[k] = act.consume[k];
  if (act.type != RV_KAKAN) {
    for (int k = 0; k < act.n_consume && k < 4; k++) hand_remove_first(g, pid, act.consume[k]);
    int m = g.n_melds[pid];
    if (m < 4) {
      uint8_t tl[4] = {RV_NONE, RV_NONE, RV_NONE, RV_NONE};
      int n = 0;
      for (int k = 0; k < act.n_consume && k < 4; k++) tl[n++] = act.consume[k];
      if (act.type == RV_ANKAN) {
        g.meld_type[pid][m] = RV_MELD_ANKAN;
        cold(g).meld_from[pid][m] = RV_NONE;
        cold(g).meld_called[pid][m] = RV_NONE;
      } else {
        if (n < 4) tl[n++] = g.last_discard_tile;
        for (int x = 1; x < n; x++)
          for (int y = x; y > 0 && tl[y - 1] > tl[y]; y--) { uint8_t t = tl[y]; tl[y] = tl[y - 1]; tl[y - 1] = t; }
        g.meld_type[pid][m] = RV_MELD_DAIMINKAN;
        cold(g).meld_from[pid][m] = g.last_discard_pid;
        cold(g).meld_called[pid][m] = g.last_discard_tile;
      }
      for (int k = 0; k < 4; k++) g.meld_tiles[pid][m][k] = tl[k];
      g.n_melds[pid] = (uint8_t)(m + 1);
    } else {
      g.overflow = 1;
    }
    if (act.type == RV_DAIMINKAN) register_pao(g, pid, g.last_discard_tile, g.last_discard_pid);
  }
  g.is_first_turn = 0;
  for (int p = 0; p < np; p++) g.flags[p] &= ~RV_F_IPPATSU_CYCLE;
  if (g.drawable_count > 0) {
    int t = cold(g).wall[g.rinshan_draw_count];   // Vec::remove(0)
    g.drawable_count--;
    hand_push(g, pid, t);
    g.drawn_tile = (uint8_t)t;
    g.rinshan_draw_count++;
    g.is_rinshan_flag = 1;
    if (act.type == RV_ANKAN) {
      int tile = act.tile != RV_NONE ? act.tile : act.consume[0];
      ev_meld(cx, g, RV_EV_ANKAN, pid, tile, c_ev[0], c_ev[1], c_ev[2], c_ev[3]);
    } else if (act.type == RV_DAIMINKAN) {
      ev_meld(cx, g, RV_EV_DAIMINKAN, pid, g.last_discard_tile, g.last_discard_pid, c_ev[0], c_ev[1], c_ev[2]);
    }
    flush_pending_kan_dora(cx, g);
    if (act.type == RV_ANKAN) reveal_kan_dora(cx, g);
    else g.pending_kan_dora_count++;
    ev_simple(cx, g, RV_EV_TSUMO, pid, t);
    g.phase = RV_WAIT_ACT;
    g.active_mask = (uint8_t)(1u << pid);
  }
  waits_update(cx.T, g, pid);
}

// ------------------------------------------------------------------ discard (state/mod.rs:1317-1413)
// `claim_seats`: seats that may have something to claim or a ron shape to miss (act_fast knows from the caches); the
// others are skipped — gen_claims would leave them with an empty list and missed = false.
__device__ __noinline__ void resolve_discard(const Ctx& cx, G& g, int pid, int tile, bool tsumogiri, uint32_t claim_seats = 0xF) {
  const int np = num_players(g);
  if (np == 3) g.pending_kan_pid = g.pending_kan_type = g.pending_kan_tile = RV_NONE;   // state_3p/mod.rs:1224-1227
  g.is_rinshan_flag = 0;
  g.flags[pid] &= ~RV_F_IPPATSU_CYCLE;
  int nr = g.n_river[pid];
  bool stage = g.flags[pid] & RV_F_RIICHI_STAGE;
  if (nr < RV_RIVER_CAP) {
    cold(g).river[pid][nr] = (uint8_t)tile;
    if (!tsumogiri) g.river_tedashi[pid] |= 1u << nr;
    if (stage) cold(g).river_riichi[pid] |= 1u << nr;
  } else {
    g.overflow = 1;
  }
  g.n_river[pid] = (uint8_t)(nr + 1);
  g.c_river_kinds[pid] |= 1ull << (tile >> 2);
  waits_update(cx.T, g, pid);   // the discarder is back to 13 tiles
  g.last_discard_pid = (uint8_t)pid;
  g.last_discard_tile = (uint8_t)tile;
  g.drawn_tile = RV_NONE;
  if (!tsumogiri) cold(g).last_tedashi[pid] = (uint8_t)tile;
  g.needs_tsumo = 1;
  if (stage) {
    g.flags[pid] |= RV_F_RIICHI_DECLARED;
    if (g.is_first_turn) g.flags[pid] |= RV_F_DOUBLE_RIICHI;
    cold(g).riichi_decl_idx[pid] = (uint8_t)nr;
    g.flags[pid] &= ~RV_F_RIICHI_STAGE;
    g.riichi_pending_acceptance = (uint8_t)pid;
  }
  flush_pending_kan_dora(cx, g);
  ev_simple(cx, g, tsumogiri ? RV_EV_DAHAI_TSUMOGIRI : RV_EV_DAHAI, pid, tile);
  g.flags[pid] &= ~RV_F_MISSED_AGARI_DOUJUN;
  if (!tid_terminal(tile)) g.flags[pid] &= ~RV_F_NAGASHI_ELIGIBLE;
  for (int i = 0; i < np; i++) g.n_claims[i] = 0;
  g.active_mask = 0;
  int claim_mask = 0;
  // seats in increasing order, as the reference walks them; the loop runs over the SET BITS so that the lanes of a warp enter
  // gen_claims together (each game usually has one candidate seat, but not the same one: a loop over 0..np-1 called
  // gen_claims three times with ~2 of 32 lanes active, 5 % of all instructions of a rollout)
  uint32_t todo = claim_seats & ((1u << np) - 1) & ~(1u << pid);
  #pragma unroll 1
  while (todo) {
    const int i = __ffs(todo) - 1;
    todo &= todo - 1;
    bool missed = gen_claims(cx, g, i, pid, tile);
    if (missed) g.flags[i] |= RV_F_MISSED_AGARI_DOUJUN;
    if (g.n_claims[i] > 0) claim_mask |= 1 << i;
  }
  if (claim_mask) {
    g.phase = RV_WAIT_RESPONSE;
    g.active_mask = (uint8_t)claim_mask;
  } else {
    accept_riichi(cx, g);
    if (!check_abortive_draw(cx, g)) {
      g.turn_count++;
      g.current_player = (uint8_t)((pid + 1) % np);
      deal_next(cx, g);
      if (g.turn_count >= (uint32_t)np) g.is_first_turn = 0;
    }
  }
}

// chankan candidates after a kakan (state/mod.rs:588-684) / kokushi-on-ankan (485-547)
// mode 0: kakan (chankan yaku), 1: kokushi-only ron on ankan, 2: ron on a kita tile (no chankan yaku, sanma.rs:66-126)
__device__ __noinline__ int chankan_ronners(const Ctx& cx, G& g, int pid, int tile, int mode) {
  const int np = num_players(g);
  int mask = 0;
  int kind = tile >> 2;
  for (int i = 0; i < np; i++) {
    if (i == pid) continue;
    uint64_t waits = g.c_waits[i];
    if (!((waits >> kind) & 1)) continue;
    uint64_t rk = g.c_river_kinds[i];
    uint32_t f = g.flags[i];
    WinRes r;
    if (mode == 1) {
      if ((rk >> kind) & 1) continue;
      uint32_t cond = RV_C_CHANKAN | ((f & RV_F_RIICHI_DECLARED) ? RV_C_RIICHI : 0);
      r = hand_calc(cx.T, g.hand[i], g.hand_len[i], g.n_melds[i], g.meld_type[i], g.meld_tiles[i], tile, g.dora_ind,
                    g.n_dora, nullptr, 0, cond, (i + np - g.oya) % np, g.round_wind % 4, 0, is_sanma(g), 0);
      if (!(r.is_win && ((r.yaku_mask >> 42) & 1 || (r.yaku_mask >> 49) & 1))) continue;
    } else {
      bool furiten = (waits & rk) != 0 || (f & (RV_F_MISSED_AGARI_RIICHI | RV_F_MISSED_AGARI_DOUJUN));
      if (furiten) continue;
      r = seat_calc(cx, g, i, tile, base_cond(g, i) | (mode == 0 ? RV_C_CHANKAN : 0), false, g.honba, mode == 2);
      if (!(r.is_win && (r.yakuman || r.han >= 1))) continue;
    }
    mask |= 1 << i;
    claim_push(g, i, pack_act(RV_RON, tile, RV_NONE, RV_NONE));   // appended to (possibly stale) current_claims
  }
  return mask;
}

// ------------------------------------------------------------------ step (state/mod.rs:330-1315)
// acts[p].type == RV_NO_ACTION  <=>  key p absent from the reference's HashMap.  No validation here.
// state_3p/sanma.rs:171-204 — rinshan draw after a kita; no new dora
__device__ __noinline__ void resolve_kita_rinshan(const Ctx& cx, G& g, int pid) {
  if (g.drawable_count > 0) {
    flush_pending_kan_dora(cx, g);
    if (g.wall_top <= g.rinshan_draw_count) return;
    int t = cold(g).wall[g.rinshan_draw_count];
    g.drawable_count--;
    hand_push(g, pid, t);
    g.drawn_tile = (uint8_t)t;
    g.rinshan_draw_count++;
    g.is_rinshan_flag = 1;
    ev_simple(cx, g, RV_EV_TSUMO, pid, t);
    g.phase = RV_WAIT_ACT;
    g.active_mask = (uint8_t)(1u << pid);
  }
  waits_update(cx.T, g, pid);
}
// state_3p/sanma.rs:9-144
__device__ __noinline__ void handle_kita(const Ctx& cx, G& g, int pid, const rv_action& act) {
  const int np = num_players(g);
  int tile = -1;
  if (act.tile != RV_NONE && (act.tile >> 2) == 30) {
    tile = act.tile;
  } else {
    for (int k = 0; k < g.hand_len[pid] && tile < 0; k++)
      if ((g.hand[pid][k] >> 2) == 30) tile = g.hand[pid][k];
    if (tile < 0) tile = act.tile != RV_NONE ? act.tile : (act.n_consume ? act.consume[0] : 0);
  }
  hand_remove_first(g, pid, tile);
  cold(g).n_kita[pid]++;
  g.is_first_turn = 0;
  ev_simple(cx, g, RV_EV_KITA, pid, tile);
  flush_pending_kan_dora(cx, g);
  waits_update(cx.T, g, pid);
  int ron_mask = chankan_ronners(cx, g, pid, tile, 2);
  if (ron_mask) {
    g.phase = RV_WAIT_RESPONSE;
    g.active_mask = (uint8_t)ron_mask;
    g.last_discard_pid = (uint8_t)pid;
    g.last_discard_tile = (uint8_t)tile;
    g.pending_kan_pid = (uint8_t)pid;
    g.pending_kan_type = RV_KITA;
    g.pending_kan_tile = (uint8_t)tile;
  } else {
    for (int p = 0; p < np; p++) g.flags[p] &= ~RV_F_IPPATSU_CYCLE;
    resolve_kita_rinshan(cx, g, pid);
  }
}

__device__ __noinline__ void step_apply_act(const Ctx& cx, G& g, const rv_action* acts) {
  const int np = num_players(g);
  int pid = g.current_player;
  if (pid >= np) return;      // nobody to act: `actions.get(&self.current_player)` finds nothing (state/mod.rs:405-407)
  const rv_action& act = acts[pid];
  switch (act.type) {
    case RV_DISCARD: {
      if (act.tile == RV_NONE) break;
      int tile = act.tile;
      bool tsumogiri = false, valid = false;
      if (g.drawn_tile != RV_NONE && g.drawn_tile == tile) { tsumogiri = true; valid = true; }
      if (hand_remove_first(g, pid, tile)) {
        hand_sort(g, pid);
        valid = true;
      }
      if (valid) resolve_discard(cx, g, pid, tile, tsumogiri);
      break;
    }
    case RV_KYUSHU_KYUHAI:
      trigger_ryukyoku(cx, g, RV_RK_KYUSHU);
      break;
    case RV_RIICHI: {
      uint32_t f = g.flags[pid];
      if (g.score[pid] >= 1000 && (np == 3 ? g.drawable_count > 0 : g.drawable_count >= 4) &&
          !(f & (RV_F_RIICHI_DECLARED | RV_F_RIICHI_STAGE))) {
        g.flags[pid] |= RV_F_RIICHI_STAGE;
        ev_simple(cx, g, RV_EV_REACH, pid);
        if (act.tile != RV_NONE) {
          int t = act.tile;
          bool tsumogiri = g.drawn_tile != RV_NONE && g.drawn_tile == t;
          cold(g).riichi_sutehai[pid] = (uint8_t)t;
          if (!tsumogiri) cold(g).last_tedashi[pid] = (uint8_t)t;
          if (hand_remove_first(g, pid, t)) hand_sort(g, pid);
          resolve_discard(cx, g, pid, t, tsumogiri);
        }
      }
      break;
    }
    case RV_ANKAN: {
      int tile = act.tile != RV_NONE ? act.tile : (act.n_consume ? act.consume[0] : 0);
      int ron_mask = 0;
      if (rule(g, RV_RULE_RON_ON_ANKAN_KOKUSHI)) ron_mask = chankan_ronners(cx, g, pid, tile, 1);
      if (ron_mask) {
        g.pending_kan_pid = (uint8_t)pid;
        g.pending_kan_type = RV_ANKAN;
        g.pending_kan_tile = (uint8_t)tile;
        g.phase = RV_WAIT_RESPONSE;
        g.active_mask = (uint8_t)ron_mask;
        g.last_discard_pid = (uint8_t)pid;
        g.last_discard_tile = (uint8_t)tile;
      } else {
        resolve_kan(cx, g, pid, act);
      }
      break;
    }
    case RV_KAKAN: {
      int tile = act.tile != RV_NONE ? act.tile : (act.n_consume ? act.consume[0] : 0);
      hand_remove_first(g, pid, tile);
      for (int m = 0; m < g.n_melds[pid]; m++)
        if (g.meld_type[pid][m] == RV_MELD_PON && (g.meld_tiles[pid][m][0] >> 2) == (tile >> 2)) {
          g.meld_type[pid][m] = RV_MELD_KAKAN;
          uint8_t* tl = g.meld_tiles[pid][m];
          tl[3] = (uint8_t)tile;
          for (int y = 3; y > 0 && tl[y - 1] > tl[y]; y--) { uint8_t t = tl[y]; tl[y] = tl[y - 1]; tl[y - 1] = t; }
          break;
        }
      {
        int c[4] = {RV_NONE, RV_NONE, RV_NONE, RV_NONE};
        for (int k = 0; k < act.n_consume && k < 4; k++) c[k] = act.consume[k];
        ev_meld(cx, g, RV_EV_KAKAN, pid, tile, c[0], c[1], c[2], c[3]);
      }
      flush_pending_kan_dora(cx, g);
      waits_update(cx.T, g, pid);
      int ron_mask = chankan_ronners(cx, g, pid, tile, 0);
      if (ron_mask) {
        g.pending_kan_pid = (uint8_t)pid;
        g.pending_kan_type = RV_KAKAN;
        g.pending_kan_tile = (uint8_t)tile;
        g.phase = RV_WAIT_RESPONSE;
        g.active_mask = (uint8_t)ron_mask;
        g.last_discard_pid = (uint8_t)pid;
        g.last_discard_tile = (uint8_t)tile;
      } else {
        resolve_kan(cx, g, pid, act);
      }
      break;
    }
    case RV_KITA:
      if (np == 3) handle_kita(cx, g, pid, act);
      break;
    case RV_TSUMO: {
      uint32_t cond = base_cond(g, pid) | RV_C_TSUMO;
      if (g.drawable_count == 0 && !g.is_rinshan_flag) cond |= RV_C_HAITEI;
      if (g.is_rinshan_flag) cond |= RV_C_RINSHAN;
      if (g.is_first_turn && all_meldless(g)) cond |= RV_C_TSUMO_FIRST_TURN;
      bool riichi = g.flags[pid] & RV_F_RIICHI_DECLARED;
      int win_tile = g.drawn_tile != RV_NONE ? g.drawn_tile : 0;
      WinRes r = seat_calc(cx, g, pid, win_tile, cond, riichi, g.honba, true);   // kita_count: state_3p/mod.rs:635
      int oya = g.oya;
      cap_double_yakuman(g, r, pid == oya, true, g.honba);
      if (r.is_win) {
        int32_t d[MAXP] = {0, 0, 0, 0};
        int32_t total_win = 0;
        int pao_payer = -1, pao_val = 0, total_val = 0;
        if (r.yakuman) {
          uint64_t m = r.yaku_mask;
          while (m) {
            int y = __ffsll((long long)m) - 1;
            m &= m - 1;
            int v = yakuman_val(g, y);
            total_val += v;
            int liable = y == 37 ? cold(g).pao[pid][0] : y == 50 ? cold(g).pao[pid][1] : RV_NONE;
            if (liable != RV_NONE) { pao_val += v; pao_payer = liable; }
          }
        }
        if (pao_val > 0) {
          int32_t unit = pid == oya ? (np - 1) * 16000 : 16000 + (np - 2) * 8000;   // state_3p/mod.rs:713-721
          int32_t honba_total = (int32_t)g.honba * (np - 1) * 100;
          if (rule(g, RV_RULE_PAO_LIABILITY_ONLY)) {
            int32_t pao_amt = pao_val * unit + honba_total;
            int non_pao = total_val - pao_val;
            d[pao_payer] -= pao_amt;
            total_win += pao_amt;
            if (non_pao > 0)
              for (int i = 0; i < np; i++)
                if (i != pid) {
                  int32_t pay = (pid == oya || i == oya) ? non_pao * 16000 : non_pao * 8000;
                  d[i] -= pay;
                  total_win += pay;
                }
          } else {
            int32_t full = total_val * unit + honba_total;
            d[pao_payer] -= full;
            total_win += full;
          }
        } else {
          for (int i = 0; i < np; i++)
            if (i != pid) {
              int32_t pay = (pid == oya || i != oya) ? (int32_t)r.ko : (int32_t)r.oya;
              d[i] = -pay;
              total_win += pay;
            }
        }
        total_win += (int32_t)(g.riichi_sticks * 1000);
        g.riichi_sticks = 0;
        d[pid] += total_win;
        for (int i = 0; i < np; i++) {
          g.score[i] += d[i];
          cold(g).score_delta[i] = d[i];
        }
        ev_hora(cx, g, pid, pid, true, r, d, riichi);
        next_round(cx, g, pid == oya, false);
      } else {
        g.current_player = (uint8_t)((g.current_player + 1) % np);
        deal_next(cx, g);
      }
      break;
    }
    default:
      break;
  }
}

// Ron settlement of a claim window (state/mod.rs:945-1142): kept out of line, a claim window rarely ends in a win.
__device__ __noinline__ void resp_ron(const Ctx& cx, G& g, int ron_mask) {
  const int np = num_players(g);
  if (np == 4 && __popc(ron_mask) >= np - 1 && rule(g, RV_RULE_SANCHAHO_IS_DRAW)) {   // no sanchaho branch in 3P
    trigger_ryukyoku(cx, g, RV_RK_SANCHAHO);
    return;
  }
  int target = g.last_discard_pid != RV_NONE ? g.last_discard_pid : g.current_player;
  int win_tile = g.last_discard_pid != RV_NONE ? g.last_discard_tile : 0;
  int32_t total[MAXP] = {0, 0, 0, 0};
  bool oya_won = false, deposit_taken = false, honba_taken = false;
  bool is_chankan = g.pending_kan_pid != RV_NONE && g.pending_kan_type != RV_KITA;   // state_3p/mod.rs:896-902
  int oya = g.oya;
  for (int dist = 1; dist < np; dist++) {   // winners sorted by distance from the discarder
    int w = (target + dist) % np;
    if (!((ron_mask >> w) & 1)) continue;
    uint32_t ron_honba = 0;
    if (!honba_taken) { honba_taken = true; ron_honba = g.honba; }
    uint32_t cond = base_cond(g, w);
    if (g.drawable_count == 0 && !g.is_rinshan_flag) cond |= RV_C_HOUTEI;
    if (is_chankan) cond |= RV_C_CHANKAN;
    bool riichi = g.flags[w] & RV_F_RIICHI_DECLARED;
    WinRes r = seat_calc(cx, g, w, win_tile, cond, riichi, ron_honba, true);   // kita_count: state_3p/mod.rs:926
    cap_double_yakuman(g, r, w == oya, false, ron_honba);
    if (r.is_win) {
      int32_t score = (int32_t)r.ron;
      int pao_payer = target;
      int32_t pao_amt = 0;
      if (r.yakuman) {
        bool has_pao = false;
        int total_val = 0, pao_val = 0;
        uint64_t m = r.yaku_mask;
        while (m) {
          int y = __ffsll((long long)m) - 1;
          m &= m - 1;
          int v = yakuman_val(g, y);
          total_val += v;
          int liable = y == 37 ? cold(g).pao[w][0] : y == 50 ? cold(g).pao[w][1] : RV_NONE;
          if (liable != RV_NONE) { has_pao = true; pao_payer = liable; pao_val += v; }
        }
        if (has_pao) {
          int32_t unit = w == oya ? 48000 : 32000;
          int32_t honba_ron = (int32_t)ron_honba * (np - 1) * 100;
          int32_t split = rule(g, RV_RULE_PAO_LIABILITY_ONLY) ? pao_val * unit : total_val * unit;
          pao_amt = split / 2 + honba_ron;
        }
      }
      int32_t td[MAXP] = {0, 0, 0, 0};
      td[w] += score;
      td[pao_payer] -= pao_amt;
      td[target] -= score - pao_amt;
      if (!deposit_taken) {
        td[w] += (int32_t)(g.riichi_sticks * 1000);
        g.riichi_sticks = 0;
        deposit_taken = true;
      }
      for (int i = 0; i < np; i++) total[i] += td[i];
      if (w == oya) oya_won = true;
      ev_hora(cx, g, w, target, false, r, td, riichi);
    }
  }
  for (int i = 0; i < np; i++) {
    g.score[i] += total[i];
    cold(g).score_delta[i] = total[i];
  }
  next_round(cx, g, oya_won, false);
}
__device__ __noinline__ void step_apply_resp(const Ctx& cx, G& g, const rv_action* acts) {
  const int np = num_players(g);
  // missed agari (state/mod.rs:902-917): every seat holding a Ron claim (stale ones included) that did not ron
  // (a Ron claim is either the first entry of a list — gen_claims pushes it before the calls — or was appended to a
  // possibly stale list by chankan_ronners: looking at the two ends is the same as scanning the list)
  uint32_t listed = 0;
  for (int p = 0; p < np; p++) listed |= g.n_claims[p] ? 1u << p : 0u;
  #pragma unroll 1
  while (listed) {        // over the set bits: the lanes of a warp read their (usually single) list together
    const int p = __ffs(listed) - 1;
    listed &= listed - 1;
    const int nc = g.n_claims[p];
    const uint32_t* cl = cold(g).claims[p];
    const uint32_t first = cl[0], last = cl[nc - 1];
    const bool has_ron = (first & 0xFF) == RV_RON || (last & 0xFF) == RV_RON;
    if (has_ron && acts[p].type != RV_RON) {
      g.flags[p] |= RV_F_MISSED_AGARI_DOUJUN;
      if (g.flags[p] & RV_F_RIICHI_DECLARED) g.flags[p] |= RV_F_MISSED_AGARI_RIICHI;
    }
  }
  int ron_mask = 0, call_pid = -1;
  for (int p = 0; p < np; p++) {
    if (!((g.active_mask >> p) & 1)) continue;
    int ty = acts[p].type;
    if (ty == RV_RON) ron_mask |= 1 << p;
    else if (ty == RV_PON || ty == RV_DAIMINKAN || ty == RV_CHI) {
      if (call_pid >= 0) {
        int oty = acts[call_pid].type;
        bool old_pon = oty == RV_PON || oty == RV_DAIMINKAN, new_pon = ty == RV_PON || ty == RV_DAIMINKAN;
        if (!old_pon && new_pon) call_pid = p;
      } else {
        call_pid = p;
      }
    }
  }
  if (ron_mask) {
    resp_ron(cx, g, ron_mask);
  } else if (call_pid >= 0) {
    int claimer = call_pid;
    const rv_action& act = acts[claimer];
    accept_riichi(cx, g);
    g.is_rinshan_flag = 0;
    g.is_first_turn = 0;
    g.flags[claimer] &= ~RV_F_MISSED_AGARI_DOUJUN;
    if (g.last_discard_pid != RV_NONE) g.flags[g.last_discard_pid] &= ~RV_F_NAGASHI_ELIGIBLE;
    for (int p = 0; p < np; p++) g.flags[p] &= ~RV_F_IPPATSU_CYCLE;
    if (act.type == RV_DAIMINKAN) {
      g.current_player = (uint8_t)claimer;
      g.active_mask = (uint8_t)(1u << claimer);
      g.forbidden[claimer][0] = g.forbidden[claimer][1] = RV_NONE;
      resolve_kan(cx, g, claimer, act);
      return;
    }
    for (int k = 0; k < act.n_consume && k < 4; k++) hand_remove_first(g, claimer, act.consume[k]);
    int discarder = g.last_discard_pid, tile = g.last_discard_tile;
    int m = g.n_melds[claimer];
    if (m < 4) {
      uint8_t tl[4] = {RV_NONE, RV_NONE, RV_NONE, RV_NONE};
      int n = 0;
      for (int k = 0; k < act.n_consume && k < 3; k++) tl[n++] = act.consume[k];
      tl[n++] = (uint8_t)tile;
      for (int x = 1; x < n; x++)
        for (int y = x; y > 0 && tl[y - 1] > tl[y]; y--) { uint8_t t = tl[y]; tl[y] = tl[y - 1]; tl[y - 1] = t; }
      for (int k = 0; k < 4; k++) g.meld_tiles[claimer][m][k] = tl[k];
      g.meld_type[claimer][m] = act.type == RV_PON ? RV_MELD_PON : RV_MELD_CHI;
      cold(g).meld_from[claimer][m] = (uint8_t)discarder;
      cold(g).meld_called[claimer][m] = (uint8_t)tile;
      g.n_melds[claimer] = (uint8_t)(m + 1);
    } else {
      g.overflow = 1;
    }
    {
      int c0 = act.n_consume > 0 ? act.consume[0] : RV_NONE, c1 = act.n_consume > 1 ? act.consume[1] : RV_NONE,
          c2 = act.n_consume > 2 ? act.consume[2] : RV_NONE;
      ev_meld(cx, g, act.type == RV_PON ? RV_EV_PON : RV_EV_CHI, claimer, tile, discarder, c0, c1, c2);
    }
    if (act.type == RV_PON) register_pao(g, claimer, tile, discarder);
    g.current_player = (uint8_t)claimer;
    g.phase = RV_WAIT_ACT;
    g.active_mask = (uint8_t)(1u << claimer);
    g.forbidden[claimer][0] = (uint8_t)tile;
    g.forbidden[claimer][1] = RV_NONE;
    if (act.type == RV_CHI && act.n_consume >= 2) {
      int t34 = tile >> 2;
      int a = act.consume[0] >> 2, b = act.consume[1] >> 2;
      if (a > b) { int t = a; a = b; b = t; }
      if (a == t34 + 1 && b == t34 + 2) {
        if (t34 % 9 <= 5) g.forbidden[claimer][1] = (uint8_t)((t34 + 3) * 4);
      } else if (t34 >= 2 && b == t34 - 1 && a == t34 - 2 && t34 % 9 >= 3) {
        g.forbidden[claimer][1] = (uint8_t)((t34 - 3) * 4);
      }
    }
    g.needs_tsumo = 0;
    g.drawn_tile = RV_NONE;
    g.c_waits[claimer] = 0;
  } else {
    for (int i = 0; i < np; i++) g.n_claims[i] = 0;
    g.active_mask = 0;
    if (g.pending_kan_pid != RV_NONE) {
      int pk = g.pending_kan_pid, pty = g.pending_kan_type;
      rv_action a = expand_act(g, pk, pack_act(pty, g.pending_kan_tile, RV_NONE, RV_NONE));
      g.pending_kan_pid = g.pending_kan_type = g.pending_kan_tile = RV_NONE;
      if (pty == RV_KITA) {   // state_3p/mod.rs:1201-1207
        for (int p = 0; p < np; p++) g.flags[p] &= ~RV_F_IPPATSU_CYCLE;
        resolve_kita_rinshan(cx, g, pk);
      } else {
        resolve_kan(cx, g, pk, a);
      }
    } else {
      accept_riichi(cx, g);
      g.turn_count++;
      g.current_player = (uint8_t)((g.current_player + 1) % np);
      deal_next(cx, g);
      if (g.turn_count >= (uint32_t)np) g.is_first_turn = 0;
    }
  }
}

__device__ inline void step_apply(const Ctx& cx, G& g, const rv_action* acts) {
  if (g.phase == RV_WAIT_ACT) step_apply_act(cx, g, acts);
  else step_apply_resp(cx, g, acts);
}

// state/mod.rs:344-392
__device__ inline bool action_matches(const rv_action& l, const rv_action& a) {
  if (l.type != a.type) return false;
  bool tiles_match = l.tile == a.tile;
  bool cons_match = l.n_consume == a.n_consume;
  if (cons_match)
    for (int k = 0; k < l.n_consume && k < 4; k++)
      if (l.consume[k] != a.consume[k]) cons_match = false;
  if (tiles_match) {
    if (cons_match) return true;
    if (a.n_consume == 0 && l.type == RV_KAKAN) return true;
    if (a.n_consume == 0 && (l.type == RV_DISCARD || l.type == RV_RIICHI || l.type == RV_TSUMO || l.type == RV_RON || l.type == RV_PASS))
      return true;
  }
  if (cons_match && (l.type == RV_ANKAN || l.type == RV_KAKAN)) return true;
  if (a.tile == RV_NONE)
    return l.type == RV_TSUMO || l.type == RV_RON || l.type == RV_RIICHI || l.type == RV_KYUSHU_KYUHAI || l.type == RV_KITA;
  return false;
}

#ifdef RV_HOSTSIM_STATS
#define RV_DECLINE(i) (RV_STAT(i), false)
#else
#define RV_DECLINE(i) false
#endif
// keyed random agent shared with the oracle (SURVEY.md §8 d)
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint32_t agent_pick(uint64_t agent_seed, uint64_t game_id, uint32_t step, int seat, uint32_t n) {
  uint64_t k = agent_seed ^ (game_id * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)step << 8) ^ (uint64_t)seat;
  return (uint32_t)(mix64(k) % n);
}

// Action::encode / ActionEncoder::encode_3p (defined in obs.cuh)
__device__ inline int action_id(const rv_action& a);
__device__ inline int action_id_3p(const rv_action& a);
// Layout per seat: three words = ids 0..95; bit 31 of the third word marks "this seat owed an action at this step".
constexpr uint32_t IDS_VALID = 0x80000000u;
__device__ __forceinline__ void ids_reset(const Ctx& cx) {        // start of an env step: no seat acts yet
  #pragma unroll
  for (int k = 0; k < 3 * MAXP; k++) cx.idbits[k] = 0;
}
__device__ __forceinline__ void ids_clear(const Ctx& cx, int seat) {
  cx.idbits[3 * seat] = cx.idbits[3 * seat + 1] = 0;
  cx.idbits[3 * seat + 2] = IDS_VALID;
}
__device__ __forceinline__ void ids_add(const Ctx& cx, const G& g, int seat, uint32_t packed) {
  // the id depends on the type, the tile kind and (chi) the two consumed kinds: all in the packed word — no expand_act.
  // (ankan / kakan: consume_tiles[0] has the kind of the packed tile, whichever physical tiles expand_act picks)
  rv_action a;
  a.type = (uint8_t)(packed & 0xFF);
  a.tile = (uint8_t)((packed >> 8) & 0xFF);
  const bool kan = a.type == RV_ANKAN || a.type == RV_KAKAN;
  a.consume[0] = kan ? a.tile : (uint8_t)((packed >> 16) & 0xFF);
  a.consume[1] = (uint8_t)(packed >> 24);
  a.n_consume = kan ? 1 : (a.type == RV_CHI ? 2 : 0);
  const bool sanma = is_sanma(g);
  const int id = sanma ? action_id_3p(a) : action_id(a);
  if (id >= 0 && id < (sanma ? 60 : 82)) cx.idbits[3 * seat + (id >> 5)] |= 1u << (id & 31);
}

// ---------------------------------------------------------------------------------------------
// Fast path of the ACT phase.  The generic turn code (legal-action enumeration, action expansion, the
// apply switch) is ~90 KB of hot SASS and a long chain of dependent loads; the overwhelmingly common
// turn — "the only legal actions are discards" — is restated here as one compact routine.  It either
// performs the WHOLE env step exactly as random_step_act would, or returns false BEFORE touching the
// state so the game is handed to the generic kernel (agari shape, riichi / kan / kyushu options, a
// declared riichi, a North tile in a sanma hand).  The discard itself is committed here; what FOLLOWS the discard is
//   * inline, when it is the plain case (nobody can claim, ordinary draw, no kan dora business), or
//   * the generic resolve_discard() (claims, first turn, after a call or a kan, last tile) — the very
//     function the generic path runs, entered with the same state.
// A hand row held in two registers: bytes 0..7 in lo, 8..15 in hi.  The fast path only runs on hands of at most 14 tiles
// (hl + 3 nm == 14), so bytes 14 and 15 are RV_NONE padding throughout; rows are 16-byte aligned (offset 192 + 16 seat).
constexpr int ACT_ROW = 14;
struct alignas(16) Row14 {
  uint64_t lo, hi;
};
__device__ __forceinline__ Row14 row_load(const uint8_t* h) {
  static_assert(sizeof(Row14) == 16 && alignof(Row14) == 16, "one 128-bit access per row");
  return *reinterpret_cast<const Row14*>(h);
}
__device__ __forceinline__ void row_store(uint8_t* h, const Row14& r) { *reinterpret_cast<Row14*>(h) = r; }
__device__ __forceinline__ int row_get(const Row14& r, int j) {
  return (int)(((j < 8 ? r.lo : r.hi) >> (8 * (j & 7))) & 0xFF);
}
__device__ __forceinline__ void row_pad(Row14& r, int j) {            // entry j := RV_NONE
  const uint64_t m = 0xFFull << (8 * (j & 7));
  if (j < 8) r.lo |= m;
  else r.hi |= m;
}
__device__ __forceinline__ void row_drop(Row14& r, int k) {           // remove entry k, shift the rest down, pad with RV_NONE
  const uint64_t below = (1ull << (8 * (k & 7))) - 1;
  if (k < 8) {
    r.lo = (r.lo & below) | ((r.lo >> 8) & ~below);
    r.lo = (r.lo & 0x00FFFFFFFFFFFFFFull) | (r.hi << 56);
    r.hi = (r.hi >> 8) | 0xFF00000000000000ull;
  } else {
    r.hi = (r.hi & below) | ((r.hi >> 8) & ~below) | 0xFF00000000000000ull;
  }
}
__device__ __forceinline__ void row_insert(Row14& r, int pos, int v) {   // entries pos.. move up by one (row holds <= 13 entries)
  const uint64_t below = (1ull << (8 * (pos & 7))) - 1, val = (uint64_t)(uint32_t)v << (8 * (pos & 7));
  if (pos < 8) {
    const uint64_t carry = r.lo >> 56;
    r.lo = (r.lo & below) | val | ((r.lo & ~below) << 8);
    r.hi = (r.hi << 8) | carry;
  } else {
    r.hi = (r.hi & below) | val | ((r.hi & ~below) << 8);
  }
}
// What follows a committed discard that is not the plain case: the generic _resolve_discard — run here, or parked for
// the TAIL class of the rollout scheduler (so the fast kernel does not carry the claim / abortive-draw / round-end code).
__device__ __forceinline__ bool act_fast_tail(const Ctx& cx, G& g, int pid, int tile, bool tsumogiri, uint32_t claim_seats) {
  RV_STAT(11);
  if (cx.defer_tail) {
    g.pending_tail[0] = (uint8_t)tile;
    g.pending_tail[1] = (uint8_t)((tsumogiri ? 1 : 0) | (claim_seats << 1));
  } else {
    resolve_discard(cx, g, pid, tile, tsumogiri, claim_seats);
  }
  return true;
}
__device__ __noinline__ void run_pending_tail(const Ctx& cx, G& g) {
  const int tile = g.pending_tail[0];
  const bool tsumogiri = (g.pending_tail[1] & 1) != 0;
  const uint32_t claim_seats = g.pending_tail[1] >> 1;
  g.pending_tail[0] = RV_NONE;
  g.pending_tail[1] = 0;
  resolve_discard(cx, g, g.current_player, tile, tsumogiri, claim_seats);
}
// NPC: seat count known at compile time (the rollout kernel is instantiated per variant so that the 4P hot path carries no
// sanma arithmetic), or 0 = read it from the record.
template <bool IDS = false, int NPC = 0>
__device__ __forceinline__ bool act_fast(const Ctx& cx, G& g, uint64_t agent_seed, uint64_t game_id) {
  const int np = NPC ? NPC : num_players(g);
  RV_STAT(10);
  const int pid = g.current_player;
  if (pid >= np) return RV_DECLINE(18);   // event-driven record between turns (current_player = RV_NONE): nobody acts
  const int drawn = g.drawn_tile;
  const int hl = g.hand_len[pid], nm = g.n_melds[pid];
  const bool sanma = np == 3;
  if (g.flags[pid] & (RV_F_RIICHI_DECLARED | RV_F_RIICHI_STAGE)) return RV_DECLINE(19);
  if (hl + 3 * nm != 14) return RV_DECLINE(21);
  const uint64_t c0 = g.c_cnt[pid][0], c1 = g.c_cnt[pid][1], c2 = g.c_cnt[pid][2], c3 = g.c_cnt[pid][3];
  if (sanma && ((c3 >> 12) & 15)) return RV_DECLINE(16);   // a North tile in hand: Kita may be legal (state_3p/sanma.rs:146-169)
  if ((c0 | c1 | c2 | c3) & 0x4444444444444444ull) return RV_DECLINE(22);             // four of a kind: ankan may be legal
  const uint32_t e0 = __ldg(&cx.T.suit_info[clamp_key9(g.c_key[pid][0])]), e1 = __ldg(&cx.T.suit_info[clamp_key9(g.c_key[pid][1])]),
                 e2 = __ldg(&cx.T.suit_info[clamp_key9(g.c_key[pid][2])]), e3 = __ldg(&cx.T.honor_info[clamp_key7(g.c_key[pid][3])]);
  // suits that are complete (M or P).  agari needs 4 of them, "some discard leaves tenpai" needs >= 2
  const int complete = ((e0 | (e0 >> 1)) & 1) + ((e1 | (e1 >> 1)) & 1) + ((e2 | (e2 >> 1)) & 1) + ((e3 | (e3 >> 1)) & 1);
  bool open_meld = false;
  #pragma unroll 1
  for (int m = 0; m < nm; m++) {
    int ty = g.meld_type[pid][m];
    if (ty != RV_MELD_ANKAN) open_meld = true;
    if (ty == RV_MELD_PON) {                                                  // kakan possible?
      int k = g.meld_tiles[pid][m][0] >> 2, su = k / 9;
      uint64_t x = su == 0 ? c0 : su == 1 ? c1 : su == 2 ? c2 : c3;
      if ((x >> (4 * (k - 9 * su))) & 15) return RV_DECLINE(23);
    }
  }
  const bool first = g.is_first_turn;
  if (nm == 0) {
    // chiitoitsu / kokushi shapes (agari now, or tenpai after a discard) need >= 5 pairs / >= 12 terminal kinds;
    // kyushu kyuhai (first turn) needs >= 9 terminal kinds
    const uint64_t L = 0x1111111111111111ull;
    int pairs = __popcll(((c0 >> 1) | (c0 >> 2)) & L) + __popcll(((c1 >> 1) | (c1 >> 2)) & L) +
                __popcll(((c2 >> 1) | (c2 >> 2)) & L) + __popcll(((c3 >> 1) | (c3 >> 2)) & L);
    if (pairs >= 5) return RV_DECLINE(24);
    const uint64_t TM = 0xF0000000Full;
    uint64_t t0 = c0 & TM, t1 = c1 & TM, t2 = c2 & TM;
    int tk = __popcll((t0 | (t0 >> 1) | (t0 >> 2)) & L) + __popcll((t1 | (t1 >> 1) | (t1 >> 2)) & L) +
             __popcll((t2 | (t2 >> 1) | (t2 >> 2)) & L) + __popcll((c3 | (c3 >> 1) | (c3 >> 2)) & L & 0xFFFFFFFull);
    if (tk >= (first ? 9 : 12)) return RV_DECLINE(25);
  }
  if (complete == 4) return RV_DECLINE(26);                                            // standard agari shape: tsumo evaluation
  if (complete >= 2 && !open_meld && g.score[pid] >= 1000 && (sanma ? g.drawable_count > 0 : g.drawable_count >= 4))
    return RV_DECLINE(27);   // riichi may be legal (3P: state_3p/legal_actions.rs:116)
  // ---- the action: every hand tile that kuikae does not forbid is a legal discard, nothing else is legal
  uint8_t* const hrow = g.hand[pid];
  Row14 hx = row_load(hrow);
  const uint32_t sc = g.step_count;
  int pick;
  const int f0 = g.forbidden[pid][0], f1 = g.forbidden[pid][1];
  uint32_t legal_rows = (1u << hl) - 1;            // hand rows that are legal discards
  if ((f0 & f1) == RV_NONE) {
    pick = (int)agent_pick(agent_seed, game_id, sc, pid, (uint32_t)hl);
  } else {
    const int k0 = f0 == RV_NONE ? 99 : f0 >> 2, k1 = f1 == RV_NONE ? 99 : f1 >> 2;
    uint32_t ok = 0;
    #pragma unroll
    for (int j = 0; j < ACT_ROW; j++) {
      int k = row_get(hx, j) >> 2;
      if (j < hl && k != k0 && k != k1) ok |= 1u << j;
    }
    const int n_ok = __popc(ok);
    if (n_ok == 0) return RV_DECLINE(20);
    legal_rows = ok;
    int r = (int)agent_pick(agent_seed, game_id, sc, pid, (uint32_t)n_ok);
    for (; r > 0; r--) ok &= ok - 1;              // r-th set bit
    pick = __ffs(ok) - 1;
  }
  const int tile = row_get(hx, pick);
  const int kind = tile >> 2, ksu = kind / 9, kr = kind - 9 * ksu;
  // ---- would anybody be offered a claim?  (legal_actions.rs:254-508, decided from the caches)
  uint32_t claim_seats = 0;                                                   // seats gen_claims has to look at
  #pragma unroll 1
  for (int d = 1; d < np; d++) {
    int i = sanma ? (pid + d) % 3 : (pid + d) & 3;
    if ((g.c_waits[i] >> kind) & 1) claim_seats |= 1u << i;                   // ron shape
    if (g.flags[i] & RV_F_RIICHI_DECLARED) continue;
    if (g.hand_len[i] < 3) continue;
    uint64_t x = g.c_cnt[i][ksu];
    if (((x >> (4 * kr)) & 15) >= 2) claim_seats |= 1u << i;                  // pon / daiminkan
    if (d == 1 && ksu < 3 && !sanma) {                                        // chi (shimocha; none in sanma)
      uint64_t y = x << 8;                                                    // nibble kr+2 of y == nibble kr of x
      int m2 = (y >> (4 * kr)) & 15, m1 = (y >> (4 * kr + 4)) & 15, p1 = (y >> (4 * kr + 12)) & 15, p2 = (y >> (4 * kr + 16)) & 15;
      if (kr >= 8) p1 = 0;
      if (kr >= 7) p2 = 0;
      if ((m2 && m1) || (m1 && p1) || (p1 && p2)) claim_seats |= 1u << i;
    }
  }
  const bool claims = claim_seats != 0;
  if (IDS) {
    // the legal list is exactly the discards of `legal_rows`: ids = their tile kinds (sanma: compact columns, action.rs:262-279)
    uint64_t kinds = 0;
    #pragma unroll
    for (int j = 0; j < ACT_ROW; j++)
      if ((legal_rows >> j) & 1) kinds |= 1ull << (row_get(hx, j) >> 2);
    if (sanma) kinds = (kinds & 1) | ((kinds >> 7) & ~1ull);
    ids_reset(cx);
    cx.idbits[3 * pid] = (uint32_t)kinds;
    cx.idbits[3 * pid + 1] = (uint32_t)(kinds >> 32);
    cx.idbits[3 * pid + 2] = IDS_VALID;
  }
  // ================= commit: nothing below can fail =================
  RV_STAT(9);
  g.step_count = sc + 1;
  const bool tsumogiri = drawn != RV_NONE && tile == drawn;
  // hand: the drawn tile (if any) sits last, the rest is sorted.  Remove the pick, re-insert the drawn tile in order.
  if (g.is_rinshan_flag) {
    // after a kan the tile drawn before it may still sit unsorted in the row: generic remove + full sort
    hand_remove_first(g, pid, tile);
    hand_sort(g, pid);
    return act_fast_tail(cx, g, pid, tile, tsumogiri, claim_seats);
  }
  if (tsumogiri) {
    row_pad(hx, hl - 1);
  } else if (drawn == RV_NONE) {
    row_drop(hx, pick);
  } else {
    row_pad(hx, hl - 1);                           // lift the drawn tile out ...
    row_drop(hx, pick);                            // ... drop the pick (pick < hl - 1 here) ...
    int pos = 0;                                   // ... and count the tiles below the drawn one (pads are 0xFF)
    #pragma unroll
    for (int j = 0; j < ACT_ROW - 2; j++) pos += row_get(hx, j) < drawn ? 1 : 0;
    row_insert(hx, pos, drawn);
  }
  row_store(hrow, hx);
  g.hand_len[pid] = (uint8_t)(hl - 1);
  cache_sub(g, pid, kind);
#ifdef RV_DEBUG_CHECK
  {
    uint64_t cc[4] = {0, 0, 0, 0};
    for (int j = 0; j < hl - 1; j++) {
      int k = hrow[j] >> 2, su = k / 9;
      cc[su] += 1ull << (4 * (k - 9 * su));
    }
    if (cc[0] != g.c_cnt[pid][0] || cc[1] != g.c_cnt[pid][1] || cc[2] != g.c_cnt[pid][2] || cc[3] != g.c_cnt[pid][3])
      printf("act_fast mismatch: pid %d hl %d nm %d drawn %d pick %d tile %d tsumogiri %d f0 %d f1 %d | row %d %d %d %d %d %d %d %d %d %d %d %d %d %d | cnt %llx %llx %llx %llx vs %llx %llx %llx %llx\n",
             pid, hl, nm, drawn, pick, tile, (int)tsumogiri, f0, f1, hrow[0], hrow[1], hrow[2], hrow[3], hrow[4], hrow[5], hrow[6], hrow[7],
             hrow[8], hrow[9], hrow[10], hrow[11], hrow[12], hrow[13], cc[0], cc[1], cc[2], cc[3], g.c_cnt[pid][0], g.c_cnt[pid][1],
             g.c_cnt[pid][2], g.c_cnt[pid][3]);
  }
#endif
  // (first turn: only a wind discard can complete a sufuurenta; 4-kan / 4-riichi draws need a kan or a riichi discard)
  // four kans (every kan draws a rinshan tile in 4P) can end the round at this discard; a pending kan dora is flipped by it
  if (claims || (first && kind >= 27 && kind <= 30) || g.drawable_count == 0 || g.n_dora >= 5 || g.rinshan_draw_count >= 4 ||
      g.pending_kan_dora_count != 0) {
    return act_fast_tail(cx, g, pid, tile, tsumogiri, claim_seats);
  }
  // _resolve_discard (state/mod.rs:1317-1413), no-claims branch
  if (sanma) g.pending_kan_pid = g.pending_kan_type = g.pending_kan_tile = RV_NONE;   // state_3p/mod.rs:1224-1227
  g.flags[pid] &= ~(RV_F_IPPATSU_CYCLE | RV_F_MISSED_AGARI_DOUJUN);
  if (!tid_terminal(tile)) g.flags[pid] &= ~RV_F_NAGASHI_ELIGIBLE;
  const int nr = g.n_river[pid];
  if (nr < RV_RIVER_CAP) {
    cold(g).river[pid][nr] = (uint8_t)tile;
    if (!tsumogiri) g.river_tedashi[pid] |= 1u << nr;
  } else {
    g.overflow = 1;
  }
  g.n_river[pid] = (uint8_t)(nr + 1);
  g.c_river_kinds[pid] |= 1ull << kind;
  {
    // waits of the 13-tile hand: only the discarded tile's suit entry changed
    const uint32_t key = g.c_key[pid][ksu];
    const uint32_t en = ksu == 3 ? __ldg(&cx.T.honor_info[clamp_key7(key)]) : __ldg(&cx.T.suit_info[clamp_key9(key)]);
    SuitInfo si;
    si.e[0] = ksu == 0 ? en : e0, si.e[1] = ksu == 1 ? en : e1, si.e[2] = ksu == 2 ? en : e2, si.e[3] = ksu == 3 ? en : e3;
    // standard-form waits only: a seven-pairs or kokushi wait after this discard needs >= 6 pairs / >= 12 terminal kinds in
    // the 14 tiles, and those hands were declined above
    g.c_waits[pid] = waits13_std(si);
  }
  g.last_discard_pid = (uint8_t)pid;
  g.last_discard_tile = (uint8_t)tile;
  if (!tsumogiri) cold(g).last_tedashi[pid] = (uint8_t)tile;
  g.n_claims[0] = g.n_claims[1] = g.n_claims[2] = g.n_claims[3] = 0;
  // the next seat draws (_deal_next, state/mod.rs:1569-1593)
  g.turn_count++;
  if (first && g.turn_count >= (uint32_t)np) g.is_first_turn = 0;
  const int nxt = sanma ? (pid + 1) % 3 : (pid + 1) & 3;
  g.current_player = (uint8_t)nxt;
  const int t2 = cold(g).wall[g.wall_top - 1];
  g.wall_top--;
  g.drawable_count--;
  {
    const int n2 = g.hand_len[nxt];
    g.hand[nxt][n2] = (uint8_t)t2;
    g.hand_len[nxt] = (uint8_t)(n2 + 1);
    cache_add(g, nxt, t2 >> 2);
  }
  g.c_waits[nxt] = 0;
  g.drawn_tile = (uint8_t)t2;
  g.needs_tsumo = 0;
  g.phase = RV_WAIT_ACT;
  g.active_mask = (uint8_t)(1u << nxt);
  g.forbidden[nxt][0] = g.forbidden[nxt][1] = RV_NONE;
  // events: dahai, tsumo
  uint32_t w[2] = {ev_w0(tsumogiri ? RV_EV_DAHAI_TSUMOGIRI : RV_EV_DAHAI, 1, pid, tile), ev_w0(RV_EV_TSUMO, 1, nxt, t2)};
  uint64_t hsh = g.ev_hash;
  uint32_t base = g.ev_words;
  hsh = (hsh ^ w[0]) * 0x100000001b3ull;
  hsh = (hsh ^ w[1]) * 0x100000001b3ull;
  if (cx.log) {
    if (base < cx.log_cap) cx.log[base] = w[0];
    if (base + 1 < cx.log_cap) cx.log[base + 1] = w[1];
  }
  g.ev_hash = hsh;
  g.ev_words = base + 2;
  g.ev_count += 2;
  return true;
}

// One env step with the on-device random agent, split by phase so that phase-sorted kernels only carry
// the code of their own phase.
template <bool IDS = false>
__device__ __noinline__ void random_step_act(const Ctx& cx, G& g, uint64_t agent_seed, uint64_t game_id) {
  const int np = num_players(g);
  rv_action acts[MAXP];
  for (int p = 0; p < np; p++) acts[p].type = RV_NO_ACTION;
  uint32_t sc = g.step_count;
  RV_STAT(3);
  RV_STAT(4);
  int pid = g.current_player;
  if (pid >= np) {            // nobody to act (an event-driven record between turns): the step is a no-op, as in the reference
    g.step_count = sc + 1;
    return;
  }
  if (IDS) {
    ids_reset(cx);
    ids_clear(cx, pid);
  }
  TurnInfo ti;
  turn_info(cx, g, pid, ti);
  uint32_t fl = g.flags[pid];
  bool plain = !(fl & (RV_F_RIICHI_DECLARED | RV_F_RIICHI_STAGE)) && g.forbidden[pid][0] == RV_NONE &&
               g.forbidden[pid][1] == RV_NONE && !(is_sanma(g) && ((ti.present >> 30) & 1));
  bool has_pon = false;
  for (int m = 0; m < g.n_melds[pid]; m++) has_pon |= g.meld_type[pid][m] == RV_MELD_PON;
  if (plain && !has_pon) {
    // common case, closed form: [Tsumo] + every hand tile + [Riichi] + ankans + [Kyushu]
    int hl = g.hand_len[pid];
    int nq = (g.drawable_count > 0 && g.drawn_tile != RV_NONE) ? __popcll(ti.quads) : 0;
    int n = (ti.can_tsumo ? 1 : 0) + hl + (ti.can_riichi ? 1 : 0) + nq + (ti.kyushu ? 1 : 0);
    int pick = (int)agent_pick(agent_seed, game_id, sc, pid, (uint32_t)n);
    uint32_t chosen;
    int k = pick - (ti.can_tsumo ? 1 : 0);
    if (k < 0) chosen = pack_act(RV_TSUMO, g.drawn_tile, RV_NONE, RV_NONE);
    else if (k < hl) chosen = pack_act(RV_DISCARD, g.hand[pid][k], RV_NONE, RV_NONE);
    else {
      k -= hl;
      if (ti.can_riichi && k == 0) chosen = pack_act(RV_RIICHI, RV_NONE, RV_NONE, RV_NONE);
      else {
        k -= ti.can_riichi ? 1 : 0;
        if (k < nq) {
          uint64_t q = ti.quads;
          for (int z = 0; z < k; z++) q &= q - 1;
          chosen = pack_act(RV_ANKAN, (__ffsll((long long)q) - 1) * 4, RV_NONE, RV_NONE);
        } else {
          chosen = pack_act(RV_KYUSHU_KYUHAI, RV_NONE, RV_NONE, RV_NONE);
        }
      }
    }
    acts[pid] = expand_act(g, pid, chosen);
    if (IDS) {
      if (ti.can_tsumo) ids_add(cx, g, pid, pack_act(RV_TSUMO, g.drawn_tile, RV_NONE, RV_NONE));
      for (int j = 0; j < hl; j++) ids_add(cx, g, pid, pack_act(RV_DISCARD, g.hand[pid][j], RV_NONE, RV_NONE));
      if (ti.can_riichi) ids_add(cx, g, pid, pack_act(RV_RIICHI, RV_NONE, RV_NONE, RV_NONE));
      uint64_t q = ti.quads;
      for (int z = 0; z < nq; z++, q &= q - 1) ids_add(cx, g, pid, pack_act(RV_ANKAN, (__ffsll((long long)q) - 1) * 4, RV_NONE, RV_NONE));
      if (ti.kyushu) ids_add(cx, g, pid, pack_act(RV_KYUSHU_KYUHAI, RV_NONE, RV_NONE, RV_NONE));
    }
  } else {
    int n = enum_turn_actions(g, pid, ti, [](uint32_t) {});
    if (n == 0) {
      // Dead end of the reference (3P: riichi declared, then kita + rinshan draws leave no tenpai-keeping discard;
      // RandomAgent's random.choice([]) raises).  The rollout retires the game: done + stalled flag (overflow bit 1).
      g.step_count = sc + 1;
      g.is_done = 1;
      g.overflow |= 2;
      return;
    }
    if (n > 0) {
      int pick = (int)agent_pick(agent_seed, game_id, sc, pid, (uint32_t)n);
      uint32_t chosen = 0;
      int idx = 0;
      enum_turn_actions(g, pid, ti, [&](uint32_t a) {
        if (idx == pick) chosen = a;
        idx++;
        if (IDS) ids_add(cx, g, pid, a);
      });
      acts[pid] = expand_act(g, pid, chosen);
    }
  }
  g.step_count = sc + 1;
  step_apply_act(cx, g, acts);
}
template <bool IDS = false>
__device__ __noinline__ void random_step_resp(const Ctx& cx, G& g, uint64_t agent_seed, uint64_t game_id) {
  const int np = num_players(g);
  rv_action acts[MAXP];
  for (int p = 0; p < np; p++) acts[p].type = RV_NO_ACTION;
  uint32_t sc = g.step_count;
  RV_STAT(3);
  if (IDS) ids_reset(cx);
  uint32_t todo = g.active_mask & ((1u << np) - 1);   // set bits in increasing order: the lanes of a warp answer their first claim together
  #pragma unroll 1
  while (todo) {
    const int p = __ffs(todo) - 1;
    todo &= todo - 1;
    if (IDS) {
      ids_clear(cx, p);
      for (int k = 0; k < g.n_claims[p]; k++) ids_add(cx, g, p, cold(g).claims[p][k]);
      ids_add(cx, g, p, pack_act(RV_PASS, RV_NONE, RV_NONE, RV_NONE));
    }
    int n = g.n_claims[p] + 1;
    int pick = (int)agent_pick(agent_seed, game_id, sc, p, (uint32_t)n);
    uint32_t chosen = pick < g.n_claims[p] ? cold(g).claims[p][pick] : pack_act(RV_PASS, RV_NONE, RV_NONE, RV_NONE);
    acts[p] = expand_act(g, p, chosen);
  }
  g.step_count = sc + 1;
  step_apply_resp(cx, g, acts);
}
template <bool IDS = false>
__device__ inline void random_step(const Ctx& cx, G& g, uint64_t agent_seed, uint64_t game_id) {
  if (g.phase == RV_WAIT_ACT) {
    if (!act_fast<IDS>(cx, g, agent_seed, game_id)) random_step_act<IDS>(cx, g, agent_seed, game_id);
  }
  else random_step_resp<IDS>(cx, g, agent_seed, game_id);
}

// ---- MJAI-driven state tracking: GameState::apply_mjai_event (state/event_handler.rs:18-330; 3P state_3p/event_handler.rs) ----
// The record follows a game that is played elsewhere (a replay, an online table): events arrive one at a time, nothing is
// drawn from the wall, and the claim windows are rebuilt from the event itself.  Only bookkeeping — but it is the
// bookkeeping the legal-action and observation paths then read, so it lives with them.  `current_player` = RV_NONE is the
// reference's u8::MAX "nobody to act" marker.
__device__ __forceinline__ void hand_remove_if_present(G& g, int p, int tile) { hand_remove_first(g, p, tile); }
__device__ __noinline__ void claims_after_tile(const Ctx& cx, G& g, int actor, int tile, bool ron_only) {
  const int np = num_players(g);
  uint32_t active = 0;
  for (int i = 0; i < MAXP; i++) g.n_claims[i] = 0;                 // current_claims.clear()
  for (int i = 0; i < np; i++) {
    if (i == actor) continue;
    gen_claims(cx, g, i, actor, tile);
    if (ron_only) {                                                 // kita: only Ron survives (3P event_handler.rs:333-341)
      int n = 0;
      for (int k = 0; k < g.n_claims[i]; k++)
        if ((cold(g).claims[i][k] & 0xFF) == RV_RON) cold(g).claims[i][n++] = cold(g).claims[i][k];
      g.n_claims[i] = (uint8_t)n;
    }
    if (g.n_claims[i] > 0) active |= 1u << i;
  }
  if (active) {
    g.phase = RV_WAIT_RESPONSE;
    g.active_mask = (uint8_t)active;
  } else {
    g.phase = RV_WAIT_ACT;
    g.active_mask = 0;
    g.current_player = RV_NONE;
  }
  g.needs_tsumo = 1;
}
__device__ __noinline__ void apply_mjai_event(const Ctx& cx, G& g, const rv_mjai_event& e) {
  const int np = num_players(g);
  const bool sanma = np == 3;
  const int actor = e.actor < np ? e.actor : 0;
  switch (e.type) {
    case RV_EV_START_GAME:
      // env.rs:56-72: reset() (logs, counters) and then the handler's own lines (event_handler.rs:21-26)
      g.ev_hash = 0xcbf29ce484222325ull;
      g.ev_count = g.ev_words = g.step_count = 0;
      cold(g).kyoku_count = 0;
      g.overflow = 0;
      g.current_player = RV_NONE;
      g.active_mask = 0;
      break;
    case RV_EV_START_KYOKU: {
      g.honba = e.honba;
      g.riichi_sticks = e.kyotaku;
      g.round_wind = e.bakaze < 4 ? e.bakaze : 0;
      g.oya = e.oya;
      g.kyoku_idx = e.kyoku > 0 ? (uint8_t)(e.kyoku - 1) : 0;
      g.current_player = RV_NONE;
      g.turn_count = 0;
      g.is_done = 0;
      g.needs_tsumo = 1;
      g.phase = RV_WAIT_ACT;
      g.active_mask = 0;
      g.last_discard_pid = g.last_discard_tile = RV_NONE;
      g.pending_kan_pid = g.pending_kan_type = g.pending_kan_tile = RV_NONE;
      g.is_rinshan_flag = 0;
      g.is_first_turn = 1;
      g.riichi_pending_acceptance = RV_NONE;
      g.drawn_tile = RV_NONE;
      g.last_error = RV_NONE;
      g.pending_init[0] = g.pending_init[1] = g.pending_init[2] = RV_NONE;
      g.pending_tail[0] = RV_NONE;
      g.pending_tail[1] = 0;
      // wall.tiles = vec![0; TILES - 13 * np]: a placeholder of the pre-draw length, nothing is ever read from it
      const int wl = sanma ? 108 : 136, left = wl - 13 * np;
      for (int i = 0; i < 136; i++) cold(g).wall[i] = i < left ? 0 : RV_NONE;
      g.wall_len = (uint8_t)wl;
      g.wall_top = (uint8_t)left;
      g.rinshan_draw_count = 0;
      g.pending_kan_dora_count = 0;
      g.drawable_count = (uint8_t)(left - 14);
      g.n_dora = 1;
      g.dora_ind[0] = e.dora_marker;
      for (int i = 1; i < 5; i++) g.dora_ind[i] = RV_NONE;
      for (int p = 0; p < MAXP; p++) {                             // PlayerState::reset_round
        for (int i = 0; i < RV_HAND_CAP; i++) g.hand[p][i] = RV_NONE;
        g.hand_len[p] = 0;
        for (int m = 0; m < 4; m++) {
          for (int k = 0; k < 4; k++) g.meld_tiles[p][m][k] = RV_NONE;
          g.meld_type[p][m] = cold(g).meld_from[p][m] = cold(g).meld_called[p][m] = RV_NONE;
        }
        g.n_melds[p] = 0;
        for (int i = 0; i < RV_RIVER_CAP; i++) cold(g).river[p][i] = RV_NONE;
        g.n_river[p] = 0;
        g.river_tedashi[p] = 0;
        cold(g).river_riichi[p] = 0;
        cold(g).riichi_decl_idx[p] = RV_NONE;
        g.flags[p] = (sanma && p == 3) ? 0 : RV_F_NAGASHI_ELIGIBLE;
        cold(g).pao[p][0] = cold(g).pao[p][1] = RV_NONE;
        g.forbidden[p][0] = g.forbidden[p][1] = RV_NONE;
        cold(g).score_delta[p] = 0;
        g.n_claims[p] = 0;
        cold(g).riichi_sutehai[p] = cold(g).last_tedashi[p] = RV_NONE;
        cold(g).n_kita[p] = 0;
        for (int k = 0; k < 4; k++) g.c_cnt[p][k] = 0, g.c_key[p][k] = 0;
        g.c_river_kinds[p] = 0;
        g.c_waits[p] = 0;
        if (p < np) {
          g.score[p] = e.scores[p];
          const int n = e.tehai_len[p] < 14 ? e.tehai_len[p] : 14;
          for (int k = 0; k < n; k++) hand_push(g, p, e.tehais[p][k] < 136 ? e.tehais[p][k] : 0);
          hand_sort(g, p);
          waits_update(cx.T, g, p);
        }
      }
      break;
    }
    case RV_EV_TSUMO: {
      const int tile = e.pai < 136 ? e.pai : 0;
      g.current_player = (uint8_t)actor;
      g.drawn_tile = (uint8_t)tile;
      hand_push(g, actor, tile);
      hand_sort(g, actor);                                          // the handler sorts after the push
      g.forbidden[actor][0] = g.forbidden[actor][1] = RV_NONE;
      if (g.wall_top > g.rinshan_draw_count) {
        g.wall_top--;
        g.drawable_count = g.drawable_count > 0 ? (uint8_t)(g.drawable_count - 1) : 0;
      }
      g.phase = RV_WAIT_ACT;
      g.active_mask = (uint8_t)(1u << actor);
      g.needs_tsumo = 0;
      waits_update(cx.T, g, actor);
      break;
    }
    case RV_EV_DAHAI:
    case RV_EV_DAHAI_TSUMOGIRI: {
      const int tile = e.pai < 136 ? e.pai : 0;
      g.current_player = (uint8_t)actor;
      hand_remove_if_present(g, actor, tile);
      const int nr = g.n_river[actor];
      if (nr < RV_RIVER_CAP) cold(g).river[actor][nr] = (uint8_t)tile;
      else g.overflow |= 1;
      g.n_river[actor] = (uint8_t)(nr + 1);
      g.c_river_kinds[actor] |= 1ull << (tile >> 2);
      g.last_discard_pid = (uint8_t)actor;
      g.last_discard_tile = (uint8_t)tile;
      g.drawn_tile = RV_NONE;
      if (g.flags[actor] & RV_F_RIICHI_STAGE) g.flags[actor] = (uint8_t)((g.flags[actor] | RV_F_RIICHI_DECLARED) & ~RV_F_RIICHI_STAGE);
      waits_update(cx.T, g, actor);
      claims_after_tile(cx, g, actor, tile, false);
      break;
    }
    case RV_EV_PON:
    case RV_EV_CHI:
    case RV_EV_DAIMINKAN: {
      const int tile = e.pai < 136 ? e.pai : 0;
      g.current_player = (uint8_t)actor;
      const int nc = e.n_consumed < (e.type == RV_EV_DAIMINKAN ? 3 : 2) ? e.n_consumed : (e.type == RV_EV_DAIMINKAN ? 3 : 2);
      for (int k = 0; k < nc; k++) hand_remove_if_present(g, actor, e.consumed[k] < 136 ? e.consumed[k] : 0);
      const int m = g.n_melds[actor];
      if (m < 4) {                                                  // tiles = [called, consumed...] — NOT sorted by the handler
        g.meld_tiles[actor][m][0] = (uint8_t)tile;
        for (int k = 0; k < 3; k++) g.meld_tiles[actor][m][1 + k] = k < nc ? (uint8_t)(e.consumed[k] < 136 ? e.consumed[k] : 0) : (uint8_t)RV_NONE;
        g.meld_type[actor][m] = e.type == RV_EV_PON ? RV_MELD_PON : e.type == RV_EV_CHI ? RV_MELD_CHI : RV_MELD_DAIMINKAN;
        cold(g).meld_from[actor][m] = RV_NONE;                      // from_who: -1
        cold(g).meld_called[actor][m] = (uint8_t)tile;
        g.n_melds[actor] = (uint8_t)(m + 1);
      } else {
        g.overflow |= 1;
      }
      g.phase = RV_WAIT_ACT;
      g.active_mask = (uint8_t)(1u << actor);
      if (e.type == RV_EV_DAIMINKAN) {
        g.needs_tsumo = 1;
      } else {
        g.drawn_tile = RV_NONE;
        g.needs_tsumo = 0;
        if (e.type == RV_EV_CHI && sanma) {
          // the 3P handler accepts a chi "gracefully" and leaves forbidden_discards alone (state_3p/event_handler.rs:182-204)
        } else {
        g.forbidden[actor][0] = g.forbidden[actor][1] = RV_NONE;
        if (rule(g, RV_RULE_KUIKAE_FORBIDDEN)) {
          g.forbidden[actor][0] = (uint8_t)tile;
          if (e.type == RV_EV_CHI && !sanma && nc == 2) {           // the 3P handler has no suji rule (chi does not exist there)
            const int t34 = tile >> 2;
            int c0 = (e.consumed[0] < 136 ? e.consumed[0] : 0) >> 2, c1 = (e.consumed[1] < 136 ? e.consumed[1] : 0) >> 2;
            if (c0 > c1) { int t = c0; c0 = c1; c1 = t; }
            if (c0 == t34 + 1 && c1 == t34 + 2) {
              if (t34 % 9 <= 5) g.forbidden[actor][1] = (uint8_t)((t34 + 3) * 4);
            } else if (t34 >= 2 && c1 == t34 - 1 && c0 == t34 - 2 && t34 % 9 >= 3) {
              g.forbidden[actor][1] = (uint8_t)((t34 - 3) * 4);
            }
          }
        }
        }
      }
      waits_update(cx.T, g, actor);
      break;
    }
    case RV_EV_ANKAN: {
      const int nc = e.n_consumed < 4 ? e.n_consumed : 4;
      const int m = g.n_melds[actor];
      for (int k = 0; k < nc; k++) hand_remove_if_present(g, actor, e.consumed[k] < 136 ? e.consumed[k] : 0);
      if (m < 4) {
        for (int k = 0; k < 4; k++) g.meld_tiles[actor][m][k] = k < nc ? (uint8_t)(e.consumed[k] < 136 ? e.consumed[k] : 0) : (uint8_t)RV_NONE;
        g.meld_type[actor][m] = RV_MELD_ANKAN;
        cold(g).meld_from[actor][m] = RV_NONE;
        cold(g).meld_called[actor][m] = RV_NONE;
        g.n_melds[actor] = (uint8_t)(m + 1);
      } else {
        g.overflow |= 1;
      }
      g.current_player = (uint8_t)actor;
      g.phase = RV_WAIT_ACT;
      g.active_mask = (uint8_t)(1u << actor);
      g.needs_tsumo = 1;
      waits_update(cx.T, g, actor);
      break;
    }
    case RV_EV_KAKAN: {
      const int tile = e.pai < 136 ? e.pai : 0;
      hand_remove_if_present(g, actor, tile);
      for (int m = 0; m < g.n_melds[actor]; m++)
        if (g.meld_type[actor][m] == RV_MELD_PON && (g.meld_tiles[actor][m][0] >> 2) == (tile >> 2)) {
          g.meld_type[actor][m] = RV_MELD_KAKAN;
          g.meld_tiles[actor][m][3] = (uint8_t)tile;                // pushed at the end, not re-sorted
          break;
        }
      g.current_player = (uint8_t)actor;
      g.phase = RV_WAIT_ACT;
      g.active_mask = (uint8_t)(1u << actor);
      g.needs_tsumo = 1;
      waits_update(cx.T, g, actor);
      break;
    }
    case RV_EV_REACH:
      g.flags[actor] |= RV_F_RIICHI_STAGE;
      break;
    case RV_EV_REACH_ACCEPTED:
      g.flags[actor] |= RV_F_RIICHI_DECLARED;
      g.riichi_sticks += 1;
      g.score[actor] -= 1000;
      break;
    case RV_EV_DORA:
      if (g.n_dora < 5) g.dora_ind[g.n_dora++] = e.pai < 136 ? e.pai : 0;
      break;
    case RV_EV_KITA:
      if (sanma) {                                                  // 4P: ignored (event_handler.rs:305-307)
        int kita = -1;
        for (int k = 0; k < g.hand_len[actor]; k++)
          if ((g.hand[actor][k] >> 2) == 30) { kita = g.hand[actor][k]; break; }
        g.current_player = (uint8_t)actor;
        if (kita >= 0) {
          hand_remove_first(g, actor, kita);
          cold(g).n_kita[actor]++;
          waits_update(cx.T, g, actor);
          claims_after_tile(cx, g, actor, kita, true);
        } else {
          for (int i = 0; i < MAXP; i++) g.n_claims[i] = 0;
          g.phase = RV_WAIT_ACT;
          g.active_mask = 0;
          g.current_player = RV_NONE;
          g.needs_tsumo = 1;
        }
      }
      break;
    case RV_EV_HORA:
    case RV_EV_RYUKYOKU:
    case RV_EV_END_KYOKU:
      g.is_done = 1;
      break;
    default:
      break;
  }
}

// ---- replay ingestion: GameState::apply_log_action (state/event_handler.rs:332-894; 3P state_3p/event_handler.rs:365-808) ----
// One record follows one kyoku of a parsed log (rv_log_kyoku / rv_log_action, include/riichienv_b200.h).  Pure bookkeeping:
// nothing is drawn from the wall and no event is pushed; the decision points between two actions are read with the
// ordinary legal-action / observation code.
__device__ __forceinline__ void log_accept_riichi(G& g) {          // `if let Some(rp) = self.riichi_pending_acceptance.take()`
  const int rp = g.riichi_pending_acceptance;
  if (rp != RV_NONE) {
    g.score[rp & 3] -= 1000;
    g.riichi_sticks += 1;
    g.riichi_pending_acceptance = RV_NONE;
  }
}
__device__ __forceinline__ int hule_yakuman_val(int yid) { return (yid >= 47 && yid <= 50) ? 2 : 1; }
__device__ __noinline__ void apply_log_hule(G& g, const rv_log_action& a) {
  const int np = num_players(g);
  const bool sanma = np == 3;
  const int nh = a.n_hule < 3 ? a.n_hule : 3;
  auto real_tsumo = [&](const rv_hule& h) { return sanma ? (h.zimo && h.seat == g.current_player) : (h.zimo != 0); };
  if (nh > 0 && !real_tsumo(a.hules[0])) g.riichi_pending_acceptance = RV_NONE;      // the deposit is void when the discard is ronned
  const int32_t honba = g.honba;
  const uint32_t sticks = g.riichi_sticks;
  bool honba_taken = false;
  for (int k = 0; k < nh; k++) {
    const rv_hule& h = a.hules[k];
    const int w = h.seat < np ? h.seat : 0;
    const bool is_oya = w == g.oya;
    // PAO: which of the winner's yakuman has a liable seat (pao map filled by ChiPengGang below)
    int pao_payer = -1, pao_val = 0, total_val = 0;
    if (h.yiman)
      for (int y = 0; y < 64; y++)
        if ((h.fans >> y) & 1) {
          const int val = hule_yakuman_val(y);
          total_val += val;
          const int liable = y == 37 ? cold(g).pao[w][0] : y == 50 ? cold(g).pao[w][1] : RV_NONE;
          if (liable != RV_NONE) {
            pao_val += val;
            pao_payer = liable;
            if (!sanma && !real_tsumo(h)) break;        // the 4P ron branch stops at the first liable yaku
          }
        }
    if (real_tsumo(h)) {
      if (pao_val > 0) {
        int32_t pao_amt, non_pao, tsumo_total = 0;
        if (sanma) {
          tsumo_total = is_oya ? (int32_t)h.point_zimo_xian * (np - 1) : (int32_t)h.point_zimo_qin + (int32_t)h.point_zimo_xian * (np - 2);
          pao_amt = total_val > 0 ? tsumo_total * pao_val / total_val : tsumo_total;
          non_pao = tsumo_total - pao_amt;
        } else {
          const int32_t unit = is_oya ? 48000 : 32000;
          pao_amt = pao_val * unit;
          non_pao = (total_val - pao_val) * unit;
        }
        if (pao_payer >= 0) {
          g.score[pao_payer & 3] -= pao_amt;
          g.score[w] += pao_amt;
        }
        if (non_pao > 0)
          for (int i = 0; i < np; i++) {
            if (i == w) continue;
            int32_t share;
            if (sanma) share = is_oya ? non_pao / (np - 1) : (i == g.oya ? (int32_t)h.point_zimo_qin : (int32_t)h.point_zimo_xian) * non_pao / (tsumo_total ? tsumo_total : 1);
            else share = is_oya ? non_pao / 3 : (i == g.oya ? non_pao / 2 : non_pao / 4);
            g.score[i] -= share;
            g.score[w] += share;
          }
        if (pao_payer >= 0) {
          const int32_t hb = sanma ? honba * (np - 1) * 100 : honba * 300;
          g.score[pao_payer & 3] -= hb;
          g.score[w] += hb;
        }
      } else {
        for (int i = 0; i < np; i++) {
          if (i == w) continue;
          const int32_t base = is_oya ? h.point_zimo_xian : (i == g.oya ? h.point_zimo_qin : h.point_zimo_xian);
          const int32_t pay = base + honba * 100;
          g.score[i] -= pay;
          g.score[w] += pay;
        }
      }
    } else if (g.last_discard_pid != RV_NONE) {
      const int d = g.last_discard_pid & 3;
      const int32_t ron_honba = honba_taken ? 0 : honba;
      honba_taken = true;
      const int32_t hb = sanma ? ron_honba * (np - 1) * 100 : ron_honba * 300;
      if (sanma ? pao_val > 0 : pao_payer >= 0) {
        if (sanma) {
          const int pp = pao_payer >= 0 ? pao_payer & 3 : d;
          const int32_t ron_total = (int32_t)h.point_rong;
          const int32_t pao_amt = ron_total * pao_val / (total_val ? total_val : 1);
          const int32_t pao_share = pao_amt / 2 + hb, disc_share = ron_total - pao_amt / 2;
          g.score[pp] -= pao_share;
          g.score[d] -= disc_share;
          g.score[w] += pao_share + disc_share;
        } else {
          const int32_t half = (int32_t)h.point_rong / 2;
          g.score[pao_payer & 3] -= half + hb;
          g.score[d] -= half;
          g.score[w] += (int32_t)h.point_rong + hb;
        }
      } else {
        const int32_t pay = (int32_t)h.point_rong + hb;
        g.score[d] -= pay;
        g.score[w] += pay;
      }
    }
  }
  if (nh > 0) {
    g.score[a.hules[0].seat < np ? a.hules[0].seat : 0] += (int32_t)sticks * 1000;
    g.riichi_sticks = 0;
  }
  g.is_done = 1;
}
__device__ __noinline__ void apply_log_action(const Ctx& cx, G& g, const rv_log_action& a) {
  const int np = num_players(g);
  const bool sanma = np == 3;
  const int s = a.seat < np ? a.seat : 0;
  auto tid = [](int t) { return t < 136 ? t : 0; };
  switch (a.type) {
    case RV_LA_DISCARD: {
      const int t = tid(a.tile);
      const bool liqi = (a.flags & 3) != 0, wliqi = (a.flags & 2) != 0;
      const bool tsumogiri = g.drawn_tile != RV_NONE && g.drawn_tile == t;
      hand_remove_first(g, s, t);
      hand_sort(g, s);
      const int nr = g.n_river[s];
      if (nr < RV_RIVER_CAP) {
        cold(g).river[s][nr] = (uint8_t)t;
        if (!tsumogiri) g.river_tedashi[s] |= 1u << nr;
        if (liqi) cold(g).river_riichi[s] |= 1u << nr;
      } else {
        g.overflow |= 1;
      }
      g.n_river[s] = (uint8_t)(nr + 1);
      g.c_river_kinds[s] |= 1ull << (t >> 2);
      g.last_discard_pid = (uint8_t)s;
      g.last_discard_tile = (uint8_t)t;
      g.drawn_tile = RV_NONE;
      g.flags[s] &= ~RV_F_MISSED_AGARI_DOUJUN;
      if (!tid_terminal(t)) g.flags[s] &= ~RV_F_NAGASHI_ELIGIBLE;
      if (liqi) {
        if (sanma) {                                  // 3P: the flags and the pending deposit are (re)written unconditionally
          g.flags[s] |= RV_F_RIICHI_DECLARED;
          if (wliqi) g.flags[s] |= RV_F_DOUBLE_RIICHI;
          g.riichi_pending_acceptance = (uint8_t)s;
        } else if (!(g.flags[s] & RV_F_RIICHI_DECLARED)) {
          g.flags[s] |= RV_F_RIICHI_DECLARED;
          if (wliqi) g.flags[s] |= RV_F_DOUBLE_RIICHI;
          g.riichi_pending_acceptance = (uint8_t)s;
        }
        cold(g).riichi_decl_idx[s] = (uint8_t)nr;
      }
      g.current_player = (uint8_t)((s + 1) % np);
      g.phase = RV_WAIT_ACT;
      g.active_mask = (uint8_t)(1u << g.current_player);
      g.needs_tsumo = 1;
      g.is_first_turn = 0;
      g.is_after_kan = 0;
      waits_update(cx.T, g, s);
      break;
    }
    case RV_LA_DEAL: {
      const int t = tid(a.tile);
      log_accept_riichi(g);
      hand_push(g, s, t);
      g.drawn_tile = (uint8_t)t;
      g.current_player = (uint8_t)s;
      g.phase = RV_WAIT_ACT;
      g.active_mask = (uint8_t)(1u << s);
      g.is_rinshan_flag = g.is_after_kan ? 1 : 0;
      g.needs_tsumo = 0;
      g.is_after_kan = 0;
      hand_sort(g, s);
      if (g.wall_top > g.rinshan_draw_count) {       // `if !self.wall.tiles.is_empty()`
        g.wall_top--;
        g.drawable_count = g.drawable_count > 0 ? (uint8_t)(g.drawable_count - 1) : 0;
      }
      waits_update(cx.T, g, s);
      break;
    }
    case RV_LA_CHI_PENG_GANG: {
      log_accept_riichi(g);
      if (g.last_discard_pid != RV_NONE) g.flags[g.last_discard_pid & 3] &= ~RV_F_NAGASHI_ELIGIBLE;
      const int n = a.n_tiles < 4 ? a.n_tiles : 4;
      int from_who = -1, called = -1;
      for (int k = 0; k < n; k++) {
        if (a.froms[k] == s) hand_remove_first(g, s, tid(a.tiles[k]));
        else if (from_who < 0) from_who = a.froms[k], called = tid(a.tiles[k]);
      }
      hand_sort(g, s);
      const int m = g.n_melds[s];
      if (m < 4) {                                    // tiles as logged ([called, consumed...]), not sorted
        for (int k = 0; k < 4; k++) g.meld_tiles[s][m][k] = k < n ? (uint8_t)tid(a.tiles[k]) : (uint8_t)RV_NONE;
        g.meld_type[s][m] = a.meld_type;
        cold(g).meld_from[s][m] = from_who < 0 ? (uint8_t)RV_NONE : (uint8_t)from_who;
        cold(g).meld_called[s][m] = called < 0 ? (uint8_t)RV_NONE : (uint8_t)called;
        g.n_melds[s] = (uint8_t)(m + 1);
        if ((a.meld_type == RV_MELD_PON || a.meld_type == RV_MELD_DAIMINKAN) && called >= 0)
          register_pao(g, s, called, from_who < 0 ? 0 : from_who);
      } else {
        g.overflow |= 1;
      }
      g.current_player = (uint8_t)s;
      g.phase = RV_WAIT_ACT;
      g.active_mask = (uint8_t)(1u << s);
      const bool gang = a.meld_type == RV_MELD_DAIMINKAN;
      g.needs_tsumo = gang ? 1 : 0;
      g.is_first_turn = 0;
      g.is_after_kan = gang ? 1 : 0;
      waits_update(cx.T, g, s);
      break;
    }
    case RV_LA_ANGANG_ADDGANG: {
      const int t0 = tid(a.n_tiles ? a.tiles[0] : 0);
      if (a.meld_type == RV_MELD_ANKAN) {
        const int kind = t0 >> 2;
        for (int r = 0; r < 4; r++)
          for (int k = 0; k < g.hand_len[s]; k++)
            if ((g.hand[s][k] >> 2) == kind) {
              hand_remove_first(g, s, g.hand[s][k]);
              break;
            }
        const int m = g.n_melds[s];
        if (m < 4) {
          for (int k = 0; k < 4; k++) g.meld_tiles[s][m][k] = (uint8_t)(kind * 4 + k);
          g.meld_type[s][m] = RV_MELD_ANKAN;
          cold(g).meld_from[s][m] = RV_NONE;
          cold(g).meld_called[s][m] = RV_NONE;
          g.n_melds[s] = (uint8_t)(m + 1);
        } else {
          g.overflow |= 1;
        }
      } else {
        hand_remove_first(g, s, t0);
        for (int m = 0; m < g.n_melds[s]; m++)
          if (g.meld_type[s][m] == RV_MELD_PON && (g.meld_tiles[s][m][0] >> 2) == (t0 >> 2)) {
            g.meld_type[s][m] = RV_MELD_KAKAN;
            uint8_t* tl = g.meld_tiles[s][m];           // push + sort
            tl[3] = (uint8_t)t0;
            for (int x = 1; x < 4; x++)
              for (int y = x; y > 0 && tl[y - 1] > tl[y]; y--) { uint8_t v = tl[y]; tl[y] = tl[y - 1]; tl[y - 1] = v; }
            break;
          }
      }
      g.last_discard_pid = (uint8_t)s;                 // chankan / kokushi-on-ankan target
      g.last_discard_tile = (uint8_t)t0;
      hand_sort(g, s);
      g.current_player = (uint8_t)s;
      g.phase = RV_WAIT_ACT;
      g.active_mask = (uint8_t)(1u << s);
      g.needs_tsumo = 1;
      g.is_first_turn = 0;
      g.is_after_kan = 1;
      waits_update(cx.T, g, s);
      break;
    }
    case RV_LA_DORA:
      if (g.n_dora < 5) g.dora_ind[g.n_dora++] = (uint8_t)tid(a.tile);
      else g.overflow |= 1;                             // the reference's Vec keeps growing (steps() preloads every marker of the kyoku)
      break;
    case RV_LA_BABEI:
      if (sanma) {                                      // 4P: `_ => {}`
        log_accept_riichi(g);
        for (int k = 0; k < g.hand_len[s]; k++)
          if ((g.hand[s][k] >> 2) == 30) {
            const int t = g.hand[s][k];
            hand_remove_first(g, s, t);
            cold(g).n_kita[s]++;
            g.last_discard_pid = (uint8_t)s;
            g.last_discard_tile = (uint8_t)t;
            break;
          }
        hand_sort(g, s);
        g.current_player = (uint8_t)s;
        g.phase = RV_WAIT_ACT;
        g.active_mask = (uint8_t)(1u << s);
        g.needs_tsumo = 1;
        g.is_first_turn = 0;
        g.is_after_kan = 1;
        waits_update(cx.T, g, s);
      }
      break;
    case RV_LA_HULE:
      apply_log_hule(g, a);
      break;
    case RV_LA_NOTILE: {
      log_accept_riichi(g);
      int nagashi = 0;
      for (int p = 0; p < np; p++)
        if (g.flags[p] & RV_F_NAGASHI_ELIGIBLE) nagashi |= 1 << p;
      if (nagashi) {
        for (int w = 0; w < np; w++) {
          if (!((nagashi >> w) & 1)) continue;
          const bool is_oya = w == g.oya;                // calculate_score(5, 30, is_oya, true, 0, np): 4000 all / 2000-4000
          for (int i = 0; i < np; i++) {
            if (i == w) continue;
            const int32_t pay = is_oya ? 4000 : (i == g.oya ? 4000 : 2000);
            g.score[i] -= pay;
            g.score[w] += pay;
          }
        }
      } else {
        int ntp = 0;
        for (int p = 0; p < np; p++) ntp += g.c_waits[p] != 0;
        if (ntp > 0 && ntp < np) {
          const int32_t pool = sanma ? 2000 : 3000;
          const int32_t pk = pool / ntp, pn = pool / (np - ntp);
          for (int p = 0; p < np; p++) g.score[p] += g.c_waits[p] != 0 ? pk : -pn;
        }
      }
      g.is_done = 1;
      break;
    }
    case RV_LA_LIUJU:
      if (!sanma) log_accept_riichi(g);
      g.is_done = 1;
      break;
    default:
      break;
  }
}
// LogKyoku::steps (replay/mod.rs:1094-1292) after `_initialize_round(oya, chang, ben, liqibang, None, scores)` has run: the
// logged hands replace the dealt ones, the dealer's 14th tile goes back to the wall when the log deals it separately, and
// EVERY dora marker the kyoku will show is installed up front (as the reference does).
__device__ __noinline__ void replay_begin_patch(const Ctx& cx, G& g, const rv_log_kyoku& k) {
  const int np = num_players(g);
  const int oya = k.oya < np ? k.oya : 0;
  for (int p = 0; p < np; p++) {
    const int n = k.hand_len[p] < 14 ? k.hand_len[p] : 14;
    for (int i = 0; i < RV_HAND_CAP; i++) g.hand[p][i] = i < n ? (uint8_t)(k.hands[p][i] < 136 ? k.hands[p][i] : 0) : (uint8_t)RV_NONE;
    g.hand_len[p] = (uint8_t)n;
  }
  if (g.hand_len[oya] == 14) {
    g.drawn_tile = k.oya_drawn_tile;
    g.needs_tsumo = 0;
  } else {
    if (g.drawn_tile != RV_NONE) {                      // wall.tiles.push(dt); drawable_count += 1
      cold(g).wall[g.wall_top] = g.drawn_tile;
      g.wall_top++;
      g.drawable_count++;
    }
    g.drawn_tile = RV_NONE;
    g.needs_tsumo = 1;
  }
  for (int p = 0; p < np; p++) hand_sort(g, p);
  const int nd = k.n_doras < 5 ? k.n_doras : 5;
  for (int i = 0; i < 5; i++) g.dora_ind[i] = i < nd ? (uint8_t)(k.doras[i] < 136 ? k.doras[i] : 0) : (uint8_t)RV_NONE;
  g.n_dora = (uint8_t)nd;
  if (k.n_doras > 5) g.overflow |= 1;
  g.is_after_kan = 0;
  refresh_caches(cx.T, g);
}

// ---- agent #1: keyed "greedy-win" (definition shared with the oracle, oracle/game.hpp greedy_pick) ----
// Test agent: takes every Tsumo / Ron, declares every Riichi, calls Pon / Kan / Kita with probability 1/4 and Chi with 1/8, and
// discards towards the lowest shanten — so rollouts end in wins (~60 % of the rounds) and exercise the settlement code that
// uniform random play reaches in 0.3 % of them.  L = the packed legal list in the reference's order.
__device__ __noinline__ int greedy_pick(const Ctx& cx, const G& g, int pid, const uint32_t* L, int n, uint64_t agent_seed,
                                        uint64_t game_id) {
  const uint64_t r = mix64(agent_seed ^ (game_id * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)g.step_count << 8) ^ (uint64_t)pid);
  for (int i = 0; i < n; i++) {
    int ty = L[i] & 0xFF;
    if (ty == RV_TSUMO || ty == RV_RON) return i;
  }
  for (int i = 0; i < n; i++)
    if ((L[i] & 0xFF) == RV_RIICHI) return i;
  const uint32_t u = (uint32_t)(r >> 40) & 0xFF, v = (uint32_t)(r >> 8);
  if (u < 96) {
    const bool calls = u < 64;   // Pon / Daiminkan / Ankan / Kakan / Kita, else Chi
    int cnt = 0;
    for (int i = 0; i < n; i++) {
      int ty = L[i] & 0xFF;
      bool k = ty == RV_PON || ty == RV_DAIMINKAN || ty == RV_ANKAN || ty == RV_KAKAN || ty == RV_KITA;
      cnt += (calls ? k : ty == RV_CHI) ? 1 : 0;
    }
    if (cnt > 0) {
      int want = (int)(v % (uint32_t)cnt);
      for (int i = 0; i < n; i++) {
        int ty = L[i] & 0xFF;
        bool k = ty == RV_PON || ty == RV_DAIMINKAN || ty == RV_ANKAN || ty == RV_KAKAN || ty == RV_KITA;
        if ((calls ? k : ty == RV_CHI) && want-- == 0) return i;
      }
    }
  }
  if (g.phase == RV_WAIT_RESPONSE) {
    for (int i = 0; i < n; i++)
      if ((L[i] & 0xFF) == RV_PASS) return i;
    return n - 1;
  }
  Cnt c = hand_cnt(g, pid);
  const int len_div3 = (g.hand_len[pid] - 1) / 3;
  const bool sanma = is_sanma(g);
  int best = 99, nbest = 0;
  int8_t sh[RV_MAX_LEGAL];
  for (int i = 0; i < n && i < RV_MAX_LEGAL; i++) {
    sh[i] = 99;
    if ((L[i] & 0xFF) != RV_DISCARD || ((L[i] >> 8) & 0xFF) == RV_NONE) continue;
    const int k = ((L[i] >> 8) & 0xFF) >> 2;
    Cnt t = c;
    cnt_sub(t, k);
    sh[i] = (int8_t)(sanma ? shanten_counts_3p(cx.T, t, len_div3) : shanten_counts(cx.T, t, len_div3));
    if (sh[i] < best) best = sh[i], nbest = 0;
    if (sh[i] == best) nbest++;
  }
  if (nbest == 0) return (int)(v % (uint32_t)n);
  int want = (int)(v % (uint32_t)nbest);
  for (int i = 0; i < n && i < RV_MAX_LEGAL; i++)
    if (sh[i] == best && want-- == 0) return i;
  return 0;
}
// One env step with agent `policy` (0 = uniform random: random_step; 1 = greedy-win) through the generic legal-list path.
__device__ __noinline__ void agent_step(const Ctx& cx, G& g, int policy, uint64_t agent_seed, uint64_t game_id) {
  if (policy == 0) {
    random_step(cx, g, agent_seed, game_id);
    return;
  }
  const int np = num_players(g);
  rv_action acts[MAXP];
  for (int p = 0; p < MAXP; p++) acts[p].type = RV_NO_ACTION;
  uint32_t L[RV_MAX_LEGAL];
  if (g.phase == RV_WAIT_ACT) {
    const int pid = g.current_player;
    if (pid >= np) {
      g.step_count++;
      return;
    }
    const int n = min(legal_actions(cx, g, pid, L, -1, nullptr), RV_MAX_LEGAL);
    if (n == 0) {   // the reference's 3P dead end, retired as in random_step_act
      g.step_count++;
      g.is_done = 1;
      g.overflow |= 2;
      return;
    }
    acts[pid] = expand_act(g, pid, L[greedy_pick(cx, g, pid, L, n, agent_seed, game_id)]);
  } else {
    for (int p = 0; p < np; p++) {
      if (!((g.active_mask >> p) & 1)) continue;
      const int n = min(legal_actions(cx, g, p, L, -1, nullptr), RV_MAX_LEGAL);
      if (n > 0) acts[p] = expand_act(g, p, L[greedy_pick(cx, g, p, L, n, agent_seed, game_id)]);
    }
  }
  g.step_count++;
  step_apply(cx, g, acts);
}

}  // namespace rv
