// Sequence features of an observation (sparse tokens, numeric, progression, candidates).
//
// Replaces riichienv-core/src/observation/sequence_features.rs:
//   tile_id_to_kan37 (44-70), encode_chi (92-133), encode_pon (144-182), relative_from (186-188),
//   encode_seq_sparse (359-406), count_tiles_remaining (409-436), get_drawn_tile (439-465),
//   encode_seq_numeric (478-503), parse_start_kyoku_info (506-523), encode_seq_progression (535-700),
//   encode_seq_candidates (725-830), find_last_discard_actor (842-852).
// The reference re-parses the seat's MJAI JSON events; here the same facts are read from the binary event stream
// (include/riichienv_b200.h, rv_event_type) in ONE forward pass over the seat's event delta:
//   * get_drawn_tile scans backwards for the seat's last tsumo and stops at dahai/chi/pon/daiminkan  ==  forward:
//     "my tsumo sets it, a dahai/chi/pon/daiminkan clears it";
//   * find_last_discard_actor = actor of the last dahai/kakan;
//   * parse_start_kyoku_info = the FIRST start_kyoku of the delta (else the current honba/sticks/scores);
//   * progression = start_kyoku markers, dahai (with the pending-reach flag), chi/pon/daiminkan/ankan/kakan tuples.
// In the live env the delta is "events since this seat's previous observation" (state/mod.rs:211-218) and the
// progression cache is off (state/mod.rs:146), which is what the cursor arguments reproduce.
#pragma once
#include "game.cuh"

namespace rv {

constexpr int SEQ_MAX_SPARSE = 25, SEQ_SPARSE_PAD = 441, SEQ_MAX_PROG = 512, SEQ_MAX_CAND = 64, SEQ_NUMERIC = 12;

__device__ __forceinline__ bool tid_is_red(int t) { return t == 16 || t == 52 || t == 88; }
__device__ __forceinline__ int kan37_of_kind(int k) { return k < 9 ? k + 1 : k < 18 ? k + 2 : k + 3; }
__device__ __forceinline__ int kan37_of_tid(int t) { return t == 16 ? 0 : t == 52 ? 10 : t == 88 ? 20 : kan37_of_kind(t >> 2); }
__device__ __forceinline__ int seq_relative_from(int actor, int target) { return (target - actor + 3) % 4; }

// 0-89: 3 suits x 30; per suit the run start s (0-6) takes 3 slots, or 6 when the run holds the five (red variants)
__device__ inline int seq_encode_chi(int c0, int c1, int called) {
  int lo = min(min(c0, c1), called) >> 2;
  int suit = lo / 9, start = lo - 9 * suit;
  int call_pos = (called >> 2) - 9 * suit - start;
  bool has_red = tid_is_red(c0) || tid_is_red(c1) || tid_is_red(called);
  bool five = start <= 4 && 4 <= start + 2;
  int off = 0;
  for (int s = 0; s < start; s++) off += (s <= 4 && 4 <= s + 2) ? 6 : 3;
  return suit * 30 + off + ((five && has_red) ? 3 + call_pos : call_pos);
}
// 0-39: 3 suits x 11 (ranks 0-3, three five-variants, ranks 5-8) + 7 honors
__device__ inline int seq_encode_pon(int c0, int c1, int called) {
  int kind = called >> 2, suit = kind / 9;
  if (suit == 3) return 33 + (kind - 27);
  int rank = kind - 9 * suit;
  if (rank == 4) return suit * 11 + 4 + (tid_is_red(called) ? 2 : (tid_is_red(c0) || tid_is_red(c1)) ? 1 : 0);
  return suit * 11 + (rank < 4 ? rank : rank + 2);
}

struct SeqOut {
  uint16_t* sparse;    // [SEQ_MAX_SPARSE], padded with SEQ_SPARSE_PAD
  float* numeric;      // [SEQ_NUMERIC]
  uint16_t* prog;      // [max_prog][5], padded with {4,276,2,2,4}
  uint16_t* cand;      // [SEQ_MAX_CAND][4], padded with {279,2,2,3}
  uint16_t* lens;      // {n_sparse, n_prog (as the reference would return, <= 512), n_cand}
  int max_prog;
};

// One observation.  `log` = the game's event words, delta = words [w0, w1).
__device__ inline void seq_encode(const Ctx& cx, const G& g, int pid, const uint32_t* log, uint32_t w0, uint32_t w1, int game_style,
                                  const SeqOut& o) {
  // ---- one pass over the delta
  int drawn = -1, last_discarder = -1, pending_reach = -1;
  bool have_start = false;
  uint32_t st_honba = 0, st_kyotaku = 0;
  int32_t st_scores[4] = {0, 0, 0, 0};
  int n_prog = 0;
  auto push_prog = [&](int a, int ty, int mo, int li, int fr) {
    if (o.prog && n_prog < o.max_prog) {
      uint16_t* r = o.prog + 5 * n_prog;
      r[0] = (uint16_t)a, r[1] = (uint16_t)ty, r[2] = (uint16_t)mo, r[3] = (uint16_t)li, r[4] = (uint16_t)fr;
    }
    n_prog++;
  };
  for (uint32_t w = w0; w < w1;) {
    const uint32_t h = log[w];
    const int ty = h & 0xFF, nw = (h >> 8) & 0xFF, a = (h >> 16) & 0xFF, b = (h >> 24) & 0xFF;
    const bool open = n_prog < SEQ_MAX_PROG;       // the reference stops parsing once 512 tuples exist (700-702)
    switch (ty) {
      case RV_EV_START_KYOKU:
        if (!have_start) {
          have_start = true;
          const uint32_t w1v = log[w + 1];
          st_honba = w1v & 0xFF;
          st_kyotaku = w1v >> 16;
          const int np = num_players(g);
          for (int s = 0; s < np; s++) st_scores[s] = (int32_t)log[w + 2 + s];
        }
        if (open) push_prog(4, 0, 2, 2, 4);
        break;
      case RV_EV_TSUMO:
        if (a == pid) drawn = b;                   // own draws are never masked
        break;
      case RV_EV_REACH:
        if (open) pending_reach = a;
        break;
      case RV_EV_DAHAI:
      case RV_EV_DAHAI_TSUMOGIRI: {
        drawn = -1;
        last_discarder = a;
        if (open) {
          int liqi = pending_reach == a ? 1 : 0;
          if (liqi) pending_reach = -1;
          push_prog(a, 1 + kan37_of_tid(b), ty == RV_EV_DAHAI_TSUMOGIRI ? 1 : 0, liqi, 4);
        }
        break;
      }
      case RV_EV_CHI:
      case RV_EV_PON: {
        drawn = -1;
        const uint32_t m = log[w + 1];
        const int target = m & 0xFF, c0 = (m >> 8) & 0xFF, c1 = (m >> 16) & 0xFF;
        if (open)
          push_prog(a, ty == RV_EV_CHI ? 38 + seq_encode_chi(c0, c1, b) : 128 + seq_encode_pon(c0, c1, b), 2, 2,
                    seq_relative_from(a, target));
        break;
      }
      case RV_EV_DAIMINKAN: {
        drawn = -1;
        const int target = log[w + 1] & 0xFF;
        if (open) push_prog(a, 168 + kan37_of_tid(b), 2, 2, seq_relative_from(a, target));
        break;
      }
      case RV_EV_ANKAN:
        if (open) push_prog(a, 205 + ((log[w + 1] & 0xFF) >> 2), 2, 2, 4);
        break;
      case RV_EV_KAKAN:
        last_discarder = a;
        if (open) push_prog(a, 239 + kan37_of_tid(b), 2, 2, 4);
        break;
      default:
        break;
    }
    w += nw ? nw : 1;
  }
  if (n_prog > SEQ_MAX_PROG) n_prog = SEQ_MAX_PROG;
  if (o.prog)
    for (int k = n_prog; k < o.max_prog; k++) {
      uint16_t* r = o.prog + 5 * k;
      r[0] = 4, r[1] = 276, r[2] = 2, r[3] = 2, r[4] = 4;
    }
  // ---- sparse (359-406)
  int n_sparse = 0;
  if (o.sparse) {
    uint16_t* s = o.sparse;
    s[n_sparse++] = (uint16_t)min(game_style, 1);
    s[n_sparse++] = (uint16_t)(2 + min(pid, 3));
    s[n_sparse++] = (uint16_t)(6 + min((int)g.round_wind, 2));
    s[n_sparse++] = (uint16_t)(9 + min((int)g.oya, 3));
    int used = g.hand_len[pid] + g.n_dora;          // other hands are masked (state/mod.rs:189-197)
    for (int p = 0; p < 4; p++) {
      used += g.n_river[p];
      for (int m = 0; m < g.n_melds[p]; m++) used += (g.meld_type[p][m] >= RV_MELD_DAIMINKAN) ? 4 : 3;
    }
    int remaining = 136 - 14 - used;
    if (remaining < 0) remaining = 0;
    s[n_sparse++] = (uint16_t)(13 + min(remaining, 69));
    for (int i = 0; i < g.n_dora && i < 5; i++) s[n_sparse++] = (uint16_t)(83 + 37 * i + kan37_of_tid(g.dora_ind[i]));
    for (int i = 0; i < g.hand_len[pid] && n_sparse < SEQ_MAX_SPARSE; i++) s[n_sparse++] = (uint16_t)(268 + g.hand[pid][i]);
    if (drawn >= 0 && n_sparse < SEQ_MAX_SPARSE) s[n_sparse++] = (uint16_t)(404 + kan37_of_tid(drawn));
    for (int k = n_sparse; k < SEQ_MAX_SPARSE; k++) s[k] = SEQ_SPARSE_PAD;
  }
  // ---- numeric (478-503)
  if (o.numeric) {
    float* f = o.numeric;
    f[0] = (float)g.honba;
    f[1] = (float)g.riichi_sticks;
    for (int i = 0; i < 4; i++) f[2 + i] = (float)g.score[(pid + i) & 3];
    f[6] = (float)(have_start ? st_honba : (uint32_t)g.honba);
    f[7] = (float)(have_start ? st_kyotaku : g.riichi_sticks);
    for (int i = 0; i < 4; i++) f[8 + i] = (float)(have_start ? st_scores[(pid + i) & 3] : g.score[(pid + i) & 3]);
  }
  // ---- candidates (725-830)
  int n_cand = 0;
  if (o.cand) {
    const bool owes = !g.is_done && ((g.phase == RV_WAIT_ACT && g.current_player == pid) ||
                                     (g.phase == RV_WAIT_RESPONSE && ((g.active_mask >> pid) & 1)));
    uint32_t packed[RV_MAX_LEGAL];
    int cnt = owes ? legal_actions(cx, g, pid, packed, -1, nullptr) : 0;
    if (cnt > RV_MAX_LEGAL) cnt = RV_MAX_LEGAL;
    for (int k = 0; k < cnt; k++) {
      const rv_action act = expand_act(g, pid, packed[k]);
      int ty = -1, mo = 2, li = 2, fr = 3;
      const int rel = last_discarder >= 0 ? seq_relative_from(pid, last_discarder) : -1;
      switch (act.type) {
        case RV_DISCARD:
          if (act.tile != RV_NONE) ty = kan37_of_tid(act.tile), mo = (drawn >= 0 && drawn == act.tile) ? 1 : 0;
          break;
        case RV_ANKAN:
          if (act.n_consume > 0) ty = 37 + (act.consume[0] >> 2);
          break;
        case RV_KAKAN: {
          int t = act.tile != RV_NONE ? act.tile : (act.n_consume > 0 ? act.consume[0] : -1);
          if (t >= 0) ty = 71 + kan37_of_tid(t);
          break;
        }
        case RV_TSUMO: ty = 108; break;
        case RV_KYUSHU_KYUHAI: ty = 109; break;
        case RV_PASS: ty = 110; break;
        case RV_CHI:
          if (act.tile != RV_NONE && act.n_consume >= 2 && rel >= 0) ty = 111 + seq_encode_chi(act.consume[0], act.consume[1], act.tile), fr = rel;
          break;
        case RV_PON:
          if (act.tile != RV_NONE && act.n_consume >= 2 && rel >= 0) ty = 201 + seq_encode_pon(act.consume[0], act.consume[1], act.tile), fr = rel;
          break;
        case RV_DAIMINKAN:
          if (act.tile != RV_NONE && rel >= 0) ty = 241 + kan37_of_tid(act.tile), fr = rel;
          break;
        case RV_RON:
          if (rel >= 0) ty = 278, fr = rel;
          break;
        default:   // Riichi (the discards that follow carry it), Kita (3P, unsupported by the reference)
          break;
      }
      if (ty >= 0) {
        if (n_cand < SEQ_MAX_CAND) {
          uint16_t* r = o.cand + 4 * n_cand;
          r[0] = (uint16_t)ty, r[1] = (uint16_t)mo, r[2] = (uint16_t)li, r[3] = (uint16_t)fr;
        }
        n_cand++;      // the reference returns every candidate; rows beyond SEQ_MAX_CAND are counted, not stored
      }
    }
    for (int k = n_cand; k < SEQ_MAX_CAND; k++) {
      uint16_t* r = o.cand + 4 * k;
      r[0] = 279, r[1] = 2, r[2] = 2, r[3] = 3;
    }
  }
  if (o.lens) o.lens[0] = (uint16_t)n_sparse, o.lens[1] = (uint16_t)n_prog, o.lens[2] = (uint16_t)n_cand;
}

}  // namespace rv
