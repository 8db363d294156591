// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// CPU restatement (plain C++17, std::vector, recursive backtracking) of the
// reference's hand evaluation, written to follow the reference's algorithm
// step by step *including its quirks*, so that it can serve as the checker
// for the CUDA path.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use anything under oracle/.
//
// Follows (paths relative to /root/reference/riichienv-core/src):
//   types.rs:12-52      Hand (34-histogram)
//   agari.rs:63-245     is_agari / find_divisions / kokushi / chiitoitsu
//   hand_evaluator.rs:24-213, 286-300   HandEvaluator::{new,calc,is_tenpai,get_waits_u8}
//   yaku.rs:232-1280    calculate_yaku, fu, pinfu, static yaku, yakuman
//   score.rs:13-99      calculate_score
//
// Parity pin: tests/golden/agari_4p.txt (816 cases from the reference's
// benches/data/agari_4p.json), score table of tests/agari_correctness.rs:286-348,
// tests/test_agari_calculator.py known answers.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <vector>

namespace orc {

constexpr int TILE_MAX = 34;

// types.rs:12-52
struct Hand {
  std::array<uint8_t, TILE_MAX> counts{};
  void add(uint8_t t) {
    if (t < TILE_MAX) counts[t] += 1;
  }
  void remove(uint8_t t) {
    if (t < TILE_MAX && counts[t] > 0) counts[t] -= 1;
  }
};

// types.rs:57-63
enum MeldType : uint8_t { Chi = 0, Pon = 1, Daiminkan = 2, Ankan = 3, Kakan = 4 };

// types.rs:98-108
struct Meld {
  MeldType meld_type = Chi;
  std::vector<uint8_t> tiles;
  bool opened = true;
  int8_t from_who = -1;
  int16_t called_tile = -1;  // -1 = None
};

// types.rs:192-233
struct Conditions {
  bool tsumo = false, riichi = false, double_riichi = false, ippatsu = false;
  bool haitei = false, houtei = false, rinshan = false;
  uint8_t player_wind = 0, round_wind = 0;
  bool chankan = false, tsumo_first_turn = false;
  uint32_t riichi_sticks = 0, honba = 0;
  uint8_t kita_count = 0;
  bool is_sanma = false;
  uint8_t num_players = 4;
};

// types.rs:281-295
struct WinResult {
  bool is_win = false, yakuman = false;
  uint32_t ron_agari = 0, tsumo_agari_oya = 0, tsumo_agari_ko = 0;
  std::vector<uint32_t> yaku;
  uint32_t han = 0, fu = 0;
  int pao_payer = -1;
  bool has_win_shape = false;
};

// ---------------------------------------------------------------- agari.rs
struct Mentsu {
  bool koutsu;  // true: Koutsu(t), false: Shuntsu(t)
  uint8_t t;
};
struct Division {
  uint8_t head = 0;
  std::vector<Mentsu> body;
};

// agari.rs:141-164
inline bool is_kokushi(const Hand& hand) {
  static const int idx[13] = {0, 8, 9, 17, 18, 26, 27, 28, 29, 30, 31, 32, 33};
  bool pair_found = false;
  for (int i : idx) {
    uint8_t c = hand.counts[i];
    if (c == 0) return false;
    if (c == 2) {
      if (pair_found) return false;
      pair_found = true;
    } else if (c > 2) {
      return false;
    }
  }
  return pair_found;
}

// agari.rs:166-177
inline bool is_chiitoitsu(const Hand& hand) {
  int pairs = 0;
  for (uint8_t c : hand.counts) {
    if (c == 2)
      pairs++;
    else if (c != 0)
      return false;
  }
  return pairs == 7;
}

inline bool valid_seq_start(int i) {
  return (i >= 0 && i <= 6) || (i >= 9 && i <= 15) || (i >= 18 && i <= 24);
}

// agari.rs:197-245
inline bool decompose(Hand& hand, int start_idx) {
  int i = start_idx;
  while (i < TILE_MAX && hand.counts[i] == 0) i++;
  if (i == TILE_MAX) return true;
  if (hand.counts[i] >= 3) {
    hand.counts[i] -= 3;
    bool ok = decompose(hand, i);
    hand.counts[i] += 3;
    if (ok) return true;
  }
  if (i < 27 && valid_seq_start(i) && hand.counts[i + 1] > 0 && hand.counts[i + 2] > 0) {
    hand.counts[i] -= 1;
    hand.counts[i + 1] -= 1;
    hand.counts[i + 2] -= 1;
    bool ok = decompose(hand, i);
    hand.counts[i] += 1;
    hand.counts[i + 1] += 1;
    hand.counts[i + 2] += 1;
    if (ok) return true;
  }
  return false;
}

// agari.rs:179-195
inline bool is_standard_agari(Hand& hand) {
  for (int i = 0; i < TILE_MAX; i++) {
    if (hand.counts[i] >= 2) {
      hand.counts[i] -= 2;
      bool ok = decompose(hand, 0);
      hand.counts[i] += 2;
      if (ok) return true;
    }
  }
  return false;
}

// agari.rs:63-71
inline bool is_agari(Hand& hand) {
  if (is_kokushi(hand)) return true;
  if (is_chiitoitsu(hand)) return true;
  return is_standard_agari(hand);
}

// agari.rs:93-139
inline void decompose_all(Hand& hand, int start_idx, std::vector<Mentsu>& cur,
                          std::vector<std::vector<Mentsu>>& results) {
  int i = start_idx;
  while (i < TILE_MAX && hand.counts[i] == 0) i++;
  if (i == TILE_MAX) {
    results.push_back(cur);
    return;
  }
  if (hand.counts[i] >= 3) {
    hand.counts[i] -= 3;
    cur.push_back({true, (uint8_t)i});
    decompose_all(hand, i, cur, results);
    cur.pop_back();
    hand.counts[i] += 3;
  }
  if (i < 27 && valid_seq_start(i) && hand.counts[i + 1] > 0 && hand.counts[i + 2] > 0) {
    hand.counts[i] -= 1;
    hand.counts[i + 1] -= 1;
    hand.counts[i + 2] -= 1;
    cur.push_back({false, (uint8_t)i});
    decompose_all(hand, i, cur, results);
    cur.pop_back();
    hand.counts[i] += 1;
    hand.counts[i + 1] += 1;
    hand.counts[i + 2] += 1;
  }
}

// agari.rs:73-91
inline std::vector<Division> find_divisions(const Hand& hand) {
  std::vector<Division> divisions;
  for (int i = 0; i < TILE_MAX; i++) {
    if (hand.counts[i] >= 2) {
      Hand h = hand;
      h.counts[i] -= 2;
      std::vector<std::vector<Mentsu>> bodies;
      std::vector<Mentsu> cur;
      decompose_all(h, 0, cur, bodies);
      for (auto& b : bodies) divisions.push_back(Division{(uint8_t)i, b});
    }
  }
  return divisions;
}

// ---------------------------------------------------------------- score.rs
struct Score {
  uint32_t total = 0, pay_ron = 0, pay_tsumo_oya = 0, pay_tsumo_ko = 0;
};
inline uint32_t ceil_100(uint32_t v) { return (v + 99) / 100 * 100; }
// score.rs:54-88
inline Score make_score_result(uint32_t base, bool is_oya, bool is_tsumo, uint32_t np) {
  uint32_t total_ron = is_oya ? ceil_100(base * 6) : ceil_100(base * 4);
  uint32_t pay_oya, pay_ko;
  if (is_oya) {
    pay_oya = 0;
    pay_ko = ceil_100(base * 2);
  } else {
    pay_oya = ceil_100(base * 2);
    pay_ko = ceil_100(base);
  }
  uint32_t total_tsumo = is_oya ? pay_ko * (np - 1) : pay_oya + pay_ko * (np - 2);
  Score s;
  if (is_tsumo) {
    s.total = total_tsumo;
    s.pay_tsumo_oya = pay_oya;
    s.pay_tsumo_ko = pay_ko;
  } else {
    s.total = total_ron;
    s.pay_ron = total_ron;
  }
  return s;
}
// score.rs:13-52
inline Score calculate_score(uint8_t han, uint8_t fu, bool is_oya, bool is_tsumo, uint32_t honba,
                             uint8_t num_players) {
  uint32_t np = num_players;
  Score s;
  if (han >= 5) {
    uint32_t base;
    if (han == 5)
      base = 2000;
    else if (han <= 7)
      base = 3000;
    else if (han <= 10)
      base = 4000;
    else if (han <= 12)
      base = 6000;
    else
      base = 8000u * (han / 13u);
    s = make_score_result(base, is_oya, is_tsumo, np);
  } else {
    uint32_t f = (fu == 25) ? 25 : (uint32_t)((fu + 9) / 10 * 10);  // score.rs:90-95
    f &= 0xFF;                                                      // u8 arithmetic
    uint32_t bp = f * (1u << (2 + han));
    s = make_score_result(bp > 2000 ? 2000 : bp, is_oya, is_tsumo, np);
  }
  if (is_tsumo) {
    s.pay_tsumo_oya += honba * 100;
    s.pay_tsumo_ko += honba * 100;
    s.total += honba * 100 * (np - 1);
  } else {
    uint32_t hr = honba * 100 * (np - 1);
    s.pay_ron += hr;
    s.total += hr;
  }
  return s;
}

// ----------------------------------------------------------------- yaku.rs
struct YakuResult {
  uint8_t han = 0, fu = 0;
  std::vector<uint32_t> yaku_ids;
  uint8_t yakuman_count = 0;
};
struct YakuContext {
  bool is_menzen = true, is_reach = false, is_ippatsu = false, is_tsumo = false;
  bool is_haitei = false, is_houtei = false, is_rinshan = false, is_chankan = false;
  bool is_tsumo_first_turn = false, is_daburu_reach = false;
  uint8_t dora_count = 0, aka_dora = 0, ura_dora_count = 0;
  uint8_t nukidora_count = 0;  // yaku_3p.rs:30 (3P only)
  uint8_t round_wind = 27, seat_wind = 27;
};

namespace yk {
inline bool is_terminal(uint8_t t) { return t >= 27 || t % 9 == 0 || t % 9 == 8; }      // yaku.rs:763
inline bool is_number_terminal(uint8_t t) { return t < 27 && (t % 9 == 0 || t % 9 == 8); }  // :766
inline bool is_honor(uint8_t t) { return t >= 27; }                                       // :769

// yaku.rs:1214-1230
inline bool is_tanyao(const Hand& hand, const std::vector<Meld>& melds) {
  static const int term[13] = {0, 8, 9, 17, 18, 26, 27, 28, 29, 30, 31, 32, 33};
  for (int t : term)
    if (hand.counts[t] > 0) return false;
  for (auto& m : melds)
    for (uint8_t t : m.tiles)
      for (int x : term)
        if (x == t) return false;
  return true;
}
// yaku.rs:807-841
inline bool is_chinitsu(const Hand& hand, const std::vector<Meld>& melds) {
  bool suits[3] = {false, false, false};
  for (int i = 0; i < TILE_MAX; i++)
    if (hand.counts[i] > 0) {
      if (i >= 27) return false;
      suits[i / 9] = true;
    }
  for (auto& m : melds)
    for (uint8_t t : m.tiles) {
      if (t >= 27) return false;
      suits[t / 9] = true;
    }
  return (int)suits[0] + suits[1] + suits[2] == 1;
}
// yaku.rs:773-805
inline bool is_honitsu(const Hand& hand, const std::vector<Meld>& melds) {
  bool suits[3] = {false, false, false};
  bool has_honor = false;
  for (int i = 0; i < TILE_MAX; i++)
    if (hand.counts[i] > 0) {
      if (i < 27)
        suits[i / 9] = true;
      else
        has_honor = true;
    }
  for (auto& m : melds)
    for (uint8_t t : m.tiles) {
      if (t < 27)
        suits[t / 9] = true;
      else
        has_honor = true;
    }
  return ((int)suits[0] + suits[1] + suits[2] == 1) && has_honor;
}
// yaku.rs:692-705
inline bool is_honroutou(const Hand& hand, const std::vector<Meld>& melds) {
  for (int i = 0; i < TILE_MAX; i++)
    if (hand.counts[i] > 0 && !is_terminal(i)) return false;
  for (auto& m : melds)
    for (uint8_t t : m.tiles)
      if (!is_terminal(t)) return false;
  return true;
}
// yaku.rs:707-731
inline bool is_junchan(const Division& div, const std::vector<Meld>& melds) {
  if (!is_number_terminal(div.head)) return false;
  for (auto& m : div.body) {
    if (m.koutsu) {
      if (!is_number_terminal(m.t)) return false;
    } else {
      if (!is_number_terminal(m.t) && !is_number_terminal(m.t + 2)) return false;
    }
  }
  for (auto& m : melds) {
    bool all_non = true;
    for (uint8_t t : m.tiles)
      if (is_number_terminal(t)) all_non = false;
    if (all_non) return false;
  }
  return true;
}
// yaku.rs:733-761
inline bool is_chantai(const Division& div, const std::vector<Meld>& melds) {
  if (!is_terminal(div.head)) return false;
  bool has_honor = is_honor(div.head);
  for (auto& m : div.body) {
    if (m.koutsu) {
      if (!is_terminal(m.t)) return false;
      if (is_honor(m.t)) has_honor = true;
    } else {
      if (!is_terminal(m.t) && !is_terminal(m.t + 2)) return false;
    }
  }
  for (auto& m : melds) {
    bool all_non = true, any_honor = false;
    for (uint8_t t : m.tiles) {
      if (is_terminal(t)) all_non = false;
      if (is_honor(t)) any_honor = true;
    }
    if (all_non) return false;
    if (any_honor) has_honor = true;
  }
  return has_honor;
}
// yaku.rs:1057-1069
inline bool is_tsuu_iisou(const Hand& hand, const std::vector<Meld>& melds) {
  for (int i = 0; i < 27; i++)
    if (hand.counts[i] > 0) return false;
  for (auto& m : melds)
    for (uint8_t t : m.tiles)
      if (t < 27) return false;
  return true;
}
// yaku.rs:1071-1083
inline bool is_chinroutou(const Hand& hand, const std::vector<Meld>& melds) {
  for (int i = 0; i < TILE_MAX; i++)
    if (hand.counts[i] > 0 && !is_number_terminal(i)) return false;
  for (auto& m : melds)
    for (uint8_t t : m.tiles)
      if (!is_number_terminal(t)) return false;
  return true;
}
// yaku.rs:1085-1099
inline bool is_green(uint8_t t) {
  return t == 19 || t == 20 || t == 21 || t == 23 || t == 25 || t == 32;
}
inline bool is_ryuu_iisou(const Hand& hand, const std::vector<Meld>& melds) {
  for (int i = 0; i < TILE_MAX; i++)
    if (hand.counts[i] > 0 && !is_green(i)) return false;
  for (auto& m : melds)
    for (uint8_t t : m.tiles)
      if (!is_green(t)) return false;
  return true;
}
// yaku.rs:1101-1132
inline bool is_chuuren_poutou(const Hand& hand) {
  uint8_t counts[9] = {0};
  int suit = -1;
  for (int i = 0; i < TILE_MAX; i++) {
    uint8_t c = hand.counts[i];
    if (c > 0) {
      if (i >= 27) return false;
      int s = i / 9;
      if (suit >= 0) {
        if (suit != s) return false;
      } else {
        suit = s;
      }
      counts[i % 9] = c;
    }
  }
  if (counts[0] < 3 || counts[8] < 3) return false;
  for (int i = 1; i < 8; i++)
    if (counts[i] == 0) return false;
  return true;
}
// yaku.rs:1134-1147
inline bool is_chuuren_9_wait(const Hand& hand, uint8_t win_tile) {
  if (win_tile >= 27) return false;
  int val = win_tile % 9;
  int base = win_tile / 9 * 9;
  if (val == 0 || val == 8) return hand.counts[base + val] == 4;
  return hand.counts[base + val] == 2;
}
// yaku.rs:1149-1187
inline bool check_ittsu(const Division& div, const std::vector<Meld>& melds) {
  for (int off : {0, 9, 18}) {
    bool a = false, b = false, c = false;
    for (auto& m : div.body)
      if (!m.koutsu) {
        if (m.t == off)
          a = true;
        else if (m.t == off + 3)
          b = true;
        else if (m.t == off + 6)
          c = true;
      }
    for (auto& m : melds)
      if (m.meld_type == Chi) {
        uint8_t t = m.tiles[0];
        if (t == off)
          a = true;
        else if (t == off + 3)
          b = true;
        else if (t == off + 6)
          c = true;
      }
    if (a && b && c) return true;
  }
  return false;
}
// yaku.rs:1189-1212
inline bool is_sanshoku_doujun(const Division& div, const std::vector<Meld>& melds) {
  for (int i = 0; i < 7; i++) {
    bool a = false, b = false, c = false;
    for (auto& m : div.body)
      if (!m.koutsu) {
        if (m.t == i) a = true;
        if (m.t == i + 9) b = true;
        if (m.t == i + 18) c = true;
      }
    for (auto& m : melds)
      if (m.meld_type == Chi) {
        uint8_t t = m.tiles[0];
        if (t == i) a = true;
        if (t == i + 9) b = true;
        if (t == i + 18) c = true;
      }
    if (a && b && c) return true;
  }
  return false;
}
// yaku.rs:1232-1280
inline bool is_sanshoku_doukou(const Division& div, const std::vector<Meld>& melds) {
  for (int i = 0; i < 9; i++) {
    bool a = false, b = false, c = false;
    for (auto& m : div.body)
      if (m.koutsu) {
        if (m.t == i) a = true;
        if (m.t == i + 9) b = true;
        if (m.t == i + 18) c = true;
      }
    for (auto& m : melds)
      if (m.meld_type != Chi) {
        uint8_t t = m.tiles[0];
        if (t == i) a = true;
        if (t == i + 9) b = true;
        if (t == i + 18) c = true;
      }
    if (a && b && c) return true;
  }
  return false;
}
inline bool is_yakuhai_tile(uint8_t t, const YakuContext& ctx) {
  return t >= 31 || t == ctx.round_wind || t == ctx.seat_wind;
}
// yaku.rs:644-686   (wg < 0 == None)
inline bool check_pinfu(const Division& div, const std::vector<Meld>& melds, const YakuContext& ctx,
                        int wg, uint8_t win_tile) {
  if (!ctx.is_menzen) return false;
  if (!melds.empty()) return false;
  for (auto& m : div.body)
    if (m.koutsu) return false;
  if (is_yakuhai_tile(div.head, ctx)) return false;
  if (wg >= 0 && !div.body[wg].koutsu) {
    uint8_t t = div.body[wg].t;
    if (win_tile == t) {
      if (t % 9 == 6) return false;
      return true;
    }
    if (win_tile == t + 2) {
      if (t % 9 == 0) return false;
      return true;
    }
  }
  return false;
}
// yaku.rs:561-642
inline uint8_t calculate_fu_with_waiting(const Division& div, const std::vector<Meld>& melds,
                                         const YakuContext& ctx, int wg, uint8_t win_tile) {
  uint8_t fu = 20;
  if (ctx.is_tsumo)
    fu += 2;
  else if (ctx.is_menzen)
    fu += 10;
  if (div.head == ctx.round_wind) fu += 2;
  if (div.head == ctx.seat_wind) fu += 2;
  if (div.head >= 31) fu += 2;
  if (wg < 0) {
    fu += 2;
  } else if (!div.body[wg].koutsu) {
    uint8_t t = div.body[wg].t;
    if (win_tile == t + 1 || (win_tile == t + 2 && (t % 9 == 0)) || (win_tile == t && (t % 9 == 6)))
      fu += 2;
  }
  for (int idx = 0; idx < (int)div.body.size(); idx++) {
    auto& m = div.body[idx];
    if (m.koutsu) {
      uint8_t f = 4;
      if (!ctx.is_tsumo && idx == wg) f = 2;
      if (is_terminal(m.t)) f *= 2;
      fu += f;
    }
  }
  for (auto& m : melds) {
    if (m.tiles.size() >= 3 && m.tiles[0] == m.tiles[1]) {
      uint8_t f = 2;
      if (!m.opened) f = 4;
      if (is_terminal(m.tiles[0])) f *= 2;
      if (m.meld_type == Daiminkan || m.meld_type == Ankan || m.meld_type == Kakan) f *= 4;
      fu += f;
    }
  }
  if (fu == 20 && !ctx.is_tsumo) fu = 30;
  return (uint8_t)((fu + 9) / 10 * 10);
}

// yaku.rs:843-890
inline void apply_static_yaku(YakuResult& res, const YakuContext& ctx) {
  if (ctx.is_reach && !ctx.is_daburu_reach) {
    res.han += 1;
    res.yaku_ids.push_back(2);
  }
  if (ctx.is_daburu_reach) {
    res.han += 2;
    res.yaku_ids.push_back(18);
  }
  if (ctx.is_ippatsu) {
    res.han += 1;
    res.yaku_ids.push_back(30);
  }
  if (ctx.is_menzen && ctx.is_tsumo) {
    res.han += 1;
    res.yaku_ids.push_back(1);
  }
  if (ctx.is_haitei && ctx.is_tsumo) {
    res.han += 1;
    res.yaku_ids.push_back(5);
  }
  if (ctx.is_houtei && !ctx.is_tsumo) {
    res.han += 1;
    res.yaku_ids.push_back(6);
  }
  if (ctx.is_rinshan && ctx.is_tsumo) {
    res.han += 1;
    res.yaku_ids.push_back(4);
  }
  if (ctx.is_chankan && !ctx.is_tsumo) {
    res.han += 1;
    res.yaku_ids.push_back(3);
  }
  if (ctx.dora_count > 0) {
    res.han += ctx.dora_count;
    res.yaku_ids.push_back(31);
  }
  if (ctx.aka_dora > 0) {
    res.han += ctx.aka_dora;
    res.yaku_ids.push_back(32);
  }
  if (ctx.ura_dora_count > 0) {
    res.han += ctx.ura_dora_count;
    res.yaku_ids.push_back(33);
  }
  if (ctx.nukidora_count > 0) {  // yaku_3p.rs:706-709
    res.han += ctx.nukidora_count;
    res.yaku_ids.push_back(34);
  }
}

inline bool body_has_koutsu(const Division& div, uint8_t t) {
  for (auto& m : div.body)
    if (m.koutsu && m.t == t) return true;
  return false;
}

// yaku.rs:892-1055
inline void apply_yakuman(YakuResult& res, const Hand& hand, const std::vector<Meld>& melds,
                          const YakuContext& ctx, const Division& div, int wg, uint8_t win_tile) {
  uint8_t yakuman_count = 0;
  if (is_tsuu_iisou(hand, melds)) {
    yakuman_count += 1;
    res.yaku_ids.push_back(39);
  }
  if (is_chinroutou(hand, melds)) {
    yakuman_count += 1;
    res.yaku_ids.push_back(41);
  }
  if (is_ryuu_iisou(hand, melds)) {
    yakuman_count += 1;
    res.yaku_ids.push_back(40);
  }
  int kans = 0;
  for (auto& m : melds)
    if (m.meld_type == Daiminkan || m.meld_type == Ankan || m.meld_type == Kakan) kans++;
  if (kans == 4) {
    yakuman_count += 1;
    res.yaku_ids.push_back(44);
  }
  if (ctx.is_menzen && (div.body.size() + melds.size()) == 4) {
    if (is_chuuren_poutou(hand)) {
      if (is_chuuren_9_wait(hand, win_tile)) {
        yakuman_count += 2;
        res.yaku_ids.push_back(47);
      } else {
        yakuman_count += 1;
        res.yaku_ids.push_back(45);
      }
    }
  }
  if (ctx.is_tsumo_first_turn && ctx.is_menzen && ctx.is_tsumo) {
    yakuman_count += 1;
    res.yaku_ids.push_back(ctx.seat_wind == 27 ? 35 : 36);
  }
  int closed = 0;
  for (int idx = 0; idx < (int)div.body.size(); idx++) {
    if (div.body[idx].koutsu) {
      if (!ctx.is_tsumo && idx == wg) continue;
      closed++;
    }
  }
  for (auto& m : melds)
    if (m.meld_type == Ankan) closed++;
  if (closed == 4) {
    if (wg < 0) {
      yakuman_count += 2;
      res.yaku_ids.push_back(48);
    } else {
      yakuman_count += 1;
      res.yaku_ids.push_back(38);
    }
  }
  auto meld_contains = [&](uint8_t t) {
    for (auto& m : melds)
      for (uint8_t x : m.tiles)
        if (x == t) return true;
    return false;
  };
  bool haku = body_has_koutsu(div, 31) || meld_contains(31);
  bool hatsu = body_has_koutsu(div, 32) || meld_contains(32);
  bool chun = body_has_koutsu(div, 33) || meld_contains(33);
  if (haku && hatsu && chun) {
    yakuman_count += 1;
    res.yaku_ids.push_back(37);
  }
  int wk = 0, wp = 0;
  for (uint8_t w = 27; w <= 30; w++) {
    bool has = body_has_koutsu(div, w);
    if (!has)
      for (auto& m : melds)
        if (m.tiles[0] == w && m.meld_type != Chi) has = true;
    if (has)
      wk++;
    else if (div.head == w)
      wp++;
  }
  if (wk == 4) {
    yakuman_count += 2;
    res.yaku_ids.push_back(50);
  } else if (wk == 3 && wp == 1) {
    yakuman_count += 1;
    res.yaku_ids.push_back(43);
  }
  if (yakuman_count > 0) {
    res.han = 13 * yakuman_count;
    res.yakuman_count = yakuman_count;
  }
}
}  // namespace yk

// yaku.rs:232-559
inline YakuResult calculate_yaku(const Hand& hand, const std::vector<Meld>& melds,
                                 const YakuContext& ctx, uint8_t win_tile) {
  using namespace yk;
  auto divisions = find_divisions(hand);
  YakuResult best;
  if (divisions.empty()) {
    if (is_kokushi(hand)) {
      if (hand.counts[win_tile] == 2) {
        best.han = 26;
        best.yakuman_count = 2;
        best.yaku_ids.push_back(49);
      } else {
        best.han = 13;
        best.yakuman_count = 1;
        best.yaku_ids.push_back(42);
      }
      return best;
    }
    if (is_chiitoitsu(hand)) {
      best.han = 2;
      best.fu = 25;
      best.yaku_ids.push_back(25);
      if (is_tanyao(hand, melds)) {
        best.han += 1;
        best.yaku_ids.push_back(12);
      }
      if (is_chinitsu(hand, melds)) {
        best.han += 6;
        best.yaku_ids.push_back(29);
      } else if (is_honitsu(hand, melds)) {
        best.han += 3;
        best.yaku_ids.push_back(27);
      }
      if (is_honroutou(hand, melds)) {
        best.han += 2;
        best.yaku_ids.push_back(24);
      }
      Division d0;
      apply_yakuman(best, hand, melds, ctx, d0, -1, win_tile);
      apply_static_yaku(best, ctx);
      return best;
    }
    return best;
  }

  for (auto& div : divisions) {
    std::vector<int> wgs;  // -1 == None (head)
    if (div.head == win_tile) wgs.push_back(-1);
    for (int idx = 0; idx < (int)div.body.size(); idx++) {
      auto& m = div.body[idx];
      if (m.koutsu) {
        if (m.t == win_tile) wgs.push_back(idx);
      } else {
        if (win_tile >= m.t && win_tile <= m.t + 2) wgs.push_back(idx);
      }
    }
    if (wgs.empty()) continue;
    for (int wg : wgs) {
      YakuResult res;
      apply_yakuman(res, hand, melds, ctx, div, wg, win_tile);
      if (res.han >= 13) {
        if (res.han > best.han) best = res;
        continue;
      }
      apply_static_yaku(res, ctx);
      if (is_tanyao(hand, melds)) {
        res.han += 1;
        res.yaku_ids.push_back(12);
      }
      if (check_pinfu(div, melds, ctx, wg, win_tile)) {
        res.han += 1;
        res.yaku_ids.push_back(14);
        res.fu = ctx.is_tsumo ? 20 : 30;
      } else {
        res.fu = calculate_fu_with_waiting(div, melds, ctx, wg, win_tile);
      }
      // Yakuhai (yaku.rs:353-386)
      uint8_t ytiles[5] = {31, 32, 33, ctx.round_wind, ctx.seat_wind};
      for (int i = 0; i < 5; i++) {
        uint8_t t = ytiles[i];
        int count = 0;
        for (auto& m : div.body)
          if (m.koutsu && m.t == t) count++;
        for (auto& m : melds)
          if (m.tiles[0] == t && m.meld_type != Chi) count++;
        if (count > 0) {
          res.han += (uint8_t)count;
          uint32_t id = t == 31 ? 7 : t == 32 ? 8 : t == 33 ? 9 : (i == 3 ? 11 : 10);
          res.yaku_ids.push_back(id);
        }
      }
      // Shousangen (yaku.rs:388-436)
      auto drag = [&](uint8_t t) {
        if (body_has_koutsu(div, t)) return true;
        for (auto& m : melds)
          if (m.tiles[0] == t && m.meld_type != Chi) return true;
        return false;
      };
      bool haku = drag(31), hatsu = drag(32), chun = drag(33);
      if (!(haku && hatsu && chun)) {
        int dk = (int)haku + hatsu + chun;
        int dp = (div.head == 31) + (div.head == 32) + (div.head == 33);
        if (dk == 2 && dp == 1) {
          res.han += 2;
          res.yaku_ids.push_back(23);
        }
      }
      // Toitoi
      int koutsu_total = 0;
      for (auto& m : div.body)
        if (m.koutsu) koutsu_total++;
      for (auto& m : melds)
        if (m.meld_type != Chi) koutsu_total++;
      if (koutsu_total == 4) {
        res.han += 2;
        res.yaku_ids.push_back(21);
      }
      // San ankou
      int closed = 0;
      for (int idx = 0; idx < (int)div.body.size(); idx++)
        if (div.body[idx].koutsu) {
          if (!ctx.is_tsumo && idx == wg) continue;
          closed++;
        }
      for (auto& m : melds)
        if (m.meld_type == Ankan) closed++;
      if (closed == 3) {
        res.han += 2;
        res.yaku_ids.push_back(22);
      }
      // San kantsu
      int kans = 0;
      for (auto& m : melds)
        if (m.meld_type == Daiminkan || m.meld_type == Ankan || m.meld_type == Kakan) kans++;
      if (kans == 3) {
        res.han += 2;
        res.yaku_ids.push_back(20);
      }
      // Iipeikou / ryanpeikou
      if (ctx.is_menzen) {
        std::vector<uint8_t> st;
        for (auto& m : div.body)
          if (!m.koutsu) st.push_back(m.t);
        std::sort(st.begin(), st.end());
        int pairs = 0;
        size_t i = 0;
        while (i + 1 < st.size()) {
          if (st[i] == st[i + 1]) {
            pairs++;
            i += 2;
          } else {
            i += 1;
          }
        }
        if (pairs == 2) {
          res.han += 3;
          res.yaku_ids.push_back(28);
        } else if (pairs == 1) {
          res.han += 1;
          res.yaku_ids.push_back(13);
        }
      }
      if (check_ittsu(div, melds)) {
        res.han += ctx.is_menzen ? 2 : 1;
        res.yaku_ids.push_back(16);
      }
      if (is_sanshoku_doujun(div, melds)) {
        res.han += ctx.is_menzen ? 2 : 1;
        res.yaku_ids.push_back(17);
      }
      if (is_sanshoku_doukou(div, melds)) {
        res.han += 2;
        res.yaku_ids.push_back(19);
      }
      if (is_chinitsu(hand, melds)) {
        res.han += ctx.is_menzen ? 6 : 5;
        res.yaku_ids.push_back(29);
      } else if (is_honitsu(hand, melds)) {
        res.han += ctx.is_menzen ? 3 : 2;
        res.yaku_ids.push_back(27);
      }
      if (is_honroutou(hand, melds)) {
        res.han += 2;
        res.yaku_ids.push_back(24);
      } else if (is_junchan(div, melds)) {
        res.han += ctx.is_menzen ? 3 : 2;
        res.yaku_ids.push_back(26);
      } else if (is_chantai(div, melds)) {
        res.han += ctx.is_menzen ? 2 : 1;
        res.yaku_ids.push_back(15);
      }
      if (res.han > best.han || (res.han == best.han && res.fu > best.fu)) best = res;
    }
  }
  return best;
}

// ------------------------------------------------------- hand_evaluator.rs
inline bool is_aka(uint8_t tid) { return tid == 16 || tid == 52 || tid == 88; }

// hand_evaluator.rs:286-300
inline uint8_t get_next_tile(uint8_t t) {
  if (t < 9) return t == 8 ? 0 : t + 1;
  if (t < 18) return t == 17 ? 9 : t + 1;
  if (t < 27) return t == 26 ? 18 : t + 1;
  if (t < 31) return t == 30 ? 27 : t + 1;
  if (t == 33) return 31;
  return t + 1;
}

// hand_evaluator_3p.rs:300-311
inline uint8_t get_next_tile_sanma(uint8_t t) {
  if (t == 0) return 8;
  if (t == 8) return 0;
  if (t >= 1 && t <= 7) return t;
  return get_next_tile(t);
}

// HandEvaluator (4P) and HandEvaluator3P (hand_evaluator_3p.rs) differ only in `calc`; `sanma` selects.
struct HandEvaluator {
  bool sanma = false;
  Hand hand, full_hand;
  std::vector<Meld> melds;  // tiles in 34-space
  uint8_t aka_dora_count = 0;

  // hand_evaluator.rs:24-75
  HandEvaluator(const std::vector<uint8_t>& tiles_136, const std::vector<Meld>& in_melds, bool sanma_ = false) : sanma(sanma_) {
    for (uint8_t t : tiles_136) {
      if (is_aka(t)) aka_dora_count++;
      full_hand.add(t / 4);
    }
    hand = full_hand;
    for (auto& meld : in_melds) {
      Meld nm = meld;
      if (nm.meld_type == Daiminkan || nm.meld_type == Ankan || nm.meld_type == Kakan) {
        uint8_t t34 = nm.tiles[0] / 4;
        if (hand.counts[t34] == 4) hand.counts[t34] = 3;
      }
      std::vector<uint8_t> m34;
      for (uint8_t t : nm.tiles) {
        if (is_aka(t)) aka_dora_count++;
        m34.push_back(t / 4);
        full_hand.add(t / 4);
      }
      nm.tiles = m34;
      if (nm.meld_type == Chi) std::sort(nm.tiles.begin(), nm.tiles.end());
      melds.push_back(nm);
    }
  }

  int current_total() const {
    int s = 0;
    for (uint8_t c : hand.counts) s += c;
    return (uint8_t)(s + (int)melds.size() * 3);
  }

  // hand_evaluator.rs:77-176
  WinResult calc(uint8_t win_tile_136, const std::vector<uint8_t>& dora_ind,
                 const std::vector<uint8_t>& ura_ind, const Conditions& cond) const {
    uint8_t win34 = win_tile_136 / 4;
    Hand hand_14 = hand, full_14 = full_hand;
    int total = current_total();
    if (total == 13) {
      hand_14.add(win34);
      full_14.add(win34);
    }
    WinResult out;
    if (!is_agari(hand_14)) return out;
    uint8_t dora = 0, ura = 0;
    for (uint8_t ind : dora_ind) {
      uint8_t nt = sanma ? get_next_tile_sanma(ind / 4) : get_next_tile(ind / 4);
      dora += full_14.counts[nt];
      if (sanma && nt == 30) dora += cond.kita_count;  // hand_evaluator_3p.rs:108-115
    }
    for (uint8_t ind : ura_ind) {
      uint8_t nt = sanma ? get_next_tile_sanma(ind / 4) : get_next_tile(ind / 4);
      ura += full_14.counts[nt];
      if (sanma && nt == 30) ura += cond.kita_count;
    }
    uint8_t aka = aka_dora_count;
    if (total == 13 && is_aka(win_tile_136)) aka++;
    YakuContext ctx;
    ctx.is_tsumo = cond.tsumo;
    ctx.is_reach = cond.riichi;
    ctx.is_daburu_reach = cond.double_riichi;
    ctx.is_ippatsu = cond.ippatsu;
    ctx.is_haitei = cond.haitei;
    ctx.is_houtei = cond.houtei;
    ctx.is_rinshan = cond.rinshan;
    ctx.is_chankan = cond.chankan;
    ctx.is_tsumo_first_turn = cond.tsumo_first_turn;
    ctx.dora_count = dora;
    ctx.aka_dora = aka;
    ctx.ura_dora_count = ura;
    ctx.nukidora_count = sanma ? cond.kita_count : 0;
    ctx.round_wind = 27 + cond.round_wind;
    ctx.seat_wind = 27 + cond.player_wind;
    ctx.is_menzen = true;
    for (auto& m : melds)
      if (m.opened) ctx.is_menzen = false;
    YakuResult yr = calculate_yaku(hand_14, melds, ctx, win34);
    bool is_oya = cond.player_wind == 0;
    uint8_t scoring_han = (yr.yakuman_count == 0 && yr.han >= 13) ? 13 : yr.han;
    Score sc = calculate_score(scoring_han, yr.fu, is_oya, cond.tsumo, cond.honba, sanma ? 3 : 4);
    bool has_yaku = false;
    for (uint32_t id : yr.yaku_ids)
      if (id != 31 && id != 32 && id != 33 && !(sanma && id == 34)) has_yaku = true;
    out.is_win = (has_yaku || yr.yakuman_count > 0) && yr.han >= 1;
    out.yakuman = yr.yakuman_count > 0;
    out.ron_agari = sc.pay_ron;
    out.tsumo_agari_oya = sc.pay_tsumo_oya;
    out.tsumo_agari_ko = sc.pay_tsumo_ko;
    out.yaku = yr.yaku_ids;
    out.han = yr.han;
    out.fu = yr.fu;
    out.has_win_shape = true;
    return out;
  }

  // hand_evaluator.rs:196-213
  std::vector<uint8_t> get_waits_u8() const {
    std::vector<uint8_t> waits;
    if (current_total() != 13) return waits;
    Hand h = hand;
    for (int i = 0; i < TILE_MAX; i++) {
      if (h.counts[i] < 4) {
        h.add(i);
        if (is_agari(h)) waits.push_back(i);
        h.remove(i);
      }
    }
    return waits;
  }
  // hand_evaluator.rs:178-194
  bool is_tenpai() const {
    if (current_total() != 13) return false;
    Hand h = hand;
    for (int i = 0; i < TILE_MAX; i++) {
      if (h.counts[i] < 4) {
        h.add(i);
        if (is_agari(h)) return true;
        h.remove(i);
      }
    }
    return false;
  }
};

}  // namespace orc
