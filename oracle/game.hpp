// ORACLE — TEST INFRASTRUCTURE ONLY (see hand.hpp header).
//
// CPU restatement of the reference's 4-player game state machine, following
// (paths relative to /root/reference/riichienv-core/src):
//   state/mod.rs:98-187     GameState::new / reset
//   state/mod.rs:330-1315   step (validation, WaitAct, WaitResponse)
//   state/mod.rs:1317-1413  _resolve_discard
//   state/mod.rs:1415-1547  _resolve_kan
//   state/mod.rs:1549-1593  _accept_riichi / _deal_next
//   state/mod.rs:1595-1844  _initialize_next_round / _initialize_round
//   state/mod.rs:1846-2081  _trigger_ryukyoku / check_abortive_draw / kan dora / ura / end game
//   state/legal_actions.rs:11-508
//   state/player.rs, state/wall.rs
// Data structures deliberately mirror the reference (Vec per player, claim map) so
// that quirks (hand order, stale current_claims, score_delta semantics) carry over.
// Events are emitted in the binary format of include/riichienv_b200.h.
#pragma once
#include <cstring>
#include <optional>

#include "../include/riichienv_b200.h"
#include "hand.hpp"
#include "json.hpp"
#include "shanten.hpp"
#include "wall.hpp"

namespace orc {

constexpr int MAXP = 4;  // array capacity; the number of seats is GameState::np (4, or 3 for sanma)

struct Action {
  uint8_t type = RV_PASS;
  int tile = -1;  // -1 == None
  std::vector<uint8_t> consume;
  int actor = -1;
  Action() {}
  // action.rs:91-105 (sorts consume_tiles)
  Action(uint8_t ty, int t, std::vector<uint8_t> c, int a) : type(ty), tile(t), consume(std::move(c)), actor(a) {
    std::sort(consume.begin(), consume.end());
  }
};

// state/player.rs:6-39
struct PlayerState {
  std::vector<uint8_t> hand;
  std::vector<Meld> melds;
  std::vector<uint8_t> discards;
  std::vector<bool> discard_from_hand, discard_is_riichi;
  int riichi_declaration_index = -1;
  int32_t score = 25000, score_delta = 0;
  bool riichi_declared = false, riichi_stage = false, double_riichi_declared = false;
  bool missed_agari_riichi = false, missed_agari_doujun = false, nagashi_eligible = true, ippatsu_cycle = false;
  int pao37 = -1, pao50 = -1;  // pao: HashMap<u8,u8> only ever holds keys 37 and 50
  std::vector<uint8_t> forbidden_discards;
  // state/player.rs:66-86
  void reset_round() {
    hand.clear(); melds.clear(); discards.clear(); discard_from_hand.clear(); discard_is_riichi.clear();
    riichi_declaration_index = -1;
    riichi_declared = riichi_stage = double_riichi_declared = false;
    missed_agari_riichi = missed_agari_doujun = false;
    nagashi_eligible = true;
    ippatsu_cycle = false;
    forbidden_discards.clear();
    score_delta = 0;
    pao37 = pao50 = -1;
  }
};

inline bool is_terminal_tile(uint8_t t) {  // types.rs:364-369
  int tt = t / 4;
  return tt / 9 == 3 || tt % 9 == 0 || tt % 9 == 8;
}

struct GameState {
  // wall (state/wall.rs:9-19)
  std::vector<uint8_t> wall_tiles;
  std::vector<uint8_t> dora_indicators;
  uint8_t rinshan_draw_count = 0, pending_kan_dora_count = 0, drawable_count = 0;
  uint64_t wall_seed = 0, hand_index = 0;
  std::vector<uint8_t> wall_abs;  // absolute (never shrinking) copy for snapshots

  PlayerState players[MAXP];
  uint8_t current_player = 0;
  uint32_t turn_count = 0;
  bool is_done = false, needs_tsumo = false;
  bool stalled = false;   // rollout only: the seat to move had no legal action (see random_step)
  uint32_t riichi_sticks = 0;
  uint8_t phase = RV_WAIT_ACT;
  std::vector<uint8_t> active_players;
  int last_discard_pid = -1, last_discard_tile = -1;
  std::vector<Action> current_claims[MAXP];
  bool has_claims_entry[MAXP] = {false, false, false, false};
  bool pending_kan = false;
  uint8_t pending_kan_pid = 0;
  Action pending_kan_act;
  uint8_t oya = 0, honba = 0, kyoku_idx = 0, round_wind = 0;
  bool is_rinshan_flag = false, is_first_turn = true;
  bool is_after_kan = false;   // state/mod.rs:87 (replay only)
  int riichi_pending_acceptance = -1;
  int drawn_tile = -1;
  uint8_t game_mode = 0;
  int np = 4;          // state_3p/mod.rs:25 (3) vs state/mod.rs:27 (4)
  bool sanma = false;
  uint8_t dora_tiles3[5] = {0, 0, 0, 0, 0}, ura_tiles3[5] = {0, 0, 0, 0, 0};   // state_3p/wall.rs:46-49
  uint8_t n_kita[MAXP] = {0, 0, 0, 0};   // kita_tiles.len() (state_3p/player.rs:38)
  uint32_t rule = RV_RULE_DEFAULT_TENHOU;
  int last_error = -1;
  int riichi_sutehais[MAXP] = {-1, -1, -1, -1}, last_tedashis[MAXP] = {-1, -1, -1, -1};

  // event stream
  bool keep_log = true;
  std::vector<uint32_t> log;
  uint32_t stat_pao = 0;        // settlements that charged a pao payer (test statistics only)
  MjaiLog text;                 // mjai_log / mjai_log_per_player as the reference writes them (json.hpp)
  uint64_t ev_hash = 0xcbf29ce484222325ull;
  uint32_t ev_count = 0, step_count = 0, kyoku_count = 0, ev_words = 0;
  // per-winner results of the last settlement (win_results, state/mod.rs:60)
  std::vector<std::pair<int, WinResult>> win_results;

  bool rb(uint32_t bit) const { return (rule & bit) != 0; }
  int32_t starting_score() const { return sanma ? 35000 : 25000; }   // state_3p/game_mode.rs:31-33
  int32_t target_score() const { return sanma ? 40000 : 30000; }     // state_3p/mod.rs:1524,1553,1561

  // ------------------------------------------------------------ events
  void push_words(const uint32_t* w, int n) {
    for (int i = 0; i < n; i++) {
      ev_hash = (ev_hash ^ w[i]) * 0x100000001b3ull;
      if (keep_log) log.push_back(w[i]);
    }
    ev_count++;
    ev_words += (uint32_t)n;
  }
  static uint32_t w0(int type, int n, int a, int b) {
    return (uint32_t)type | ((uint32_t)n << 8) | ((uint32_t)(a & 0xFF) << 16) | ((uint32_t)(b & 0xFF) << 24);
  }
  void ev_simple(int type, int a = 0, int b = 0) {
    uint32_t w = w0(type, 1, a, b);
    push_words(&w, 1);
    if (!keep_log) return;
    text.np = np;
    JsonObj ev;
    switch (type) {
      case RV_EV_START_GAME: text.push(ev.str("type", "start_game")); break;                            // mod.rs:158-162
      case RV_EV_END_KYOKU: text.push(ev.str("type", "end_kyoku")); break;                              // mod.rs:1674, 2075
      case RV_EV_END_GAME: text.push(ev.str("type", "end_game")); break;                                // mod.rs:2079
      case RV_EV_TSUMO: text.push(ev.str("type", "tsumo").num("actor", a).str("pai", mjai_tile(b)), nullptr, a); break;
      case RV_EV_DAHAI:                                                                                 // mod.rs:1365-1368
      case RV_EV_DAHAI_TSUMOGIRI:
        text.push(ev.str("type", "dahai").num("actor", a).str("pai", mjai_tile(b)).boolean("tsumogiri", type == RV_EV_DAHAI_TSUMOGIRI));
        break;
      case RV_EV_REACH: text.push(ev.str("type", "reach").num("actor", a)); break;                      // mod.rs:454-455
      case RV_EV_REACH_ACCEPTED: text.push(ev.str("type", "reach_accepted").num("actor", a)); break;    // mod.rs:1556-1563
      case RV_EV_DORA: text.push(ev.str("type", "dora").str("dora_marker", mjai_tile(b))); break;       // mod.rs:2031-2036
      case RV_EV_KITA: text.push(ev.str("type", "kita").num("actor", a).str("pai", mjai_tile(b))); break;   // sanma.rs:47-54
      default: break;
    }
  }
  void ev_deltas(int type, int a, const int32_t* d) {   // 1 + np words
    uint32_t w[5] = {w0(type, 1 + np, a, 0), (uint32_t)d[0], (uint32_t)d[1], (uint32_t)d[2], (uint32_t)d[3]};
    push_words(w, 1 + np);
    if (!keep_log) return;
    static const char* reasons[] = {"exhaustive_draw", "nagashimangan", "kyushu_kyuhai", "sufuurenta", "suukansansen",
                                    "suucha_riichi", "sanchaho"};
    std::string reason = a < RV_RK_ILLEGAL_BASE ? reasons[a] : "Error: Illegal Action by Player " + std::to_string(a - RV_RK_ILLEGAL_BASE);
    text.np = np;
    text.push(JsonObj().str("type", "ryukyoku").str("reason", reason).nums("deltas", d, d + np));       // mod.rs:1957-1963
  }
  // pon / chi / daiminkan / ankan / kakan (mod.rs:1195-1224, 1498-1517, 571-581)
  void text_meld(const char* type, int actor, int target, int tile, const std::vector<uint8_t>& consumed) {
    if (!keep_log) return;
    JsonObj ev;
    ev.str("type", type).num("actor", actor);
    if (target >= 0) ev.num("target", target);
    if (tile >= 0) ev.str("pai", mjai_tile(tile));
    std::vector<std::string> c;
    for (uint8_t t : consumed) c.push_back(mjai_tile(t));
    text.np = np;
    text.push(ev.strs("consumed", c));
  }

  // ------------------------------------------------------------ ctor / reset
  // state/mod.rs:98-167
  GameState(uint8_t mode, uint64_t seed, uint8_t rw, uint32_t rule_bits, bool log_events = true)
      : wall_seed(seed), round_wind(rw), game_mode(mode), rule(rule_bits), keep_log(log_events) {
    sanma = mode >= 3;    // game_variant.rs:12-36
    np = sanma ? 3 : 4;
    for (auto& p : players) p.score = starting_score();
    ev_simple(RV_EV_START_GAME);
    _initialize_round(0, rw, 0, 0, nullptr, nullptr);
  }
  // state/mod.rs:171-187 + env.rs:799-851
  void reset(uint8_t oya_ = 0, uint8_t rw = 0, uint8_t honba_ = 0, uint32_t kyotaku = 0,
             const std::vector<uint8_t>* wall = nullptr, const int32_t* scores = nullptr) {
    log.clear();
    text.clear();
    stalled = false;
    ev_hash = 0xcbf29ce484222325ull;
    ev_count = 0;
    ev_words = 0;
    step_count = 0;
    kyoku_count = 0;
    ev_simple(RV_EV_START_GAME);
    int32_t def[MAXP] = {starting_score(), starting_score(), starting_score(), starting_score()};
    _initialize_round(oya_, rw, honba_, kyotaku, wall, scores ? scores : def);
  }

  // state/wall.rs:36-67 / 69-80
  void wall_shuffle() {
    wall_tiles = wall_from_seed(wall_seed, hand_index, sanma ? 108 : 136);
    hand_index++;
    wall_loaded();
  }
  void load_wall(const std::vector<uint8_t>& t) {
    wall_tiles.assign(t.rbegin(), t.rend());
    wall_loaded();
  }
  void wall_loaded() {
    wall_abs = wall_tiles;
    dora_indicators.clear();
    if (sanma) {
      // state_3p/wall.rs:104-112 (shuffle) and 141-146 (load_wall: t[99],t[97].. == reversed index 8+2i)
      if (wall_tiles.size() == 108)
        for (int i = 0; i < 5; i++) {
          dora_tiles3[i] = wall_tiles[8 + 2 * i];
          ura_tiles3[i] = wall_tiles[9 + 2 * i];
        }
      dora_indicators.push_back(dora_tiles3[0]);
    } else if (wall_tiles.size() > 5) {
      dora_indicators.push_back(wall_tiles[4]);
    }
    rinshan_draw_count = 0;
    pending_kan_dora_count = 0;
    drawable_count = 0;
  }

  // state/mod.rs:1695-1844
  void _initialize_round(uint8_t oya_, uint8_t rw, uint8_t honba_, uint32_t kyotaku,
                         const std::vector<uint8_t>* wall, const int32_t* scores) {
    oya = oya_;
    kyoku_idx = oya_;
    current_player = oya_;
    honba = honba_;
    riichi_sticks = kyotaku;
    round_wind = rw;
    for (auto& p : players) p.reset_round();
    for (int i = 0; i < MAXP; i++) n_kita[i] = 0;
    is_done = false;
    for (int i = 0; i < np; i++) {
      current_claims[i].clear();
      has_claims_entry[i] = false;
    }
    pending_kan = false;
    is_rinshan_flag = false;
    rinshan_draw_count = 0;
    pending_kan_dora_count = 0;
    is_first_turn = true;
    riichi_pending_acceptance = -1;
    turn_count = 0;
    needs_tsumo = true;
    last_discard_pid = last_discard_tile = -1;
    win_results.clear();
    for (int i = 0; i < np; i++) riichi_sutehais[i] = last_tedashis[i] = -1;
    if (scores)
      for (int i = 0; i < np; i++) players[i].score = scores[i];
    if (wall)
      load_wall(*wall);
    else
      wall_shuffle();
    kyoku_count++;
    for (int r = 0; r < 3; r++)
      for (int idx = 0; idx < np; idx++) {
        int p = (idx + oya) % np;
        for (int k = 0; k < 4; k++)
          if (!wall_tiles.empty()) {
            players[p].hand.push_back(wall_tiles.back());
            wall_tiles.pop_back();
          }
      }
    for (int idx = 0; idx < np; idx++) {
      int p = (idx + oya) % np;
      if (!wall_tiles.empty()) {
        players[p].hand.push_back(wall_tiles.back());
        wall_tiles.pop_back();
      }
    }
    for (auto& p : players) std::sort(p.hand.begin(), p.hand.end());
    drawable_count = (uint8_t)(wall_tiles.size() - 14);

    {  // start_kyoku (state/mod.rs:1785-1820); 4P: 19 words, 3P: 15 words (3 scores, 39 tehai bytes)
      uint32_t w[19];
      int nb = 13 * np, nwords = 2 + np + (nb + 3) / 4;
      w[0] = w0(RV_EV_START_KYOKU, nwords, round_wind % 4, oya);
      w[1] = (uint32_t)honba | ((uint32_t)dora_indicators[0] << 8) | ((kyotaku & 0xFFFF) << 16);
      for (int i = 0; i < np; i++) w[2 + i] = (uint32_t)players[i].score;
      uint8_t th[52];
      memset(th, 0xFF, sizeof th);
      for (int i = 0; i < np; i++)
        for (size_t k = 0; k < players[i].hand.size() && k < 13; k++) th[i * 13 + k] = players[i].hand[k];
      memcpy(&w[2 + np], th, (size_t)((nb + 3) / 4) * 4);
      push_words(w, nwords);
      if (keep_log) {   // mod.rs:1785-1819
        static const char* winds[4] = {"E", "S", "W", "N"};
        JsonObj ev;
        ev.str("type", "start_kyoku").str("bakaze", winds[round_wind % 4]).num("kyoku", oya + 1).num("honba", honba);
        ev.num("kyotaku", riichi_sticks).num("oya", oya).str("dora_marker", mjai_tile(dora_indicators[0]));
        std::vector<int32_t> sc;
        std::vector<std::vector<std::string>> tehais;
        for (int i = 0; i < np; i++) {
          sc.push_back(players[i].score);
          tehais.emplace_back();
          for (uint8_t t : players[i].hand) tehais.back().push_back(mjai_tile(t));
        }
        ev.nums("scores", sc.begin(), sc.end());
        text.np = np;
        text.push(ev, &tehais);
      }
    }
    current_player = oya;
    phase = RV_WAIT_ACT;
    active_players = {oya};
    if (!wall_tiles.empty()) {
      uint8_t t = wall_tiles.back();
      wall_tiles.pop_back();
      drawable_count -= 1;
      players[oya].hand.push_back(t);
      drawn_tile = t;
      needs_tsumo = false;
      ev_simple(RV_EV_TSUMO, oya, t);
    } else {
      needs_tsumo = true;
      drawn_tile = -1;
    }
  }

  // ------------------------------------------------------------ helpers
  Conditions base_cond(int pid) const {
    Conditions c;
    c.riichi = players[pid].riichi_declared;
    c.double_riichi = players[pid].double_riichi_declared;
    c.ippatsu = players[pid].ippatsu_cycle;
    c.player_wind = (uint8_t)((pid + np - oya) % np);
    c.round_wind = (uint8_t)(round_wind % 4);
    c.riichi_sticks = riichi_sticks;
    c.honba = honba;
    c.is_sanma = sanma;
    c.num_players = (uint8_t)np;
    return c;
  }
  // state/mod.rs:2048-2069
  std::vector<uint8_t> _get_ura_indicators() const {
    std::vector<uint8_t> out;
    if (sanma) {  // state_3p/mod.rs:1925-1931
      for (size_t i = 0; i < dora_indicators.size() && i < 5; i++) out.push_back(ura_tiles3[i]);
      return out;
    }
    for (size_t i = 0; i < dora_indicators.size(); i++) {
      size_t raw = 5 + 2 * i;
      size_t idx = raw >= rinshan_draw_count ? raw - rinshan_draw_count : 0;
      if (idx < wall_tiles.size()) out.push_back(wall_tiles[idx]);
    }
    return out;
  }
  // state/mod.rs:2021-2046 ; 3P: state_3p/mod.rs:1893-1915
  void _reveal_kan_dora() {
    size_t count = dora_indicators.size();
    if (count < 5) {
      if (sanma) {
        dora_indicators.push_back(dora_tiles3[count]);
        ev_simple(RV_EV_DORA, 0, dora_indicators.back());
        return;
      }
      size_t raw = 4 + 2 * count;
      size_t base = raw >= rinshan_draw_count ? raw - rinshan_draw_count : 0;
      if (base < wall_tiles.size()) {
        dora_indicators.push_back(wall_tiles[base]);
        ev_simple(RV_EV_DORA, 0, dora_indicators.back());
      }
    }
  }
  static bool vec_remove_first(std::vector<uint8_t>& v, uint8_t t) {
    for (size_t i = 0; i < v.size(); i++)
      if (v[i] == t) {
        v.erase(v.begin() + i);
        return true;
      }
    return false;
  }

  // ------------------------------------------------------------ legal actions
  // state/legal_actions.rs:11-252
  std::vector<Action> _get_legal_actions_internal(int pid) const {
    std::vector<Action> legals;
    const PlayerState& P = players[pid];
    if (is_done) return legals;
    if (phase == RV_WAIT_ACT) {
      if (pid != current_player) return legals;
      // 1. Tsumo
      if (drawn_tile >= 0 && !P.riichi_stage) {
        Conditions cond = base_cond(pid);
        cond.tsumo = true;
        cond.haitei = drawable_count == 0 && !is_rinshan_flag;
        cond.rinshan = is_rinshan_flag;
        cond.tsumo_first_turn = is_first_turn && P.discards.empty();
        std::vector<uint8_t> hand = P.hand;
        for (int i = (int)hand.size() - 1; i >= 0; i--)
          if (hand[i] == drawn_tile) {
            hand.erase(hand.begin() + i);
            break;
          }
        HandEvaluator calc(hand, P.melds, sanma);
        WinResult res = calc.calc((uint8_t)drawn_tile, dora_indicators, {}, cond);
        if (res.is_win && (res.yakuman || res.han >= 1)) legals.emplace_back(RV_TSUMO, drawn_tile, std::vector<uint8_t>{}, pid);
      }
      // 2. Discard / Riichi
      auto forbidden = [&](uint8_t t) {
        for (uint8_t f : P.forbidden_discards)
          if (f / 4 == t / 4) return true;
        return false;
      };
      if (P.riichi_declared) {
        if (drawn_tile >= 0) legals.emplace_back(RV_DISCARD, drawn_tile, std::vector<uint8_t>{}, pid);
      } else if (P.riichi_stage) {
        for (uint8_t t : P.hand) {
          if (forbidden(t)) continue;
          std::vector<uint8_t> tmp = P.hand;
          vec_remove_first(tmp, t);
          HandEvaluator calc(tmp, P.melds, sanma);
          if (calc.is_tenpai()) legals.emplace_back(RV_DISCARD, t, std::vector<uint8_t>{}, pid);
        }
      } else {
        for (uint8_t t : P.hand)
          if (!forbidden(t)) legals.emplace_back(RV_DISCARD, t, std::vector<uint8_t>{}, pid);
        bool all_closed = true;
        for (auto& m : P.melds)
          if (m.opened) all_closed = false;
        if (P.score >= 1000 && (sanma ? drawable_count > 0 : drawable_count >= 4) && all_closed) {   // state_3p/legal_actions.rs:116
          bool can = false;
          for (size_t skip = 0; skip < P.hand.size(); skip++) {
            std::vector<uint8_t> tmp = P.hand;
            tmp.erase(tmp.begin() + skip);
            HandEvaluator calc(tmp, P.melds, sanma);
            if (calc.is_tenpai()) {
              can = true;
              break;
            }
          }
          if (can) legals.emplace_back(RV_RIICHI, -1, std::vector<uint8_t>{}, pid);
        }
      }
      // 3. Kan
      if (drawable_count > 0 && drawn_tile >= 0) {
        int counts[34] = {0};
        for (uint8_t t : P.hand) counts[t / 4]++;
        if (!P.riichi_declared && !P.riichi_stage) {
          for (int tv = 0; tv < 34; tv++)
            if (counts[tv] == 4) {
              uint8_t lo = (uint8_t)(tv * 4);
              legals.emplace_back(RV_ANKAN, lo, std::vector<uint8_t>{lo, (uint8_t)(lo + 1), (uint8_t)(lo + 2), (uint8_t)(lo + 3)}, pid);
            }
          for (auto& m : P.melds)
            if (m.meld_type == Pon) {
              uint8_t target = m.tiles[0] / 4;
              for (uint8_t t : P.hand)
                if (t / 4 == target) legals.emplace_back(RV_KAKAN, t, m.tiles, pid);
            }
        } else if (P.riichi_declared) {
          uint8_t t = (uint8_t)drawn_tile, t34 = t / 4;
          if (counts[t34] == 4) {
            std::vector<uint8_t> pre = P.hand;
            vec_remove_first(pre, t);
            HandEvaluator cpre(pre, P.melds, sanma);
            auto wpre = cpre.get_waits_u8();
            std::vector<uint8_t> post;
            for (uint8_t x : P.hand)
              if (x / 4 != t34) post.push_back(x);
            std::vector<Meld> mpost = P.melds;
            uint8_t lo = t34 * 4;
            Meld am;
            am.meld_type = Ankan;
            am.tiles = {lo, (uint8_t)(lo + 1), (uint8_t)(lo + 2), (uint8_t)(lo + 3)};
            am.opened = false;
            mpost.push_back(am);
            HandEvaluator cpost(post, mpost, sanma);
            auto wpost = cpost.get_waits_u8();
            if (wpre == wpost && !wpre.empty())
              legals.emplace_back(RV_ANKAN, lo, am.tiles, pid);
          }
        }
      }
      // 4. Kyushu kyuhai
      bool no_calls = true;
      for (auto& p : players)
        if (!p.melds.empty()) no_calls = false;
      if (is_first_turn && no_calls && !P.riichi_stage) {
        bool seen[34] = {false};
        int distinct = 0;
        for (uint8_t t : P.hand)
          if (is_terminal_tile(t) && !seen[t / 4]) {
            seen[t / 4] = true;
            distinct++;
          }
        if (distinct >= 9) legals.emplace_back(RV_KYUSHU_KYUHAI, -1, std::vector<uint8_t>{}, pid);
      }
      // 5. Kita (state_3p/legal_actions.rs:240-243, sanma.rs:146-169)
      if (sanma && drawn_tile >= 0 && drawable_count > 0)
        for (uint8_t t : P.hand)
          if (t / 4 == 30) legals.emplace_back(RV_KITA, t, std::vector<uint8_t>{}, pid);
    } else {
      if (has_claims_entry[pid])
        for (auto& a : current_claims[pid]) legals.push_back(a);
      legals.emplace_back(RV_PASS, -1, std::vector<uint8_t>{}, pid);
    }
    return legals;
  }

  // state/legal_actions.rs:254-508
  std::pair<std::vector<Action>, bool> _get_claim_actions_for_player(int i, int pid, uint8_t tile) const {
    std::vector<Action> legals;
    bool missed_agari = false;
    const PlayerState& P = players[i];
    const auto& hand = P.hand;
    uint8_t tile_class = tile / 4;
    bool in_discards = false;
    for (uint8_t d : P.discards)
      if (d / 4 == tile_class) in_discards = true;
    bool in_missed = P.missed_agari_doujun || (P.riichi_declared && P.missed_agari_riichi);
    if (!in_discards && !in_missed) {
      HandEvaluator calc(hand, P.melds, sanma);
      Conditions cond = base_cond(i);
      cond.houtei = drawable_count == 0 && !is_rinshan_flag;
      bool furiten = false;
      for (uint8_t w : calc.get_waits_u8()) {
        for (uint8_t d : P.discards)
          if (d / 4 == w) furiten = true;
        if (furiten) break;
      }
      if (P.missed_agari_riichi || P.missed_agari_doujun) furiten = true;
      if (!furiten) {
        WinResult res = calc.calc(tile, dora_indicators, {}, cond);
        if (res.is_win)
          legals.emplace_back(RV_RON, tile, std::vector<uint8_t>{}, i);
        else if (res.has_win_shape)
          missed_agari = true;
      }
    }
    // 2. Pon / Kan
    if (!P.riichi_declared && drawable_count > 0) {
      std::vector<uint8_t> matching;
      for (uint8_t t : hand)
        if (t / 4 == tile / 4) matching.push_back(t);
      size_t count = matching.size();
      if (count >= 2 && hand.size() >= 3) {
        auto check_pon_kuikae = [&](const std::vector<uint8_t>& consumes) {
          std::vector<bool> used(consumes.size(), false);
          for (uint8_t t : hand) {
            bool consumed_this = false;
            for (size_t k = 0; k < consumes.size(); k++)
              if (!used[k] && consumes[k] == t) {
                used[k] = true;
                consumed_this = true;
                break;
              }
            if (consumed_this) continue;
            bool forb = rb(RV_RULE_KUIKAE_FORBIDDEN) && (t / 4 == tile / 4);
            if (!forb) return true;
          }
          return false;
        };
        for (size_t a = 0; a < matching.size(); a++)
          for (size_t b = a + 1; b < matching.size(); b++) {
            std::vector<uint8_t> consumes = {matching[a], matching[b]};
            if (check_pon_kuikae(consumes)) legals.emplace_back(RV_PON, tile, consumes, i);
          }
      }
      if (count >= 3) {
        std::vector<uint8_t> consumes(matching.begin(), matching.begin() + 3);
        legals.emplace_back(RV_DAIMINKAN, tile, consumes, i);
      }
    }
    // 3. Chi
    bool is_shimocha = !sanma && i == (pid + 1) % 4;   // no chi in 3P (state_3p/legal_actions.rs:386)
    if (!P.riichi_declared && drawable_count > 0 && is_shimocha && hand.size() >= 3) {
      int t_val = tile / 4;
      if (t_val < 27) {
        auto check_chi_kuikae = [&](uint8_t c1, uint8_t c2) {
          int forb[2] = {-1, -1};
          if (rb(RV_RULE_KUIKAE_FORBIDDEN)) {
            forb[0] = t_val;
            int a = c1 / 4, b = c2 / 4;
            if (a > b) std::swap(a, b);
            if (a == t_val + 1 && b == t_val + 2) {
              if (t_val % 9 <= 5) forb[1] = t_val + 3;
            } else if (t_val >= 2 && b == t_val - 1 && a == t_val - 2 && t_val % 9 >= 3) {
              forb[1] = t_val - 3;
            }
          }
          bool used1 = false, used2 = false;
          for (uint8_t t : hand) {
            if (!used1 && t == c1) {
              used1 = true;
              continue;
            }
            if (!used2 && t == c2) {
              used2 = true;
              continue;
            }
            int tt = t / 4;
            if (tt != forb[0] && tt != forb[1]) return true;
          }
          return false;
        };
        auto pattern = [&](int ta, int tb) {
          std::vector<uint8_t> o1, o2;
          for (uint8_t t : hand) {
            if (t / 4 == ta) o1.push_back(t);
          }
          for (uint8_t t : hand) {
            if (t / 4 == tb) o2.push_back(t);
          }
          for (uint8_t c1 : o1)
            for (uint8_t c2 : o2)
              if (check_chi_kuikae(c1, c2)) legals.emplace_back(RV_CHI, tile, std::vector<uint8_t>{c1, c2}, i);
        };
        if (t_val % 9 >= 2) pattern(t_val - 2, t_val - 1);
        if (t_val % 9 >= 1 && t_val % 9 <= 7) pattern(t_val - 1, t_val + 1);
        if (t_val % 9 <= 6) pattern(t_val + 1, t_val + 2);
      }
    }
    return {legals, missed_agari};
  }

  // ------------------------------------------------------------ step
  // state/mod.rs:344-392: fuzzy legality match
  static bool action_matches(const Action& l, const Action& act) {
    if (l.type != act.type) return false;
    bool tiles_match = l.tile == act.tile;
    bool consumes_match = l.consume == act.consume;
    if (tiles_match) {
      if (consumes_match) return true;
      if (act.consume.empty() && l.type == RV_KAKAN) return true;
      if (act.consume.empty() &&
          (l.type == RV_DISCARD || l.type == RV_RIICHI || l.type == RV_TSUMO || l.type == RV_RON || l.type == RV_PASS))
        return true;
    }
    if (consumes_match && (l.type == RV_ANKAN || l.type == RV_KAKAN)) return true;
    if (act.tile < 0)
      return l.type == RV_TSUMO || l.type == RV_RON || l.type == RV_RIICHI || l.type == RV_KYUSHU_KYUHAI || l.type == RV_KITA;
    return false;
  }

  void cap_double_yakuman(WinResult& res, bool is_oya, bool tsumo, uint32_t hb) const {
    // state/mod.rs:720-745 / 1009-1034
    if (res.yakuman && res.han > 13) {
      uint32_t cap = 0;
      for (uint32_t y : res.yaku) {
        if (y == 47 && !rb(RV_RULE_JUNSEI_CHUUREN_DOUBLE)) cap += 13;
        if (y == 48 && !rb(RV_RULE_SUUANKOU_TANKI_DOUBLE)) cap += 13;
        if (y == 49 && !rb(RV_RULE_KOKUSHI13_DOUBLE)) cap += 13;
        if (y == 50 && !rb(RV_RULE_DAISUUSHII_DOUBLE)) cap += 13;
      }
      if (cap > 0) {
        uint32_t h = res.han > cap ? res.han - cap : 0;
        res.han = std::max<uint32_t>(h, 13);
        Score c = calculate_score((uint8_t)res.han, 0, is_oya, tsumo, hb, np);
        res.ron_agari = c.pay_ron;
        res.tsumo_agari_oya = c.pay_tsumo_oya;
        res.tsumo_agari_ko = c.pay_tsumo_ko;
      }
    }
  }
  int yakuman_val(uint32_t yid) const {
    if (yid == 47 && rb(RV_RULE_JUNSEI_CHUUREN_DOUBLE)) return 2;
    if (yid == 48 && rb(RV_RULE_SUUANKOU_TANKI_DOUBLE)) return 2;
    if (yid == 49 && rb(RV_RULE_KOKUSHI13_DOUBLE)) return 2;
    if (yid == 50 && rb(RV_RULE_DAISUUSHII_DOUBLE)) return 2;
    return 1;
  }
  int pao_for(int pid, uint32_t yid) const {
    if (yid == 37) return players[pid].pao37;
    if (yid == 50) return players[pid].pao50;
    return -1;
  }
  void ev_hora(int actor, int target, bool tsumo, const WinResult& res, const int32_t* deltas, bool with_ura) {
    std::vector<uint8_t> ura;
    if (with_ura) ura = _get_ura_indicators();
    uint32_t w[10];
    int nw = 6 + np;   // 4P: 10 words, 3P: 9 words (np deltas)
    w[0] = w0(RV_EV_HORA, nw, actor, target);
    w[1] = (tsumo ? 1u : 0u) | ((uint32_t)ura.size() << 8) | ((res.han & 0xFF) << 16) | ((res.fu & 0xFF) << 24);
    uint8_t ub[8];
    memset(ub, 0xFF, 8);
    for (size_t i = 0; i < ura.size() && i < 5; i++) ub[i] = ura[i];
    ub[5] = res.yakuman ? 1 : 0;
    ub[6] = ub[7] = 0;
    memcpy(&w[2], ub, 8);
    for (int i = 0; i < np; i++) w[4 + i] = (uint32_t)deltas[i];
    uint64_t mask = 0;
    for (uint32_t y : res.yaku) mask |= 1ull << y;
    w[4 + np] = (uint32_t)mask;
    w[5 + np] = (uint32_t)(mask >> 32);
    push_words(w, nw);
    if (keep_log) {   // mod.rs:867-884 (tsumo), 1115-1131 (ron)
      JsonObj ev;
      ev.str("type", "hora").num("actor", actor).num("target", target).nums("deltas", deltas, deltas + np);
      if (tsumo) ev.boolean("tsumo", true);
      std::vector<std::string> um;
      for (uint8_t t : ura) um.push_back(mjai_tile(t));
      text.np = np;
      text.push(ev.strs("ura_markers", um));
    }
  }

  // acts[pid].has_value() <=> key present in the reference's HashMap
  void step(const std::optional<Action> acts[MAXP]) {
    if (is_done) return;
    step_count++;
    // validation (state/mod.rs:340-402)
    for (int pid = 0; pid < np; pid++) {
      if (!acts[pid]) continue;
      auto legals = _get_legal_actions_internal(pid);
      bool ok = false;
      for (auto& l : legals)
        if (action_matches(l, *acts[pid])) {
          ok = true;
          break;
        }
      if (!ok) {
        last_error = pid;
        _trigger_ryukyoku(RV_RK_ILLEGAL_BASE + pid);
        return;
      }
    }
    if (phase == RV_WAIT_ACT) {
      int pid = current_player;
      if (!acts[pid]) return;
      const Action& act = *acts[pid];
      PlayerState& P = players[pid];
      switch (act.type) {
        case RV_DISCARD: {
          if (act.tile < 0) break;
          uint8_t tile = (uint8_t)act.tile;
          bool tsumogiri = false, valid = false;
          if (drawn_tile >= 0 && drawn_tile == tile) {
            tsumogiri = true;
            valid = true;
          }
          if (vec_remove_first(P.hand, tile)) {
            std::sort(P.hand.begin(), P.hand.end());
            valid = true;
            if (drawn_tile >= 0 && drawn_tile == tile) tsumogiri = true;
          }
          if (valid) _resolve_discard(pid, tile, tsumogiri);
          break;
        }
        case RV_KYUSHU_KYUHAI:
          _trigger_ryukyoku(RV_RK_KYUSHU);
          break;
        case RV_RIICHI: {
          if (P.score >= 1000 && (sanma ? drawable_count > 0 : drawable_count >= 4) && !P.riichi_declared && !P.riichi_stage) {
            P.riichi_stage = true;
            ev_simple(RV_EV_REACH, pid);
            if (act.tile >= 0) {
              uint8_t t = (uint8_t)act.tile;
              bool tsumogiri = drawn_tile >= 0 && drawn_tile == t;
              riichi_sutehais[pid] = t;
              if (!tsumogiri) last_tedashis[pid] = t;
              if (vec_remove_first(P.hand, t)) std::sort(P.hand.begin(), P.hand.end());
              _resolve_discard(pid, t, tsumogiri);
            }
          }
          break;
        }
        case RV_ANKAN: {
          uint8_t tile = act.tile >= 0 ? (uint8_t)act.tile : (act.consume.empty() ? 0 : act.consume[0]);
          std::vector<uint8_t> ronners;
          if (rb(RV_RULE_RON_ON_ANKAN_KOKUSHI)) {
            for (int i = 0; i < np; i++) {
              if (i == pid) continue;
              bool in_disc = false;
              for (uint8_t d : players[i].discards)
                if (d / 4 == tile / 4) in_disc = true;
              if (in_disc) continue;
              Conditions cond;
              cond.riichi = players[i].riichi_declared;
              cond.chankan = true;
              cond.player_wind = (uint8_t)((i + np - oya) % np);
              cond.round_wind = (uint8_t)(round_wind % 4);
              cond.is_sanma = sanma;
              cond.num_players = (uint8_t)np;
              HandEvaluator calc(players[i].hand, players[i].melds, sanma);
              WinResult res = calc.calc(tile, dora_indicators, {}, cond);
              bool kok = false;
              for (uint32_t y : res.yaku)
                if (y == 42 || y == 49) kok = true;
              if (res.is_win && kok) {
                ronners.push_back((uint8_t)i);
                has_claims_entry[i] = true;
                current_claims[i].emplace_back(RV_RON, tile, std::vector<uint8_t>{}, i);
              }
            }
          }
          if (!ronners.empty()) {
            pending_kan = true;
            pending_kan_pid = (uint8_t)pid;
            pending_kan_act = act;
            phase = RV_WAIT_RESPONSE;
            active_players = ronners;
            last_discard_pid = pid;
            last_discard_tile = tile;
          } else {
            _resolve_kan(pid, act);
          }
          break;
        }
        case RV_KAKAN: {
          uint8_t tile = act.tile >= 0 ? (uint8_t)act.tile : (act.consume.empty() ? 0 : act.consume[0]);
          vec_remove_first(P.hand, tile);
          for (auto& m : P.melds)
            if (m.meld_type == Pon && m.tiles[0] / 4 == tile / 4) {
              m.meld_type = Kakan;
              m.tiles.push_back(tile);
              std::sort(m.tiles.begin(), m.tiles.end());
              break;
            }
          {
            uint8_t c[4] = {0xFF, 0xFF, 0xFF, 0xFF};
            for (size_t k = 0; k < act.consume.size() && k < 4; k++) c[k] = act.consume[k];
            uint32_t w[2] = {w0(RV_EV_KAKAN, 2, pid, tile), 0};
            memcpy(&w[1], c, 4);
            push_words(w, 2);
            text_meld("kakan", pid, -1, tile, act.consume);
          }
          while (pending_kan_dora_count > 0) {
            pending_kan_dora_count--;
            _reveal_kan_dora();
          }
          std::vector<uint8_t> ronners;
          for (int i = 0; i < np; i++) {
            if (i == pid) continue;
            Conditions cond = base_cond(i);
            cond.chankan = true;
            HandEvaluator calc(players[i].hand, players[i].melds, sanma);
            bool furiten = false;
            for (uint8_t w : calc.get_waits_u8()) {
              for (uint8_t d : players[i].discards)
                if (d / 4 == w) furiten = true;
              if (furiten) break;
            }
            if (players[i].missed_agari_riichi || players[i].missed_agari_doujun) furiten = true;
            WinResult res;
            if (!furiten) res = calc.calc(tile, dora_indicators, {}, cond);
            if (res.is_win && (res.yakuman || res.han >= 1)) {
              ronners.push_back((uint8_t)i);
              has_claims_entry[i] = true;
              current_claims[i].emplace_back(RV_RON, tile, std::vector<uint8_t>{}, i);
            }
          }
          if (!ronners.empty()) {
            pending_kan = true;
            pending_kan_pid = (uint8_t)pid;
            pending_kan_act = act;
            phase = RV_WAIT_RESPONSE;
            active_players = ronners;
            last_discard_pid = pid;
            last_discard_tile = tile;
          } else {
            _resolve_kan(pid, act);
          }
          break;
        }
        case RV_TSUMO: {
          Conditions cond = base_cond(pid);
          cond.tsumo = true;
          cond.haitei = drawable_count == 0 && !is_rinshan_flag;
          cond.rinshan = is_rinshan_flag;
          bool all_meldless = true;
          for (auto& p : players)
            if (!p.melds.empty()) all_meldless = false;
          cond.tsumo_first_turn = is_first_turn && all_meldless;
          cond.kita_count = n_kita[pid];   // state_3p/mod.rs:635
          HandEvaluator calc(P.hand, P.melds, sanma);
          uint8_t win_tile = drawn_tile >= 0 ? (uint8_t)drawn_tile : 0;
          std::vector<uint8_t> ura;
          if (P.riichi_declared) ura = _get_ura_indicators();
          WinResult res = calc.calc(win_tile, dora_indicators, ura, cond);
          cap_double_yakuman(res, pid == oya, true, cond.honba);
          if (res.is_win) {
            int32_t deltas[MAXP] = {0, 0, 0, 0};
            int32_t total_win = 0;
            int pao_payer = -1;
            int pao_val = 0, total_val = 0;
            if (res.yakuman)
              for (uint32_t y : res.yaku) {
                int v = yakuman_val(y);
                total_val += v;
                int liable = pao_for(pid, y);
                if (liable >= 0) {
                  pao_val += v;
                  pao_payer = liable;
                }
              }
            if (pao_val > 0) {
              stat_pao++;
              // state_3p/mod.rs:713-721: per-yakuman tsumo total depends on the player count
              int32_t unit = pid == oya ? (np - 1) * 16000 : 16000 + (np - 2) * 8000;
              int32_t honba_total = (int32_t)honba * (np - 1) * 100;
              if (pao_payer >= 0) {
                if (rb(RV_RULE_PAO_LIABILITY_ONLY)) {
                  int32_t pao_amt = pao_val * unit + honba_total;
                  int non_pao = total_val - pao_val;
                  deltas[pao_payer] -= pao_amt;
                  total_win += pao_amt;
                  if (non_pao > 0) {
                    if (pid == oya) {
                      int32_t share = non_pao * 16000;
                      for (int i = 0; i < np; i++)
                        if (i != pid) {
                          deltas[i] -= share;
                          total_win += share;
                        }
                    } else {
                      for (int i = 0; i < np; i++)
                        if (i != pid) {
                          int32_t pay = (i == oya) ? non_pao * 16000 : non_pao * 8000;
                          deltas[i] -= pay;
                          total_win += pay;
                        }
                    }
                  }
                } else {
                  int32_t full = total_val * unit + honba_total;
                  deltas[pao_payer] -= full;
                  total_win += full;
                }
              }
            } else {
              for (int i = 0; i < np; i++)
                if (i != pid) {
                  int32_t pay = (pid == oya || i != oya) ? (int32_t)res.tsumo_agari_ko : (int32_t)res.tsumo_agari_oya;
                  deltas[i] = -pay;
                  total_win += pay;
                }
            }
            total_win += (int32_t)(riichi_sticks * 1000);
            riichi_sticks = 0;
            deltas[pid] += total_win;
            for (int i = 0; i < np; i++) {
              players[i].score += deltas[i];
              players[i].score_delta = deltas[i];
            }
            for (uint32_t y : res.yaku)
              if (pao_for(pid, y) >= 0) {
                res.pao_payer = pao_for(pid, y);
                break;
              }
            win_results.push_back({pid, res});
            ev_hora(pid, pid, true, res, deltas, P.riichi_declared);
            _initialize_next_round(pid == oya, false);
          } else {
            current_player = (current_player + 1) % np;
            _deal_next();
          }
          break;
        }
        case RV_KITA:
          if (sanma) handle_kita(pid, act);
          break;
        default:
          break;
      }
    } else {  // WaitResponse (state/mod.rs:900-1314)
      for (int pid = 0; pid < np; pid++) {
        if (!has_claims_entry[pid]) continue;
        bool has_ron = false;
        for (auto& a : current_claims[pid])
          if (a.type == RV_RON) has_ron = true;
        if (has_ron) {
          bool roned = acts[pid] && acts[pid]->type == RV_RON;
          if (!roned) {
            players[pid].missed_agari_doujun = true;
            if (players[pid].riichi_declared) players[pid].missed_agari_riichi = true;
          }
        }
      }
      std::vector<uint8_t> ron_claims;
      int call_pid = -1;
      Action call_act;
      for (uint8_t pid : active_players) {
        if (!acts[pid]) continue;
        const Action& act = *acts[pid];
        if (act.type == RV_RON) {
          ron_claims.push_back(pid);
        } else if (act.type == RV_PON || act.type == RV_DAIMINKAN || act.type == RV_CHI) {
          if (call_pid >= 0) {
            bool old_is_pon = call_act.type == RV_PON || call_act.type == RV_DAIMINKAN;
            bool new_is_pon = act.type == RV_PON || act.type == RV_DAIMINKAN;
            if (!old_is_pon && new_is_pon) {
              call_pid = pid;
              call_act = act;
            }
          } else {
            call_pid = pid;
            call_act = act;
          }
        }
      }
      if (!ron_claims.empty()) {
        if (!sanma && (int)ron_claims.size() >= np - 1 && rb(RV_RULE_SANCHAHO_IS_DRAW)) {   // 3P has no sanchaho branch
          _trigger_ryukyoku(RV_RK_SANCHAHO);
          return;
        }
        int target_pid = last_discard_pid >= 0 ? last_discard_pid : current_player;
        uint8_t win_tile = last_discard_pid >= 0 ? (uint8_t)last_discard_tile : 0;
        std::stable_sort(ron_claims.begin(), ron_claims.end(), [&](uint8_t a, uint8_t b) {
          return (a + np - target_pid) % np < (b + np - target_pid) % np;
        });
        int32_t total_deltas[MAXP] = {0, 0, 0, 0};
        bool oya_won = false, deposit_taken = false, honba_taken = false;
        for (uint8_t w_pid : ron_claims) {
          PlayerState& W = players[w_pid];
          bool is_chankan = pending_kan && pending_kan_act.type != RV_KITA;   // state_3p/mod.rs:896-902
          uint32_t ron_honba = 0;
          if (!honba_taken) {
            honba_taken = true;
            ron_honba = honba;
          }
          Conditions cond = base_cond(w_pid);
          cond.houtei = drawable_count == 0 && !is_rinshan_flag;
          cond.chankan = is_chankan;
          cond.honba = ron_honba;
          cond.kita_count = n_kita[w_pid];   // state_3p/mod.rs:926
          HandEvaluator calc(W.hand, W.melds, sanma);
          std::vector<uint8_t> ura;
          if (W.riichi_declared) ura = _get_ura_indicators();
          WinResult res = calc.calc(win_tile, dora_indicators, ura, cond);
          cap_double_yakuman(res, w_pid == oya, false, ron_honba);
          if (res.is_win) {
            int32_t score = (int32_t)res.ron_agari;
            int pao_payer = target_pid;
            int32_t pao_amt = 0;
            if (res.yakuman) {
              bool has_pao = false;
              int total_val = 0, pao_val = 0;
              for (uint32_t y : res.yaku) {
                int v = yakuman_val(y);
                total_val += v;
                int liable = pao_for(w_pid, y);
                if (liable >= 0) {
                  has_pao = true;
                  pao_payer = liable;
                  pao_val += v;
                }
              }
              if (has_pao) {
                stat_pao++;
                int32_t unit = (w_pid == oya) ? 48000 : 32000;
                int32_t honba_ron = (int32_t)ron_honba * (np - 1) * 100;
                int32_t split_base = rb(RV_RULE_PAO_LIABILITY_ONLY) ? pao_val * unit : total_val * unit;
                pao_amt = split_base / 2 + honba_ron;
              }
            }
            int32_t this_d[MAXP] = {0, 0, 0, 0};
            this_d[w_pid] += score;
            this_d[pao_payer] -= pao_amt;
            this_d[target_pid] -= score - pao_amt;
            total_deltas[w_pid] += score;
            total_deltas[pao_payer] -= pao_amt;
            total_deltas[target_pid] -= score - pao_amt;
            if (!deposit_taken) {
              int32_t stick = (int32_t)(riichi_sticks * 1000);
              total_deltas[w_pid] += stick;
              this_d[w_pid] += stick;
              riichi_sticks = 0;
              deposit_taken = true;
            }
            for (uint32_t y : res.yaku)
              if (pao_for(w_pid, y) >= 0) {
                res.pao_payer = pao_for(w_pid, y);
                break;
              }
            win_results.push_back({w_pid, res});
            if (w_pid == oya) oya_won = true;
            ev_hora(w_pid, target_pid, false, res, this_d, W.riichi_declared);
          }
        }
        for (int i = 0; i < np; i++) {
          players[i].score += total_deltas[i];
          players[i].score_delta = total_deltas[i];
        }
        _initialize_next_round(oya_won, false);
      } else if (call_pid >= 0) {
        int claimer = call_pid;
        const Action& action = call_act;
        PlayerState& C = players[claimer];
        _accept_riichi();
        is_rinshan_flag = false;
        is_first_turn = false;
        C.missed_agari_doujun = false;
        if (last_discard_pid >= 0) players[last_discard_pid].nagashi_eligible = false;
        for (auto& p : players) p.ippatsu_cycle = false;
        if (action.type == RV_DAIMINKAN) {
          current_player = (uint8_t)claimer;
          active_players = {(uint8_t)claimer};
          C.forbidden_discards.clear();
          _resolve_kan(claimer, action);
          return;
        }
        for (uint8_t t : action.consume) vec_remove_first(C.hand, t);
        int discarder = last_discard_pid;
        uint8_t tile = (uint8_t)last_discard_tile;
        std::vector<uint8_t> tiles = action.consume;
        tiles.push_back(tile);
        std::sort(tiles.begin(), tiles.end());
        Meld m;
        m.meld_type = action.type == RV_PON ? Pon : Chi;
        m.tiles = tiles;
        m.opened = true;
        m.from_who = (int8_t)discarder;
        m.called_tile = tile;
        C.melds.push_back(m);
        {
          uint8_t c[4] = {(uint8_t)discarder, 0xFF, 0xFF, 0xFF};
          for (size_t k = 0; k < action.consume.size() && k < 3; k++) c[1 + k] = action.consume[k];
          uint32_t w[2] = {w0(action.type == RV_PON ? RV_EV_PON : RV_EV_CHI, 2, claimer, tile), 0};
          memcpy(&w[1], c, 4);
          push_words(w, 2);
          text_meld(action.type == RV_PON ? "pon" : "chi", claimer, discarder, tile, action.consume);
        }
        if (m.meld_type == Pon) register_pao(claimer, tile, discarder);
        current_player = (uint8_t)claimer;
        phase = RV_WAIT_ACT;
        active_players = {(uint8_t)claimer};
        C.forbidden_discards.clear();
        if (action.type == RV_PON) {
          C.forbidden_discards.push_back(tile);
        } else {
          C.forbidden_discards.push_back(tile);
          int t34 = tile / 4;
          std::vector<int> c34;
          for (uint8_t x : action.consume) c34.push_back(x / 4);
          std::sort(c34.begin(), c34.end());
          if (c34[0] == t34 + 1 && c34[1] == t34 + 2) {
            if (t34 % 9 <= 5) C.forbidden_discards.push_back((uint8_t)((t34 + 3) * 4));
          } else if (t34 >= 2 && c34[1] == t34 - 1 && c34[0] == t34 - 2 && t34 % 9 >= 3) {
            C.forbidden_discards.push_back((uint8_t)((t34 - 3) * 4));
          }
        }
        needs_tsumo = false;
        drawn_tile = -1;
      } else {
        for (int i = 0; i < np; i++) {
          current_claims[i].clear();
          has_claims_entry[i] = false;
        }
        active_players.clear();
        if (pending_kan) {
          pending_kan = false;
          Action a = pending_kan_act;
          if (a.type == RV_KITA) {   // state_3p/mod.rs:1201-1207
            for (auto& p : players) p.ippatsu_cycle = false;
            resolve_kita_rinshan(pending_kan_pid);
          } else {
            _resolve_kan(pending_kan_pid, a);
          }
        } else {
          _accept_riichi();
          turn_count += 1;
          current_player = (current_player + 1) % np;
          _deal_next();
          if (turn_count >= (uint32_t)np) is_first_turn = false;
        }
      }
    }
  }

  // state/mod.rs:1229-1259 / 1444-1472
  void register_pao(int claimer, uint8_t tile, int discarder) {
    int tv = tile / 4;
    auto& ms = players[claimer].melds;
    if (tv >= 31 && tv <= 33) {
      int n = 0;
      for (auto& m : ms) {
        int t = m.tiles[0] / 4;
        if (t >= 31 && t <= 33 && m.meld_type != Chi) n++;
      }
      if (n == 3) players[claimer].pao37 = discarder;
    } else if (tv >= 27 && tv <= 30) {
      int n = 0;
      for (auto& m : ms) {
        int t = m.tiles[0] / 4;
        if (t >= 27 && t <= 30 && m.meld_type != Chi) n++;
      }
      if (n == 4) players[claimer].pao50 = discarder;
    }
  }

  // state/mod.rs:1317-1413
  void _resolve_discard(int pid, uint8_t tile, bool tsumogiri) {
    PlayerState& P = players[pid];
    if (sanma) pending_kan = false;   // state_3p/mod.rs:1224-1227
    is_rinshan_flag = false;
    P.ippatsu_cycle = false;
    P.discards.push_back(tile);
    last_discard_pid = pid;
    last_discard_tile = tile;
    drawn_tile = -1;
    P.discard_from_hand.push_back(!tsumogiri);
    P.discard_is_riichi.push_back(P.riichi_stage);
    if (!tsumogiri) last_tedashis[pid] = tile;
    needs_tsumo = true;
    if (P.riichi_stage) {
      P.riichi_declared = true;
      if (is_first_turn) P.double_riichi_declared = true;
      P.riichi_declaration_index = (int)P.discards.size() - 1;
      P.riichi_stage = false;
      riichi_pending_acceptance = pid;
    }
    while (pending_kan_dora_count > 0) {
      pending_kan_dora_count--;
      _reveal_kan_dora();
    }
    ev_simple(tsumogiri ? RV_EV_DAHAI_TSUMOGIRI : RV_EV_DAHAI, pid, tile);
    P.missed_agari_doujun = false;
    P.nagashi_eligible = P.nagashi_eligible && is_terminal_tile(tile);
    for (int i = 0; i < np; i++) {
      current_claims[i].clear();
      has_claims_entry[i] = false;
    }
    active_players.clear();
    bool has_claims = false;
    std::vector<uint8_t> claim_active;
    for (int i = 0; i < np; i++) {
      if (i == pid) continue;
      auto [legals, missed] = _get_claim_actions_for_player(i, pid, tile);
      if (missed) players[i].missed_agari_doujun = true;
      if (!legals.empty()) {
        has_claims = true;
        claim_active.push_back((uint8_t)i);
        current_claims[i] = legals;
        has_claims_entry[i] = true;
      }
    }
    if (has_claims) {
      phase = RV_WAIT_RESPONSE;
      active_players = claim_active;
    } else {
      if (riichi_pending_acceptance >= 0) _accept_riichi();
      if (!check_abortive_draw()) {
        turn_count += 1;
        current_player = (uint8_t)((pid + 1) % np);
        _deal_next();
        if (turn_count >= (uint32_t)np) is_first_turn = false;
      }
    }
  }

  // state/mod.rs:1415-1547
  void _resolve_kan(int pid, const Action& action) {
    PlayerState& P = players[pid];
    if (action.type != RV_KAKAN) {
      for (uint8_t t : action.consume) vec_remove_first(P.hand, t);
      Meld m;
      if (action.type == RV_ANKAN) {
        m.meld_type = Ankan;
        m.tiles = action.consume;
        m.from_who = -1;
        m.called_tile = -1;
        m.opened = false;
      } else {
        m.meld_type = Daiminkan;
        m.tiles = action.consume;
        m.tiles.push_back((uint8_t)last_discard_tile);
        std::sort(m.tiles.begin(), m.tiles.end());
        m.from_who = (int8_t)last_discard_pid;
        m.called_tile = last_discard_tile;
        m.opened = true;
      }
      P.melds.push_back(m);
      if (action.type == RV_DAIMINKAN) register_pao(pid, (uint8_t)last_discard_tile, last_discard_pid);
    }
    is_first_turn = false;
    for (auto& p : players) p.ippatsu_cycle = false;
    if (drawable_count > 0) {
      uint8_t t = wall_tiles.front();
      wall_tiles.erase(wall_tiles.begin());
      drawable_count -= 1;
      P.hand.push_back(t);
      drawn_tile = t;
      rinshan_draw_count += 1;
      is_rinshan_flag = true;
      if (action.type == RV_ANKAN) {
        uint8_t tile = action.tile >= 0 ? (uint8_t)action.tile : action.consume[0];
        uint8_t c[4] = {0xFF, 0xFF, 0xFF, 0xFF};
        for (size_t k = 0; k < action.consume.size() && k < 4; k++) c[k] = action.consume[k];
        uint32_t w[2] = {w0(RV_EV_ANKAN, 2, pid, tile), 0};
        memcpy(&w[1], c, 4);
        push_words(w, 2);
        text_meld("ankan", pid, -1, tile, action.consume);
      } else if (action.type == RV_DAIMINKAN) {
        uint8_t c[4] = {(uint8_t)last_discard_pid, 0xFF, 0xFF, 0xFF};
        for (size_t k = 0; k < action.consume.size() && k < 3; k++) c[1 + k] = action.consume[k];
        uint32_t w[2] = {w0(RV_EV_DAIMINKAN, 2, pid, last_discard_tile), 0};
        memcpy(&w[1], c, 4);
        push_words(w, 2);
        text_meld("daiminkan", pid, last_discard_pid, last_discard_tile, action.consume);
      }
      while (pending_kan_dora_count > 0) {
        pending_kan_dora_count--;
        _reveal_kan_dora();
      }
      if (action.type == RV_ANKAN)
        _reveal_kan_dora();
      else
        pending_kan_dora_count += 1;
      ev_simple(RV_EV_TSUMO, pid, t);
      phase = RV_WAIT_ACT;
      active_players = {(uint8_t)pid};
    }
  }

  // state_3p/sanma.rs:9-144
  void handle_kita(int pid, const Action& act) {
    PlayerState& P = players[pid];
    int tile = -1;
    if (act.tile >= 0 && act.tile / 4 == 30) {
      tile = act.tile;
    } else {
      for (uint8_t t : P.hand)
        if (t / 4 == 30) {
          tile = t;
          break;
        }
      if (tile < 0) tile = act.tile >= 0 ? act.tile : (act.consume.empty() ? 0 : act.consume[0]);
    }
    vec_remove_first(P.hand, (uint8_t)tile);
    n_kita[pid]++;
    is_first_turn = false;
    ev_simple(RV_EV_KITA, pid, tile);
    while (pending_kan_dora_count > 0) {
      pending_kan_dora_count--;
      _reveal_kan_dora();
    }
    std::vector<uint8_t> ronners;
    for (int i = 0; i < np; i++) {
      if (i == pid) continue;
      HandEvaluator calc(players[i].hand, players[i].melds, sanma);
      bool furiten = false;
      for (uint8_t w : calc.get_waits_u8()) {
        for (uint8_t d : players[i].discards)
          if (d / 4 == w) furiten = true;
        if (furiten) break;
      }
      if (players[i].missed_agari_riichi || players[i].missed_agari_doujun) furiten = true;
      if (furiten) continue;
      Conditions cond = base_cond(i);   // chankan stays false: kita does not award chankan
      cond.kita_count = n_kita[i];
      WinResult res = calc.calc((uint8_t)tile, dora_indicators, {}, cond);
      if (res.is_win && (res.yakuman || res.han >= 1)) {
        ronners.push_back((uint8_t)i);
        has_claims_entry[i] = true;
        current_claims[i].emplace_back(RV_RON, tile, std::vector<uint8_t>{}, i);
      }
    }
    if (!ronners.empty()) {
      phase = RV_WAIT_RESPONSE;
      active_players = ronners;
      last_discard_pid = pid;
      last_discard_tile = tile;
      pending_kan = true;
      pending_kan_pid = (uint8_t)pid;
      pending_kan_act = Action(RV_KITA, tile, {}, pid);
    } else {
      for (auto& p : players) p.ippatsu_cycle = false;
      resolve_kita_rinshan(pid);
    }
  }
  // state_3p/sanma.rs:171-204
  void resolve_kita_rinshan(int pid) {
    if (drawable_count > 0) {
      while (pending_kan_dora_count > 0) {
        pending_kan_dora_count--;
        _reveal_kan_dora();
      }
      if (wall_tiles.empty()) return;
      uint8_t t = wall_tiles.front();
      wall_tiles.erase(wall_tiles.begin());
      drawable_count = drawable_count > 0 ? drawable_count - 1 : 0;
      players[pid].hand.push_back(t);
      drawn_tile = t;
      rinshan_draw_count += 1;
      is_rinshan_flag = true;
      ev_simple(RV_EV_TSUMO, pid, t);
      phase = RV_WAIT_ACT;
      active_players = {(uint8_t)pid};
    }
  }

  // state/mod.rs:1549-1567
  void _accept_riichi() {
    if (riichi_pending_acceptance >= 0) {
      int p = riichi_pending_acceptance;
      players[p].score -= 1000;
      players[p].score_delta -= 1000;
      riichi_sticks += 1;
      players[p].riichi_declared = true;
      players[p].ippatsu_cycle = true;
      ev_simple(RV_EV_REACH_ACCEPTED, p);
      riichi_pending_acceptance = -1;
    }
  }

  // state/mod.rs:1569-1593
  void _deal_next() {
    is_rinshan_flag = false;
    if (drawable_count == 0) {
      _trigger_ryukyoku(RV_RK_EXHAUSTIVE);
      return;
    }
    if (!wall_tiles.empty()) {
      uint8_t t = wall_tiles.back();
      wall_tiles.pop_back();
      drawable_count -= 1;
      int pid = current_player;
      players[pid].hand.push_back(t);
      drawn_tile = t;
      needs_tsumo = false;
      phase = RV_WAIT_ACT;
      active_players = {(uint8_t)pid};
      ev_simple(RV_EV_TSUMO, pid, t);
      players[pid].forbidden_discards.clear();
    }
  }

  // state/mod.rs:2071-2081
  void _process_end_game() {
    is_done = true;
    ev_simple(RV_EV_END_KYOKU);
    ev_simple(RV_EV_END_GAME);
  }

  // state/mod.rs:1595-1688
  void _initialize_next_round(bool oya_won, bool is_draw) {
    if (is_done) return;
    for (int i = 0; i < np; i++)
      if (players[i].score < 0) {
        _process_end_game();
        return;
      }
    int32_t dealer_score = players[oya].score;
    bool dealer_is_top = true;
    for (int seat = 0; seat < np; seat++) {
      bool ok = seat == oya || dealer_score > players[seat].score || (dealer_score == players[seat].score && oya <= seat);
      if (!ok) dealer_is_top = false;
    }
    bool is_last_regular = false;
    if (game_mode == 1 || game_mode == 4) is_last_regular = round_wind == 0 && oya == np - 1;
    if (game_mode == 2 || game_mode == 5) is_last_regular = round_wind == 1 && oya == np - 1;
    if (oya_won && is_last_regular && dealer_is_top && dealer_score >= target_score()) {
      _process_end_game();
      return;
    }
    uint8_t next_honba = honba, next_oya = oya, next_rw = round_wind;
    if (oya_won) {
      next_honba = next_honba == 255 ? 255 : next_honba + 1;
    } else if (is_draw) {
      next_honba = next_honba == 255 ? 255 : next_honba + 1;
      next_oya = (next_oya + 1) % np;
      if (next_oya == 0) next_rw += 1;
    } else {
      next_honba = 0;
      next_oya = (next_oya + 1) % np;
      if (next_oya == 0) next_rw += 1;
    }
    int32_t max_score = players[0].score;
    for (int i = 0; i < np; i++) max_score = std::max(max_score, players[i].score);
    switch (game_mode) {
      case 1:
      case 4:
        if (next_rw >= 1 && (max_score >= target_score() || next_rw > 1)) {
          _process_end_game();
          return;
        }
        break;
      case 2:
      case 5:
        if (next_rw >= 2 && (max_score >= target_score() || next_rw > 2)) {
          _process_end_game();
          return;
        }
        break;
      case 0:
      case 3:
        _process_end_game();
        return;
      default:
        if (next_rw >= 1) {
          _process_end_game();
          return;
        }
    }
    ev_simple(RV_EV_END_KYOKU);
    int32_t sc[MAXP];
    for (int i = 0; i < np; i++) sc[i] = players[i].score;
    _initialize_round(next_oya, next_rw, next_honba, riichi_sticks, nullptr, sc);
  }

  // state/mod.rs:1846-1968   (reason: rv_ryukyoku_reason code)
  void _trigger_ryukyoku(int reason) {
    _accept_riichi();
    bool tenpai[MAXP] = {false, false, false, false};
    int final_reason = reason;
    std::vector<uint8_t> nagashi;
    if (reason == RV_RK_EXHAUSTIVE) {
      for (int i = 0; i < np; i++) {
        HandEvaluator calc(players[i].hand, players[i].melds, sanma);
        if (calc.is_tenpai()) tenpai[i] = true;
      }
      for (int i = 0; i < np; i++)
        if (players[i].nagashi_eligible) nagashi.push_back((uint8_t)i);
      if (!nagashi.empty()) {
        final_reason = RV_RK_NAGASHI;
        for (uint8_t w : nagashi) {
          bool is_oya = w == oya;
          Score sr = calculate_score(5, 30, is_oya, true, 0, np);
          for (int i = 0; i < np; i++) {
            if (i == w) continue;
            int32_t pay = (is_oya || i != oya) ? (int32_t)sr.pay_tsumo_ko : (int32_t)sr.pay_tsumo_oya;
            players[i].score -= pay;
            players[i].score_delta -= pay;
            players[w].score += pay;
            players[w].score_delta += pay;
          }
        }
      } else {
        int num_tp = 0;
        for (bool t : tenpai) num_tp += t;
        if (num_tp > 0 && num_tp < np) {
          int32_t pool = sanma ? 2000 : 3000;   // state_3p/game_mode.rs:39-41
          int32_t pk = pool / num_tp, pn = pool / (np - num_tp);
          for (int i = 0; i < np; i++) {
            int32_t d = tenpai[i] ? pk : -pn;
            players[i].score += d;
            players[i].score_delta = d;
          }
        }
      }
    } else if (reason >= RV_RK_ILLEGAL_BASE && reason - RV_RK_ILLEGAL_BASE < np) {
      int pid = reason - RV_RK_ILLEGAL_BASE;
      if (pid == oya) {
        int32_t penalty = 4000 * (np - 1), each = penalty / (np - 1);
        for (int i = 0; i < np; i++) {
          if (i == pid) {
            players[i].score -= penalty;
            players[i].score_delta = -penalty;
          } else {
            players[i].score += each;
            players[i].score_delta = each;
          }
        }
      } else {
        int32_t total = 4000 + 2000 * (np - 2);
        for (int i = 0; i < np; i++) {
          if (i == pid) {
            players[i].score -= total;
            players[i].score_delta = -total;
          } else if (i == oya) {
            players[i].score += 4000;
            players[i].score_delta = 4000;
          } else {
            players[i].score += 2000;
            players[i].score_delta = 2000;
          }
        }
      }
    }
    bool is_renchan;
    if (final_reason == RV_RK_EXHAUSTIVE)
      is_renchan = tenpai[oya];
    else if (final_reason == RV_RK_NAGASHI)
      is_renchan = std::find(nagashi.begin(), nagashi.end(), oya) != nagashi.end();
    else
      is_renchan = true;
    int32_t d[MAXP];
    for (int i = 0; i < np; i++) d[i] = players[i].score_delta;
    ev_deltas(RV_EV_RYUKYOKU, final_reason, d);
    _initialize_next_round(is_renchan, true);
  }

  // state/mod.rs:1970-2019
  bool check_abortive_draw() {
    bool turns_ok = !sanma, melds_empty = true;   // sufuurenta is disabled in 3P (state_3p/mod.rs:1860-1863)
    for (auto& p : players) {
      if (p.discards.size() != 1) turns_ok = false;
      if (!p.melds.empty()) melds_empty = false;
    }
    if (turns_ok && melds_empty && !players[0].discards.empty()) {
      int first = players[0].discards[0] / 4;
      if (first >= 27 && first <= 30) {
        bool all = true;
        for (auto& p : players)
          if (p.discards.empty() || p.discards[0] / 4 != first) all = false;
        if (all) {
          _trigger_ryukyoku(RV_RK_SUFUURENTA);
          return true;
        }
      }
    }
    std::vector<int> owners;
    for (int pid = 0; pid < np; pid++)
      for (auto& m : players[pid].melds)
        if (m.meld_type == Daiminkan || m.meld_type == Ankan || m.meld_type == Kakan) owners.push_back(pid);
    if (owners.size() == 4) {
      bool same = true;
      for (int o : owners)
        if (o != owners[0]) same = false;
      if (!same) {
        _trigger_ryukyoku(RV_RK_SUUKANSANSEN);
        return true;
      }
    }
    bool all_riichi = !sanma;   // suucha riichi is disabled in 3P (state_3p/mod.rs:1886-1888)
    for (int i = 0; i < np; i++)
      if (!players[i].riichi_declared) all_riichi = false;
    if (all_riichi) {
      _trigger_ryukyoku(RV_RK_SUUCHA_RIICHI);
      return true;
    }
    return false;
  }

  // env.rs:673-689
  void ranks(uint8_t out[MAXP]) const {
    int idx[MAXP] = {0, 1, 2, 3};
    std::stable_sort(idx, idx + np, [&](int a, int b) { return players[a].score > players[b].score; });
    for (int r = 0; r < np; r++) out[idx[r]] = (uint8_t)(r + 1);
  }

  // ------------------------------------------------------------ snapshot
  static uint32_t pack_claim(const Action& a) {
    uint32_t c0 = a.consume.size() > 0 ? a.consume[0] : 0xFF, c1 = a.consume.size() > 1 ? a.consume[1] : 0xFF;
    return (uint32_t)a.type | ((uint32_t)(a.tile & 0xFF) << 8) | (c0 << 16) | (c1 << 24);
  }
  void to_snapshot(rv_game_state& s) const {
    memset(&s, 0, sizeof s);
    memset(s.wall, 0xFF, sizeof s.wall);
    for (size_t i = 0; i < wall_abs.size() && i < 136; i++) s.wall[i] = wall_abs[i];
    s.wall_len = (uint8_t)wall_abs.size();
    s.wall_top = (uint8_t)(wall_tiles.size() + rinshan_draw_count);
    s.rinshan_draw_count = rinshan_draw_count;
    s.pending_kan_dora_count = pending_kan_dora_count;
    s.drawable_count = drawable_count;
    s.n_dora = (uint8_t)std::min<size_t>(dora_indicators.size(), 5);   // the record holds five; a replay can push more (overflow bit 0)
    if (dora_indicators.size() > 5) s.overflow |= 1;
    s.is_after_kan = is_after_kan;
    memset(s.dora_ind, 0xFF, 5);
    for (size_t i = 0; i < dora_indicators.size() && i < 5; i++) s.dora_ind[i] = dora_indicators[i];
    s.phase = phase;
    memset(s.hand, 0xFF, sizeof s.hand);
    memset(s.meld_tiles, 0xFF, sizeof s.meld_tiles);
    memset(s.meld_type, 0xFF, sizeof s.meld_type);
    memset(s.meld_from, 0xFF, sizeof s.meld_from);
    memset(s.meld_called, 0xFF, sizeof s.meld_called);
    memset(s.river, 0xFF, sizeof s.river);
    memset(s.forbidden, 0xFF, sizeof s.forbidden);
    memset(s.riichi_decl_idx, 0xFF, sizeof s.riichi_decl_idx);   // unused 3P seat slot reads as None everywhere
    memset(s.pao, 0xFF, sizeof s.pao);
    memset(s.riichi_sutehai, 0xFF, sizeof s.riichi_sutehai);
    memset(s.last_tedashi, 0xFF, sizeof s.last_tedashi);
    memset(s.claims, 0, sizeof s.claims);
    for (int p = 0; p < np; p++) {
      const PlayerState& P = players[p];
      s.hand_len[p] = (uint8_t)P.hand.size();
      for (size_t k = 0; k < P.hand.size() && k < RV_HAND_CAP; k++) s.hand[p][k] = P.hand[k];
      s.n_melds[p] = (uint8_t)P.melds.size();
      for (size_t m = 0; m < P.melds.size() && m < 4; m++) {
        for (size_t k = 0; k < P.melds[m].tiles.size() && k < 4; k++) s.meld_tiles[p][m][k] = P.melds[m].tiles[k];
        s.meld_type[p][m] = P.melds[m].meld_type;
        s.meld_from[p][m] = (uint8_t)P.melds[m].from_who;
        s.meld_called[p][m] = (uint8_t)(P.melds[m].called_tile < 0 ? 0xFF : P.melds[m].called_tile);
      }
      s.n_river[p] = (uint8_t)P.discards.size();
      for (size_t k = 0; k < P.discards.size(); k++) {
        if (k < RV_RIVER_CAP) {
          s.river[p][k] = P.discards[k];
          if (k < P.discard_from_hand.size() && P.discard_from_hand[k]) s.river_tedashi[p] |= 1u << k;   // (the event handler pushes discards only)
          if (k < P.discard_is_riichi.size() && P.discard_is_riichi[k]) s.river_riichi[p] |= 1u << k;
        } else {
          s.overflow = 1;
        }
      }
      s.riichi_decl_idx[p] = (uint8_t)(P.riichi_declaration_index < 0 ? 0xFF : P.riichi_declaration_index);
      s.flags[p] = (P.riichi_declared ? RV_F_RIICHI_DECLARED : 0) | (P.riichi_stage ? RV_F_RIICHI_STAGE : 0) |
                   (P.double_riichi_declared ? RV_F_DOUBLE_RIICHI : 0) | (P.missed_agari_riichi ? RV_F_MISSED_AGARI_RIICHI : 0) |
                   (P.missed_agari_doujun ? RV_F_MISSED_AGARI_DOUJUN : 0) | (P.nagashi_eligible ? RV_F_NAGASHI_ELIGIBLE : 0) |
                   (P.ippatsu_cycle ? RV_F_IPPATSU_CYCLE : 0);
      s.pao[p][0] = (uint8_t)(P.pao37 < 0 ? 0xFF : P.pao37);
      s.pao[p][1] = (uint8_t)(P.pao50 < 0 ? 0xFF : P.pao50);
      for (size_t k = 0; k < P.forbidden_discards.size() && k < 2; k++) s.forbidden[p][k] = P.forbidden_discards[k];
      s.riichi_sutehai[p] = (uint8_t)(riichi_sutehais[p] < 0 ? 0xFF : riichi_sutehais[p]);
      s.last_tedashi[p] = (uint8_t)(last_tedashis[p] < 0 ? 0xFF : last_tedashis[p]);
      s.score[p] = P.score;
      s.score_delta[p] = P.score_delta;
      s.n_claims[p] = has_claims_entry[p] ? (uint8_t)current_claims[p].size() : 0;
      for (size_t k = 0; k < current_claims[p].size() && k < RV_MAX_CLAIMS; k++) s.claims[p][k] = pack_claim(current_claims[p][k]);
    }
    s.current_player = current_player;
    s.oya = oya;
    s.honba = honba;
    s.kyoku_idx = kyoku_idx;
    s.round_wind = round_wind;
    s.is_done = is_done;
    s.needs_tsumo = needs_tsumo;
    s.is_first_turn = is_first_turn;
    s.is_rinshan_flag = is_rinshan_flag;
    s.riichi_pending_acceptance = (uint8_t)(riichi_pending_acceptance < 0 ? 0xFF : riichi_pending_acceptance);
    s.drawn_tile = (uint8_t)(drawn_tile < 0 ? 0xFF : drawn_tile);
    s.last_discard_pid = (uint8_t)(last_discard_pid < 0 ? 0xFF : last_discard_pid);
    s.last_discard_tile = (uint8_t)(last_discard_tile < 0 ? 0xFF : last_discard_tile);
    s.pending_kan_pid = pending_kan ? pending_kan_pid : 0xFF;
    s.pending_kan_type = pending_kan ? pending_kan_act.type : 0xFF;
    s.pending_kan_tile = pending_kan ? (uint8_t)(pending_kan_act.tile >= 0 ? pending_kan_act.tile : pending_kan_act.consume[0]) : 0xFF;
    for (int p = 0; p < np; p++) s.n_kita[p] = n_kita[p];
    s.active_mask = 0;
    for (uint8_t a : active_players) s.active_mask |= (uint8_t)(1u << a);
    s.last_error = (uint8_t)(last_error < 0 ? 0xFF : last_error);
    s.pending_init[0] = s.pending_init[1] = s.pending_init[2] = 0xFF;
    s.pending_tail[0] = 0xFF;
    s.pending_tail[1] = 0;
    s.game_mode = game_mode;
    s.overflow = stalled ? 2 : s.overflow;
    s.rule_bits = (uint8_t)rule;
    s.riichi_sticks = riichi_sticks;
    s.turn_count = turn_count;
    s.seed = wall_seed;
    s.hand_index = hand_index;
    s.step_count = step_count;
    s.kyoku_count = kyoku_count;
    s.ev_count = ev_count;
    s.ev_words = ev_words;
    // derived caches recomputed from scratch (the CUDA path maintains them incrementally)
    for (int p = 0; p < np; p++) {
      const PlayerState& P = players[p];
      static const uint32_t P5[9] = {1, 5, 25, 125, 625, 3125, 15625, 78125, 390625};
      for (uint8_t t : P.hand) {
        int kind = t / 4, su = kind / 9, pos = kind % 9;
        s.c_cnt[p][su] += 1ull << (4 * pos);
        s.c_key[p][su] += P5[pos];
      }
      for (uint8_t d : P.discards) s.c_river_kinds[p] |= 1ull << (d / 4);
      HandEvaluator he(P.hand, P.melds, sanma);
      for (uint8_t w : he.get_waits_u8()) s.c_waits[p] |= 1ull << w;
    }
    s.ev_hash = ev_hash;
  }
};

// keyed random agent shared with the CUDA path (SURVEY.md §8 d)
inline uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
inline uint32_t agent_pick(uint64_t agent_seed, uint64_t game_id, uint32_t step_count, int seat, uint32_t n_legal) {
  uint64_t k = agent_seed ^ (game_id * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)step_count << 8) ^ (uint64_t)seat;
  return (uint32_t)(mix64(k) % n_legal);
}

// One env step with the keyed agent. Returns false if the game is done.
inline bool random_step(GameState& g, uint64_t agent_seed, uint64_t game_id) {
  if (g.is_done) return false;
  std::optional<Action> acts[MAXP];
  uint32_t sc = g.step_count;
  if (g.phase == RV_WAIT_ACT) {
    int pid = g.current_player;
    auto legals = g._get_legal_actions_internal(pid);
    if (legals.empty()) {
      // Dead end of the reference (3P: riichi declared, then kita + rinshan draws leave no tenpai-keeping discard;
      // `random.choice([])` raises in RandomAgent).  The rollout retires the game: done + stalled flag.
      g.step_count++;
      g.is_done = true;
      g.stalled = true;
      return true;
    }
    acts[pid] = legals[agent_pick(agent_seed, game_id, sc, pid, (uint32_t)legals.size())];
  } else {
    for (uint8_t pid : g.active_players) {
      auto legals = g._get_legal_actions_internal(pid);
      if (!legals.empty()) acts[pid] = legals[agent_pick(agent_seed, game_id, sc, pid, (uint32_t)legals.size())];
    }
  }
  g.step(acts);
  return true;
}

// ---- GameState::apply_mjai_event (state/event_handler.rs:18-330; sanma: state_3p/event_handler.rs:19-362) ----
// `e` is the parsed event (include/riichienv_b200.h rv_mjai_event: tiles already through mjai_to_tid, "?" -> 0).
inline void apply_mjai_event(GameState& g, const rv_mjai_event& e) {
  const int np = g.np;
  const int actor = e.actor < np ? e.actor : 0;
  auto tid = [](uint8_t t) -> uint8_t { return t < 136 ? t : 0; };
  auto remove_first = [](std::vector<uint8_t>& v, uint8_t t) {
    for (size_t i = 0; i < v.size(); i++)
      if (v[i] == t) {
        v.erase(v.begin() + i);
        return;
      }
  };
  auto claims_after = [&](uint8_t tile, bool ron_only) {          // event_handler.rs:112-138 / 3P 316-356
    for (int i = 0; i < MAXP; i++) {
      g.current_claims[i].clear();
      g.has_claims_entry[i] = false;
    }
    g.active_players.clear();
    std::vector<uint8_t> claim_active;
    for (int i = 0; i < np; i++) {
      if (i == actor) continue;
      auto [legals, missed] = g._get_claim_actions_for_player(i, actor, tile);
      (void)missed;
      if (ron_only) {
        std::vector<Action> r;
        for (auto& a : legals)
          if (a.type == RV_RON) r.push_back(a);
        legals = r;
      }
      if (!legals.empty()) {
        claim_active.push_back((uint8_t)i);
        g.current_claims[i] = legals;
        g.has_claims_entry[i] = true;
      }
    }
    if (!claim_active.empty()) {
      g.phase = RV_WAIT_RESPONSE;
      g.active_players = claim_active;
    } else {
      g.phase = RV_WAIT_ACT;
      g.active_players.clear();
      g.current_player = 0xFF;
    }
    g.needs_tsumo = true;
  };
  switch (e.type) {
    case RV_EV_START_GAME:     // env.rs:56-72 (reset(): logs and counters) + event_handler.rs:21-26
      g.log.clear();
      g.text.clear();
      g.ev_hash = 0xcbf29ce484222325ull;
      g.ev_count = g.ev_words = g.step_count = g.kyoku_count = 0;
      g.stalled = false;
      g.current_player = 0xFF;
      g.active_players.clear();
      break;
    case RV_EV_START_KYOKU: {
      g.honba = e.honba;
      g.riichi_sticks = e.kyotaku;
      g.round_wind = e.bakaze < 4 ? e.bakaze : 0;
      g.oya = e.oya;
      g.kyoku_idx = e.kyoku > 0 ? e.kyoku - 1 : 0;
      g.current_player = 0xFF;
      g.turn_count = 0;
      g.is_done = false;
      g.needs_tsumo = true;
      g.phase = RV_WAIT_ACT;
      g.active_players.clear();
      g.last_discard_pid = g.last_discard_tile = -1;
      for (int i = 0; i < MAXP; i++) {
        g.current_claims[i].clear();
        g.has_claims_entry[i] = false;
      }
      g.pending_kan = false;
      g.is_rinshan_flag = false;
      g.is_first_turn = true;
      g.riichi_pending_acceptance = -1;
      g.drawn_tile = -1;
      g.win_results.clear();
      g.last_error = -1;
      for (int p = 0; p < MAXP; p++) g.riichi_sutehais[p] = g.last_tedashis[p] = -1;
      const int wl = g.sanma ? 108 : 136, left = wl - 13 * np;
      g.wall_tiles.assign(left, 0);
      g.wall_abs.assign(wl, 0xFF);
      for (int i = 0; i < left; i++) g.wall_abs[i] = 0;
      g.dora_indicators = {tid(e.dora_marker)};
      g.rinshan_draw_count = 0;
      g.pending_kan_dora_count = 0;
      g.drawable_count = (uint8_t)(left - 14);
      for (int p = 0; p < MAXP; p++) {
        g.players[p].reset_round();
        g.n_kita[p] = 0;
      }
      for (int p = 0; p < np; p++) {
        g.players[p].score = e.scores[p];
        std::vector<uint8_t> hand;
        for (int k = 0; k < e.tehai_len[p] && k < 14; k++) hand.push_back(tid(e.tehais[p][k]));
        std::sort(hand.begin(), hand.end());
        g.players[p].hand = hand;
      }
      break;
    }
    case RV_EV_TSUMO: {
      const uint8_t tile = tid(e.pai);
      g.current_player = (uint8_t)actor;
      g.drawn_tile = tile;
      g.players[actor].hand.push_back(tile);
      std::sort(g.players[actor].hand.begin(), g.players[actor].hand.end());
      g.players[actor].forbidden_discards.clear();
      if (!g.wall_tiles.empty()) {
        g.wall_tiles.pop_back();
        g.drawable_count = g.drawable_count > 0 ? g.drawable_count - 1 : 0;
      }
      g.phase = RV_WAIT_ACT;
      g.active_players = {(uint8_t)actor};
      g.needs_tsumo = false;
      break;
    }
    case RV_EV_DAHAI:
    case RV_EV_DAHAI_TSUMOGIRI: {
      const uint8_t tile = tid(e.pai);
      g.current_player = (uint8_t)actor;
      remove_first(g.players[actor].hand, tile);
      g.players[actor].discards.push_back(tile);
      g.last_discard_pid = actor;
      g.last_discard_tile = tile;
      g.drawn_tile = -1;
      if (g.players[actor].riichi_stage) {
        g.players[actor].riichi_declared = true;
        g.players[actor].riichi_stage = false;
      }
      claims_after(tile, false);
      break;
    }
    case RV_EV_PON:
    case RV_EV_CHI:
    case RV_EV_DAIMINKAN: {
      const uint8_t tile = tid(e.pai);
      g.current_player = (uint8_t)actor;
      const int want = e.type == RV_EV_DAIMINKAN ? 3 : 2;
      const int nc = e.n_consumed < want ? e.n_consumed : want;
      Meld m;
      m.meld_type = e.type == RV_EV_PON ? Pon : e.type == RV_EV_CHI ? Chi : Daiminkan;
      m.tiles.push_back(tile);
      for (int k = 0; k < nc; k++) {
        m.tiles.push_back(tid(e.consumed[k]));
        remove_first(g.players[actor].hand, tid(e.consumed[k]));
      }
      m.opened = true;
      m.from_who = -1;
      m.called_tile = tile;
      g.players[actor].melds.push_back(m);
      g.phase = RV_WAIT_ACT;
      g.active_players = {(uint8_t)actor};
      if (e.type == RV_EV_DAIMINKAN) {
        g.needs_tsumo = true;
        break;
      }
      g.drawn_tile = -1;
      g.needs_tsumo = false;
      if (e.type == RV_EV_CHI && g.sanma) break;      // the 3P handler leaves forbidden_discards alone on a chi
      g.players[actor].forbidden_discards.clear();
      if (g.rb(RV_RULE_KUIKAE_FORBIDDEN)) {
        g.players[actor].forbidden_discards.push_back(tile);
        if (e.type == RV_EV_CHI && nc == 2) {
          const int t34 = tile / 4;
          int c0 = tid(e.consumed[0]) / 4, c1 = tid(e.consumed[1]) / 4;
          if (c0 > c1) std::swap(c0, c1);
          if (c0 == t34 + 1 && c1 == t34 + 2) {
            if (t34 % 9 <= 5) g.players[actor].forbidden_discards.push_back((uint8_t)((t34 + 3) * 4));
          } else if (t34 >= 2 && c1 == t34 - 1 && c0 == t34 - 2 && t34 % 9 >= 3) {
            g.players[actor].forbidden_discards.push_back((uint8_t)((t34 - 3) * 4));
          }
        }
      }
      break;
    }
    case RV_EV_ANKAN: {
      Meld m;
      m.meld_type = Ankan;
      for (int k = 0; k < e.n_consumed && k < 4; k++) {
        m.tiles.push_back(tid(e.consumed[k]));
        remove_first(g.players[actor].hand, tid(e.consumed[k]));
      }
      m.opened = false;
      m.from_who = -1;
      m.called_tile = -1;
      g.players[actor].melds.push_back(m);
      g.current_player = (uint8_t)actor;
      g.phase = RV_WAIT_ACT;
      g.active_players = {(uint8_t)actor};
      g.needs_tsumo = true;
      break;
    }
    case RV_EV_KAKAN: {
      const uint8_t tile = tid(e.pai);
      remove_first(g.players[actor].hand, tile);
      for (auto& m : g.players[actor].melds)
        if (m.meld_type == Pon && m.tiles[0] / 4 == tile / 4) {
          m.meld_type = Kakan;
          m.tiles.push_back(tile);
          break;
        }
      g.current_player = (uint8_t)actor;
      g.phase = RV_WAIT_ACT;
      g.active_players = {(uint8_t)actor};
      g.needs_tsumo = true;
      break;
    }
    case RV_EV_REACH:
      g.players[actor].riichi_stage = true;
      break;
    case RV_EV_REACH_ACCEPTED:
      g.players[actor].riichi_declared = true;
      g.riichi_sticks += 1;
      g.players[actor].score -= 1000;
      break;
    case RV_EV_DORA:
      g.dora_indicators.push_back(tid(e.pai));
      break;
    case RV_EV_KITA:
      if (g.sanma) {
        int kita = -1;
        for (uint8_t t : g.players[actor].hand)
          if (t / 4 == 30) {
            kita = t;
            break;
          }
        g.current_player = (uint8_t)actor;
        if (kita >= 0) {
          remove_first(g.players[actor].hand, (uint8_t)kita);
          g.n_kita[actor]++;
          claims_after((uint8_t)kita, true);
        } else {
          for (int i = 0; i < MAXP; i++) {
            g.current_claims[i].clear();
            g.has_claims_entry[i] = false;
          }
          g.phase = RV_WAIT_ACT;
          g.active_players.clear();
          g.current_player = 0xFF;
          g.needs_tsumo = true;
        }
      }
      break;
    case RV_EV_HORA:
    case RV_EV_RYUKYOKU:
    case RV_EV_END_KYOKU:
      g.is_done = true;
      break;
    default:
      break;
  }
}

// ---- replay ingestion: GameState::apply_log_action (state/event_handler.rs:332-894; sanma state_3p/event_handler.rs:365-808) ----
// `a` is one replay `Action` (replay/mod.rs:35-80) in the fixed layout of include/riichienv_b200.h (rv_log_action).
inline void apply_log_action(GameState& g, const rv_log_action& a) {
  const int np = g.np;
  const bool sanma = g.sanma;
  const int seat = a.seat < np ? a.seat : 0;
  auto tid = [](uint8_t t) -> uint8_t { return t < 136 ? t : 0; };
  auto remove_first = [](std::vector<uint8_t>& v, uint8_t t) {
    for (size_t i = 0; i < v.size(); i++)
      if (v[i] == t) {
        v.erase(v.begin() + i);
        return;
      }
  };
  auto accept_pending = [&]() {                               // `if let Some(rp) = self.riichi_pending_acceptance.take()`
    if (g.riichi_pending_acceptance >= 0) {
      g.players[g.riichi_pending_acceptance].score -= 1000;
      g.riichi_sticks += 1;
      g.riichi_pending_acceptance = -1;
    }
  };
  PlayerState& P = g.players[seat];
  switch (a.type) {
    case RV_LA_DISCARD: {                                     // event_handler.rs:334-416 / 3P 368-414
      const uint8_t t = tid(a.tile);
      const bool is_liqi = a.flags & 1, is_wliqi = a.flags & 2;
      const bool tsumogiri = g.drawn_tile >= 0 && g.drawn_tile == t;
      remove_first(P.hand, t);
      std::sort(P.hand.begin(), P.hand.end());
      P.discards.push_back(t);
      P.discard_from_hand.push_back(!tsumogiri);
      P.discard_is_riichi.push_back(is_liqi || is_wliqi);
      g.last_discard_pid = seat;
      g.last_discard_tile = t;
      g.drawn_tile = -1;
      P.missed_agari_doujun = false;
      if (!sanma) {
        P.nagashi_eligible = P.nagashi_eligible && is_terminal_tile(t);
        if (is_liqi || is_wliqi) {
          if (!P.riichi_declared) {
            P.riichi_declared = true;
            if (is_wliqi) P.double_riichi_declared = true;
            g.riichi_pending_acceptance = seat;
          }
          P.riichi_declaration_index = (int)P.discards.size() - 1;
        }
      } else {
        P.riichi_declared = P.riichi_declared || is_liqi || is_wliqi;
        if (is_wliqi) P.double_riichi_declared = true;
        if (is_liqi || is_wliqi) {
          P.riichi_declaration_index = (int)P.discards.size() - 1;
          g.riichi_pending_acceptance = seat;
        }
        P.nagashi_eligible = P.nagashi_eligible && is_terminal_tile(t);
      }
      g.current_player = (uint8_t)((seat + 1) % np);
      g.phase = RV_WAIT_ACT;
      g.active_players = {g.current_player};
      g.needs_tsumo = true;
      g.is_first_turn = false;
      g.is_after_kan = false;
      break;
    }
    case RV_LA_DEAL: {                                        // 417-436
      const uint8_t t = tid(a.tile);
      accept_pending();
      P.hand.push_back(t);
      g.drawn_tile = t;
      g.current_player = (uint8_t)seat;
      g.phase = RV_WAIT_ACT;
      g.active_players = {g.current_player};
      g.is_rinshan_flag = g.is_after_kan && seat == g.current_player;
      g.needs_tsumo = false;
      g.is_after_kan = false;
      std::sort(P.hand.begin(), P.hand.end());
      if (!g.wall_tiles.empty()) {
        g.wall_tiles.pop_back();
        g.drawable_count = g.drawable_count > 0 ? g.drawable_count - 1 : 0;
      }
      break;
    }
    case RV_LA_CHI_PENG_GANG: {                               // 437-567
      accept_pending();
      if (g.last_discard_pid >= 0) g.players[g.last_discard_pid].nagashi_eligible = false;
      const int n = a.n_tiles < 4 ? a.n_tiles : 4;
      for (int i = 0; i < n; i++)
        if (a.froms[i] == seat) remove_first(P.hand, tid(a.tiles[i]));
      std::sort(P.hand.begin(), P.hand.end());
      int from_who = -1, ct = -1;
      for (int i = 0; i < n; i++)
        if (a.froms[i] != seat) {
          from_who = a.froms[i];
          ct = tid(a.tiles[i]);
          break;
        }
      const uint8_t discarder = (uint8_t)(from_who < 0 ? 0 : from_who);
      Meld m;
      m.meld_type = (MeldType)a.meld_type;
      for (int i = 0; i < n; i++) m.tiles.push_back(tid(a.tiles[i]));
      m.opened = true;
      m.from_who = (int8_t)from_who;
      m.called_tile = (int16_t)ct;
      P.melds.push_back(m);
      if ((m.meld_type == Pon || m.meld_type == Daiminkan) && ct >= 0) {
        const int tv = ct / 4;
        auto count = [&](int lo, int hi) {
          int c = 0;
          for (auto& x : P.melds) {
            const int t = x.tiles[0] / 4;
            if (t >= lo && t <= hi && x.meld_type != Chi) c++;
          }
          return c;
        };
        if (tv >= 31 && tv <= 33) {
          if (count(31, 33) == 3) P.pao37 = discarder;
        } else if (tv >= 27 && tv <= 30) {
          if (count(27, 30) == 4) P.pao50 = discarder;
        }
      }
      g.current_player = (uint8_t)seat;
      g.phase = RV_WAIT_ACT;
      g.active_players = {g.current_player};
      const bool gang = m.meld_type == Daiminkan;
      g.needs_tsumo = gang;
      g.is_first_turn = false;
      g.is_after_kan = gang;
      break;
    }
    case RV_LA_ANGANG_ADDGANG: {                              // 568-663
      const uint8_t t0 = tid(a.n_tiles ? a.tiles[0] : 0);
      if (a.meld_type == RV_MELD_ANKAN) {
        const uint8_t tv = t0 / 4;
        for (int r = 0; r < 4; r++)
          for (size_t i = 0; i < P.hand.size(); i++)
            if (P.hand[i] / 4 == tv) {
              P.hand.erase(P.hand.begin() + i);
              break;
            }
        Meld m;
        m.meld_type = Ankan;
        m.tiles = {(uint8_t)(tv * 4), (uint8_t)(tv * 4 + 1), (uint8_t)(tv * 4 + 2), (uint8_t)(tv * 4 + 3)};
        m.opened = false;
        m.from_who = -1;
        m.called_tile = -1;
        P.melds.push_back(m);
      } else {
        remove_first(P.hand, t0);
        for (auto& m : P.melds)
          if (m.meld_type == Pon && m.tiles[0] / 4 == t0 / 4) {
            m.meld_type = Kakan;
            m.tiles.push_back(t0);
            std::sort(m.tiles.begin(), m.tiles.end());
            break;
          }
        g.last_discard_pid = seat;                            // chankan target
        g.last_discard_tile = t0;
      }
      std::sort(P.hand.begin(), P.hand.end());
      g.current_player = (uint8_t)seat;
      g.phase = RV_WAIT_ACT;
      g.active_players = {g.current_player};
      g.needs_tsumo = true;
      g.is_first_turn = false;
      g.is_after_kan = true;
      if (sanma || a.meld_type == RV_MELD_ANKAN) {            // kokushi can ron a closed kan (3P: set for both kinds)
        g.last_discard_pid = seat;
        g.last_discard_tile = t0;
      }
      break;
    }
    case RV_LA_DORA:                                          // 664-666
      g.dora_indicators.push_back(tid(a.tile));
      break;
    case RV_LA_BABEI:                                         // 3P 573-595; the 4P match has no arm for it
      if (sanma) {
        accept_pending();
        for (size_t i = 0; i < P.hand.size(); i++)
          if (P.hand[i] / 4 == 30) {
            const uint8_t t = P.hand[i];
            P.hand.erase(P.hand.begin() + i);
            g.n_kita[seat]++;
            g.last_discard_pid = seat;
            g.last_discard_tile = t;
            break;
          }
        std::sort(P.hand.begin(), P.hand.end());
        g.current_player = (uint8_t)seat;
        g.phase = RV_WAIT_ACT;
        g.active_players = {g.current_player};
        g.needs_tsumo = true;
        g.is_first_turn = false;
        g.is_after_kan = true;
      }
      break;
    case RV_LA_HULE: {                                        // 667-818 / 3P 596-735
      const int nh = a.n_hule < 3 ? a.n_hule : 3;
      auto real_tsumo = [&](const rv_hule& h) { return sanma ? (h.zimo && h.seat == g.current_player) : (h.zimo != 0); };
      if (nh > 0 && !real_tsumo(a.hules[0])) g.riichi_pending_acceptance = -1;
      const int32_t honba = g.honba;
      const uint32_t sticks = g.riichi_sticks;
      bool honba_taken = false;
      for (int k = 0; k < nh; k++) {
        const rv_hule& h = a.hules[k];
        const int w = h.seat < np ? h.seat : 0;
        const bool is_oya = w == g.oya;
        const bool tsumo = real_tsumo(h);
        int pao_payer = -1, pao_val = 0, total_val = 0;
        if (h.yiman)
          for (int y = 0; y < 64; y++) {
            if (!((h.fans >> y) & 1)) continue;
            const int val = (y >= 47 && y <= 50) ? 2 : 1;
            total_val += val;
            const int liable = y == 37 ? g.players[w].pao37 : y == 50 ? g.players[w].pao50 : -1;
            if (liable >= 0) {
              pao_val += val;
              pao_payer = liable;
              if (!sanma && !tsumo) break;
            }
          }
        if (tsumo) {
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
              g.players[pao_payer].score -= pao_amt;
              g.players[w].score += pao_amt;
            }
            if (non_pao > 0)
              for (int i = 0; i < np; i++) {
                if (i == w) continue;
                int32_t share;
                if (sanma)
                  share = is_oya ? non_pao / (np - 1)
                                 : (i == g.oya ? (int32_t)h.point_zimo_qin : (int32_t)h.point_zimo_xian) * non_pao / (tsumo_total ? tsumo_total : 1);
                else
                  share = is_oya ? non_pao / 3 : (i == g.oya ? non_pao / 2 : non_pao / 4);
                g.players[i].score -= share;
                g.players[w].score += share;
              }
            if (pao_payer >= 0) {
              const int32_t hb = sanma ? honba * (np - 1) * 100 : honba * 300;
              g.players[pao_payer].score -= hb;
              g.players[w].score += hb;
            }
          } else {
            for (int i = 0; i < np; i++) {
              if (i == w) continue;
              const int32_t base = is_oya ? h.point_zimo_xian : (i == g.oya ? h.point_zimo_qin : h.point_zimo_xian);
              const int32_t pay = base + honba * 100;
              g.players[i].score -= pay;
              g.players[w].score += pay;
            }
          }
        } else if (g.last_discard_pid >= 0) {
          const int d = g.last_discard_pid;
          const int32_t ron_honba = honba_taken ? 0 : honba;
          honba_taken = true;
          const int32_t hb = sanma ? ron_honba * (np - 1) * 100 : ron_honba * 300;
          if (sanma ? pao_val > 0 : pao_payer >= 0) {
            if (sanma) {
              const int pp = pao_payer >= 0 ? pao_payer : d;
              const int32_t ron_total = (int32_t)h.point_rong;
              const int32_t pao_amt = ron_total * pao_val / (total_val ? total_val : 1);
              const int32_t pao_share = pao_amt / 2 + hb, disc_share = ron_total - pao_amt / 2;
              g.players[pp].score -= pao_share;
              g.players[d].score -= disc_share;
              g.players[w].score += pao_share + disc_share;
            } else {
              const int32_t half = (int32_t)h.point_rong / 2;
              g.players[pao_payer].score -= half + hb;
              g.players[d].score -= half;
              g.players[w].score += (int32_t)h.point_rong + hb;
            }
          } else {
            const int32_t pay = (int32_t)h.point_rong + hb;
            g.players[d].score -= pay;
            g.players[w].score += pay;
          }
        }
      }
      if (nh > 0) {
        g.players[a.hules[0].seat < np ? a.hules[0].seat : 0].score += (int32_t)sticks * 1000;
        g.riichi_sticks = 0;
      }
      g.is_done = true;
      break;
    }
    case RV_LA_NOTILE: {                                      // 819-882
      accept_pending();
      std::vector<int> nagashi;
      for (int i = 0; i < np; i++)
        if (g.players[i].nagashi_eligible) nagashi.push_back(i);
      if (!nagashi.empty()) {
        for (int w : nagashi) {
          const bool is_oya = w == g.oya;
          Score sc = calculate_score(5, 30, is_oya, true, 0, (uint8_t)np);
          for (int i = 0; i < np; i++) {
            if (i == w) continue;
            const int32_t pay = is_oya ? (int32_t)sc.pay_tsumo_ko : (i == g.oya ? (int32_t)sc.pay_tsumo_oya : (int32_t)sc.pay_tsumo_ko);
            g.players[i].score -= pay;
            g.players[w].score += pay;
          }
        }
      } else {
        bool tenpai[MAXP] = {false, false, false, false};
        int ntp = 0;
        for (int i = 0; i < np; i++) {
          HandEvaluator calc(g.players[i].hand, g.players[i].melds, sanma);
          tenpai[i] = calc.is_tenpai();
          ntp += tenpai[i];
        }
        if (ntp > 0 && ntp < np) {
          const int32_t pool = sanma ? 2000 : 3000;
          const int32_t pk = pool / ntp, pn = pool / (np - ntp);
          for (int i = 0; i < np; i++) g.players[i].score += tenpai[i] ? pk : -pn;
        }
      }
      g.is_done = true;
      break;
    }
    case RV_LA_LIUJU:                                         // 883-890 (the 3P arm does not settle the deposit)
      if (!sanma) accept_pending();
      g.is_done = true;
      break;
    default:
      break;
  }
}
// KyokuStepIterator::_collect_pass_observations (replay/mod.rs:130-181) asks `_get_claim_actions_for_player(i, discarder, tile)`
// of every other seat after a logged discard: here for all seats at once, left in current_claims / active_players / phase
// (the caller snapshots the record, reads the lists and puts the record back).
inline void claims_for_last_discard(GameState& g) {
  if (g.last_discard_pid < 0) return;
  const int actor = g.last_discard_pid;
  const uint8_t tile = (uint8_t)g.last_discard_tile;
  for (int i = 0; i < MAXP; i++) {
    g.current_claims[i].clear();
    g.has_claims_entry[i] = false;
  }
  g.active_players.clear();
  for (int i = 0; i < g.np; i++) {
    if (i == actor) continue;
    auto [legals, missed] = g._get_claim_actions_for_player(i, actor, tile);
    (void)missed;
    if (!legals.empty()) {
      g.active_players.push_back((uint8_t)i);
      g.current_claims[i] = legals;
      g.has_claims_entry[i] = true;
    }
  }
  if (!g.active_players.empty()) {
    g.phase = RV_WAIT_RESPONSE;
  } else {
    g.phase = RV_WAIT_ACT;
    g.current_player = 0xFF;
  }
  g.needs_tsumo = true;
}
// LogKyoku::steps (replay/mod.rs:1094-1292), the part after `_initialize_round(oya, bakaze, ben, liqibang, None, scores)`
inline void replay_begin_patch(GameState& g, const rv_log_kyoku& k) {
  const int np = g.np;
  const int oya = k.oya < np ? k.oya : 0;
  for (int i = 0; i < np; i++) {
    std::vector<uint8_t> h;
    for (int j = 0; j < k.hand_len[i] && j < 14; j++) h.push_back(k.hands[i][j] < 136 ? k.hands[i][j] : 0);
    g.players[i].hand = h;
  }
  if (g.players[oya].hand.size() == 14) {
    g.drawn_tile = k.oya_drawn_tile == 0xFF ? -1 : k.oya_drawn_tile;
    g.needs_tsumo = false;
  } else {
    if (g.drawn_tile >= 0) {
      const size_t top = g.wall_tiles.size() + g.rinshan_draw_count;
      g.wall_tiles.push_back((uint8_t)g.drawn_tile);
      if (top < g.wall_abs.size()) g.wall_abs[top] = (uint8_t)g.drawn_tile;
      g.drawable_count += 1;
    }
    g.drawn_tile = -1;
    g.needs_tsumo = true;
  }
  for (int i = 0; i < np; i++) std::sort(g.players[i].hand.begin(), g.players[i].hand.end());
  g.dora_indicators.clear();
  for (int i = 0; i < k.n_doras && i < RV_LOG_MAX_DORAS; i++) g.dora_indicators.push_back(k.doras[i] < 136 ? k.doras[i] : 0);
  g.is_after_kan = false;
}

// ---- the keyed "greedy-win" agent (test agent #1; shared definition with the kernel, csrc/game.cuh greedy_pick) ----
// Uniform random play almost never completes a hand (~0.3 % of rounds end in a win), which leaves the settlement code
// (state/mod.rs:685-893, 919-1142) statistically untested.  This agent plays towards wins:
//   r = mix64(agent_seed ^ game_id * 0x9E3779B97F4A7C15 ^ step_count << 8 ^ seat)
//   1. the first Tsumo / Ron of the legal list, if any;  2. else the first Riichi;
//   3. else, u = (r >> 40) & 0xFF: with u < 64 a uniformly keyed ((r >> 8) mod count) action among Pon / Daiminkan / Ankan /
//      Kakan / Kita if there is one; with 64 <= u < 96 likewise among the Chi actions;
//   4. else, in a claim window: Pass;  5. else, on the own turn: among the Discard actions the ones that leave the lowest
//      shanten (calculate_shanten / calculate_shanten_3p of the remaining tiles), one of them keyed by (r >> 8) mod count.
inline int greedy_pick(const GameState& g, int pid, const std::vector<Action>& L, uint64_t agent_seed, uint64_t game_id) {
  const uint64_t r = mix64(agent_seed ^ (game_id * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)g.step_count << 8) ^ (uint64_t)pid);
  const int n = (int)L.size();
  for (int i = 0; i < n; i++)
    if (L[i].type == RV_TSUMO || L[i].type == RV_RON) return i;
  for (int i = 0; i < n; i++)
    if (L[i].type == RV_RIICHI) return i;
  const uint32_t u = (uint32_t)(r >> 40) & 0xFF, v = (uint32_t)(r >> 8);
  auto nth_of = [&](auto pred) -> int {
    int cnt = 0;
    for (int i = 0; i < n; i++) cnt += pred(L[i]) ? 1 : 0;
    if (cnt == 0) return -1;
    int want = (int)(v % (uint32_t)cnt);
    for (int i = 0; i < n; i++)
      if (pred(L[i]) && want-- == 0) return i;
    return -1;
  };
  auto is_k = [](const Action& a) { return a.type == RV_PON || a.type == RV_DAIMINKAN || a.type == RV_ANKAN || a.type == RV_KAKAN || a.type == RV_KITA; };
  auto is_c = [](const Action& a) { return a.type == RV_CHI; };
  if (u < 64) {
    int i = nth_of(is_k);
    if (i >= 0) return i;
  } else if (u < 96) {
    int i = nth_of(is_c);
    if (i >= 0) return i;
  }
  if (g.phase == RV_WAIT_RESPONSE) {
    for (int i = 0; i < n; i++)
      if (L[i].type == RV_PASS) return i;
    return n - 1;
  }
  // discards by shanten of what remains
  uint8_t cnt[34] = {0};
  const auto& hand = g.players[pid].hand;
  for (uint8_t t : hand)
    if (t / 4 < 34) cnt[t / 4]++;
  const int len_div3 = ((int)hand.size() - 1) / 3;
  int best = 99, nbest = 0;
  std::vector<int> sh(n, 99);
  for (int i = 0; i < n; i++) {
    if (L[i].type != RV_DISCARD || L[i].tile < 0) continue;
    const int k = L[i].tile / 4;
    cnt[k]--;
    sh[i] = g.sanma ? shanten_from_counts_3p(cnt, len_div3) : shanten_from_counts(cnt, len_div3);
    cnt[k]++;
    if (sh[i] < best) best = sh[i], nbest = 0;
    if (sh[i] == best) nbest++;
  }
  if (nbest == 0) return (int)(v % (uint32_t)n);   // no discard in the list (cannot happen on a live turn)
  int want = (int)(v % (uint32_t)nbest);
  for (int i = 0; i < n; i++)
    if (sh[i] == best && want-- == 0) return i;
  return 0;
}

// One env step with agent `policy` (0 = uniform random, 1 = greedy-win).  Returns false if the game is done.
inline bool agent_step(GameState& g, int policy, uint64_t agent_seed, uint64_t game_id) {
  if (policy == 0) return random_step(g, agent_seed, game_id);
  if (g.is_done) return false;
  std::optional<Action> acts[MAXP];
  if (g.phase == RV_WAIT_ACT) {
    int pid = g.current_player;
    auto legals = g._get_legal_actions_internal(pid);
    if (legals.empty()) {   // the reference's 3P dead end, as in random_step
      g.step_count++;
      g.is_done = true;
      g.stalled = true;
      return true;
    }
    acts[pid] = legals[greedy_pick(g, pid, legals, agent_seed, game_id)];
  } else {
    for (uint8_t pid : g.active_players) {
      auto legals = g._get_legal_actions_internal(pid);
      if (!legals.empty()) acts[pid] = legals[greedy_pick(g, pid, legals, agent_seed, game_id)];
    }
  }
  g.step(acts);
  return true;
}

}  // namespace orc
