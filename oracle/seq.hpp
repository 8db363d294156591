// ORACLE — TEST INFRASTRUCTURE ONLY (see hand.hpp header).
//
// Sequence features restated from riichienv-core/src/observation/sequence_features.rs, function by function, over
// the seat's event delta (state/mod.rs:211-218: the events an Observation carries are those pushed since the seat's
// previous observation; the progression cache is off in the live env, state/mod.rs:146).  The reference parses MJAI
// JSON strings; the oracle keeps the same events as binary records (game.hpp, same words the JSON renderer reads), so
// "parse the event" here means decoding one record.  Known answers pinned in tests/test_oracle_golden.py from the
// reference's own unit tests (sequence_features.rs:845-930).
#pragma once
#include "game.hpp"

namespace orc {

struct SeqEvent {      // one decoded MJAI event of the delta
  int type, actor, tile;            // rv_event_type, "actor", "pai" (tid)
  bool tsumogiri = false;
  int target = -1;
  std::vector<uint8_t> consumed;
  uint32_t honba = 0, kyotaku = 0;  // start_kyoku
  int32_t scores[4] = {0, 0, 0, 0};
};

inline std::vector<SeqEvent> seq_decode(const std::vector<uint32_t>& log, uint32_t w0, uint32_t w1, int np) {
  std::vector<SeqEvent> evs;
  uint32_t w = w0;
  while (w < w1 && w < log.size()) {
    uint32_t h = log[w];
    int nw = (h >> 8) & 0xFF;
    SeqEvent e;
    e.type = h & 0xFF;
    e.actor = (h >> 16) & 0xFF;
    e.tile = (h >> 24) & 0xFF;
    switch (e.type) {
      case RV_EV_START_KYOKU:
        e.honba = log[w + 1] & 0xFF;
        e.kyotaku = log[w + 1] >> 16;
        for (int s = 0; s < np; s++) e.scores[s] = (int32_t)log[w + 2 + s];
        break;
      case RV_EV_DAHAI_TSUMOGIRI:
        e.tsumogiri = true;
        e.type = RV_EV_DAHAI;
        break;
      case RV_EV_PON:
      case RV_EV_CHI:
        e.target = log[w + 1] & 0xFF;
        e.consumed = {(uint8_t)(log[w + 1] >> 8), (uint8_t)(log[w + 1] >> 16)};
        break;
      case RV_EV_DAIMINKAN:
        e.target = log[w + 1] & 0xFF;
        e.consumed = {(uint8_t)(log[w + 1] >> 8), (uint8_t)(log[w + 1] >> 16), (uint8_t)(log[w + 1] >> 24)};
        break;
      case RV_EV_ANKAN:
        e.consumed = {(uint8_t)log[w + 1], (uint8_t)(log[w + 1] >> 8), (uint8_t)(log[w + 1] >> 16), (uint8_t)(log[w + 1] >> 24)};
        break;
      default:
        break;
    }
    evs.push_back(e);
    w += nw ? nw : 1;
  }
  return evs;
}

// sequence_features.rs:44-70
inline int seq_tile_type_to_kan37(int tile_type) {
  if (tile_type <= 8) return tile_type + 1;
  if (tile_type <= 17) return tile_type + 2;
  if (tile_type <= 26) return tile_type + 3;
  if (tile_type <= 33) return tile_type + 3;
  return 0;
}
inline int seq_tile_id_to_kan37(int tile_id) {
  if (tile_id == 16) return 0;
  if (tile_id == 52) return 10;
  if (tile_id == 88) return 20;
  return seq_tile_type_to_kan37(tile_id / 4);
}
inline bool seq_red(int t) { return t == 16 || t == 52 || t == 88; }
// sequence_features.rs:92-133
inline int seq_encode_chi(const std::vector<uint8_t>& consumed, int called_tile) {
  std::vector<int> all_tiles = {called_tile};
  for (uint8_t c : consumed) all_tiles.push_back(c);
  std::sort(all_tiles.begin(), all_tiles.end());
  int first_type = all_tiles[0] / 4, suit = first_type / 9, suit_base = suit * 9, seq_start = first_type - suit_base;
  int call_pos = called_tile / 4 - suit_base - seq_start;
  bool has_red = false;
  for (int t : all_tiles) has_red |= seq_red(t);
  bool five_in_seq = (suit_base + 4) >= (suit_base + seq_start) && (suit_base + 4) <= (suit_base + seq_start + 2);
  bool involves_five = five_in_seq && (seq_start <= 4 && 4 <= seq_start + 2);
  int offset = 0;
  for (int s = 0; s < seq_start; s++) offset += (s <= 4 && 4 <= s + 2) ? 6 : 3;
  int sub_idx = (involves_five && has_red) ? 3 + call_pos : call_pos;
  return suit * 30 + offset + sub_idx;
}
// sequence_features.rs:144-182
inline int seq_encode_pon(const std::vector<uint8_t>& consumed, int called_tile) {
  int called_type = called_tile / 4, suit = called_type / 9;
  if (suit == 3) return 33 + (called_type - 27);
  int rank = called_type - suit * 9, suit_offset = suit * 11;
  if (rank == 4) {
    bool consumed_has_red = false;
    for (uint8_t t : consumed) consumed_has_red |= seq_red(t);
    int sub_idx = seq_red(called_tile) ? 2 : consumed_has_red ? 1 : 0;
    return suit_offset + 4 + sub_idx;
  }
  return suit_offset + (rank < 4 ? rank : rank + 2);
}
inline int seq_relative_from(int actor, int target) { return (target - actor + 3) % 4; }   // 186-188

struct SeqFeatures {
  std::vector<uint16_t> sparse;
  float numeric[12];
  std::vector<std::array<uint16_t, 5>> prog;
  std::vector<std::array<uint16_t, 4>> cand;
};

// get_drawn_tile (439-465): walk the delta backwards
inline int seq_get_drawn_tile(const std::vector<SeqEvent>& evs, int pid) {
  for (int i = (int)evs.size() - 1; i >= 0; i--) {
    const SeqEvent& e = evs[i];
    if (e.type == RV_EV_TSUMO && e.actor == pid) return e.tile;      // own tsumo is never masked
    if (e.type == RV_EV_DAHAI || e.type == RV_EV_CHI || e.type == RV_EV_PON || e.type == RV_EV_DAIMINKAN) break;
  }
  return -1;
}
// find_last_discard_actor (842-852)
inline int seq_find_last_discard_actor(const std::vector<SeqEvent>& evs) {
  for (int i = (int)evs.size() - 1; i >= 0; i--)
    if (evs[i].type == RV_EV_DAHAI || evs[i].type == RV_EV_KAKAN) return evs[i].actor;
  return -1;
}

inline SeqFeatures encode_seq(const GameState& g, int pid, uint32_t w0, uint32_t w1, int game_style) {
  SeqFeatures f;
  std::vector<SeqEvent> evs = seq_decode(g.log, w0, w1, g.np);
  // ---- encode_seq_sparse (359-406)
  f.sparse.push_back((uint16_t)std::min(game_style, 1));
  f.sparse.push_back((uint16_t)(2 + std::min(pid, 3)));
  f.sparse.push_back((uint16_t)(6 + std::min<int>(g.round_wind, 2)));
  f.sparse.push_back((uint16_t)(9 + std::min<int>(g.oya, 3)));
  {
    // count_tiles_remaining (409-436); the other seats' hands are masked to empty in the Observation
    uint32_t used = (uint32_t)g.players[pid].hand.size();
    for (int i = 0; i < 4; i++) used += (uint32_t)g.players[i].discards.size();
    for (int i = 0; i < 4; i++)
      for (auto& m : g.players[i].melds) used += (uint32_t)m.tiles.size();
    used += (uint32_t)g.dora_indicators.size();
    uint32_t wall_size = 136u > 14u + used ? 136u - (14u + used) : 0u;
    f.sparse.push_back((uint16_t)(13 + std::min<uint32_t>(wall_size, 69)));
  }
  for (size_t i = 0; i < g.dora_indicators.size() && i < 5; i++)
    f.sparse.push_back((uint16_t)(83 + 37 * i + seq_tile_id_to_kan37(g.dora_indicators[i])));
  for (uint8_t tid : g.players[pid].hand)
    if (tid < 136) f.sparse.push_back((uint16_t)(268 + tid));
  int drawn = seq_get_drawn_tile(evs, pid);
  if (drawn >= 0) f.sparse.push_back((uint16_t)(404 + seq_tile_id_to_kan37(drawn)));
  // ---- encode_seq_numeric (478-503) + parse_start_kyoku_info (506-523)
  f.numeric[0] = (float)g.honba;
  f.numeric[1] = (float)g.riichi_sticks;
  for (int i = 0; i < 4; i++) f.numeric[2 + i] = (float)g.players[(pid + i) % 4].score;
  {
    uint32_t sh = g.honba, sr = g.riichi_sticks;
    int32_t ss[4];
    for (int i = 0; i < 4; i++) ss[i] = g.players[i].score;
    for (auto& e : evs)
      if (e.type == RV_EV_START_KYOKU) {
        sh = e.honba;
        sr = e.kyotaku;
        for (int i = 0; i < 4; i++) ss[i] = e.scores[i];
        break;
      }
    f.numeric[6] = (float)sh;
    f.numeric[7] = (float)sr;
    for (int i = 0; i < 4; i++) f.numeric[8 + i] = (float)ss[(pid + i) % 4];
  }
  // ---- encode_seq_progression (535-700), JSON fallback path
  {
    int pending_reach_actor = -1;
    for (auto& e : evs) {
      switch (e.type) {
        case RV_EV_START_KYOKU: f.prog.push_back({4, 0, 2, 2, 4}); break;
        case RV_EV_REACH: pending_reach_actor = e.actor; break;
        case RV_EV_DAHAI: {
          uint16_t liqi = 0;
          if (pending_reach_actor == e.actor) {
            pending_reach_actor = -1;
            liqi = 1;
          }
          f.prog.push_back({(uint16_t)e.actor, (uint16_t)(1 + seq_tile_id_to_kan37(e.tile)), (uint16_t)(e.tsumogiri ? 1 : 0), liqi, 4});
          break;
        }
        case RV_EV_CHI:
          f.prog.push_back({(uint16_t)e.actor, (uint16_t)(38 + seq_encode_chi(e.consumed, e.tile)), 2, 2,
                            (uint16_t)seq_relative_from(e.actor, e.target)});
          break;
        case RV_EV_PON:
          f.prog.push_back({(uint16_t)e.actor, (uint16_t)(128 + seq_encode_pon(e.consumed, e.tile)), 2, 2,
                            (uint16_t)seq_relative_from(e.actor, e.target)});
          break;
        case RV_EV_DAIMINKAN:
          f.prog.push_back({(uint16_t)e.actor, (uint16_t)(168 + seq_tile_id_to_kan37(e.tile)), 2, 2,
                            (uint16_t)seq_relative_from(e.actor, e.target)});
          break;
        case RV_EV_ANKAN:
          f.prog.push_back({(uint16_t)e.actor, (uint16_t)(205 + e.consumed[0] / 4), 2, 2, 4});
          break;
        case RV_EV_KAKAN:
          f.prog.push_back({(uint16_t)e.actor, (uint16_t)(239 + seq_tile_id_to_kan37(e.tile)), 2, 2, 4});
          break;
        default: break;
      }
      if (f.prog.size() >= 512) break;
    }
  }
  // ---- encode_seq_candidates (725-830)
  {
    bool owes = !g.is_done && ((g.phase == RV_WAIT_ACT && g.current_player == pid) ||
                               (g.phase == RV_WAIT_RESPONSE &&
                                std::find(g.active_players.begin(), g.active_players.end(), (uint8_t)pid) != g.active_players.end()));
    std::vector<Action> legal;
    if (owes) legal = g._get_legal_actions_internal(pid);
    int target = seq_find_last_discard_actor(evs);
    for (auto& a : legal) {
      switch (a.type) {
        case RV_DISCARD:
          if (a.tile >= 0) f.cand.push_back({(uint16_t)seq_tile_id_to_kan37(a.tile), (uint16_t)((drawn >= 0 && drawn == a.tile) ? 1 : 0), 2, 3});
          break;
        case RV_ANKAN:
          if (!a.consume.empty()) f.cand.push_back({(uint16_t)(37 + a.consume[0] / 4), 2, 2, 3});
          break;
        case RV_KAKAN: {
          int t = a.tile >= 0 ? a.tile : (a.consume.empty() ? -1 : a.consume[0]);
          if (t >= 0) f.cand.push_back({(uint16_t)(71 + seq_tile_id_to_kan37(t)), 2, 2, 3});
          break;
        }
        case RV_TSUMO: f.cand.push_back({108, 2, 2, 3}); break;
        case RV_KYUSHU_KYUHAI: f.cand.push_back({109, 2, 2, 3}); break;
        case RV_PASS: f.cand.push_back({110, 2, 2, 3}); break;
        case RV_CHI:
          if (a.tile >= 0 && a.consume.size() >= 2 && target >= 0)
            f.cand.push_back({(uint16_t)(111 + seq_encode_chi(a.consume, a.tile)), 2, 2, (uint16_t)seq_relative_from(pid, target)});
          break;
        case RV_PON:
          if (a.tile >= 0 && a.consume.size() >= 2 && target >= 0)
            f.cand.push_back({(uint16_t)(201 + seq_encode_pon(a.consume, a.tile)), 2, 2, (uint16_t)seq_relative_from(pid, target)});
          break;
        case RV_DAIMINKAN:
          if (a.tile >= 0 && target >= 0)
            f.cand.push_back({(uint16_t)(241 + seq_tile_id_to_kan37(a.tile)), 2, 2, (uint16_t)seq_relative_from(pid, target)});
          break;
        case RV_RON:
          if (target >= 0) f.cand.push_back({278, 2, 2, (uint16_t)seq_relative_from(pid, target)});
          break;
        default: break;   // Riichi, Kita
      }
    }
  }
  return f;
}

}  // namespace orc
