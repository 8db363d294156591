// ORACLE — TEST INFRASTRUCTURE ONLY (see hand.hpp header).
//
// Observation::encode / mask restated from observation/python.rs:457-806, 98-111 and
// action.rs:158-227, on top of the snapshot GameState::get_observation builds
// (state/mod.rs:189-263).  Channel spec: docs/FEATURE_ENCODING.md:8-82.
#pragma once
#include "game.hpp"

namespace orc {

inline int obs_next_tile(int tile) {  // observation/helpers.rs:25-50 (tid in, tid of copy 0 out)
  int ty = (tile / 4) / 9, num = (tile / 4) % 9;
  if (ty < 3) return (ty * 9 + (num == 8 ? 0 : num + 1)) * 4;
  int base = tile / 4;
  if (base >= 27 && base < 31) return (27 + (base - 27 + 1) % 4) * 4;
  if (base >= 31 && base < 34) return (31 + (base - 31 + 1) % 3) * 4;
  return tile;
}

// out: 74*34 floats, channel-major
inline void encode_obs(const GameState& g, int pid, float* arr) {
  auto A = [&](int ch, int col) -> float& { return arr[ch * 34 + col]; };
  for (int i = 0; i < 74 * 34; i++) arr[i] = 0.0f;
  const auto& hand = g.players[pid].hand;
  HandEvaluator he(hand, g.players[pid].melds);
  std::vector<uint8_t> waits = he.get_waits_u8();
  bool is_tenpai = !waits.empty();
  int rel[4] = {pid, (pid + 1) % 4, (pid + 2) % 4, (pid + 3) % 4};
  uint8_t counts[34] = {0};
  for (uint8_t t : hand) {
    int idx = t / 4;
    counts[idx]++;
    if (t == 16 || t == 52 || t == 88) A(4, idx) = 1.0f;
  }
  for (int i = 0; i < 34; i++) {
    if (counts[i] >= 1) A(0, i) = 1.0f;
    if (counts[i] >= 2) A(1, i) = 1.0f;
    if (counts[i] >= 3) A(2, i) = 1.0f;
    if (counts[i] >= 4) A(3, i) = 1.0f;
  }
  {
    int m_idx = 0;
    for (auto& m : g.players[pid].melds) {
      if (m_idx >= 4) break;
      for (uint8_t t : m.tiles) A(5 + m_idx, t / 4) = 1.0f;
      m_idx++;
    }
  }
  for (uint8_t t : g.dora_indicators) A(9, t / 4) = 1.0f;
  auto tail = [&](int p, int skip, int take, int ch_base) {
    const auto& d = g.players[p].discards;
    int i = 0;
    for (int k = (int)d.size() - 1 - skip; k >= 0 && i < take; k--, i++) A(ch_base + i, d[k] / 4) = 1.0f;
  };
  tail(pid, 0, 4, 10);
  for (int i = 1; i < 4; i++) tail((pid + i) % 4, 0, 4, 14 + (i - 1) * 4);
  for (int c = 0; c < 4; c++) {
    float v = (float)g.players[rel[c]].discards.size() / 24.0f;
    for (int k = 0; k < 34; k++) A(26 + c, k) = v;
  }
  {
    int used = 0;
    for (auto& p : g.players) used += (int)p.discards.size();
    for (auto& p : g.players)
      for (auto& m : p.melds) used += (int)m.tiles.size();
    used += (int)hand.size();
    used += (int)g.dora_indicators.size();
    float left = (float)std::max(136 - used, 0);
    float v = left / 70.0f;
    for (int k = 0; k < 34; k++) A(30, k) = v;
  }
  if (g.players[pid].riichi_declared)
    for (int k = 0; k < 34; k++) A(31, k) = 1.0f;
  for (int i = 1; i < 4; i++)
    if (g.players[(pid + i) % 4].riichi_declared)
      for (int k = 0; k < 34; k++) A(32 + i - 1, k) = 1.0f;
  if (27 + g.round_wind < 34) A(35, 27 + g.round_wind) = 1.0f;
  A(36, 27 + (pid + 4 - g.oya) % 4) = 1.0f;
  {
    float hn = (float)g.honba / 10.0f, sn = (float)g.riichi_sticks / 5.0f;
    for (int k = 0; k < 34; k++) {
      A(37, k) = hn;
      A(38, k) = sn;
    }
  }
  for (int c = 0; c < 4; c++) {
    int s = g.players[rel[c]].score;
    float v1 = (float)std::min(std::max(s, 0), 100000) / 100000.0f;
    float v2 = (float)std::min(std::max(s, 0), 30000) / 30000.0f;
    for (int k = 0; k < 34; k++) {
      A(39 + c, k) = v1;
      A(43 + c, k) = v2;
    }
  }
  for (uint8_t t : waits) A(47, t) = 1.0f;
  for (int k = 0; k < 34; k++) A(48, k) = is_tenpai ? 1.0f : 0.0f;
  {
    int rank = 0;
    for (auto& p : g.players)
      if (p.score > g.players[pid].score) rank++;
    if (rank < 4)
      for (int k = 0; k < 34; k++) A(49 + rank, k) = 1.0f;
  }
  for (int k = 0; k < 34; k++) {
    A(53, k) = (float)g.kyoku_idx / 8.0f;
    A(54, k) = ((float)g.round_wind * 4.0f + (float)g.kyoku_idx) / 7.0f;
  }
  {
    uint8_t dc[4] = {0, 0, 0, 0};
    for (int p = 0; p < 4; p++) {
      for (auto& m : g.players[p].melds)
        for (uint8_t t : m.tiles)
          for (uint8_t di : g.dora_indicators)
            if (t / 4 == obs_next_tile(di) / 4) dc[p]++;
      for (uint8_t t : g.players[p].discards)
        for (uint8_t di : g.dora_indicators)
          if (t / 4 == obs_next_tile(di) / 4) dc[p]++;
    }
    for (uint8_t t : hand)
      for (uint8_t di : g.dora_indicators)
        if (t / 4 == obs_next_tile(di) / 4) dc[pid]++;
    for (int c = 0; c < 4; c++) {
      float v = (float)dc[rel[c]] / 12.0f;
      for (int k = 0; k < 34; k++) A(55 + c, k) = v;
    }
  }
  for (int c = 0; c < 4; c++) {
    float v = (float)g.players[rel[c]].melds.size() / 4.0f;
    for (int k = 0; k < 34; k++) A(59 + c, k) = v;
  }
  {
    uint8_t seen[34] = {0};
    for (uint8_t t : hand) seen[t / 4]++;
    for (auto& p : g.players)
      for (auto& m : p.melds)
        for (uint8_t t : m.tiles) seen[t / 4]++;
    for (auto& p : g.players)
      for (uint8_t t : p.discards) seen[t / 4]++;
    for (uint8_t t : g.dora_indicators) seen[t / 4]++;
    for (int i = 0; i < 34; i++) A(63, i) = (float)seen[i] / 4.0f;
  }
  tail(pid, 4, 4, 64);
  tail((pid + 1) % 4, 4, 2, 68);
  // 70-73: tsumogiri_flags is always empty in the live env (observation/mod.rs:105)
}

// Action::encode (action.rs:158-227); -1 on error
inline int action_encode(const Action& a) {
  switch (a.type) {
    case RV_DISCARD: return a.tile < 0 ? -1 : a.tile / 4;
    case RV_RIICHI: return 37;
    case RV_CHI: {
      if (a.tile < 0) return -1;
      std::vector<int> t34;
      for (uint8_t x : a.consume) t34.push_back(x / 4);
      t34.push_back(a.tile / 4);
      std::sort(t34.begin(), t34.end());
      t34.erase(std::unique(t34.begin(), t34.end()), t34.end());
      if (t34.size() != 3) return -1;
      int tg = a.tile / 4;
      return tg == t34[0] ? 38 : tg == t34[1] ? 39 : 40;
    }
    case RV_PON: return 41;
    case RV_DAIMINKAN: return a.tile < 0 ? -1 : 42 + a.tile / 4;
    case RV_ANKAN:
    case RV_KAKAN: return a.consume.empty() ? -1 : 42 + a.consume[0] / 4;
    case RV_RON:
    case RV_TSUMO: return 79;
    case RV_KYUSHU_KYUHAI: return 80;
    case RV_PASS: return 81;
  }
  return -1;
}
// Observation::mask (observation/python.rs:98-111) for a seat that owes an action
inline void encode_mask(const GameState& g, int pid, uint8_t* out82) {
  memset(out82, 0, 82);
  for (auto& a : g._get_legal_actions_internal(pid)) {
    int id = action_encode(a);
    if (id >= 0 && id < 82) out82[id] = 1;
  }
}

}  // namespace orc
