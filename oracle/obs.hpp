// ORACLE — TEST INFRASTRUCTURE ONLY (see hand.hpp header).
//
// Observation::encode / mask restated from observation/python.rs:457-806, 98-111 and
// action.rs:158-227, on top of the snapshot GameState::get_observation builds
// (state/mod.rs:189-263).  Channel spec: docs/FEATURE_ENCODING.md:8-82.
#pragma once
#include "game.hpp"
#include "shanten.hpp"
#include <cmath>

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
// ---- sanma: Observation3P::encode (observation_3p/python.rs:402-708), 74 channels x 27 compact columns ----
inline int compact3(int t34) {  // observation_3p/helpers.rs:8-15; -1 = None
  if (t34 == 0) return 0;
  if (t34 >= 8 && t34 <= 33) return t34 - 7;
  return -1;
}
inline int obs_next_tile_sanma(int tile) {  // observation_3p/helpers.rs:41-50
  int t34 = tile / 4;
  if (t34 == 0) return 8 * 4;
  if (t34 == 8) return 0;
  if (t34 >= 1 && t34 <= 7) return tile;
  return obs_next_tile(tile);
}
// out: 74*27 floats, channel-major
inline void encode_obs_3p(const GameState& g, int pid, float* arr) {
  const int W = 27, NP3 = 3;
  auto A = [&](int ch, int col) -> float& { return arr[ch * W + col]; };
  auto set = [&](int ch, int tile) {
    int c = compact3(tile / 4);
    if (c >= 0) A(ch, c) = 1.0f;
  };
  for (int i = 0; i < 74 * W; i++) arr[i] = 0.0f;
  const auto& hand = g.players[pid].hand;
  HandEvaluator he(hand, g.players[pid].melds, true);   // state_3p/mod.rs:189-194
  std::vector<uint8_t> waits = he.get_waits_u8();
  bool is_tenpai = !waits.empty();
  int rel[3] = {pid, (pid + 1) % 3, (pid + 2) % 3};      // observation_3p/mod.rs:120-123
  uint8_t counts[27] = {0};
  for (uint8_t t : hand) {
    int idx = compact3(t / 4);
    if (idx < 0) continue;
    counts[idx]++;
    if (t == 16 || t == 52 || t == 88) A(4, idx) = 1.0f;
  }
  for (int i = 0; i < W; i++)
    for (int k = 1; k <= 4; k++)
      if (counts[i] >= k) A(k - 1, i) = 1.0f;
  {
    int m_idx = 0;
    for (auto& m : g.players[pid].melds) {
      if (m_idx >= 4) break;
      for (uint8_t t : m.tiles) set(5 + m_idx, t);
      m_idx++;
    }
  }
  for (uint8_t t : g.dora_indicators) set(9, t);
  auto tail = [&](int p, int skip, int take, int ch_base) {
    const auto& d = g.players[p].discards;
    int i = 0;
    for (int k = (int)d.size() - 1 - skip; k >= 0 && i < take; k--, i++) set(ch_base + i, d[k]);
  };
  tail(pid, 0, 4, 10);
  for (int i = 1; i < NP3; i++) tail((pid + i) % NP3, 0, 4, 14 + (i - 1) * 4);
  for (int c = 0; c < NP3; c++) {
    float v = (float)g.players[rel[c]].discards.size() / 24.0f;
    for (int k = 0; k < W; k++) A(26 + c, k) = v;
  }
  {
    int used = 0;
    for (int p = 0; p < NP3; p++) used += (int)g.players[p].discards.size();
    for (int p = 0; p < NP3; p++)
      for (auto& m : g.players[p].melds) used += (int)m.tiles.size();
    used += (int)hand.size();
    used += (int)g.dora_indicators.size();
    float v = (float)std::max(108 - used, 0) / 70.0f;
    for (int k = 0; k < W; k++) A(30, k) = v;
  }
  if (g.players[pid].riichi_declared)
    for (int k = 0; k < W; k++) A(31, k) = 1.0f;
  for (int i = 1; i < NP3; i++)
    if (g.players[(pid + i) % NP3].riichi_declared)
      for (int k = 0; k < W; k++) A(32 + i - 1, k) = 1.0f;
  {
    int c = compact3(27 + g.round_wind);
    if (c >= 0) A(35, c) = 1.0f;
    c = compact3(27 + (pid + NP3 - g.oya) % NP3);
    if (c >= 0) A(36, c) = 1.0f;
  }
  for (int k = 0; k < W; k++) {
    A(37, k) = (float)g.honba / 10.0f;
    A(38, k) = (float)g.riichi_sticks / 5.0f;
  }
  for (int c = 0; c < NP3; c++) {
    int s = g.players[rel[c]].score;
    float v1 = (float)std::min(std::max(s, 0), 100000) / 100000.0f;
    float v2 = (float)std::min(std::max(s, 0), 30000) / 30000.0f;
    for (int k = 0; k < W; k++) {
      A(39 + c, k) = v1;
      A(43 + c, k) = v2;
    }
  }
  for (uint8_t t : waits) {
    int c = compact3(t);
    if (c >= 0) A(47, c) = 1.0f;
  }
  for (int k = 0; k < W; k++) A(48, k) = is_tenpai ? 1.0f : 0.0f;
  {
    int rank = 0;
    for (int p = 0; p < NP3; p++)
      if (g.players[p].score > g.players[pid].score) rank++;
    if (rank < NP3)
      for (int k = 0; k < W; k++) A(49 + rank, k) = 1.0f;
  }
  for (int k = 0; k < W; k++) {
    A(53, k) = (float)g.kyoku_idx / 8.0f;
    A(54, k) = ((float)g.round_wind * 4.0f + (float)g.kyoku_idx) / 7.0f;
  }
  {
    uint8_t dc[3] = {0, 0, 0};
    for (int p = 0; p < NP3; p++) {
      for (auto& m : g.players[p].melds)
        for (uint8_t t : m.tiles)
          for (uint8_t di : g.dora_indicators)
            if (t / 4 == obs_next_tile_sanma(di) / 4) dc[p]++;
      for (uint8_t t : g.players[p].discards)
        for (uint8_t di : g.dora_indicators)
          if (t / 4 == obs_next_tile_sanma(di) / 4) dc[p]++;
    }
    for (uint8_t t : hand)
      for (uint8_t di : g.dora_indicators)
        if (t / 4 == obs_next_tile_sanma(di) / 4) dc[pid]++;
    for (int c = 0; c < NP3; c++) {
      float v = (float)dc[rel[c]] / 12.0f;
      for (int k = 0; k < W; k++) A(55 + c, k) = v;
    }
  }
  for (int c = 0; c < NP3; c++) {
    float v = (float)g.players[rel[c]].melds.size() / 4.0f;
    for (int k = 0; k < W; k++) A(59 + c, k) = v;
  }
  {
    uint8_t seen[27] = {0};
    auto see = [&](uint8_t t) {
      int c = compact3(t / 4);
      if (c >= 0) seen[c]++;
    };
    for (uint8_t t : hand) see(t);
    for (int p = 0; p < NP3; p++)
      for (auto& m : g.players[p].melds)
        for (uint8_t t : m.tiles) see(t);
    for (int p = 0; p < NP3; p++)
      for (uint8_t t : g.players[p].discards) see(t);
    for (uint8_t t : g.dora_indicators) see(t);
    for (int i = 0; i < W; i++) A(63, i) = (float)seen[i] / 4.0f;
  }
  tail(pid, 4, 4, 64);
  tail((pid + 1) % NP3, 4, 2, 68);
  // 70-72: tsumogiri_flags is always empty in the live env (observation_3p/mod.rs:100)
}
// ActionEncoder::encode_3p (action.rs:262-346); -1 on error
inline int action_encode_3p(const Action& a) {
  switch (a.type) {
    case RV_DISCARD: return a.tile < 0 ? -1 : compact3(a.tile / 4);
    case RV_RIICHI: return 27;
    case RV_CHI: return -1;
    case RV_PON: return 28;
    case RV_DAIMINKAN: {
      int c = a.tile < 0 ? -1 : compact3(a.tile / 4);
      return c < 0 ? -1 : 29 + c;
    }
    case RV_ANKAN:
    case RV_KAKAN: {
      int c = a.consume.empty() ? -1 : compact3(a.consume[0] / 4);
      return c < 0 ? -1 : 29 + c;
    }
    case RV_RON:
    case RV_TSUMO: return 56;
    case RV_KYUSHU_KYUHAI: return 57;
    case RV_PASS: return 58;
    case RV_KITA: return 59;
  }
  return -1;
}
// ---- extended encoders: Observation::encode_extended (observation/python.rs:1272-1294), 215 channels x 34 -------------
// shanten.rs:250-261 on a tid list
inline int shanten_tiles(const std::vector<int>& tiles) {
  uint8_t cnt[34] = {0};
  int n = 0;
  for (int t : tiles)
    if (t / 4 < 34) cnt[t / 4]++, n++;
  return shanten_from_counts(cnt, n / 3);
}
// shanten.rs:265-296 (caller guarantees a 3n+1 hand)
inline int effective_tiles(const std::vector<int>& hand) {
  int cur = shanten_tiles(hand), eff = 0;
  uint8_t hc[34] = {0};
  for (int t : hand)
    if (t / 4 < 34) hc[t / 4]++;
  for (int k = 0; k < 34; k++) {
    if (hc[k] >= 4) continue;
    std::vector<int> nh = hand;
    nh.push_back(k * 4);
    if (shanten_tiles(nh) < cur) eff++;
  }
  return eff;
}
// shanten.rs:301-327
inline int effective_tiles_with_discard(const std::vector<int>& hand) {
  if (hand.size() % 3 == 1) return effective_tiles(hand);
  int sh = shanten_tiles(hand), best = 0;
  for (size_t idx = 0; idx < hand.size(); idx++) {
    std::vector<int> sub;
    for (size_t i = 0; i < hand.size(); i++)
      if (i != idx) sub.push_back(hand[i]);
    if (shanten_tiles(sub) <= sh) best = std::max(best, effective_tiles(sub));
  }
  return best;
}
// shanten.rs:331-393
inline int best_ukeire(const std::vector<int>& hand, const std::vector<int>& visible) {
  int best = 0;
  uint32_t vis[34] = {0};
  for (int t : visible)
    if (t / 4 < 34) vis[t / 4]++;
  int cur = shanten_tiles(hand);
  uint8_t base[34] = {0};
  for (int t : hand)
    if (t / 4 < 34) base[t / 4]++;
  for (size_t idx = 0; idx < hand.size(); idx++) {
    std::vector<int> sub;
    for (size_t i = 0; i < hand.size(); i++)
      if (i != idx) sub.push_back(hand[i]);
    uint8_t nc[34];
    memcpy(nc, base, 34);
    nc[hand[idx] / 4]--;
    int ns = shanten_tiles(sub);
    if (ns > cur) continue;
    uint32_t uke = 0;
    for (int k = 0; k < 34; k++) {
      if (nc[k] >= 4) continue;
      std::vector<int> th = sub;
      th.push_back(k * 4);
      if (shanten_tiles(th) < ns) {
        uint32_t rem = 4u > vis[k] ? 4u - vis[k] : 0u;          // saturating_sub twice
        rem = rem > nc[k] ? rem - nc[k] : 0u;
        uke += rem;
      }
    }
    best = std::max(best, (int)uke);
  }
  return best;
}

// ---- 3P variants (shanten.rs:470-615), groundwork for Observation3P::encode_extended: draws range over the 27 tile kinds
// that exist in sanma (SANMA_VALID_TILE_TYPES, shanten.rs:241-247)
inline int shanten_tiles_3p(const std::vector<int>& tiles) {
  uint8_t cnt[34] = {0};
  int n = 0;
  for (int t : tiles)
    if (t / 4 < 34) cnt[t / 4]++, n++;
  return shanten_from_counts_3p(cnt, n / 3);
}
inline bool sanma_kind(int k) { return k == 0 || (k >= 8 && k < 34); }
inline int effective_tiles_3p(const std::vector<int>& hand) {
  int cur = shanten_tiles_3p(hand), eff = 0;
  uint8_t hc[34] = {0};
  for (int t : hand)
    if (t / 4 < 34) hc[t / 4]++;
  for (int k = 0; k < 34; k++) {
    if (!sanma_kind(k) || hc[k] >= 4) continue;
    std::vector<int> nh = hand;
    nh.push_back(k * 4);
    if (shanten_tiles_3p(nh) < cur) eff++;
  }
  return eff;
}
inline int effective_tiles_3p_with_discard(const std::vector<int>& hand) {
  if (hand.size() % 3 == 1) return effective_tiles_3p(hand);
  int sh = shanten_tiles_3p(hand), best = 0;
  for (size_t idx = 0; idx < hand.size(); idx++) {
    std::vector<int> sub;
    for (size_t i = 0; i < hand.size(); i++)
      if (i != idx) sub.push_back(hand[i]);
    if (shanten_tiles_3p(sub) <= sh) best = std::max(best, effective_tiles_3p(sub));
  }
  return best;
}
inline int best_ukeire_3p(const std::vector<int>& hand, const std::vector<int>& visible) {
  int best = 0;
  uint32_t vis[34] = {0};
  for (int t : visible)
    if (t / 4 < 34) vis[t / 4]++;
  int cur = shanten_tiles_3p(hand);
  uint8_t base[34] = {0};
  for (int t : hand)
    if (t / 4 < 34) base[t / 4]++;
  for (size_t idx = 0; idx < hand.size(); idx++) {
    std::vector<int> sub;
    for (size_t i = 0; i < hand.size(); i++)
      if (i != idx) sub.push_back(hand[i]);
    uint8_t nc[34];
    memcpy(nc, base, 34);
    nc[hand[idx] / 4]--;
    int ns = shanten_tiles_3p(sub);
    if (ns > cur) continue;
    uint32_t uke = 0;
    for (int k = 0; k < 34; k++) {
      if (!sanma_kind(k) || nc[k] >= 4) continue;
      std::vector<int> th = sub;
      th.push_back(k * 4);
      if (shanten_tiles_3p(th) < ns) {
        uint32_t rem = 4u > vis[k] ? 4u - vis[k] : 0u;
        rem = rem > nc[k] ? rem - nc[k] : 0u;
        uke += rem;
      }
    }
    best = std::max(best, (int)uke);
  }
  return best;
}

// out: 215*34 floats, channel-major.  Channel blocks (python.rs:1281-1290): base 0, decay 74, shanten 78, ankan 94,
// fuuro 98, action availability 178, discard candidates 189, pass context 194, last tedashi 197, riichi sutehai 206.
inline void encode_obs_extended(const GameState& g, int pid, float* arr) {
  auto A = [&](int ch, int col) -> float& { return arr[ch * 34 + col]; };
  auto B = [&](int ch, float v) { for (int k = 0; k < 34; k++) arr[ch * 34 + k] = v; };
  for (int i = 0; i < 215 * 34; i++) arr[i] = 0.0f;
  encode_obs(g, pid, arr);   // observation/encode.rs:12-291 == python.rs:457-806 except channel 30 (below)
  const int rel[4] = {pid, (pid + 1) % 4, (pid + 2) % 4, (pid + 3) % 4};
  std::vector<int> hand(g.players[pid].hand.begin(), g.players[pid].hand.end());
  {
    // encode.rs:94-110: encode_base_into counts a called meld one tile short ("already counted in discards");
    // Observation::encode (python.rs) does not.  Both kept as they are.
    int used = 0;
    for (auto& p : g.players) used += (int)p.discards.size();
    for (auto& p : g.players)
      for (auto& m : p.melds) used += (int)m.tiles.size() - (m.called_tile >= 0 ? 1 : 0);
    used += (int)hand.size() + (int)g.dora_indicators.size();
    B(30, (float)std::max(136 - used, 0) / 70.0f);
  }
  // decay (encode.rs:295-313)
  for (int c = 0; c < 4; c++) {
    const auto& d = g.players[rel[c]].discards;
    for (size_t turn = 0; turn < d.size(); turn++) {
      float age = (float)(d.size() - 1 - turn);
      A(74 + c, d[turn] / 4) += expf(-0.2f * age);
    }
  }
  // shanten efficiency (encode.rs:317-350)
  {
    std::vector<int> vis;
    for (auto& p : g.players)
      for (uint8_t t : p.discards) vis.push_back(t);
    for (auto& p : g.players)
      for (auto& m : p.melds)
        for (uint8_t t : m.tiles) vis.push_back(t);
    for (uint8_t t : g.dora_indicators) vis.push_back(t);
    for (int c = 0; c < 4; c++) {
      int b = 78 + c * 4;
      if (c == 0) {
        int sv = shanten_tiles(hand);
        B(b, std::max((float)sv, 0.0f) / 8.0f);
        B(b + 1, (float)effective_tiles_with_discard(hand) / 34.0f);
        B(b + 2, (float)best_ukeire(hand, vis) / 80.0f);
      } else {
        B(b, 0.5f), B(b + 1, 0.5f), B(b + 2, 0.5f);
      }
      B(b + 3, std::min((float)g.players[rel[c]].discards.size() / 18.0f, 1.0f));
    }
  }
  // ankan (encode.rs:354-368), fuuro (encode.rs:372-396)
  for (int c = 0; c < 4; c++) {
    int mi = 0;
    for (auto& m : g.players[rel[c]].melds) {
      if (m.meld_type == Ankan && !m.tiles.empty()) A(94 + c, m.tiles[0] / 4) = 1.0f;
      if (mi < 4) {
        int slot = 0;
        for (uint8_t t : m.tiles) {
          if (slot >= 4) break;
          A(98 + c * 20 + mi * 5 + slot, t / 4) = 1.0f;
          if (t == 16 || t == 52 || t == 88) A(98 + c * 20 + mi * 5 + 4, t / 4) = 1.0f;
          slot++;
        }
      }
      mi++;
    }
  }
  // action availability (encode.rs:399-430) over the seat's legal list (empty unless the seat owes an action)
  {
    bool may = !g.is_done && ((g.phase == RV_WAIT_ACT && g.current_player == pid) ||
                              (g.phase == RV_WAIT_RESPONSE &&
                               std::find(g.active_players.begin(), g.active_players.end(), (uint8_t)pid) != g.active_players.end()));
    if (may)
      for (auto& a : g._get_legal_actions_internal(pid)) {
        switch (a.type) {
          case RV_RIICHI: B(178, 1.0f); break;
          case RV_CHI:
            if (a.consume.size() == 2) {
              int t0 = a.consume[0] / 4, t1 = a.consume[1] / 4, diff = std::abs(t1 - t0);
              if (diff == 1) B(t0 < t1 ? 179 : 181, 1.0f);
              else if (diff == 2) B(180, 1.0f);
            }
            break;
          case RV_PON: B(182, 1.0f); break;
          case RV_DAIMINKAN: B(183, 1.0f); break;
          case RV_ANKAN: B(184, 1.0f); break;
          case RV_KAKAN: B(185, 1.0f); break;
          case RV_TSUMO:
          case RV_RON: B(186, 1.0f); break;
          case RV_KYUSHU_KYUHAI: B(187, 1.0f); break;
          case RV_PASS: B(188, 1.0f); break;
          default: break;
        }
      }
  }
  // discard candidates (encode.rs:433-476)
  {
    int cur = shanten_tiles(hand), keep = 0, inc = 0;
    B(189, (float)hand.size() / 34.0f);
    for (size_t idx = 0; idx < hand.size(); idx++) {
      std::vector<int> sub;
      for (size_t i = 0; i < hand.size(); i++)
        if (i != idx) sub.push_back(hand[i]);
      int ns = shanten_tiles(sub);
      if (ns == cur) keep++;
      else if (ns > cur) inc++;
    }
    if (!hand.empty()) {
      B(190, (float)keep / (float)hand.size());
      B(191, (float)inc / (float)hand.size());
    }
    B(192, cur == -1 ? 1.0f : 0.0f);
    B(193, g.players[pid].riichi_declared ? 1.0f : 0.0f);
  }
  std::vector<uint8_t> dora_tiles;   // tids of copy 0 of each dora kind (helpers.rs:25-50 returns a tid)
  for (uint8_t di : g.dora_indicators) dora_tiles.push_back((uint8_t)obs_next_tile(di));
  auto is_dora_tid = [&](int tid) { return std::find(dora_tiles.begin(), dora_tiles.end(), (uint8_t)tid) != dora_tiles.end(); };
  auto tile_ctx = [&](int ch, int tile) {
    B(ch, (float)(tile / 4) / 33.0f);
    B(ch + 1, (tile == 16 || tile == 52 || tile == 88) ? 1.0f : 0.0f);
    B(ch + 2, is_dora_tid(tile) ? 1.0f : 0.0f);   // compares a TID with kind*4: only copy 0 of a dora kind matches
  };
  // pass context (encode.rs:479-510).  state/mod.rs:252 builds Observation.last_discard with
  // `self.last_discard.map(|(tile, _pid)| tile as u32)` on a tuple stored as (pid, tile) (state/mod.rs:1329): the value
  // the encoder sees is the DISCARDER'S SEAT, not the tile.  Restated as is.
  if (g.last_discard_pid >= 0) tile_ctx(194, g.last_discard_pid);
  // last tedashi (encode.rs:513-547), riichi sutehai (encode.rs:550-584): opponents in ABSOLUTE seat order
  {
    int opp = 0;
    for (int p = 0; p < 4; p++) {
      if (p == pid) continue;
      if (g.last_tedashis[p] >= 0) tile_ctx(197 + opp * 3, g.last_tedashis[p]);
      if (g.riichi_sutehais[p] >= 0) tile_ctx(206 + opp * 3, g.riichi_sutehais[p]);
      opp++;
    }
  }
}

// Observation::encode_kawa_overview (observation/python.rs:881-930): (4, 7, 34) floats, seats in ABSOLUTE order.
// Channels 0-3: the n-th copy of a kind a seat discarded (n capped at 4); 4-6: "aka" flags, which the reference raises for
// tile ids 20 / 24 / 28 (not 16 / 52 / 88) at columns 5 / 14 / 23 — restated as is.
inline void encode_kawa_overview(const GameState& g, float* arr) {
  for (int i = 0; i < 4 * 7 * 34; i++) arr[i] = 0.0f;
  for (int p = 0; p < 4; p++) {
    uint8_t cnt[34] = {0};
    bool aka[3] = {false, false, false};
    for (uint8_t t : g.players[p].discards) {
      int k = t / 4;
      if (k < 34) {
        arr[(p * 7 + std::min<int>(cnt[k], 3)) * 34 + k] = 1.0f;
        if (cnt[k] < 255) cnt[k]++;
      }
      if (t == 20) aka[0] = true;
      else if (t == 24) aka[1] = true;
      else if (t == 28) aka[2] = true;
    }
    for (int i = 0; i < 3; i++)
      if (aka[i]) arr[(p * 7 + 4 + i) * 34 + 5 + i * 9] = 1.0f;
  }
}

// Observation3P::encode_kawa_overview (observation_3p/python.rs:760-808): (3, 7, 27) floats over the compact columns.  The aka
// flags test tile ids 20 / 24 / 28 as in 4P; id 20's kind (5m... the reference's comment) has no sanma column, ids 24 / 28 raise
// channel 5 at compact(13) = 6 and channel 6 at compact(22) = 15 — restated as is (no sanma wall holds those ids).
inline void encode_kawa_overview_3p(const GameState& g, float* arr) {
  const int W = 27;
  auto compact = [](int k) { return k == 0 ? 0 : (k >= 8 && k < 34 ? k - 7 : -1); };
  for (int i = 0; i < 3 * 7 * W; i++) arr[i] = 0.0f;
  for (int p = 0; p < 3; p++) {
    uint8_t cnt[27] = {0};
    bool aka[3] = {false, false, false};
    for (uint8_t t : g.players[p].discards) {
      int idx = compact(t / 4);
      if (idx >= 0) {
        arr[(p * 7 + std::min<int>(cnt[idx], 3)) * W + idx] = 1.0f;
        if (cnt[idx] < 255) cnt[idx]++;
      }
      if (t == 20) aka[0] = true;
      else if (t == 24) aka[1] = true;
      else if (t == 28) aka[2] = true;
    }
    if (aka[1]) arr[(p * 7 + 5) * W + 6] = 1.0f;
    if (aka[2]) arr[(p * 7 + 6) * W + 15] = 1.0f;
  }
}

// Observation3P::encode_extended (observation_3p/python.rs:1117-1140; blocks observation_3p/encode.rs:22-620): 215 x 27 floats.
// Same block offsets as 4P; three relative seats per group (the fourth channel of a group stays zero), two opponents in the
// absolute-order blocks, compact columns, tile kind / 26, effective tiles / 27, the sanma dora successor, and the same two
// quirks (called melds count one tile short in channel 30; the pass context is fed the discarder's seat,
// state_3p/mod.rs:220).  A statement-level diff of encode_base_into against Observation3P::encode shows channel 30 as the only
// difference.  (Oracle only so far: the device kernel for these rows is round-2 work.)
inline void encode_obs_3p_extended(const GameState& g, int pid, float* arr) {
  const int W = 27;
  auto A = [&](int ch, int col) -> float& { return arr[ch * W + col]; };
  auto B = [&](int ch, float v) { for (int k = 0; k < W; k++) arr[ch * W + k] = v; };
  for (int i = 0; i < 215 * W; i++) arr[i] = 0.0f;
  encode_obs_3p(g, pid, arr);
  const int rel[3] = {pid, (pid + 1) % 3, (pid + 2) % 3};
  std::vector<int> hand(g.players[pid].hand.begin(), g.players[pid].hand.end());
  {
    int used = 0;
    for (int p = 0; p < 3; p++) used += (int)g.players[p].discards.size();
    for (int p = 0; p < 3; p++)
      for (auto& m : g.players[p].melds) used += (int)m.tiles.size() - (m.called_tile >= 0 ? 1 : 0);
    used += (int)hand.size() + (int)g.dora_indicators.size();
    B(30, (float)std::max(108 - used, 0) / 70.0f);
  }
  for (int c = 0; c < 3; c++) {   // decay (encode.rs:315-333)
    const auto& d = g.players[rel[c]].discards;
    for (size_t turn = 0; turn < d.size(); turn++) {
      int col = compact3(d[turn] / 4);
      if (col >= 0) A(74 + c, col) += expf(-0.2f * (float)(d.size() - 1 - turn));
    }
  }
  {   // shanten efficiency (encode.rs:337-370)
    std::vector<int> vis;
    for (int p = 0; p < 3; p++)
      for (uint8_t t : g.players[p].discards) vis.push_back(t);
    for (int p = 0; p < 3; p++)
      for (auto& m : g.players[p].melds)
        for (uint8_t t : m.tiles) vis.push_back(t);
    for (uint8_t t : g.dora_indicators) vis.push_back(t);
    for (int c = 0; c < 3; c++) {
      int b = 78 + c * 4;
      if (c == 0) {
        B(b, std::max((float)shanten_tiles_3p(hand), 0.0f) / 8.0f);
        B(b + 1, (float)effective_tiles_3p_with_discard(hand) / 27.0f);
        B(b + 2, (float)best_ukeire_3p(hand, vis) / 80.0f);
      } else {
        B(b, 0.5f), B(b + 1, 0.5f), B(b + 2, 0.5f);
      }
      B(b + 3, std::min((float)g.players[rel[c]].discards.size() / 18.0f, 1.0f));
    }
  }
  for (int c = 0; c < 3; c++) {   // ankan (encode.rs:374-387), fuuro (encode.rs:392-418)
    int mi = 0;
    for (auto& m : g.players[rel[c]].melds) {
      if (m.meld_type == Ankan && !m.tiles.empty() && compact3(m.tiles[0] / 4) >= 0) A(94 + c, compact3(m.tiles[0] / 4)) = 1.0f;
      if (mi < 4) {
        int slot = 0;
        for (uint8_t t : m.tiles) {
          if (slot >= 4) break;
          int col = compact3(t / 4);
          if (col >= 0) A(98 + c * 20 + mi * 5 + slot, col) = 1.0f;
          if ((t == 16 || t == 52 || t == 88) && col >= 0) A(98 + c * 20 + mi * 5 + 4, col) = 1.0f;
          slot++;
        }
      }
      mi++;
    }
  }
  {   // action availability (encode.rs:420-452)
    bool may = !g.is_done && ((g.phase == RV_WAIT_ACT && g.current_player == pid) ||
                              (g.phase == RV_WAIT_RESPONSE &&
                               std::find(g.active_players.begin(), g.active_players.end(), (uint8_t)pid) != g.active_players.end()));
    if (may)
      for (auto& a : g._get_legal_actions_internal(pid)) {
        switch (a.type) {
          case RV_RIICHI: B(178, 1.0f); break;
          case RV_CHI:
            if (a.consume.size() == 2) {
              int t0 = a.consume[0] / 4, t1 = a.consume[1] / 4, diff = std::abs(t1 - t0);
              if (diff == 1) B(t0 < t1 ? 179 : 181, 1.0f);
              else if (diff == 2) B(180, 1.0f);
            }
            break;
          case RV_PON: B(182, 1.0f); break;
          case RV_DAIMINKAN: B(183, 1.0f); break;
          case RV_ANKAN: B(184, 1.0f); break;
          case RV_KAKAN: B(185, 1.0f); break;
          case RV_TSUMO:
          case RV_RON: B(186, 1.0f); break;
          case RV_KYUSHU_KYUHAI: B(187, 1.0f); break;
          case RV_PASS: B(188, 1.0f); break;
          default: break;   // Kita raises no channel
        }
      }
  }
  {   // discard candidates (encode.rs:455-498)
    int cur = shanten_tiles_3p(hand), keep = 0, inc = 0;
    B(189, (float)hand.size() / 34.0f);
    for (size_t idx = 0; idx < hand.size(); idx++) {
      std::vector<int> sub;
      for (size_t i = 0; i < hand.size(); i++)
        if (i != idx) sub.push_back(hand[i]);
      int ns = shanten_tiles_3p(sub);
      if (ns == cur) keep++;
      else if (ns > cur) inc++;
    }
    if (!hand.empty()) {
      B(190, (float)keep / (float)hand.size());
      B(191, (float)inc / (float)hand.size());
    }
    B(192, cur == -1 ? 1.0f : 0.0f);
    B(193, g.players[pid].riichi_declared ? 1.0f : 0.0f);
  }
  std::vector<uint8_t> dora_tiles;
  for (uint8_t di : g.dora_indicators) dora_tiles.push_back((uint8_t)obs_next_tile_sanma(di));
  auto tile_ctx = [&](int ch, int tile) {
    int col = compact3(tile / 4);
    if (col >= 0) B(ch, (float)col / 26.0f);
    B(ch + 1, (tile == 16 || tile == 52 || tile == 88) ? 1.0f : 0.0f);
    B(ch + 2, std::find(dora_tiles.begin(), dora_tiles.end(), (uint8_t)tile) != dora_tiles.end() ? 1.0f : 0.0f);
  };
  if (g.last_discard_pid >= 0) tile_ctx(194, g.last_discard_pid);   // the discarder's seat, as in 4P (state_3p/mod.rs:220)
  {
    int opp = 0;
    for (int p = 0; p < 3; p++) {
      if (p == pid) continue;
      if (g.last_tedashis[p] >= 0) tile_ctx(197 + opp * 3, g.last_tedashis[p]);
      if (g.riichi_sutehais[p] >= 0) tile_ctx(206 + opp * 3, g.riichi_sutehais[p]);
      opp++;
    }
  }
}

// Observation::mask (observation/python.rs:98-111; 3P: observation_3p/python.rs:102-114) for a seat that owes an action:
// 82 bytes (4P) or 60 bytes (3P)
inline void encode_mask(const GameState& g, int pid, uint8_t* out) {
  const bool sanma = g.np == 3;
  const int ids = sanma ? 60 : 82;
  memset(out, 0, ids);
  for (auto& a : g._get_legal_actions_internal(pid)) {
    int id = sanma ? action_encode_3p(a) : action_encode(a);
    if (id >= 0 && id < ids) out[id] = 1;
  }
}

}  // namespace orc
