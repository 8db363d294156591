// ORACLE — TEST INFRASTRUCTURE ONLY (see hand.hpp header).
//
// Shanten number.  The reference (shanten.rs:163-239) uses Cryolite's "nyanten"
// perfect-hash tables (data/*.bin), which encode the exact replacement number of the
// normal form.  The tables are data of a third-party algorithm; this oracle restates
// the DEFINITION those tables encode — the minimum number of tiles that must be
// exchanged to reach `m` mentsu + one pair, where a target hand may not use more than
// four copies of any tile — by exhaustive enumeration of per-suit targets, then the
// reference's closed forms for chiitoitsu / kokushi and its combination rule
// (shanten.rs:198-239).  Pinned against the reference's own tables through
// tests/golden/shanten_golden.txt (generated from data/*.bin by
// tests/golden/make_golden.py) and tests/test_shanten.py's known answers.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <unordered_map>

namespace orc {

struct SuitCost {
  uint8_t c[5][2];  // [mentsu][has_head] -> tiles missing
};

namespace sh {
inline void rec(const uint8_t* cnt, int n, bool honors, int start_type, int k, uint8_t* target, SuitCost& out) {
  int cost = 0;
  for (int i = 0; i < n; i++)
    if (target[i] > cnt[i]) cost += target[i] - cnt[i];
  out.c[k][0] = std::min<uint8_t>(out.c[k][0], (uint8_t)cost);
  for (int p = 0; p < n; p++) {
    if (target[p] + 2 > 4) continue;
    int add = 0;
    for (int j = 1; j <= 2; j++)
      if (target[p] + j > cnt[p]) add++;
    out.c[k][1] = std::min<uint8_t>(out.c[k][1], (uint8_t)(cost + add));
  }
  if (k == 4) return;
  int n_types = honors ? n : n + 7;
  for (int ty = start_type; ty < n_types; ty++) {
    if (ty < n) {  // koutsu
      if (target[ty] + 3 > 4) continue;
      target[ty] += 3;
      rec(cnt, n, honors, ty, k + 1, target, out);
      target[ty] -= 3;
    } else {  // shuntsu starting at s
      int s = ty - n;
      if (target[s] + 1 > 4 || target[s + 1] + 1 > 4 || target[s + 2] + 1 > 4) continue;
      target[s]++;
      target[s + 1]++;
      target[s + 2]++;
      rec(cnt, n, honors, ty, k + 1, target, out);
      target[s]--;
      target[s + 1]--;
      target[s + 2]--;
    }
  }
}
inline SuitCost suit_cost(const uint8_t* cnt, int n, bool honors) {
  static std::unordered_map<uint64_t, SuitCost> memo[2];
  static std::mutex mu;
  uint64_t key = 0;
  for (int i = 0; i < n; i++) key = key * 5 + cnt[i];
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = memo[honors].find(key);
    if (it != memo[honors].end()) return it->second;
  }
  SuitCost out;
  memset(&out, 99, sizeof out);
  uint8_t target[9] = {0};
  rec(cnt, n, honors, 0, 0, target, out);
  std::lock_guard<std::mutex> g(mu);
  memo[honors][key] = out;
  return out;
}
}  // namespace sh

// replacement number of the normal form minus one (shanten.rs:186-196 computes the same value by table lookup)
inline int shanten_normal(const uint8_t* t, int m) {
  SuitCost sc[4] = {sh::suit_cost(t, 9, false), sh::suit_cost(t + 9, 9, false), sh::suit_cost(t + 18, 9, false),
                    sh::suit_cost(t + 27, 7, true)};
  int best[5][2];
  for (int k = 0; k < 5; k++) best[k][0] = sc[0].c[k][0], best[k][1] = sc[0].c[k][1];
  for (int s = 1; s < 4; s++) {
    int nb[5][2];
    for (int k = 0; k < 5; k++) nb[k][0] = nb[k][1] = 999;
    for (int a = 0; a < 5; a++)
      for (int b = 0; a + b < 5; b++) {
        nb[a + b][0] = std::min(nb[a + b][0], best[a][0] + sc[s].c[b][0]);
        nb[a + b][1] = std::min(nb[a + b][1], std::min(best[a][1] + sc[s].c[b][0], best[a][0] + sc[s].c[b][1]));
      }
    memcpy(best, nb, sizeof best);
  }
  return best[m][1] - 1;
}
// shanten.rs:198-211
inline int shanten_chitoi(const uint8_t* t) {
  int pairs = 0, kinds = 0;
  for (int i = 0; i < 34; i++)
    if (t[i] > 0) {
      kinds++;
      if (t[i] >= 2) pairs++;
    }
  int red = kinds < 7 ? 7 - kinds : 0;
  return 7 - pairs + red - 1;
}
// shanten.rs:213-226
inline int shanten_kokushi(const uint8_t* t) {
  static const int T[13] = {0, 8, 9, 17, 18, 26, 27, 28, 29, 30, 31, 32, 33};
  int kinds = 0, pair = 0;
  for (int i : T)
    if (t[i] > 0) {
      kinds++;
      if (t[i] >= 2) pair = 1;
    }
  return 14 - kinds - pair - 1;
}
// shanten.rs:228-239
inline int shanten_from_counts(const uint8_t* t, int len_div3) {
  if (len_div3 > 4) len_div3 = 4;
  int s = shanten_normal(t, len_div3);
  if (s <= 0 || len_div3 < 4) return s;
  s = std::min(s, shanten_chitoi(t));
  if (s > 0) return std::min(s, shanten_kokushi(t));
  return s;
}

// shanten.rs:407-435 — 3P normal form: 1m / 9m cannot form sequences, so their counts are relocated into EMPTY honor slots
// (the honor part of the tables is koutsu / pair only and position independent) before the 4P lookup; when no empty slot is
// left the count stays in manzu (the reference's overflow fallback, kept)
inline int shanten_normal_3p(const uint8_t* tiles, int m) {
  uint8_t t[34];
  memcpy(t, tiles, 34);
  const uint8_t mc[2] = {t[0], t[8]};
  const int mp[2] = {0, 8};
  t[0] = t[8] = 0;
  int next_slot = 27;
  for (int i = 0; i < 2; i++) {
    if (mc[i] == 0) continue;
    while (next_slot < 34 && t[next_slot] != 0) next_slot++;
    if (next_slot < 34) t[next_slot++] = mc[i];
    else t[mp[i]] = mc[i];
  }
  return shanten_normal(t, m);
}
// shanten.rs:437-453 — 2m-8m (kinds 1..7) do not exist in 3P and are skipped
inline int shanten_chitoi_3p(const uint8_t* t) {
  int pairs = 0, kinds = 0;
  for (int i = 0; i < 34; i++) {
    if (i >= 1 && i <= 7) continue;
    if (t[i] > 0) {
      kinds++;
      if (t[i] >= 2) pairs++;
    }
  }
  int red = kinds < 7 ? 7 - kinds : 0;
  return 7 - pairs + red - 1;
}
// shanten.rs:455-468
inline int shanten_from_counts_3p(const uint8_t* t, int len_div3) {
  if (len_div3 > 4) len_div3 = 4;
  int s = shanten_normal_3p(t, len_div3);
  if (s <= 0 || len_div3 < 4) return s;
  s = std::min(s, shanten_chitoi_3p(t));
  if (s > 0) return std::min(s, shanten_kokushi(t));
  return s;
}

}  // namespace orc
