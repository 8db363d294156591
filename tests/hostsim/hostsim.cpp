// TEST INFRASTRUCTURE ONLY (see cuda_shim.h).  Host compile of the device sources with a
// C surface parallel to oracle/capi.cpp, so tests can diff kernel logic vs oracle on CPU.
#include "cuda_shim.h"
#include "../../include/riichienv_b200.h"
// Emulation of the rollout kernels' shared-memory staging: the random-step entry points run the game code on a
// copy of the record's hot prefix followed by poison, with cold(g) redirected to the real record — a device-code
// access to a cold array that forgets cold(g) reads poison (parity with the oracle breaks) or trips the poison check.
static thread_local rv_game_state* hs_staged = nullptr;
static thread_local rv_game_state* hs_home = nullptr;
#define RV_COLD_HOOK(g) (&(g) == hs_staged ? *hs_home : (g))

#include <atomic>
#include <string>
#include <unistd.h>
#include <cstdio>
#include <thread>
#include <vector>

#include "../../riichienv_b200/csrc/obs.cuh"
#include "../../riichienv_b200/csrc/obs_ext.cuh"
#include "../../riichienv_b200/csrc/obs_ext3.cuh"
#include <cmath>
#include "../../riichienv_b200/csrc/seq.cuh"
#include "../../riichienv_b200/csrc/validate.h"

using namespace rv;

static std::vector<uint32_t> g_suit_info, g_honor_info;
static std::vector<uint64_t> g_suit_cost, g_honor_cost;
static Tables g_T;
static bool g_ready = false;

template <int N, bool SEQ>
static void gen(std::vector<uint64_t>& cost, std::vector<uint32_t>& info, int n_keys, int threads) {
  cost.assign(n_keys, 0xFFFFFFFFFFull);
  info.assign(n_keys, 0);
  std::atomic<int> next{0};
  auto work = [&] {
    while (true) {
      int b = next.fetch_add(1024);
      if (b >= n_keys) break;
      for (int key = b; key < std::min(n_keys, b + 1024); key++) {
        uint8_t c[9];
        int k = key, sum = 0;
        for (int i = 0; i < N; i++) {
          c[i] = k % 5;
          k /= 5;
          sum += c[i];
        }
        if (sum > 14) continue;
        uint8_t best[5][2];
        suit_cost_dp<N, SEQ>(c, best);
        uint64_t packed = 0;
        for (int m = 0; m < 5; m++) {
          uint64_t a = best[m][0] > 15 ? 15 : best[m][0], b2 = best[m][1] > 15 ? 15 : best[m][1];
          packed |= a << (8 * m);
          packed |= b2 << (8 * m + 4);
        }
        cost[key] = packed;
        uint32_t e = 0;
        if (sum % 3 == 0 && best[sum / 3][0] == 0) e |= 1u;
        if (sum % 3 == 2 && best[sum / 3][1] == 0) e |= 2u;
        info[key] = e;
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 0; t < threads; t++) th.emplace_back(work);
  for (auto& t : th) t.join();
  std::vector<uint32_t> base = info;
  for (int key = 0; key < n_keys; key++) {
    uint8_t c[9];
    int k = key, sum = 0;
    for (int i = 0; i < N; i++) {
      c[i] = k % 5;
      k /= 5;
      sum += c[i];
    }
    if (sum > 13) continue;
    uint32_t e = base[key] & 3u;
    for (int i = 0; i < N; i++) {
      if (c[i] >= 4) continue;
      uint32_t o = base[key + pow5(i)] & 3u;
      if (o & 1u) e |= 1u << (2 + i);
      if (o & 2u) e |= 1u << (11 + i);
    }
    info[key] = e;
  }
  std::vector<uint32_t> base2 = info;
  for (int key = 0; key < n_keys; key++) {
    uint8_t c[9];
    int k = key, sum = 0;
    for (int i = 0; i < N; i++) {
      c[i] = k % 5;
      k /= 5;
      sum += c[i];
    }
    if (sum > 14 || sum == 0) continue;
    uint32_t e = base2[key];
    for (int i = 0; i < N; i++)
      if (c[i] > 0) e |= discard_bits(base2[key - pow5(i)]);
    info[key] = e;
  }
}

static uint64_t tables_checksum() {   // FNV-1a over the four tables: a stale or damaged cache file is regenerated
  uint64_t h = 0xcbf29ce484222325ull;
  auto eat = [&](const void* p, size_t n) {
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 0x100000001b3ull;
  };
  eat(g_suit_info.data(), 4 * g_suit_info.size());
  eat(g_honor_info.data(), 4 * g_honor_info.size());
  eat(g_suit_cost.data(), 8 * g_suit_cost.size());
  eat(g_honor_cost.data(), 8 * g_honor_cost.size());
  return h;
}
extern "C" {
int hs_init(const char* cache_path, int threads) {
  if (g_ready) return 0;
  bool loaded = false;
  if (cache_path) {
    FILE* f = fopen(cache_path, "rb");
    if (f) {
      g_suit_info.resize(SUIT_KEYS);
      g_honor_info.resize(HONOR_KEYS);
      g_suit_cost.resize(SUIT_KEYS);
      g_honor_cost.resize(HONOR_KEYS);
      size_t ok = fread(g_suit_info.data(), 4, SUIT_KEYS, f) + fread(g_honor_info.data(), 4, HONOR_KEYS, f) +
                  fread(g_suit_cost.data(), 8, SUIT_KEYS, f) + fread(g_honor_cost.data(), 8, HONOR_KEYS, f);
      uint64_t sum = 0;
      size_t got = fread(&sum, 8, 1, f);
      fclose(f);
      loaded = ok == (size_t)2 * SUIT_KEYS + 2 * HONOR_KEYS && got == 1 && sum == tables_checksum();
    }
  }
  if (!loaded) {
    gen<9, true>(g_suit_cost, g_suit_info, SUIT_KEYS, threads);
    gen<7, false>(g_honor_cost, g_honor_info, HONOR_KEYS, threads);
    if (cache_path) {
      std::string tmp = std::string(cache_path) + ".tmp" + std::to_string((long long)getpid());
      FILE* f = fopen(tmp.c_str(), "wb");   // written beside the cache and renamed into place: readers never see half a file
      if (f) {
        fwrite(g_suit_info.data(), 4, SUIT_KEYS, f);
        fwrite(g_honor_info.data(), 4, HONOR_KEYS, f);
        fwrite(g_suit_cost.data(), 8, SUIT_KEYS, f);
        fwrite(g_honor_cost.data(), 8, HONOR_KEYS, f);
        uint64_t sum = tables_checksum();
        fwrite(&sum, 8, 1, f);
        fclose(f);
        rename(tmp.c_str(), cache_path);
      }
    }
  }
  g_T.suit_info = g_suit_info.data();
  g_T.honor_info = g_honor_info.data();
  g_T.suit_cost = g_suit_cost.data();
  g_T.honor_cost = g_honor_cost.data();
  g_ready = true;
  return 0;
}

int hs_hand_eval(const rv_hand_query* q, rv_hand_result* out, int64_t n) {
  for (int64_t i = 0; i < n; i++) hand_eval_one(g_T, q[i], out[i]);
  return 0;
}
int hs_shanten_counts(const uint8_t* cnt34, int len_div3) {
  Cnt c;
  cnt_zero(c);
  for (int i = 0; i < 34; i++) cnt_add(c, i, cnt34[i]);
  return shanten_counts(g_T, c, len_div3);
}
int hs_shanten_counts_3p(const uint8_t* cnt34, int len_div3) {
  Cnt c;
  cnt_zero(c);
  for (int i = 0; i < 34; i++) cnt_add(c, i, cnt34[i]);
  return shanten_counts_3p(g_T, c, len_div3);
}
int hs_is_agari(const uint8_t* cnt34) {
  Cnt c;
  cnt_zero(c);
  for (int i = 0; i < 34; i++) cnt_add(c, i, cnt34[i]);
  return agari14(g_T, c) ? 1 : 0;
}
uint64_t hs_waits(const uint8_t* cnt34) {
  Cnt c;
  cnt_zero(c);
  for (int i = 0; i < 34; i++) cnt_add(c, i, cnt34[i]);
  return waits13(g_T, c);
}

// one game
struct HS { G g; std::vector<uint32_t> log; };
void* hs_game_new(int mode, uint64_t seed, uint32_t rule, uint32_t log_cap) {
  HS* h = new HS();
  memset(&h->g, 0, sizeof(G));
  h->g.game_mode = (uint8_t)mode;
  h->g.rule_bits = (uint8_t)rule;
  h->g.seed = seed;
  h->g.hand_index = 1;
  h->g.last_error = RV_NONE;
  h->g.pending_init[0] = h->g.pending_init[1] = h->g.pending_init[2] = RV_NONE;
  h->g.pending_tail[0] = RV_NONE;
  h->g.is_done = 1;
  for (int s = 0; s < MAXP; s++) h->g.score[s] = mode >= 3 ? (s < 3 ? 35000 : 0) : 25000;
  h->log.assign(log_cap, 0);
  return h;
}
static Ctx hs_ctx(HS* h) {
  Ctx cx;
  cx.T = g_T;
  cx.log = h->log.empty() ? nullptr : h->log.data();
  cx.log_cap = (uint32_t)h->log.size();
  cx.defer_init = false;
  cx.defer_tail = false;
  return cx;
}
void hs_game_free(void* p) { delete (HS*)p; }
void hs_game_reset(void* p, int oya, int rw, int honba, uint32_t kyotaku, const uint8_t* wall, const int32_t* scores) {
  HS* h = (HS*)p;
  Ctx cx = hs_ctx(h);
  game_reset(cx, h->g, oya, rw, honba, kyotaku, wall, scores);
}
int hs_game_legal(void* p, int pid, rv_action* out) {
  HS* h = (HS*)p;
  G& g = h->g;
  Ctx cx = hs_ctx(h);
  bool owes = !g.is_done && ((g.phase == RV_WAIT_ACT && g.current_player == pid) ||
                             (g.phase == RV_WAIT_RESPONSE && ((g.active_mask >> pid) & 1)));
  if (!owes) return 0;
  uint32_t packed[RV_MAX_LEGAL];
  int n = legal_actions(cx, g, pid, packed, -1, nullptr);
  if (n > RV_MAX_LEGAL) n = RV_MAX_LEGAL;
  for (int k = 0; k < n; k++) out[k] = expand_act(g, pid, packed[k]);
  return n;
}
void hs_game_step(void* p, const rv_action* in) {
  HS* h = (HS*)p;
  G& g = h->g;
  if (g.is_done) return;
  Ctx cx = hs_ctx(h);
  rv_action acts[MAXP];
  for (int s = 0; s < MAXP; s++) {
    acts[s] = in[s];
    int nc = acts[s].n_consume > 4 ? 4 : acts[s].n_consume;
    if (acts[s].type != RV_NO_ACTION) std::sort(acts[s].consume, acts[s].consume + nc);
  }
  g.step_count++;
  for (int s = 0; s < MAXP; s++) {
    if (acts[s].type == RV_NO_ACTION) continue;
    uint32_t packed[RV_MAX_LEGAL];
    int cnt = legal_actions(cx, g, s, packed, -1, nullptr);
    if (cnt > RV_MAX_LEGAL) cnt = RV_MAX_LEGAL;
    bool ok = false;
    for (int k = 0; k < cnt && !ok; k++) ok = action_matches(expand_act(g, s, packed[k]), acts[s]);
    if (!ok) {
      g.last_error = (uint8_t)s;
      trigger_ryukyoku(cx, g, RV_RK_ILLEGAL_BASE + s);
      return;
    }
  }
  step_apply(cx, g, acts);
}
struct Staged {
  alignas(16) unsigned char buf[sizeof(G) + 64];
  G* home;
  explicit Staged(G& real) : home(&real) {
    memset(buf, 0xAB, sizeof buf);
    memcpy(buf, &real, RV_HOT_BYTES);
    hs_staged = reinterpret_cast<G*>(buf);
    hs_home = home;
  }
  G& g() { return *reinterpret_cast<G*>(buf); }
  ~Staged() {
    for (size_t k = RV_HOT_BYTES; k < sizeof buf; k++)
      if (buf[k] != 0xAB) {
        fprintf(stderr, "hostsim: device code wrote a cold field through the staged copy (offset %zu)\n", k);
        abort();
      }
    memcpy(home, buf, RV_HOT_BYTES);
    hs_staged = hs_home = nullptr;
  }
};
void hs_game_random_step(void* p, uint64_t agent_seed, uint64_t game_id) {
  HS* h = (HS*)p;
  Ctx cx = hs_ctx(h);
  if (h->g.is_done) return;
  Staged st(h->g);
  random_step(cx, st.g(), agent_seed, game_id);
}
// one random step as the lock-step kernels take it: the round's deal is parked and then run by the warp-cooperative routine
void hs_game_random_step_coopdeal(void* p, uint64_t agent_seed, uint64_t game_id) {
  HS* h = (HS*)p;
  Ctx cx = hs_ctx(h);
  if (h->g.is_done) return;
  cx.defer_init = true;
  random_step(cx, h->g, agent_seed, game_id);
  if (h->g.pending_init[0] != RV_NONE) {
    DealScratch S;
    run_pending_init_coop(cx, h->g, S);
  }
}
void hs_game_apply_event(void* p, const rv_mjai_event* e) {
  HS* h = (HS*)p;
  Ctx cx = hs_ctx(h);
  cx.log = nullptr;
  cx.log_cap = 0;
  if (e->type != RV_EV_NONE) apply_mjai_event(cx, h->g, *e);
}
void hs_game_apply_log_action(void* p, const rv_log_action* a) {
  HS* h = (HS*)p;
  Ctx cx = hs_ctx(h);
  cx.log = nullptr;
  cx.log_cap = 0;
  if (a->type != RV_LA_NONE) apply_log_action(cx, h->g, *a);
}
void hs_game_replay_begin(void* p, const rv_log_kyoku* k) {
  HS* h = (HS*)p;
  Ctx cx = hs_ctx(h);
  const int np = h->g.game_mode >= 3 ? 3 : 4;
  int32_t sc[4] = {k->scores[0], k->scores[1], k->scores[2], k->scores[3]};
  game_reset(cx, h->g, k->oya < np ? k->oya : 0, k->chang < 4 ? k->chang : 0, k->ben, k->liqibang, nullptr, sc);
  cx.log = nullptr;
  cx.log_cap = 0;
  replay_begin_patch(cx, h->g, *k);
}
void hs_game_agent_step(void* p, int policy, uint64_t agent_seed, uint64_t game_id) {
  HS* h = (HS*)p;
  Ctx cx = hs_ctx(h);
  if (h->g.is_done) return;
  agent_step(cx, h->g, policy, agent_seed, game_id);
}
// One scheduler visit as the rollout kernels make it: a parked discard tail or a parked deal is run on its own visit,
// otherwise the game takes one random step with both deferrals armed.  Returns 1 if an env step was taken.
int hs_game_random_step_deferred(void* p, uint64_t agent_seed, uint64_t game_id) {
  HS* h = (HS*)p;
  Ctx cx = hs_ctx(h);
  cx.defer_init = true;
  cx.defer_tail = true;
  Staged st(h->g);
  G& g = st.g();
  if (g.pending_tail[0] != RV_NONE) {
    run_pending_tail(cx, g);
    return 0;
  }
  if (g.pending_init[0] != RV_NONE) {
    run_pending_init(cx, g);
    return 0;
  }
  if (g.is_done) return 0;
  random_step(cx, g, agent_seed, game_id);
  return 1;
}
void hs_game_snapshot(void* p, rv_game_state* out) { *out = ((HS*)p)->g; }
int hs_game_load_snapshot(void* p, const rv_game_state* in) {   // as rv_vec_set_state: -1 for a record that is not a position
  if (rv_state_defect(*in)) return -1;
  ((HS*)p)->g = *in;
  refresh_caches(g_T, ((HS*)p)->g);
  return 0;
}
const char* hs_state_defect(const rv_game_state* in) { return rv_state_defect(*in); }
int hs_game_call(void* p, int op, uint8_t* out) {   // env.rs:624-631 test hooks, as orc_game_call
  HS* h = (HS*)p;
  Ctx cx = hs_ctx(h);
  if (op == 0) {
    reveal_kan_dora(cx, h->g);
    return h->g.n_dora;
  }
  if (op == 1) return ura_indicators(h->g, out);
  if (op == 2) {
    trigger_ryukyoku(cx, h->g, RV_RK_EXHAUSTIVE);
    return h->g.is_done;
  }
  if (op == 6) {
    if (h->g.last_discard_pid != RV_NONE) claims_after_tile(cx, h->g, h->g.last_discard_pid, h->g.last_discard_tile, false);
    return __builtin_popcount(h->g.active_mask);
  }
  if (op >= 3 && op <= 5) {
    next_round(cx, h->g, op == 4, op == 5);
    return h->g.is_done;
  }
  return -1;
}
void hs_game_copy_log(void* dst, void* src) { ((HS*)dst)->log = ((HS*)src)->log; }
uint32_t hs_game_events(void* p, uint32_t* out, uint32_t cap) {
  HS* h = (HS*)p;
  uint32_t n = std::min<uint32_t>(h->g.ev_words, (uint32_t)h->log.size());
  if (out) memcpy(out, h->log.data(), 4 * std::min(n, cap));
  return h->g.ev_words;
}
void hs_game_encode(void* p, int pid, float* obs, uint8_t* mask) {
  HS* h = (HS*)p;
  const G& g = h->g;
  bool sanma = is_sanma(g);
  if (obs) {
    int seen[34];
    for (int k = 0; k < 34; k++) seen[k] = obs_seen(g, pid, k);
    for (int ch = 0; ch < OBS_CH; ch++) {
      int kind;
      uint64_t m;
      float v;
      if (sanma) {
        obs_channel<true>(g, pid, ch, kind, m, v);
        for (int col = 0; col < OBS_W3; col++) {
          int k34 = obs_col_kind3(col);
          obs[ch * OBS_W3 + col] = obs_value(kind, m, v, seen[k34], k34);
        }
      } else {
        obs_channel<false>(g, pid, ch, kind, m, v);
        for (int col = 0; col < OBS_W; col++) obs[ch * OBS_W + col] = obs_value(kind, m, v, seen[col], col);
      }
    }
  }
  if (mask) {
    memset(mask, 0, sanma ? OBS_IDS3 : OBS_IDS);
    Ctx cx = hs_ctx(h);
    uint32_t packed[RV_MAX_LEGAL];
    int cnt = legal_actions(cx, g, pid, packed, -1, nullptr);
    if (cnt > RV_MAX_LEGAL) cnt = RV_MAX_LEGAL;
    for (int k = 0; k < cnt; k++) {
      rv_action a = expand_act(g, pid, packed[k]);
      int id = sanma ? action_id_3p(a) : action_id(a);
      if (id >= 0 && id < (sanma ? OBS_IDS3 : OBS_IDS)) mask[id] = 1;
    }
  }
}
// Observation::encode_extended through the scalar definitions of obs_ext.cuh (4P): 215 x 34 floats
// sanma rows (Observation3P::encode_extended, 215 x 27) through the scalar definitions of obs_ext3.cuh
static void hs_encode_ext3(HS* h, int pid, float* obs) {
  const G& g = h->g;
  int seen[34], vis[34], called = 0;
  for (int k = 0; k < 34; k++) {
    seen[k] = obs_seen(g, pid, k);
    vis[k] = seen[k] - (int)((g.c_cnt[pid][k / 9] >> (4 * (k % 9))) & 15);
  }
  for (int q = 0; q < 3; q++)
    for (int m = 0; m < g.n_melds[q]; m++) called += g.meld_called[q][m] != RV_NONE;
  for (int ch = 0; ch < OBS_CH; ch++) {
    int kind;
    uint64_t m;
    float v;
    obs_channel<true>(g, pid, ch, kind, m, v);
    for (int col = 0; col < OBS_W3; col++) {
      int k34 = obs_col_kind3(col);
      obs[ch * OBS_W3 + col] = obs_value(kind, m, v, seen[k34], k34);
    }
  }
  {
    int used = 0;
    for (int k = 0; k < 34; k++) used += seen[k];
    int left = 108 - used + called;
    for (int col = 0; col < OBS_W3; col++) obs[30 * OBS_W3 + col] = (float)(left < 0 ? 0 : left) / 70.0f;
  }
  DecayTab D;
  for (int age = 0; age < RV_RIVER_CAP; age++) D.w[age] = expf(-0.2f * (float)age);
  for (int r = 0; r < 4; r++) {
    float* row = obs + (74 + r) * OBS_W3;
    for (int col = 0; col < OBS_W3; col++) row[col] = 0.0f;
    if (r < 3) obs_ext3_decay_row(g, &g.river[0][0], (pid + r) % 3, D, row);
  }
  ObsExtInfo I;
  Ctx cx = hs_ctx(h);
  obs_ext3_shanten_scalar(g_T, g, pid, vis, I);
  I.avail = 0;
  I.dora_kinds = obs_ext3_dora_kinds(g);
  if (!g.is_done && ((g.active_mask >> pid) & 1)) {
    uint32_t packed[RV_MAX_LEGAL];
    int cnt = legal_actions(cx, g, pid, packed, -1, nullptr);
    if (cnt > RV_MAX_LEGAL) cnt = RV_MAX_LEGAL;
    for (int k = 0; k < cnt; k++) I.avail |= obs_avail_bit(expand_act(g, pid, packed[k]));
  }
  for (int ch = 78; ch < OBSX_CH; ch++) {
    uint64_t m;
    float v;
    obs_ext3_channel(g, g, pid, ch, I, m, v);
    for (int col = 0; col < OBS_W3; col++) obs[ch * OBS_W3 + col] = ((m >> col) & 1) ? v : 0.0f;
  }
}
void hs_game_encode_ext(void* p, int pid, float* obs) {
  HS* h = (HS*)p;
  if (is_sanma(h->g)) {
    hs_encode_ext3(h, pid, obs);
    return;
  }
  const G& g = h->g;
  int seen[34], vis[34], called = 0;
  for (int k = 0; k < 34; k++) {
    seen[k] = obs_seen(g, pid, k);
    vis[k] = seen[k] - (int)((g.c_cnt[pid][k / 9] >> (4 * (k % 9))) & 15);
  }
  for (int q = 0; q < 4; q++)
    for (int m = 0; m < g.n_melds[q]; m++) called += g.meld_called[q][m] != RV_NONE;
  for (int ch = 0; ch < OBS_CH; ch++) {
    int kind;
    uint64_t m;
    float v;
    obs_channel<false>(g, pid, ch, kind, m, v);
    for (int col = 0; col < OBS_W; col++) obs[ch * OBS_W + col] = obs_value(kind, m, v, seen[col], col);
  }
  {
    int used = 0;
    for (int k = 0; k < 34; k++) used += seen[k];
    int left = 136 - used + called;
    for (int col = 0; col < OBS_W; col++) obs[30 * OBS_W + col] = (float)(left < 0 ? 0 : left) / 70.0f;
  }
  DecayTab D;
  for (int age = 0; age < RV_RIVER_CAP; age++) D.w[age] = expf(-0.2f * (float)age);
  for (int r = 0; r < 4; r++) {
    float* row = obs + (74 + r) * OBS_W;
    for (int col = 0; col < OBS_W; col++) row[col] = 0.0f;
    obs_ext_decay_row(g, &g.river[0][0], (pid + r) & 3, D, row);
  }
  ObsExtInfo I;
  Ctx cx = hs_ctx(h);
  obs_ext_shanten_scalar(g_T, g, pid, vis, I);
  {
    // the incremental evaluation the warp kernel uses (sh_* in obs_ext.cuh) must give the same five numbers
    ObsExtInfo F;
    obs_ext_shanten_fast_scalar(g_T, g, pid, vis, F);
    if (F.shanten != I.shanten || F.eff != I.eff || F.uke != I.uke || F.keep != I.keep || F.inc != I.inc) I.shanten = -77;
  }
  I.avail = 0;
  I.dora_kinds = obs_ext_dora_kinds(g);
  if (!g.is_done && ((g.active_mask >> pid) & 1)) {
    uint32_t packed[RV_MAX_LEGAL];
    int cnt = legal_actions(cx, g, pid, packed, -1, nullptr);
    if (cnt > RV_MAX_LEGAL) cnt = RV_MAX_LEGAL;
    for (int k = 0; k < cnt; k++) I.avail |= obs_avail_bit(expand_act(g, pid, packed[k]));
  }
  for (int ch = 78; ch < OBSX_CH; ch++) {
    uint64_t m;
    float v;
    obs_ext_channel(g, g, pid, ch, I, m, v);
    for (int col = 0; col < OBS_W; col++) obs[ch * OBS_W + col] = ((m >> col) & 1) ? v : 0.0f;
  }
}
void hs_game_encode_kawa(void* p, float* out) {
  const G& g = ((HS*)p)->g;
  if (is_sanma(g)) {
    for (int i = 0; i < KAWA_FLOATS3; i++) out[i] = 0.0f;
    for (int q = 0; q < 3; q++) obs_kawa_seat<true>(g, &g.river[0][0], q, out + q * 7 * OBS_W3);
    return;
  }
  for (int i = 0; i < KAWA_FLOATS; i++) out[i] = 0.0f;
  for (int q = 0; q < 4; q++) obs_kawa_seat<false>(g, &g.river[0][0], q, out + q * 7 * OBS_W);
}
void hs_game_encode_seq(void* p, int pid, uint32_t w0, uint32_t w1, int game_style, uint16_t* sparse, float* numeric, uint16_t* prog,
                        int max_prog, uint16_t* cand, uint16_t* lens) {
  HS* h = (HS*)p;
  Ctx cx = hs_ctx(h);
  SeqOut o{sparse, numeric, prog, cand, lens, max_prog};
  seq_encode(cx, h->g, pid, h->log.data(), w0, w1, game_style, o);
}
int hs_wall_from_seed(uint64_t seed, uint64_t hand_index, int n, uint8_t* out) {
  wall_from_seed(seed, hand_index, n, out);
  return n;
}
}
