// ORACLE — TEST INFRASTRUCTURE ONLY (see hand.hpp header).
// extern "C" surface used by tests/ (ctypes), __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs.  Speaks the same structs as
// include/riichienv_b200.h so results can be compared byte for byte.
#include <atomic>
#include <mutex>
#include <thread>

#include "../include/rv_synth.h"
#include "game.hpp"
#include "obs.hpp"
#include "seq.hpp"
#include "shanten.hpp"

using namespace orc;

static void to_rv_action(const Action& a, rv_action& o) {
  o.type = a.type;
  o.tile = (uint8_t)(a.tile < 0 ? 0xFF : a.tile);
  o.n_consume = (uint8_t)std::min<size_t>(a.consume.size(), 4);
  for (int k = 0; k < 4; k++) o.consume[k] = k < o.n_consume ? a.consume[k] : 0xFF;
  o.actor = (uint8_t)(a.actor < 0 ? 0xFF : a.actor);
}
static Action from_rv_action(const rv_action& a) {
  std::vector<uint8_t> c;
  for (int k = 0; k < a.n_consume && k < 4; k++) c.push_back(a.consume[k]);
  return Action(a.type, a.tile == 0xFF ? -1 : a.tile, c, a.actor == 0xFF ? -1 : a.actor);
}

static void load_snapshot(GameState& g, const rv_game_state& s) {
  g.game_mode = s.game_mode;
  g.sanma = s.game_mode >= 3;
  g.np = g.sanma ? 3 : 4;
  const int NP = g.np;
  if (g.sanma && s.wall_len == 108)
    for (int i = 0; i < 5; i++) {
      g.dora_tiles3[i] = s.wall[8 + 2 * i];
      g.ura_tiles3[i] = s.wall[9 + 2 * i];
    }
  for (int p = 0; p < 4; p++) g.n_kita[p] = s.n_kita[p];
  g.wall_abs.assign(s.wall, s.wall + s.wall_len);
  g.wall_tiles.assign(s.wall + s.rinshan_draw_count, s.wall + s.wall_top);
  g.rinshan_draw_count = s.rinshan_draw_count;
  g.pending_kan_dora_count = s.pending_kan_dora_count;
  g.drawable_count = s.drawable_count;
  g.dora_indicators.assign(s.dora_ind, s.dora_ind + s.n_dora);
  g.phase = s.phase;
  for (int p = 0; p < NP; p++) {
    PlayerState& P = g.players[p];
    P.hand.assign(s.hand[p], s.hand[p] + s.hand_len[p]);
    P.melds.clear();
    for (int m = 0; m < s.n_melds[p]; m++) {
      Meld M;
      M.meld_type = (MeldType)s.meld_type[p][m];
      for (int k = 0; k < 4; k++)
        if (s.meld_tiles[p][m][k] != 0xFF) M.tiles.push_back(s.meld_tiles[p][m][k]);
      M.opened = M.meld_type != Ankan;
      M.from_who = (int8_t)s.meld_from[p][m];
      M.called_tile = s.meld_called[p][m] == 0xFF ? -1 : s.meld_called[p][m];
      P.melds.push_back(M);
    }
    P.discards.assign(s.river[p], s.river[p] + std::min<int>(s.n_river[p], RV_RIVER_CAP));
    P.discard_from_hand.clear();
    P.discard_is_riichi.clear();
    for (size_t k = 0; k < P.discards.size(); k++) {
      P.discard_from_hand.push_back((s.river_tedashi[p] >> k) & 1);
      P.discard_is_riichi.push_back((s.river_riichi[p] >> k) & 1);
    }
    P.riichi_declaration_index = s.riichi_decl_idx[p] == 0xFF ? -1 : s.riichi_decl_idx[p];
    uint8_t f = s.flags[p];
    P.riichi_declared = f & RV_F_RIICHI_DECLARED;
    P.riichi_stage = f & RV_F_RIICHI_STAGE;
    P.double_riichi_declared = f & RV_F_DOUBLE_RIICHI;
    P.missed_agari_riichi = f & RV_F_MISSED_AGARI_RIICHI;
    P.missed_agari_doujun = f & RV_F_MISSED_AGARI_DOUJUN;
    P.nagashi_eligible = f & RV_F_NAGASHI_ELIGIBLE;
    P.ippatsu_cycle = f & RV_F_IPPATSU_CYCLE;
    P.pao37 = s.pao[p][0] == 0xFF ? -1 : s.pao[p][0];
    P.pao50 = s.pao[p][1] == 0xFF ? -1 : s.pao[p][1];
    P.forbidden_discards.clear();
    for (int k = 0; k < 2; k++)
      if (s.forbidden[p][k] != 0xFF) P.forbidden_discards.push_back(s.forbidden[p][k]);
    g.riichi_sutehais[p] = s.riichi_sutehai[p] == 0xFF ? -1 : s.riichi_sutehai[p];
    g.last_tedashis[p] = s.last_tedashi[p] == 0xFF ? -1 : s.last_tedashi[p];
    P.score = s.score[p];
    P.score_delta = s.score_delta[p];
    g.current_claims[p].clear();
    g.has_claims_entry[p] = s.n_claims[p] > 0;
    for (int k = 0; k < s.n_claims[p]; k++) {
      uint32_t c = s.claims[p][k];
      uint8_t type = c & 0xFF, tile = (c >> 8) & 0xFF, c0 = (c >> 16) & 0xFF, c1 = (c >> 24) & 0xFF;
      std::vector<uint8_t> cons;
      if (c0 != 0xFF) cons.push_back(c0);
      if (c1 != 0xFF) cons.push_back(c1);
      if (type == RV_DAIMINKAN) {  // third tile: next matching tile in hand order
        int n = 0;
        cons.clear();
        for (uint8_t t : P.hand)
          if (t / 4 == tile / 4 && n < 3) {
            cons.push_back(t);
            n++;
          }
      }
      g.current_claims[p].emplace_back(type, tile == 0xFF ? -1 : tile, cons, p);
    }
  }
  g.current_player = s.current_player;
  g.oya = s.oya;
  g.honba = s.honba;
  g.kyoku_idx = s.kyoku_idx;
  g.round_wind = s.round_wind;
  g.is_done = s.is_done;
  g.needs_tsumo = s.needs_tsumo;
  g.is_first_turn = s.is_first_turn;
  g.is_rinshan_flag = s.is_rinshan_flag;
  g.is_after_kan = s.is_after_kan;
  g.riichi_pending_acceptance = s.riichi_pending_acceptance == 0xFF ? -1 : s.riichi_pending_acceptance;
  g.drawn_tile = s.drawn_tile == 0xFF ? -1 : s.drawn_tile;
  g.last_discard_pid = s.last_discard_pid == 0xFF ? -1 : s.last_discard_pid;
  g.last_discard_tile = s.last_discard_pid == 0xFF ? -1 : s.last_discard_tile;
  g.pending_kan = s.pending_kan_pid != 0xFF;
  if (g.pending_kan) {
    g.pending_kan_pid = s.pending_kan_pid;
    std::vector<uint8_t> cons;
    uint8_t t = s.pending_kan_tile;
    if (s.pending_kan_type == RV_ANKAN) {
      uint8_t lo = t / 4 * 4;
      cons = {lo, (uint8_t)(lo + 1), (uint8_t)(lo + 2), (uint8_t)(lo + 3)};
    }
    g.pending_kan_act = Action(s.pending_kan_type, t, cons, s.pending_kan_pid);
  }
  g.active_players.clear();
  for (int p = 0; p < NP; p++)
    if (s.active_mask & (1u << p)) g.active_players.push_back((uint8_t)p);
  g.last_error = s.last_error == 0xFF ? -1 : s.last_error;
  g.stalled = (s.overflow & 2) != 0;
  g.game_mode = s.game_mode;
  g.rule = s.rule_bits;
  g.riichi_sticks = s.riichi_sticks;
  g.turn_count = s.turn_count;
  g.wall_seed = s.seed;
  g.hand_index = s.hand_index;
  g.step_count = s.step_count;
  g.kyoku_count = s.kyoku_count;
  g.ev_count = s.ev_count;
  g.ev_words = s.ev_words;
  g.ev_hash = s.ev_hash;
}

static void eval_one(const rv_hand_query& q, rv_hand_result& r) {
  memset(&r, 0, sizeof r);
  std::vector<uint8_t> tiles(q.tiles, q.tiles + q.n_tiles);
  std::vector<Meld> melds;
  for (int m = 0; m < q.n_melds; m++) {
    Meld M;
    M.meld_type = (MeldType)q.meld_type[m];
    for (int k = 0; k < 4; k++)
      if (q.meld_tiles[m][k] != 0xFF) M.tiles.push_back(q.meld_tiles[m][k]);
    M.opened = M.meld_type != Ankan;
    melds.push_back(M);
  }
  Conditions c;
  c.tsumo = q.cond & RV_C_TSUMO;
  c.riichi = q.cond & RV_C_RIICHI;
  c.double_riichi = q.cond & RV_C_DOUBLE_RIICHI;
  c.ippatsu = q.cond & RV_C_IPPATSU;
  c.haitei = q.cond & RV_C_HAITEI;
  c.houtei = q.cond & RV_C_HOUTEI;
  c.rinshan = q.cond & RV_C_RINSHAN;
  c.chankan = q.cond & RV_C_CHANKAN;
  c.tsumo_first_turn = q.cond & RV_C_TSUMO_FIRST_TURN;
  c.player_wind = q.player_wind;
  c.round_wind = q.round_wind;
  c.honba = q.honba;
  bool sanma = q.sanma & 1;
  c.kita_count = q.kita_count;
  c.is_sanma = sanma;
  c.num_players = sanma ? 3 : 4;
  HandEvaluator he(tiles, melds, sanma);
  std::vector<uint8_t> dora(q.dora_ind, q.dora_ind + q.n_dora), ura(q.ura_ind, q.ura_ind + q.n_ura);
  WinResult w = he.calc(q.win_tile, dora, ura, c);
  r.is_win = w.is_win;
  r.yakuman = w.yakuman;
  r.has_win_shape = w.has_win_shape;
  r.han = (uint8_t)w.han;
  r.fu = (uint8_t)w.fu;
  r.ron_agari = w.ron_agari;
  r.tsumo_agari_oya = w.tsumo_agari_oya;
  r.tsumo_agari_ko = w.tsumo_agari_ko;
  r.n_yaku = (uint8_t)w.yaku.size();
  for (uint32_t y : w.yaku) r.yaku_mask |= 1ull << y;
  // waits of the 3n+1 hand
  std::vector<uint8_t> t13 = tiles;
  int total = he.current_total();
  bool ok13 = total == 13;
  if (total == 14) {
    for (int i = (int)t13.size() - 1; i >= 0; i--)
      if (t13[i] / 4 == q.win_tile / 4) {
        t13.erase(t13.begin() + i);
        ok13 = true;
        break;
      }
  }
  if (ok13) {
    HandEvaluator h13(t13, melds, sanma);
    for (uint8_t x : h13.get_waits_u8()) r.wait_mask |= 1ull << x;
  }
  // shanten (shanten.rs:250-261) over the concealed tiles (+ win tile when 3n+1)
  {
    uint8_t cnt[34] = {0};
    int n = 0;
    for (uint8_t t : tiles) {
      cnt[t / 4]++;
      n++;
    }
    if (total == 13) {
      cnt[q.win_tile / 4]++;
      n++;
    }
    // sanma queries: calculate_shanten_3p (shanten.rs:470-484)
    r.shanten = (int8_t)(sanma ? shanten_from_counts_3p(cnt, n / 3) : shanten_from_counts(cnt, n / 3));
    uint8_t c13[34] = {0};
    int n13 = 0;
    for (uint8_t t : t13) {
      c13[t / 4]++;
      n13++;
    }
    r.shanten13 = ok13 ? (int8_t)(sanma ? shanten_from_counts_3p(c13, n13 / 3) : shanten_from_counts(c13, n13 / 3)) : (int8_t)127;
  }
}

extern "C" {

int orc_hand_eval(const rv_hand_query* q, rv_hand_result* out, int64_t n) {
  for (int64_t i = 0; i < n; i++) eval_one(q[i], out[i]);
  return 0;
}
int orc_hand_eval_mt(const rv_hand_query* q, rv_hand_result* out, int64_t n, int threads) {
  if (threads <= 1) return orc_hand_eval(q, out, n);
  std::vector<std::thread> th;
  std::atomic<int64_t> next{0};
  for (int t = 0; t < threads; t++)
    th.emplace_back([&] {
      while (true) {
        int64_t b = next.fetch_add(4096);
        if (b >= n) break;
        int64_t e = std::min(n, b + 4096);
        for (int64_t i = b; i < e; i++) eval_one(q[i], out[i]);
      }
    });
  for (auto& t : th) t.join();
  return 0;
}
// the seeded synthetic hand stream of BASELINE.json configs[1] (input data; one definition in include/rv_synth.h)
int orc_hand_queries_seeded(rv_hand_query* q, uint64_t first, int64_t n) {
  for (int64_t i = 0; i < n; i++) rv_synth_hand(first + (uint64_t)i, &q[i]);
  return 0;
}
int orc_is_agari(const uint8_t* counts34) {
  Hand h;
  for (int i = 0; i < 34; i++) h.counts[i] = counts34[i];
  return is_agari(h) ? 1 : 0;
}
int orc_is_tenpai_counts(const uint8_t* counts34) {  // agari.rs:15-61 semantics (13-tile histogram)
  Hand h;
  for (int i = 0; i < 34; i++) h.counts[i] = counts34[i];
  for (int i = 0; i < 34; i++)
    if (h.counts[i] < 4) {
      h.add(i);
      bool a = is_agari(h);
      h.remove(i);
      if (a) return 1;
    }
  return 0;
}
int orc_shanten_counts(const uint8_t* counts34, int len_div3) { return shanten_from_counts(counts34, len_div3); }
int orc_shanten_counts_3p(const uint8_t* counts34, int len_div3) { return shanten_from_counts_3p(counts34, len_div3); }
int orc_calculate_score(int han, int fu, int is_oya, int is_tsumo, uint32_t honba, int np, uint32_t out[4]) {
  Score s = calculate_score((uint8_t)han, (uint8_t)fu, is_oya, is_tsumo, honba, (uint8_t)np);
  out[0] = s.pay_ron;
  out[1] = s.pay_tsumo_oya;
  out[2] = s.pay_tsumo_ko;
  out[3] = s.total;
  return 0;
}
// known-answer hook: n output words of ChaCha(rounds) keyed by 8 raw words, counter 0
void orc_chacha_words(const uint32_t* key, int rounds, int n, uint32_t* out) {
  ChaCha12 rng(key, rounds);
  for (int i = 0; i < n; i++) out[i] = rng.next_u32();
}
int orc_wall_from_seed(uint64_t seed, uint64_t hand_index, int n_tiles, uint8_t* out) {
  auto w = wall_from_seed(seed, hand_index, n_tiles);
  memcpy(out, w.data(), w.size());
  return (int)w.size();
}

void* orc_game_new(int mode, uint64_t seed, int round_wind, uint32_t rule, int keep_log) {
  return new GameState((uint8_t)mode, seed, (uint8_t)round_wind, rule, keep_log != 0);
}
void orc_game_free(void* h) { delete (GameState*)h; }
void orc_game_reset(void* h, int oya, int rw, int honba, uint32_t kyotaku, const uint8_t* wall, const int32_t* scores) {
  GameState* g = (GameState*)h;
  std::vector<uint8_t> w;
  if (wall) w.assign(wall, wall + (g->sanma ? 108 : 136));
  g->reset((uint8_t)oya, (uint8_t)rw, (uint8_t)honba, kyotaku, wall ? &w : nullptr, scores);
}
int orc_game_legal(void* h, int pid, rv_action* out) {
  GameState* g = (GameState*)h;
  bool owes = !g->is_done && ((g->phase == RV_WAIT_ACT && g->current_player == pid) ||
                              (g->phase == RV_WAIT_RESPONSE &&
                               std::find(g->active_players.begin(), g->active_players.end(), (uint8_t)pid) != g->active_players.end()));
  if (!owes) return 0;  // state/mod.rs:200-208
  auto l = g->_get_legal_actions_internal(pid);
  int n = (int)std::min<size_t>(l.size(), RV_MAX_LEGAL);
  for (int i = 0; i < n; i++) to_rv_action(l[i], out[i]);
  return n;
}
// batch helpers for the lock-step parity gate (tests/test_gpu_parity.py): legal lists of every seat of n games as
// [n][4][RV_MAX_LEGAL] rv_action + counts [n][4]; one keyed random step of every game (game id = seed_base + i)
void orc_games_legal_batch(void** hs, int64_t n, rv_action* out, uint8_t* counts) {
  for (int64_t i = 0; i < n; i++)
    for (int p = 0; p < 4; p++) counts[i * 4 + p] = (uint8_t)orc_game_legal(hs[i], p, out + ((size_t)i * 4 + p) * RV_MAX_LEGAL);
}
void orc_game_step(void* h, const rv_action* acts) {
  GameState* g = (GameState*)h;
  std::optional<Action> a[MAXP];
  for (int p = 0; p < g->np; p++)
    if (acts[p].type != RV_NO_ACTION) a[p] = from_rv_action(acts[p]);
  g->step(a);
}
int orc_game_random_step(void* h, uint64_t agent_seed, uint64_t game_id) {
  return random_step(*(GameState*)h, agent_seed, game_id) ? 1 : 0;
}
void orc_games_random_step_batch(void** hs, int64_t n, uint64_t agent_seed, uint64_t seed_base) {
  for (int64_t i = 0; i < n; i++) orc_game_random_step(hs[i], agent_seed, seed_base + (uint64_t)i);
}
void orc_game_snapshot(void* h, rv_game_state* out) { ((GameState*)h)->to_snapshot(*out); }
void orc_game_load_snapshot(void* h, const rv_game_state* in) { load_snapshot(*(GameState*)h, *in); }
// test hooks of the PyO3 class (env.rs:624-631): op 0 = _reveal_kan_dora() -> number of indicators;
// op 1 = _get_ura_markers() -> tile ids into out[5], returns how many
int orc_game_call(void* h, int op, uint8_t* out) {
  GameState* g = (GameState*)h;
  if (op == 0) {
    g->_reveal_kan_dora();
    return (int)g->dora_indicators.size();
  }
  if (op == 1) {
    auto u = g->_get_ura_indicators();
    for (size_t i = 0; i < u.size() && i < 5; i++) out[i] = u[i];
    return (int)std::min<size_t>(u.size(), 5);
  }
  if (op == 2) {        // tests.rs:172-262
    g->_trigger_ryukyoku(RV_RK_EXHAUSTIVE);
    return g->is_done ? 1 : 0;
  }
  if (op == 6) {        // replay: claim lists of every seat against the last discard
    claims_for_last_discard(*g);
    return (int)g->active_players.size();
  }
  if (op >= 3 && op <= 5) {   // tests.rs:375-428
    g->_initialize_next_round(op == 4, op == 5);
    return g->is_done ? 1 : 0;
  }
  return -1;
}
void orc_game_copy_log(void* dst, void* src) {
  ((GameState*)dst)->log = ((GameState*)src)->log;
  ((GameState*)dst)->text = ((GameState*)src)->text;
}
// the oracle's own MJAI text log (json.hpp): viewer -1 = mjai_log, 0..3 = mjai_log_per_player[viewer]; lines joined by '\n'.
// Returns the byte length (without terminator); copies at most cap - 1 bytes.
uint32_t orc_game_mjai_log(void* h, int viewer, char* out, uint32_t cap) {
  GameState* g = (GameState*)h;
  const std::vector<std::string>& v = viewer < 0 ? g->text.all : g->text.seat[viewer & 3];
  std::string s;
  for (size_t i = 0; i < v.size(); i++) s += (i ? "\n" : "") + v[i];
  if (out && cap) {
    size_t n = std::min<size_t>(s.size(), cap - 1);
    memcpy(out, s.data(), n);
    out[n] = 0;
  }
  return (uint32_t)s.size();
}
uint32_t orc_game_events(void* h, uint32_t* out, uint32_t cap) {
  GameState* g = (GameState*)h;
  uint32_t n = (uint32_t)g->log.size();
  if (out) memcpy(out, g->log.data(), sizeof(uint32_t) * std::min(n, cap));
  return n;
}

// Run n seeded games (game g: seed = seed_base + g, RiichiEnv(seed).reset() then the
// keyed agent) for at most max_steps env steps each.  All outputs optional.
// policy: 0 = uniform random agent, 1 = greedy-win agent (game.hpp greedy_pick).
// hist (optional, 128 x u64, summed over all games from their event logs): [y] for y < 64 = hora events whose yaku set
// holds id y; [64] hora events, [65] of them tsumo, [66] ron, [67] discards / kans that were ronned by two or more seats,
// [68] hora settlements with a pao payer, [69] rounds dealt, [70] ryukyoku events, [71 + reason] ryukyoku by reason code
// (0..7), [80] yakuman hora, [81] kazoe (>= 13 han, no yakuman), [82] games finished.
static int64_t run_agent(int policy, int mode, uint32_t rule, uint64_t seed_base, int64_t n, uint64_t agent_seed, uint32_t max_steps,
                         int threads, int32_t* scores, uint8_t* ranks, uint8_t* done, uint32_t* steps, uint32_t* kyoku,
                         uint32_t* evcount, uint64_t* hash, uint32_t* max_river, uint64_t* hist, const uint8_t* walls = nullptr) {
  std::atomic<int64_t> next{0};
  std::atomic<int64_t> total{0};
  std::mutex hist_mu;
  auto work = [&] {
    uint64_t local[128] = {0};
    while (true) {
      int64_t g = next.fetch_add(1);
      if (g >= n) break;
      GameState gs((uint8_t)mode, seed_base + (uint64_t)g, 0, rule, hist != nullptr);
      if (walls) {              // reset(wall=...) -> load_wall (state/wall.rs:69-80): the first round is dealt from the caller's tiles
        const int wl = mode >= 3 ? 108 : 136;
        std::vector<uint8_t> w(walls + (size_t)g * wl, walls + (size_t)(g + 1) * wl);
        gs.reset(0, 0, 0, 0, &w, nullptr);
      } else {
        gs.reset();
      }
      uint32_t mr = 0;
      while (!gs.is_done && gs.step_count < max_steps) {
        agent_step(gs, policy, agent_seed, seed_base + (uint64_t)g);
        if (max_river)
          for (auto& p : gs.players) mr = std::max<uint32_t>(mr, (uint32_t)p.discards.size());
      }
      if (hist) {
        const std::vector<uint32_t>& w = gs.log;
        int run = 0;   // consecutive hora events (a multi-ron settles in one step)
        for (size_t i = 0; i < w.size();) {
          int type = w[i] & 0xFF, nw = std::max<int>(1, (w[i] >> 8) & 0xFF);
          if (type == RV_EV_HORA) {
            uint64_t mask = (uint64_t)w[i + 4 + gs.np] | ((uint64_t)w[i + 5 + gs.np] << 32);
            for (int y = 0; y < 64; y++)
              if ((mask >> y) & 1) local[y]++;
            local[64]++;
            local[(w[i + 1] & 0xFF) ? 65 : 66]++;
            int han = (w[i + 1] >> 16) & 0xFF;
            bool yakuman = (w[i + 3] >> 8) & 0xFF;
            if (yakuman) local[80]++;
            else if (han >= 13) local[81]++;
            if (++run == 2) local[67]++;
          } else {
            run = 0;
            if (type == RV_EV_START_KYOKU) local[69]++;
            if (type == RV_EV_RYUKYOKU) {
              local[70]++;
              int reason = (w[i] >> 16) & 0xFF;
              local[71 + std::min(reason, 8)]++;
            }
          }
          i += nw;
        }
        local[68] += gs.stat_pao;
        local[82] += gs.is_done ? 1 : 0;
      }
      total += gs.step_count;
      if (scores)
        for (int i = 0; i < gs.np; i++) scores[g * MAXP + i] = gs.players[i].score;
      if (ranks) gs.ranks(ranks + g * MAXP);
      if (done) done[g] = gs.is_done;
      if (steps) steps[g] = gs.step_count;
      if (kyoku) kyoku[g] = gs.kyoku_count;
      if (evcount) evcount[g] = gs.ev_count;
      if (hash) hash[g] = gs.ev_hash;
      if (max_river) max_river[g] = mr;
    }
    if (hist) {
      std::lock_guard<std::mutex> lk(hist_mu);
      for (int i = 0; i < 128; i++) hist[i] += local[i];
    }
  };
  if (threads <= 1) {
    work();
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) th.emplace_back(work);
    for (auto& t : th) t.join();
  }
  return total.load();
}
int64_t orc_run_random(int mode, uint32_t rule, uint64_t seed_base, int64_t n, uint64_t agent_seed, uint32_t max_steps,
                       int threads, int32_t* scores, uint8_t* ranks, uint8_t* done, uint32_t* steps, uint32_t* kyoku,
                       uint32_t* evcount, uint64_t* hash, uint32_t* max_river) {
  return run_agent(0, mode, rule, seed_base, n, agent_seed, max_steps, threads, scores, ranks, done, steps, kyoku, evcount, hash,
                   max_river, nullptr);
}
int64_t orc_run_agent(int policy, int mode, uint32_t rule, uint64_t seed_base, int64_t n, uint64_t agent_seed, uint32_t max_steps,
                      int threads, int32_t* scores, uint8_t* ranks, uint8_t* done, uint32_t* steps, uint32_t* kyoku,
                      uint32_t* evcount, uint64_t* hash, uint64_t* hist) {
  return run_agent(policy, mode, rule, seed_base, n, agent_seed, max_steps, threads, scores, ranks, done, steps, kyoku, evcount,
                   hash, nullptr, hist);
}
// the same with explicit walls (n x 136 / 108 tiles): single-round modes never reach the seeded shuffle
int64_t orc_run_agent_walls(int policy, int mode, uint32_t rule, uint64_t seed_base, int64_t n, uint64_t agent_seed, uint32_t max_steps,
                            int threads, const uint8_t* walls, int32_t* scores, uint8_t* ranks, uint8_t* done, uint32_t* steps,
                            uint32_t* kyoku, uint32_t* evcount, uint64_t* hash) {
  return run_agent(policy, mode, rule, seed_base, n, agent_seed, max_steps, threads, scores, ranks, done, steps, kyoku, evcount,
                   hash, nullptr, nullptr, walls);
}
void orc_game_apply_event(void* h, const rv_mjai_event* e) { apply_mjai_event(*(GameState*)h, *e); }
// replay ingestion (test oracle of rv_vec_replay_begin / rv_vec_apply_log_actions)
void orc_game_apply_log_action(void* h, const rv_log_action* a) {
  if (a->type != RV_LA_NONE) apply_log_action(*(GameState*)h, *a);
}
void orc_game_replay_begin(void* h, const rv_log_kyoku* k) {
  GameState* g = (GameState*)h;
  int32_t sc[4] = {k->scores[0], k->scores[1], k->scores[2], k->scores[3]};
  g->reset((uint8_t)(k->oya < g->np ? k->oya : 0), (uint8_t)(k->chang < 4 ? k->chang : 0), k->ben, k->liqibang, nullptr, sc);
  replay_begin_patch(*g, *k);
}
int orc_game_agent_step(void* h, int policy, uint64_t agent_seed, uint64_t game_id) {
  return agent_step(*(GameState*)h, policy, agent_seed, game_id) ? 1 : 0;
}
}
extern "C" void orc_game_encode(void* h, int pid, float* obs, uint8_t* mask) {
  GameState* g = (GameState*)h;
  if (obs) g->np == 3 ? encode_obs_3p(*g, pid, obs) : encode_obs(*g, pid, obs);   // 74x34 (4P) / 74x27 (3P) floats
  if (mask) encode_mask(*g, pid, mask);
}
extern "C" void orc_game_encode(void* h, int pid, float* obs, uint8_t* mask);
// encode() + mask() rows of every seat that owes an action, n games, ascending (game, seat) order — the row order of
// rv_vec_encode.  obs [max_rows][74][W], mask [max_rows][IDS], index [max_rows] = game * 4 + seat.  Returns the row count.
extern "C" int64_t orc_games_encode_batch(void** hs, int64_t n, float* obs, uint8_t* mask, int32_t* index, int64_t max_rows) {
  int64_t row = 0;
  for (int64_t i = 0; i < n; i++) {
    GameState* g = (GameState*)hs[i];
    if (g->is_done) continue;
    const int W = g->np == 3 ? 27 : 34, IDS = g->np == 3 ? 60 : 82;
    for (int p = 0; p < g->np; p++) {
      bool owes = (g->phase == RV_WAIT_ACT && g->current_player == p) ||
                  (g->phase == RV_WAIT_RESPONSE &&
                   std::find(g->active_players.begin(), g->active_players.end(), (uint8_t)p) != g->active_players.end());
      if (!owes || row >= max_rows) continue;
      orc_game_encode(g, p, obs + row * 74 * W, mask + row * IDS);
      index[row] = (int32_t)(i * 4 + p);
      row++;
    }
  }
  return row;
}
// Observation::encode_extended: 215x34 floats (4P only)
extern "C" void orc_game_encode_ext(void* h, int pid, float* obs) {
  GameState* g = (GameState*)h;
  g->np == 3 ? encode_obs_3p_extended(*g, pid, obs) : encode_obs_extended(*g, pid, obs);   // 215x34 (4P) / 215x27 (3P)
}
// Observation::encode_kawa_overview: 4x7x34 floats (4P), 3x7x27 (sanma)
extern "C" void orc_game_encode_kawa(void* h, float* out) {
  GameState* g = (GameState*)h;
  if (g->sanma) encode_kawa_overview_3p(*g, out);
  else encode_kawa_overview(*g, out);
}
// shanten.rs:250-393 on tid lists (known-answer hooks): out = {shanten, effective_with_discard, best_ukeire}
extern "C" void orc_ukeire(const int* hand, int n, const int* visible, int nv, int* out) {
  std::vector<int> h(hand, hand + n), v(visible, visible + nv);
  out[0] = shanten_tiles(h);
  out[1] = effective_tiles_with_discard(h);
  out[2] = best_ukeire(h, v);
}
extern "C" void orc_ukeire_3p(const int* hand, int n, const int* visible, int nv, int* out) {   // shanten.rs:470-615
  std::vector<int> h(hand, hand + n), v(visible, visible + nv);
  out[0] = shanten_tiles_3p(h);
  out[1] = effective_tiles_3p_with_discard(h);
  out[2] = best_ukeire_3p(h, v);
}
// sequence features of seat pid over the event delta [w0, w1) (words); fixed-size outputs padded like the device path
extern "C" void orc_game_encode_seq(void* h, int pid, uint32_t w0, uint32_t w1, int game_style, uint16_t* sparse, float* numeric,
                                    uint16_t* prog, int max_prog, uint16_t* cand, uint16_t* lens) {
  GameState* g = (GameState*)h;
  SeqFeatures f = encode_seq(*g, pid, w0, w1, game_style);
  for (int k = 0; k < 25; k++) sparse[k] = k < (int)f.sparse.size() ? f.sparse[k] : 441;
  for (int k = 0; k < 12; k++) numeric[k] = f.numeric[k];
  static const uint16_t PP[5] = {4, 276, 2, 2, 4}, CP[4] = {279, 2, 2, 3};
  for (int k = 0; k < max_prog; k++)
    for (int j = 0; j < 5; j++) prog[5 * k + j] = k < (int)f.prog.size() ? f.prog[k][j] : PP[j];
  for (int k = 0; k < 64; k++)
    for (int j = 0; j < 4; j++) cand[4 * k + j] = k < (int)f.cand.size() ? f.cand[k][j] : CP[j];
  lens[0] = (uint16_t)f.sparse.size();
  lens[1] = (uint16_t)f.prog.size();
  lens[2] = (uint16_t)f.cand.size();
}
extern "C" int orc_seq_encode_chi(int c0, int c1, int called) { return seq_encode_chi({(uint8_t)c0, (uint8_t)c1}, called); }
extern "C" int orc_seq_encode_pon(int c0, int c1, int called) { return seq_encode_pon({(uint8_t)c0, (uint8_t)c1}, called); }
extern "C" int orc_seq_kan37(int tid) { return seq_tile_id_to_kan37(tid); }
extern "C" int orc_seq_relative_from(int actor, int target) { return seq_relative_from(actor, target); }
extern "C" int orc_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(rv_game_state);
    case 1: return (int)sizeof(rv_hand_query);
    case 2: return (int)sizeof(rv_hand_result);
    case 3: return (int)sizeof(rv_action);
  }
  return -1;
}
