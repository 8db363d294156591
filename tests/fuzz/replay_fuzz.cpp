// TEST INFRASTRUCTURE: mutation fuzzer for the host-side log readers (csrc/replay.cpp), built with -fsanitize=address,undefined by
// tests/test_replay.py::test_reader_fuzz_under_sanitizers.  Mutates digits / suit letters / a few key letters of a paifu and of an
// MJAI log, reads the result with every rv_replay_* entry point and checks the win-context queries stay inside their arrays.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <fstream>
#include <sstream>
#include <random>
#include "../../include/riichienv_b200.h"
int rv_internal_fail(int code, const std::string& msg) { (void)msg; return code; }
static std::string slurp(const char* p) { std::ifstream f(p); std::stringstream ss; ss << f.rdbuf(); return ss.str(); }
int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: replay_fuzz <paifu.json> <log.jsonl> <iterations>\n"); return 2; }
  std::string paifu = slurp(argv[1]), mjai = slurp(argv[2]);
  std::mt19937 rng(7);
  long ok = 0, bad = 0, ctxs = 0, refused = 0, prog = 0;
  const char* tiles = "0123456789mpsz";
  for (int it = 0; it < atoi(argv[3]); it++) {
    const bool use_paifu = it % 3 != 0;
    std::string t = use_paifu ? paifu : mjai;
    if (!use_paifu) { size_t cut = t.find('\n', rng() % t.size()); t = t.substr(0, cut == std::string::npos ? t.size() : cut + 1); }
    int flips = it == 0 ? 0 : 1 + rng() % 40;
    for (int k = 0; k < flips; k++) {
      size_t p = rng() % t.size();
      if (t[p] >= '0' && t[p] <= '9') t[p] = (char)('0' + rng() % 10);
      else if (t[p] == 'm' || t[p] == 'p' || t[p] == 's' || t[p] == 'z') t[p] = tiles[10 + rng() % 4];
      else if (t[p] >= 'a' && t[p] <= 'z' && (rng() % 8 == 0)) t[p] = (char)('a' + rng() % 26);
    }
    rv_replay* r = nullptr;
    int rc = use_paifu ? rv_replay_from_mjsoul_text(t.data(), t.size(), 0xC0, &r) : rv_replay_from_text(t.data(), t.size(), 0xC0, &r);
    if (rc != 0) { bad++; continue; }
    ok++;
    int n = rv_replay_num_rounds(r);
    for (int i = 0; i < n; i++) {
      rv_log_kyoku k; rv_replay_kyoku(r, i, &k);
      std::vector<rv_log_action> a(k.n_actions + 1); int m = 0; rv_replay_actions(r, i, a.data(), k.n_actions, &m);
      std::vector<rv_log_action_aux> x(k.n_actions + 1); rv_replay_actions_aux(r, i, x.data(), k.n_actions, &m);
      int len = 0; rv_replay_paishan(r, i, nullptr, 0, &len);
      if (len > 0) { std::vector<char> buf(len); rv_replay_paishan(r, i, buf.data(), len, &len); }
      int nc = 0;
      if (rv_replay_win_contexts(r, i, nullptr, 0, &nc) == 0) {
        std::vector<rv_win_context> c(nc + 1); rv_replay_win_contexts(r, i, c.data(), nc, &nc); ctxs += nc;
        for (int j = 0; j < nc; j++) if (c[j].query.n_tiles > 14 || c[j].query.n_melds > 4 || c[j].query.n_dora > 5 || c[j].query.n_ura > 5) { printf("bad query\n"); return 1; }
      } else refused++;
      std::vector<uint8_t> tg(k.n_actions + 1);
      for (auto& b : tg) b = rng() & 1;
      int np = 0; rv_replay_progression(a.data(), tg.data(), k.n_actions, nullptr, 0, &np);
      std::vector<uint16_t> out(5 * (np + 1)); rv_replay_progression(a.data(), tg.data(), k.n_actions, out.data(), np, &np); prog += np;
    }
    int nc = 0; rv_replay_win_contexts(r, -1, nullptr, 0, &nc);
    rv_replay_free(r);
  }
  printf("ok %ld bad %ld win contexts %ld refused rounds %ld progression tuples %ld\n", ok, bad, ctxs, refused, prog);
}
