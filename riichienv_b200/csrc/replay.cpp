// Host-side MJAI log reader of the replay-ingestion path: JSON lines (optionally gzip) -> one rv_log_kyoku
// plus a list of rv_log_action per round.  Replaces MjaiReplay::from_jsonl and KyokuBuilder
// (riichienv-core/src/replay/mjai_replay.rs:184-633; tile names: parser.rs:336-395 `mjai_to_tid`).  The records feed
// rv_vec_replay_begin / rv_vec_apply_log_actions, which track the kyoku on the device.
#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/riichienv_b200.h"

int rv_internal_fail(int code, const std::string& msg);   // riichienv_b200.cu: sets rv_last_error()

namespace {
// ---------------------------------------------------------------- a small JSON reader (objects, arrays, strings, numbers, literals)
struct JVal {
  enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
  bool b = false;
  double num = 0;
  std::string str;
  std::vector<JVal> arr;
  std::vector<std::pair<std::string, JVal>> obj;
  const JVal* get(const char* key) const {
    for (auto& kv : obj)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
};
struct JParser {
  const char* p;
  const char* end;
  std::string err;
  void ws() {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n')) p++;
  }
  bool fail(const char* m) {
    if (err.empty()) err = m;
    return false;
  }
  bool str(std::string& out) {
    if (p >= end || *p != '"') return fail("expected string");
    p++;
    while (p < end && *p != '"') {
      if (*p == '\\') {
        if (++p >= end) return fail("bad escape");
        switch (*p) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'u': {                                     // names only: keep the code point as UTF-8 (BMP)
            if (end - p < 5) return fail("bad \\u escape");
            unsigned cp = 0;
            for (int k = 1; k <= 4; k++) {
              char c = p[k];
              cp = cp * 16 + (c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : 0);
            }
            p += 4;
            if (cp < 0x80) out += (char)cp;
            else if (cp < 0x800) out += (char)(0xC0 | (cp >> 6)), out += (char)(0x80 | (cp & 0x3F));
            else out += (char)(0xE0 | (cp >> 12)), out += (char)(0x80 | ((cp >> 6) & 0x3F)), out += (char)(0x80 | (cp & 0x3F));
            break;
          }
          default: out += *p;
        }
        p++;
      } else {
        out += *p++;
      }
    }
    if (p >= end) return fail("unterminated string");
    p++;
    return true;
  }
  bool value(JVal& v, int depth = 0) {
    if (depth > 32) return fail("nesting too deep");
    ws();
    if (p >= end) return fail("EOF while parsing a value");
    if (*p == '{') {
      v.kind = JVal::Obj;
      p++;
      ws();
      if (p < end && *p == '}') return p++, true;
      v.obj.reserve(8);
      while (true) {
        ws();
        std::string k;
        if (!str(k)) return false;
        ws();
        if (p >= end || *p != ':') return fail("expected ':'");
        p++;
        JVal c;
        if (!value(c, depth + 1)) return false;
        v.obj.emplace_back(std::move(k), std::move(c));
        ws();
        if (p < end && *p == ',') { p++; continue; }
        if (p < end && *p == '}') return p++, true;
        return fail("expected ',' or '}'");
      }
    }
    if (*p == '[') {
      v.kind = JVal::Arr;
      p++;
      ws();
      if (p < end && *p == ']') return p++, true;
      v.arr.reserve(16);
      while (true) {
        JVal c;
        if (!value(c, depth + 1)) return false;
        v.arr.push_back(std::move(c));
        ws();
        if (p < end && *p == ',') { p++; continue; }
        if (p < end && *p == ']') return p++, true;
        return fail("expected ',' or ']'");
      }
    }
    if (*p == '"') {
      v.kind = JVal::Str;
      return str(v.str);
    }
    if (end - p >= 4 && !strncmp(p, "true", 4)) return v.kind = JVal::Bool, v.b = true, p += 4, true;
    if (end - p >= 5 && !strncmp(p, "false", 5)) return v.kind = JVal::Bool, v.b = false, p += 5, true;
    if (end - p >= 4 && !strncmp(p, "null", 4)) return v.kind = JVal::Null, p += 4, true;
    {                                                     // plain integers (nearly every number of a log) without strtod
      const char* q = p;
      const bool neg = q < end && *q == '-';
      if (neg) q++;
      const char* d0 = q;
      double acc = 0;
      while (q < end && *q >= '0' && *q <= '9' && q - d0 < 15) acc = acc * 10 + (*q++ - '0');
      if (q > d0 && (q >= end || (*q != '.' && *q != 'e' && *q != 'E' && !(*q >= '0' && *q <= '9') && *q != 'x' && *q != 'X'))) {
        v.kind = JVal::Num;
        v.num = neg ? -acc : acc;
        p = q;
        return true;
      }
    }
    char* e = nullptr;
    std::string tmp(p, (size_t)std::min<ptrdiff_t>(end - p, 40));
    double d = strtod(tmp.c_str(), &e);
    if (e == tmp.c_str()) return fail("expected value");
    v.kind = JVal::Num;
    v.num = d;
    p += e - tmp.c_str();
    return true;
  }
};

// parser.rs:336-395: "5mr" red fives are copy 0 of the five (16 / 52 / 88), a plain five is copy 1, everything else copy 0;
// honors by letter (E S W N P F C) or as 1z..7z.  Unknown strings -> None -> the caller's unwrap_or(0).
int mjai_to_tid(const std::string& s) {
  static const char* honors[7] = {"E", "S", "W", "N", "P", "F", "C"};
  for (int k = 0; k < 7; k++)
    if (s == honors[k]) return 108 + 4 * k;
  if (s == "5mr") return 16;
  if (s == "5pr") return 52;
  if (s == "5sr") return 88;
  if (s.size() < 2 || s[0] < '0' || s[0] > '9') return -1;   // only the first two characters are looked at
  const int num = s[0] - '0';
  const int suit = s[1] == 'm' ? 0 : s[1] == 'p' ? 1 : s[1] == 's' ? 2 : s[1] == 'z' ? 3 : -1;
  if (suit < 0) return -1;
  if (num == 0) return suit < 3 ? suit * 36 + 16 : -1;
  if (suit == 3) return 108 + (num - 1) * 4;                  // 8z / 9z give 136 / 140 as in the reference; consumers clamp
  return suit * 36 + (num - 1) * 4 + (num == 5 ? 1 : 0);
}
uint8_t tile_of(const JVal* v) {
  if (!v || v->kind != JVal::Str) return 0;
  int t = mjai_to_tid(v->str);
  return (uint8_t)(t < 0 ? 0 : t);   // parse_mjai_tile: unwrap_or(0)
}

// what the reference's Action carries besides the fields of rv_log_action (WinResultContextIterator and Kyoku.events() read them)
struct ActAux : rv_log_action_aux {
  ActAux() {
    memset(static_cast<rv_log_action_aux*>(this), 0, sizeof(rv_log_action_aux));
    n_doras = 0xFF;                       // DiscardTile / DealTile `doras`: None
    left_tile_count = 0xFF;               // DealTile `left_tile_count`: None
    tile_raw_id = 0;                      // AnGangAddGang `tile_raw_id` (0 in MJAI logs, mjai_replay.rs:507,518)
  }
};
struct Kyoku {
  rv_log_kyoku k;
  std::vector<rv_log_action> actions;
  std::vector<ActAux> aux;                // empty (all None) for MJAI logs, else one per action
  std::vector<uint8_t> wall;              // LogKyoku.paishan as tids; empty = None
  std::string paishan;                    // the string itself (Kyoku.paishan, the NewRound record of Kyoku.events())
  bool has_paishan = false;
};
// KyokuBuilder (mjai_replay.rs:159-270)
struct Builder {
  Kyoku out;
  int np = 4;
  bool liqi[4] = {}, wliqi[4] = {}, reach_accepted[4] = {}, reached[4] = {}, first_discard[4] = {true, true, true, true};
  bool has_calls = false;
  std::vector<rv_hule> pending_hule;
  void flush_hule() {
    if (pending_hule.empty()) return;
    rv_log_action a;
    memset(&a, 0, sizeof a);
    a.type = RV_LA_HULE;
    a.n_hule = (uint8_t)std::min<size_t>(pending_hule.size(), 3);
    for (int i = 0; i < a.n_hule; i++) a.hules[i] = pending_hule[i];
    out.actions.push_back(a);
    pending_hule.clear();
  }
};
rv_log_action blank(int type, int seat) {
  rv_log_action a;
  memset(&a, 0, sizeof a);
  a.type = (uint8_t)type;
  a.seat = (uint8_t)seat;
  a.tile = RV_NONE;
  memset(a.tiles, RV_NONE, 4);
  memset(a.froms, RV_NONE, 4);
  return a;
}
int geti(const JVal& o, const char* key, int dflt = 0) {
  const JVal* v = o.get(key);
  return v && v->kind == JVal::Num ? (int)v->num : dflt;
}
}  // namespace

struct rv_replay {
  std::vector<Kyoku> rounds;
};

namespace {
// serde's derive: a missing required field or a wrong type is a parse error, unknown `type` strings are MjaiEvent::Other
bool require(const JVal& o, const char* key, JVal::Kind kind, std::string& err) {
  const JVal* v = o.get(key);
  if (!v) return err = std::string("missing field `") + key + "`", false;
  if (v->kind != kind) return err = std::string("invalid type for field `") + key + "`", false;
  return true;
}
bool finish_builder(std::unique_ptr<Builder>& b, rv_replay* r) {
  if (!b) return false;
  b->flush_hule();
  rv_log_kyoku& k = b->out.k;
  for (int p = 0; p < 4; p++) k.wliqi[p] = b->wliqi[p];
  k.n_actions = (int32_t)b->out.actions.size();
  // LogKyoku::steps (replay/mod.rs:1125-1132, 1214-1247): the dealer and, for a 14-tile deal, the tile it holds as "drawn"
  int oya = k.ju % k.np;
  for (int p = 0; p < k.np; p++)
    if (k.hand_len[p] == 14) { oya = p; break; }
  k.oya = (uint8_t)oya;
  k.oya_drawn_tile = RV_NONE;
  if (k.hand_len[oya] == 14) {
    int dt = k.hands[oya][13];
    if (!b->out.actions.empty()) {
      const rv_log_action& a = b->out.actions[0];
      if (a.type == RV_LA_HULE) {
        for (int i = 0; i < a.n_hule; i++)
          if (a.hules[i].seat == oya && a.hules[i].zimo) { dt = a.hules[i].hu_tile; break; }
      } else if (a.type == RV_LA_DISCARD) {
        if (a.seat == oya) dt = a.tile;
      } else if (a.type == RV_LA_ANGANG_ADDGANG) {
        if (a.seat == oya && a.n_tiles) dt = a.tiles[0];
      }
    }
    k.oya_drawn_tile = (uint8_t)dt;
  }
  r->rounds.push_back(std::move(b->out));
  b.reset();
  return true;
}
// the two events that make up most of a log (mjai_replay.rs:392-433), shared by the DOM path and the flat-line scanner
void on_tsumo(Builder& b, int s, uint8_t tile) {
  rv_log_action a = blank(RV_LA_DEAL, s);
  a.tile = tile;
  b.out.actions.push_back(a);
  if (b.out.k.left_tile_count > 0) b.out.k.left_tile_count--;
}
void on_dahai(Builder& b, int s, uint8_t tile) {
  rv_log_action a = blank(RV_LA_DISCARD, s);
  a.tile = tile;
  const bool is_liqi = b.liqi[s];
  const bool is_wliqi = is_liqi && b.first_discard[s] && !b.has_calls;
  if (is_wliqi) b.wliqi[s] = true;
  a.flags = (uint8_t)((is_liqi ? 1 : 0) | (is_wliqi ? 2 : 0));
  b.out.actions.push_back(a);
  b.first_discard[s] = false;
  if (is_liqi) b.liqi[s] = false;
}
// A line that is a flat object of short plain strings, small non-negative integers and booleans — `tsumo` and `dahai`, most of
// a log — read without building the DOM.  Returns false for anything else (escapes, nesting, negative or long numbers, floats,
// a missing or mistyped field, another event type): the caller then takes the DOM path, which also words the errors.
bool flat_tsumo_dahai(const char* p, const char* z, Builder& b) {
  struct V { const char* k; int kl; int kind; const char* s; int sl; long num; } v[8];   // kind: 0 string, 1 integer, 2 bool
  int n = 0;
  auto ws = [&]() { while (p < z && (*p == ' ' || *p == '\t' || *p == '\r')) p++; };
  ws();
  if (p >= z || *p++ != '{') return false;
  while (true) {
    ws();
    if (p >= z || *p++ != '"' || n == 8) return false;
    V& x = v[n];
    x.k = p;
    while (p < z && *p != '"' && *p != '\\') p++;
    if (p >= z || *p != '"') return false;
    x.kl = (int)(p++ - x.k);
    ws();
    if (p >= z || *p++ != ':') return false;
    ws();
    if (p >= z) return false;
    if (*p == '"') {
      x.kind = 0;
      x.s = ++p;
      while (p < z && *p != '"' && *p != '\\') p++;
      if (p >= z || *p != '"') return false;
      x.sl = (int)(p++ - x.s);
    } else if (*p >= '0' && *p <= '9') {
      x.kind = 1;
      x.num = 0;
      const char* d0 = p;
      while (p < z && *p >= '0' && *p <= '9' && p - d0 < 9) x.num = x.num * 10 + (*p++ - '0');
      if (p < z && ((*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E' || *p == 'x' || *p == 'X')) return false;
    } else if (z - p >= 4 && !strncmp(p, "true", 4)) {
      x.kind = 2, x.num = 1, p += 4;
    } else if (z - p >= 5 && !strncmp(p, "false", 5)) {
      x.kind = 2, x.num = 0, p += 5;
    } else {
      return false;
    }
    n++;
    ws();
    if (p < z && *p == ',') { p++; continue; }
    if (p < z && *p == '}') { p++; break; }
    return false;
  }
  ws();
  if (p != z) return false;
  auto get = [&](const char* key) -> const V* {
    const int kl = (int)strlen(key);
    for (int i = 0; i < n; i++)
      if (v[i].kl == kl && !memcmp(v[i].k, key, (size_t)kl)) return &v[i];
    return nullptr;
  };
  const V* ty = get("type");
  if (!ty || ty->kind != 0) return false;
  const bool tsumo = ty->sl == 5 && !memcmp(ty->s, "tsumo", 5), dahai = ty->sl == 5 && !memcmp(ty->s, "dahai", 5);
  if (!tsumo && !dahai) return false;
  const V* actor = get("actor");
  const V* pai = get("pai");
  if (!actor || actor->kind != 1 || !pai || pai->kind != 0) return false;
  if (dahai) {
    const V* tg = get("tsumogiri");
    if (!tg || tg->kind != 2) return false;
  }
  const int s = actor->num >= b.np ? 0 : (int)actor->num;
  const int t = mjai_to_tid(std::string(pai->s, (size_t)pai->sl));
  const uint8_t tile = (uint8_t)(t < 0 ? 0 : t);
  b.flush_hule();
  if (tsumo) on_tsumo(b, s, tile);
  else on_dahai(b, s, tile);
  return true;
}
// MjaiReplay::process_event (mjai_replay.rs:388-632)
void process_event(Builder& b, const std::string& type, const JVal& e) {
  if (type != "hora") b.flush_hule();
  const int np = b.np;
  auto seat = [&](const char* key) { int a = geti(e, key); return a < 0 || a >= np ? 0 : a; };
  rv_log_kyoku& k = b.out.k;
  if (type == "tsumo") {
    on_tsumo(b, seat("actor"), tile_of(e.get("pai")));
  } else if (type == "dahai") {
    on_dahai(b, seat("actor"), tile_of(e.get("pai")));
  } else if (type == "reach") {
    const int s = seat("actor");
    b.liqi[s] = true;
    b.reached[s] = true;
  } else if (type == "reach_accepted") {
    b.reach_accepted[seat("actor")] = true;
  } else if (type == "chi" || type == "pon" || type == "kan" || type == "daiminkan") {
    b.has_calls = true;
    const int s = seat("actor"), target = seat("target");
    rv_log_action a = blank(RV_LA_CHI_PENG_GANG, s);
    a.meld_type = type == "chi" ? RV_MELD_CHI : type == "pon" ? RV_MELD_PON : RV_MELD_DAIMINKAN;
    a.tiles[0] = tile_of(e.get("pai"));
    a.froms[0] = (uint8_t)target;
    int n = 1;
    if (const JVal* c = e.get("consumed"))
      for (size_t i = 0; i < c->arr.size() && n < 4; i++) a.tiles[n] = tile_of(&c->arr[i]), a.froms[n] = (uint8_t)s, n++;
    a.n_tiles = (uint8_t)n;
    a.tile = a.tiles[0];
    b.out.actions.push_back(a);
  } else if (type == "ankan") {
    b.has_calls = true;
    rv_log_action a = blank(RV_LA_ANGANG_ADDGANG, seat("actor"));
    a.meld_type = RV_MELD_ANKAN;
    int n = 0;
    if (const JVal* c = e.get("consumed"))
      for (size_t i = 0; i < c->arr.size() && n < 4; i++) a.tiles[n++] = tile_of(&c->arr[i]);
    a.n_tiles = (uint8_t)n;
    b.out.actions.push_back(a);
  } else if (type == "kakan") {
    b.has_calls = true;
    rv_log_action a = blank(RV_LA_ANGANG_ADDGANG, seat("actor"));
    a.meld_type = RV_MELD_KAKAN;
    a.tiles[0] = tile_of(e.get("pai"));
    a.n_tiles = 1;
    b.out.actions.push_back(a);
  } else if (type == "dora") {
    const uint8_t m = tile_of(e.get("dora_marker"));
    if (k.n_doras < RV_LOG_MAX_DORAS) k.doras[k.n_doras++] = m;
    rv_log_action a = blank(RV_LA_DORA, 0);
    a.tile = m;
    b.out.actions.push_back(a);
  } else if (type == "hora") {
    const int actor = seat("actor"), target = seat("target");
    rv_hule h;
    memset(&h, 0, sizeof h);
    const JVal* pai = e.get("pai");
    if (pai && pai->kind == JVal::Str) {
      h.hu_tile = tile_of(pai);
    } else if (!b.out.actions.empty()) {              // infer from the last action
      const rv_log_action& last = b.out.actions.back();
      if (last.type == RV_LA_DEAL || last.type == RV_LA_DISCARD) h.hu_tile = last.tile;
      else if (last.type == RV_LA_ANGANG_ADDGANG) h.hu_tile = last.tiles[0];
    }
    h.seat = (uint8_t)actor;
    h.zimo = actor == target;
    h.count = (uint32_t)geti(e, "han");
    h.fu = (uint32_t)geti(e, "fu");
    h.n_li_doras = 0xFF;
    const JVal* ura = e.get("uradora_markers");
    if (!ura) ura = e.get("ura_markers");
    if (ura && ura->kind == JVal::Arr) {
      h.n_li_doras = (uint8_t)std::min<size_t>(ura->arr.size(), 5);
      k.n_ura_doras = 0;
      for (size_t i = 0; i < ura->arr.size(); i++) {
        const uint8_t t = tile_of(&ura->arr[i]);
        if (i < 5) h.li_doras[i] = t;
        if (k.n_ura_doras < RV_LOG_MAX_DORAS) k.ura_doras[k.n_ura_doras++] = t;
      }
    }
    const JVal* scores = e.get("scores");
    const JVal* delta = e.get("delta");
    if (!delta) delta = e.get("deltas");
    if (scores && scores->kind == JVal::Arr) {
      for (size_t i = 0; i < scores->arr.size() && i < 4; i++) k.end_scores[i] = (int32_t)scores->arr[i].num;
    } else if (delta && delta->kind == JVal::Arr) {
      const bool first = b.pending_hule.empty();
      for (size_t i = 0; i < delta->arr.size() && i < (size_t)np; i++) {
        const int32_t d = (int32_t)delta->arr[i].num;
        if (first) k.end_scores[i] = k.scores[i] + d - (b.reach_accepted[i] ? 1000 : 0);
        else k.end_scores[i] += d;
      }
    }
    b.pending_hule.push_back(h);
  } else if (type == "kita") {
    b.out.actions.push_back(blank(RV_LA_BABEI, seat("actor")));
  } else if (type == "ryukyoku") {
    const JVal* scores = e.get("scores");
    const JVal* delta = e.get("delta");
    if (!delta) delta = e.get("deltas");
    if (scores && scores->kind == JVal::Arr) {
      for (size_t i = 0; i < scores->arr.size() && i < 4; i++) k.end_scores[i] = (int32_t)scores->arr[i].num;
    } else if (delta && delta->kind == JVal::Arr) {
      for (size_t i = 0; i < delta->arr.size() && i < (size_t)np; i++)
        k.end_scores[i] = k.scores[i] + (int32_t)delta->arr[i].num - (b.reached[i] ? 1000 : 0);
    }
    b.out.actions.push_back(blank(RV_LA_NOTILE, 0));
  }
}

int parse_lines(const char* text, size_t len, uint32_t rule_bits, rv_replay** out) {
  std::unique_ptr<rv_replay> r(new rv_replay);
  std::unique_ptr<Builder> b;
  const char* p = text;
  const char* end = text + len;
  int line_no = 0;
  while (p < end) {
    const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
    const char* le = nl ? nl : end;
    line_no++;
    const char* a = p;
    const char* z = le;
    while (a < z && (*a == ' ' || *a == '\t' || *a == '\r')) a++;
    while (z > a && (z[-1] == ' ' || z[-1] == '\t' || z[-1] == '\r')) z--;
    p = nl ? nl + 1 : end;
    if (a == z) continue;
    if (b && flat_tsumo_dahai(a, z, *b)) continue;
    JParser jp{a, z, {}};
    JVal e;
    bool ok = jp.value(e);
    if (ok) {
      jp.ws();
      if (jp.p != z) ok = jp.fail("trailing characters");
    }
    if (ok && e.kind != JVal::Obj) ok = jp.fail("invalid type: expected an event object");
    std::string type;
    if (ok) {
      const JVal* t = e.get("type");
      if (!t || t->kind != JVal::Str) ok = jp.fail("missing field `type`");
      else type = t->str;
    }
    std::string err = jp.err;
    if (ok) {
      // field checks of the serde enum (mjai_replay.rs:67-156) for the variants this reader consumes
      if (type == "start_kyoku") {
        ok = require(e, "bakaze", JVal::Str, err) && require(e, "kyoku", JVal::Num, err) && require(e, "honba", JVal::Num, err) &&
             (e.get("kyoutaku") ? require(e, "kyoutaku", JVal::Num, err) : require(e, "kyotaku", JVal::Num, err)) &&
             require(e, "oya", JVal::Num, err) && require(e, "scores", JVal::Arr, err) && require(e, "dora_marker", JVal::Str, err) &&
             require(e, "tehais", JVal::Arr, err);
      } else if (type == "tsumo") {
        ok = require(e, "actor", JVal::Num, err) && require(e, "pai", JVal::Str, err);
      } else if (type == "dahai") {
        ok = require(e, "actor", JVal::Num, err) && require(e, "pai", JVal::Str, err) && require(e, "tsumogiri", JVal::Bool, err);
      } else if (type == "pon" || type == "chi" || type == "kan" || type == "daiminkan") {
        ok = require(e, "actor", JVal::Num, err) && require(e, "target", JVal::Num, err) && require(e, "pai", JVal::Str, err) &&
             require(e, "consumed", JVal::Arr, err);
      } else if (type == "kakan") {
        ok = require(e, "actor", JVal::Num, err) && require(e, "pai", JVal::Str, err);
      } else if (type == "ankan") {
        ok = require(e, "actor", JVal::Num, err) && require(e, "consumed", JVal::Arr, err);
      } else if (type == "dora") {
        ok = require(e, "dora_marker", JVal::Str, err);
      } else if (type == "reach" || type == "reach_accepted" || type == "kita") {
        ok = require(e, "actor", JVal::Num, err);
      } else if (type == "hora") {
        ok = require(e, "actor", JVal::Num, err) && require(e, "target", JVal::Num, err);
      }
    }
    if (!ok) return rv_internal_fail(RV_ERR_INVALID, "Parse error: " + err + " at line " + std::to_string(line_no));
    if (type == "start_kyoku") {
      if (b) finish_builder(b, r.get());
      b.reset(new Builder);
      b->out.actions.reserve(192);
      rv_log_kyoku& k = b->out.k;
      memset(&k, 0, sizeof k);
      const JVal& scores = *e.get("scores");
      const int np = (int)scores.arr.size();
      if (np != 3 && np != 4)
        return rv_internal_fail(RV_ERR_INVALID, "Parse error: start_kyoku with " + std::to_string(np) + " scores at line " + std::to_string(line_no));
      b->np = np;
      k.np = (uint8_t)np;
      const std::string& bk = e.get("bakaze")->str;
      k.chang = bk == "S" ? 1 : bk == "W" ? 2 : bk == "N" ? 3 : 0;
      k.ju = (uint8_t)(geti(e, "kyoku") - 1);
      k.ben = (uint8_t)geti(e, "honba");
      k.liqibang = (uint8_t)geti(e, e.get("kyoutaku") ? "kyoutaku" : "kyotaku");
      k.left_tile_count = np == 3 ? 55 : 70;
      k.rule_bits = rule_bits;
      for (int p = 0; p < np; p++) k.scores[p] = k.end_scores[p] = (int32_t)scores.arr[p].num;
      memset(k.doras, RV_NONE, sizeof k.doras);
      memset(k.ura_doras, RV_NONE, sizeof k.ura_doras);
      memset(k.hands, RV_NONE, sizeof k.hands);
      k.doras[0] = tile_of(e.get("dora_marker"));
      k.n_doras = 1;
      const JVal& th = *e.get("tehais");
      for (size_t p = 0; p < th.arr.size() && p < (size_t)np; p++) {
        int n = 0;
        for (size_t i = 0; i < th.arr[p].arr.size() && n < 14; i++) k.hands[p][n++] = tile_of(&th.arr[p].arr[i]);
        k.hand_len[p] = (uint8_t)n;
      }
    } else if (type == "end_kyoku" || type == "end_game") {
      if (b) finish_builder(b, r.get());
    } else if (b) {
      process_event(*b, type, e);
    }
  }
  if (b) finish_builder(b, r.get());
  // the next round's start scores are the authoritative post-round scores of every non-final round
  for (size_t i = 0; i + 1 < r->rounds.size(); i++)
    for (int p = 0; p < 4; p++) r->rounds[i].k.end_scores[p] = r->rounds[i + 1].k.scores[p];
  *out = r.release();
  return RV_OK;
}

// ---------------------------------------------------------------- MjSoul paifu (replay/mjsoul_replay.rs)
// TileConverter::parse_tile / parse_tile_136 (replay/mod.rs:2184-2222): "<digit><suit>", digit 0 = red five
uint8_t paifu_tile(const JVal* v) {
  if (!v || v->kind != JVal::Str || v->str.empty()) return 0;
  const std::string& t = v->str;
  int num = (t[0] >= '0' && t[0] <= '9') ? t[0] - '0' : 0;
  const bool aka = num == 0;
  if (aka) num = 5;
  const std::string suit = t.substr(1);
  int id34 = suit == "m" ? num - 1 : suit == "p" ? 9 + num - 1 : suit == "s" ? 18 + num - 1 : suit == "z" ? 27 + num - 1 : 0;
  id34 &= 0xFF;
  if (aka) return id34 == 4 ? 16 : id34 == 13 ? 52 : id34 == 22 ? 88 : (uint8_t)(id34 * 4);
  if (id34 == 4 || id34 == 13 || id34 == 22) return (uint8_t)(id34 * 4 + 1);
  return (uint8_t)(id34 * 4);
}
bool getb(const JVal& o, const char* key) {
  const JVal* v = o.get(key);
  return v && v->kind == JVal::Bool && v->b;
}
// MjSoulReplay::parse_raw_action (mjsoul_replay.rs:541-687); `name` / `data` as serde's adjacently tagged enum reads them
rv_log_action paifu_action(const std::string& name, const JVal& d) {
  static const JVal empty;
  auto seat = [&]() { int s = geti(d, "seat"); return s < 0 || s > 3 ? 0 : s; };
  if (name == "DiscardTile") {
    rv_log_action a = blank(RV_LA_DISCARD, seat());
    a.tile = paifu_tile(d.get("tile"));
    a.flags = (uint8_t)((getb(d, "is_liqi") ? 1 : 0) | (getb(d, "is_wliqi") ? 2 : 0));
    return a;
  }
  if (name == "DealTile") {
    rv_log_action a = blank(RV_LA_DEAL, seat());
    a.tile = paifu_tile(d.get("tile"));
    return a;
  }
  if (name == "ChiPengGang") {
    rv_log_action a = blank(RV_LA_CHI_PENG_GANG, seat());
    const int mt = geti(d, "type");
    a.meld_type = mt == 0 ? RV_MELD_CHI : mt == 1 ? RV_MELD_PON : mt == 2 ? RV_MELD_DAIMINKAN : mt == 3 ? RV_MELD_ANKAN : RV_MELD_CHI;
    int n = 0;
    const JVal* tiles = d.get("tiles");
    const JVal* froms = d.get("froms");
    if (tiles && tiles->kind == JVal::Arr)
      for (size_t i = 0; i < tiles->arr.size() && n < 4; i++) {
        a.tiles[n] = paifu_tile(&tiles->arr[i]);
        a.froms[n] = froms && i < froms->arr.size() ? (uint8_t)froms->arr[i].num : (uint8_t)RV_NONE;
        n++;
      }
    a.n_tiles = (uint8_t)n;
    a.tile = n ? a.tiles[0] : (uint8_t)RV_NONE;
    return a;
  }
  if (name == "AnGangAddGang") {
    rv_log_action a = blank(RV_LA_ANGANG_ADDGANG, seat());
    a.meld_type = geti(d, "type") == 3 ? RV_MELD_ANKAN : RV_MELD_KAKAN;
    a.tiles[0] = paifu_tile(d.get("tiles"));
    a.n_tiles = 1;
    return a;
  }
  if (name == "Hule") {
    rv_log_action a = blank(RV_LA_HULE, 0);
    const JVal* hs = d.get("hules");
    int n = 0;
    if (hs && hs->kind == JVal::Arr)
      for (size_t i = 0; i < hs->arr.size() && n < 3; i++, n++) {
        const JVal& h = hs->arr[i];
        rv_hule& o = a.hules[n];
        memset(&o, 0, sizeof o);
        o.seat = (uint8_t)geti(h, "seat");
        o.hu_tile = paifu_tile(h.get("hu_tile"));
        o.zimo = getb(h, "zimo");
        o.yiman = getb(h, "yiman");
        o.count = (uint32_t)geti(h, "count");
        o.fu = (uint32_t)geti(h, "fu");
        o.point_rong = (uint32_t)geti(h, "point_rong");
        o.point_zimo_qin = (uint32_t)geti(h, "point_zimo_qin");
        o.point_zimo_xian = (uint32_t)geti(h, "point_zimo_xian");
        if (const JVal* fans = h.get("fans"))
          for (auto& f : fans->arr) {
            const int id = geti(f, "id"), val = geti(f, "val");     // fans with val 0 are dropped
            if (val > 0 && id >= 0 && id < 64) o.fans |= 1ull << id;
          }
        const JVal* li = h.get("ura_dora_indicators");
        if (!li || li->kind != JVal::Arr) li = h.get("li_doras");
        o.n_li_doras = 0xFF;
        if (li && li->kind == JVal::Arr) {
          o.n_li_doras = (uint8_t)std::min<size_t>(li->arr.size(), 5);
          for (int k = 0; k < o.n_li_doras; k++) o.li_doras[k] = paifu_tile(&li->arr[k]);
        }
      }
    a.n_hule = (uint8_t)n;
    return a;
  }
  if (name == "dora") {
    rv_log_action a = blank(RV_LA_DORA, 0);
    a.tile = paifu_tile(d.get("dora_marker"));
    return a;
  }
  if (name == "NoTile") return blank(RV_LA_NOTILE, 0);
  if (name == "BaBei") {
    rv_log_action a = blank(RV_LA_BABEI, seat());
    a.flags = getb(d, "moqie") ? 1 : 0;
    return a;
  }
  if (name == "LiuJu") {
    rv_log_action a = blank(RV_LA_LIUJU, seat());
    a.tile = (uint8_t)geti(d, "type");
    int n = 0;
    if (const JVal* tiles = d.get("tiles"))
      for (size_t i = 0; i < tiles->arr.size() && n < 4; i++) a.tiles[n++] = paifu_tile(&tiles->arr[i]);   // (the record keeps four)
    a.n_tiles = (uint8_t)n;
    return a;
  }
  return blank(RV_LA_NONE, 0);                            // NewRound and unknown names: Action::Other
}
// the Option fields of DiscardTile / DealTile / AnGangAddGang (mjsoul_replay.rs:547-640)
ActAux paifu_aux(const std::string& name, const JVal& d) {
  ActAux x;
  auto list = [&](const JVal* v) {
    if (!v || v->kind != JVal::Arr || v->arr.empty()) return false;
    x.n_doras = (uint8_t)std::min<size_t>(v->arr.size(), RV_LOG_MAX_DORAS);
    for (int i = 0; i < x.n_doras; i++) x.doras[i] = paifu_tile(&v->arr[i]);
    return true;
  };
  if (name == "DiscardTile") {
    list(d.get("doras"));
  } else if (name == "DealTile") {
    if (!list(d.get("doras"))) {
      const JVal* dm = d.get("dora_marker");
      if (dm && dm->kind == JVal::Str) x.n_doras = 1, x.doras[0] = paifu_tile(dm);
    }
    const JVal* lc = d.get("left_tile_count");
    if (lc && lc->kind == JVal::Num) x.left_tile_count = (uint8_t)std::min(254, std::max(0, (int)lc->num));
  } else if (name == "AnGangAddGang") {
    x.tile_raw_id = (uint8_t)(paifu_tile(d.get("tiles")) >> 2);   // TileConverter::parse_tile_34(&tiles).0
  }
  return x;
}
// MjSoulReplay::kyoku_from_raw_actions (mjsoul_replay.rs:448-539) + the oya / drawn-tile derivation of LogKyoku::steps
bool paifu_round(const JVal& round, uint32_t rule_bits, Kyoku& out, std::string& err) {
  if (round.kind != JVal::Arr || round.arr.empty()) return err = "a round is not a list of actions", false;
  rv_log_kyoku& k = out.k;
  memset(&k, 0, sizeof k);
  memset(k.doras, RV_NONE, sizeof k.doras);
  memset(k.ura_doras, RV_NONE, sizeof k.ura_doras);
  memset(k.hands, RV_NONE, sizeof k.hands);
  k.np = 4;
  k.left_tile_count = 70;
  k.rule_bits = rule_bits;
  auto name_of = [](const JVal& a) { const JVal* n = a.get("name"); return n && n->kind == JVal::Str ? n->str : std::string(); };
  static const JVal none;
  const JVal& first = round.arr[0];
  if (name_of(first) == "NewRound") {
    const JVal* dp = first.get("data");
    const JVal& d = dp ? *dp : none;
    const JVal* sc = d.get("scores");
    if (!sc || sc->kind != JVal::Arr) return err = "NewRound: missing field `scores`", false;
    k.np = sc->arr.size() == 3 ? 3 : 4;
    for (size_t p = 0; p < sc->arr.size() && p < 4; p++) k.scores[p] = k.end_scores[p] = (int32_t)sc->arr[p].num;
    const JVal* da = d.get("dora_indicators");
    if (!da || da->kind != JVal::Arr) da = d.get("doras");
    if (da && da->kind == JVal::Arr) {
      for (size_t i = 0; i < da->arr.size() && k.n_doras < RV_LOG_MAX_DORAS; i++) k.doras[k.n_doras++] = paifu_tile(&da->arr[i]);
    } else if (const JVal* dm = d.get("dora_marker")) {
      if (dm->kind == JVal::Str) k.doras[k.n_doras++] = paifu_tile(dm);
    }
    static const char* tk[4] = {"tiles0", "tiles1", "tiles2", "tiles3"};
    for (int p = 0; p < 4; p++) {
      const JVal* t = d.get(tk[p]);
      if (!t || t->kind != JVal::Arr) {
        if (p < k.np) return err = std::string("NewRound: missing field `") + tk[p] + "`", false;
        continue;
      }
      int n = 0;
      for (size_t i = 0; i < t->arr.size() && n < 14; i++) k.hands[p][n++] = paifu_tile(&t->arr[i]);
      k.hand_len[p] = (uint8_t)n;
    }
    k.chang = (uint8_t)geti(d, "chang");
    k.ju = (uint8_t)geti(d, "ju");
    k.ben = (uint8_t)(d.get("ben") && d.get("ben")->kind == JVal::Num ? geti(d, "ben") : geti(d, "honba"));
    k.liqibang = (uint8_t)geti(d, "liqibang");
    k.left_tile_count = (uint8_t)geti(d, "left_tile_count", 70);
    if (const JVal* ud = d.get("ura_doras"))
      for (size_t i = 0; i < ud->arr.size() && k.n_ura_doras < RV_LOG_MAX_DORAS; i++) k.ura_doras[k.n_ura_doras++] = paifu_tile(&ud->arr[i]);
    if (const JVal* ps = d.get("paishan"))
      if (ps->kind == JVal::Str) {                         // parse_paishan (replay/mod.rs:1631-1641): two characters per tile
        out.paishan = ps->str;
        out.has_paishan = true;
        for (size_t i = 0; i + 1 < ps->str.size(); i += 2) {
          JVal t;
          t.kind = JVal::Str;
          t.str = ps->str.substr(i, 2);
          out.wall.push_back(paifu_tile(&t));
        }
      }
  }
  for (auto& a : round.arr) {
    const JVal* dp = a.get("data");
    out.actions.push_back(paifu_action(name_of(a), dp ? *dp : none));
    out.aux.push_back(paifu_aux(name_of(a), dp ? *dp : none));
  }
  for (auto& a : out.actions)
    if (a.type == RV_LA_DISCARD && (a.flags & 2) && a.seat < 4) k.wliqi[a.seat] = 1;
  k.n_actions = (int32_t)out.actions.size();
  int oya = k.ju % k.np;
  for (int p = 0; p < k.np; p++)
    if (k.hand_len[p] == 14) { oya = p; break; }
  k.oya = (uint8_t)oya;
  k.oya_drawn_tile = RV_NONE;
  if (k.hand_len[oya] == 14) {
    int dt = k.hands[oya][13];
    for (auto& a : out.actions) {                          // `self.actions.first()` is the NewRound placeholder: no match, dt stays
      (void)a;
      break;
    }
    k.oya_drawn_tile = (uint8_t)dt;
  }
  return true;
}
int parse_paifu(const char* text, size_t len, uint32_t rule_bits, rv_replay** out) {
  JParser jp{text, text + len, {}};
  JVal v;
  if (!jp.value(v)) return rv_internal_fail(RV_ERR_INVALID, "Failed to parse JSON: " + jp.err);
  const JVal* rounds = nullptr;
  if (v.kind == JVal::Obj) {
    rounds = v.get("rounds");                              // GameLog { rounds } (from_json)
    if (!rounds) rounds = v.get("data");                   // Paifu { header, data } (from_dict)
    if (!rounds) return rv_internal_fail(RV_ERR_INVALID, "Invalid dict format: missing 'data'");
  } else if (v.kind == JVal::Arr) {
    rounds = &v;
  } else {
    return rv_internal_fail(RV_ERR_INVALID, "Invalid input format: expected dict or list");
  }
  if (rounds->kind != JVal::Arr) return rv_internal_fail(RV_ERR_INVALID, "Failed to parse rounds list");
  std::unique_ptr<rv_replay> r(new rv_replay);
  for (auto& rd : rounds->arr) {
    Kyoku k;
    std::string err;
    if (!paifu_round(rd, rule_bits, k, err)) return rv_internal_fail(RV_ERR_INVALID, "Failed to parse rounds: " + err);
    r->rounds.push_back(std::move(k));
  }
  for (size_t i = 0; i + 1 < r->rounds.size(); i++)
    for (int p = 0; p < 4; p++) r->rounds[i].k.end_scores[p] = r->rounds[i + 1].k.scores[p];
  *out = r.release();
  return RV_OK;
}
// ---------------------------------------------------------------- the progression cache of replay mode
// GameState::apply_log_action with enable_seq_caching (state/event_handler.rs:351-379, 443-486, 573-606) feeds a synthesized
// reach / dahai / chi / pon / daiminkan / ankan / kakan event to process_single_event_progression
// (observation/sequence_features.rs:212-313); a replay observation's encode_seq_progression returns that list as it stands
// (sequence_features.rs:505-507: no start marker, no 512 cap).  The tile codes are the ones of seq.cuh.
bool prog_red(int t) { return t == 16 || t == 52 || t == 88; }
int prog_kan37(int t) {
  if (t == 16) return 0;
  if (t == 52) return 10;
  if (t == 88) return 20;
  const int k = t >> 2;
  return k < 9 ? k + 1 : k < 18 ? k + 2 : k + 3;
}
int prog_chi(const std::vector<int>& consumed, int called) {            // encode_chi (93-133)
  std::vector<int> all(consumed);
  all.push_back(called);
  std::sort(all.begin(), all.end());
  const int first = all[0] / 4, suit = first / 9, start = first - 9 * suit, call_pos = called / 4 - 9 * suit - start;
  bool red = false;
  for (int t : all) red |= prog_red(t);
  const bool five = start <= 4 && 4 <= start + 2;
  int off = 0;
  for (int q = 0; q < start; q++) off += (q <= 4 && 4 <= q + 2) ? 6 : 3;
  return suit * 30 + off + (five && red ? 3 + call_pos : call_pos);
}
int prog_pon(const std::vector<int>& consumed, int called) {            // encode_pon (144-183)
  const int kind = called / 4, suit = kind / 9;
  if (suit == 3) return 33 + kind - 27;
  const int rank = kind - 9 * suit;
  if (rank == 4) {
    bool red = false;
    for (int t : consumed) red |= prog_red(t);
    return suit * 11 + 4 + (prog_red(called) ? 2 : red ? 1 : 0);
  }
  return suit * 11 + (rank < 4 ? rank : rank + 2);
}
}  // namespace

extern "C" int rv_replay_progression(const rv_log_action* actions, const uint8_t* tsumogiri, int n, uint16_t* out, int cap, int* n_out) {
  if (n < 0 || cap < 0 || (n > 0 && (!actions || !tsumogiri)) || (cap > 0 && !out) || !n_out)
    return rv_internal_fail(RV_ERR_INVALID, "rv_replay_progression: bad arguments");
  int m = 0;
  auto push = [&](int actor, int type, int moqie, int liqi, int from) {
    if (m < cap) {
      uint16_t* r = out + 5 * m;
      r[0] = (uint16_t)actor, r[1] = (uint16_t)type, r[2] = (uint16_t)moqie, r[3] = (uint16_t)liqi, r[4] = (uint16_t)from;
    }
    m++;
  };
  for (int i = 0; i < n; i++) {
    const rv_log_action& a = actions[i];
    const int s = a.seat;
    if (a.type == RV_LA_DISCARD) {
      if (a.tile < 136) push(s, 1 + prog_kan37(a.tile), tsumogiri[i] ? 1 : 0, (a.flags & 3) ? 1 : 0, 4);
    } else if (a.type == RV_LA_CHI_PENG_GANG) {
      if (a.meld_type != RV_MELD_CHI && a.meld_type != RV_MELD_PON && a.meld_type != RV_MELD_DAIMINKAN) continue;
      int target = 0, called = -1;
      bool have_target = false;
      std::vector<int> consumed;
      for (int k = 0; k < a.n_tiles && k < 4; k++) {
        if (a.froms[k] != s) {
          if (!have_target) have_target = true, target = a.froms[k], called = a.tiles[k];
        } else if (a.tiles[k] < 136) {
          consumed.push_back(a.tiles[k]);
        }
      }
      if (called < 0 || called >= 136) continue;                        // pai "" does not parse: no entry
      const int rel = ((target - s + 3) % 4 + 4) % 4;
      if (a.meld_type == RV_MELD_DAIMINKAN) push(s, 168 + prog_kan37(called), 2, 2, rel);
      else if (consumed.size() >= 2) push(s, a.meld_type == RV_MELD_CHI ? 38 + prog_chi(consumed, called) : 128 + prog_pon(consumed, called), 2, 2, rel);
    } else if (a.type == RV_LA_ANGANG_ADDGANG && a.n_tiles > 0 && a.tiles[0] < 136) {
      if (a.meld_type == RV_MELD_ANKAN) push(s, 205 + a.tiles[0] / 4, 2, 2, 4);
      else push(s, 239 + prog_kan37(a.tiles[0]), 2, 2, 4);
    }
  }
  *n_out = m;
  return RV_OK;
}
namespace {
// ---------------------------------------------------------------- WinResultContextIterator (replay/mod.rs:1594-2093)
// The walk over one kyoku that tracks hands, melds and the win conditions up to every Hule and hands HandEvaluator::calc its
// arguments.  Here it only BUILDS the queries: the evaluation is one rv_hand_eval_batch over every context of a log.
struct WinWalk {
  struct M { uint8_t type, n, t[4]; int8_t from; uint8_t called; };
  const Kyoku& ky;
  int np;
  std::vector<uint8_t> hand[4];
  std::vector<M> melds[4];
  bool liqi[4] = {}, wliqi[4] = {}, ippatsu[4] = {}, rinshan[4] = {}, first_turn[4] = {true, true, true, true};
  bool ippatsu_before_babei[4] = {};
  bool last_kakan = false, last_babei = false;
  int kakan_tile = -1;
  std::vector<uint8_t> doras;
  int left, dora_count = 1, pending_minkan = 0;
  int kita[4] = {};
  explicit WinWalk(const Kyoku& k) : ky(k), np(k.k.np), left(k.k.left_tile_count) {
    for (int p = 0; p < np; p++) hand[p].assign(k.k.hands[p], k.k.hands[p] + k.k.hand_len[p]);
    doras.assign(k.k.doras, k.k.doras + k.k.n_doras);
  }
  static void all(bool (&a)[4], bool v) { a[0] = a[1] = a[2] = a[3] = v; }
  static bool match_and_remove(std::vector<uint8_t>& h, uint8_t t) {   // TileConverter::match_and_remove_u8 (2243-2255)
    auto it = std::find(h.begin(), h.end(), t);
    if (it == h.end()) it = std::find_if(h.begin(), h.end(), [&](uint8_t x) { return x / 4 == t / 4; });
    if (it == h.end()) return false;
    h.erase(it);
    return true;
  }
  void recalc_doras() {                                                // 1679-1696
    const size_t len = ky.wall.size();
    if (!len) return;
    doras.clear();
    for (int i = 0; i < dora_count; i++)
      if (len >= 5 + 2 * (size_t)i) doras.push_back(ky.wall[len - 5 - 2 * i]);
  }
  void sync_doras() {                                                  // 1698-1712
    if (ky.wall.empty()) return;
    if ((int)doras.size() > dora_count) dora_count = (int)doras.size(), pending_minkan = 0;
    else if (dora_count > (int)doras.size()) recalc_doras();
  }
  std::vector<uint8_t> ura_from_wall() const {                         // 1714-1728
    std::vector<uint8_t> u;
    const size_t len = ky.wall.size();
    for (int i = 0; i < dora_count; i++)
      if (len >= 6 + 2 * (size_t)i) u.push_back(ky.wall[len - 6 - 2 * i]);
    return u;
  }
  void after_kakan_reset() {
    if (!last_kakan) return;
    all(ippatsu, false), all(first_turn, false);
    last_kakan = last_babei = false;
    kakan_tile = -1;
  }
  void flush_pending() {
    if (pending_minkan > 0) dora_count += pending_minkan, pending_minkan = 0;
  }
  bool run(int round, std::vector<rv_win_context>& out, std::string& err) {
    static const ActAux none;
    for (size_t ai = 0; ai < ky.actions.size(); ai++) {
      const rv_log_action& a = ky.actions[ai];
      const ActAux& x = ai < ky.aux.size() ? ky.aux[ai] : none;
      if (a.type != RV_LA_HULE) {
        all(rinshan, false);
        if (a.type != RV_LA_BABEI) last_babei = false;
      }
      const int s = a.seat;
      if ((a.type == RV_LA_DISCARD || a.type == RV_LA_DEAL || a.type == RV_LA_CHI_PENG_GANG || a.type == RV_LA_ANGANG_ADDGANG ||
           a.type == RV_LA_BABEI) && s >= np)
        return err = "action of a seat the kyoku does not have", false;
      switch (a.type) {
        case RV_LA_DISCARD:
          after_kakan_reset();
          if (a.flags & 2) wliqi[s] = ippatsu[s] = true;
          if (a.flags & 1) liqi[s] = ippatsu[s] = true;
          if (!(a.flags & 1)) ippatsu[s] = false;
          first_turn[s] = false;
          match_and_remove(hand[s], a.tile);
          if (x.n_doras != 0xFF) doras.assign(x.doras, x.doras + x.n_doras);
          flush_pending();
          sync_doras();
          break;
        case RV_LA_DEAL:
          after_kakan_reset();
          hand[s].push_back(a.tile);
          if (x.left_tile_count != 0xFF) left = x.left_tile_count;
          else if (left > 0) left--;
          if (x.n_doras != 0xFF) doras.assign(x.doras, x.doras + x.n_doras), rinshan[s] = true;
          sync_doras();
          break;
        case RV_LA_CHI_PENG_GANG: {
          all(rinshan, false), all(ippatsu, false), all(first_turn, false);
          last_kakan = last_babei = false;
          kakan_tile = -1;
          M m{a.meld_type, a.n_tiles, {RV_NONE, RV_NONE, RV_NONE, RV_NONE}, -1, RV_NONE};
          for (int i = 0; i < a.n_tiles; i++) {
            m.t[i] = a.tiles[i];
            if (a.froms[i] == s) match_and_remove(hand[s], a.tiles[i]);
            else if (m.from < 0) m.from = (int8_t)a.froms[i], m.called = a.tiles[i];   // the first tile of another seat
          }
          melds[s].push_back(m);
          if (a.meld_type == RV_MELD_DAIMINKAN) {
            rinshan[s] = true;
            flush_pending();
            pending_minkan++;
          }
          break;
        }
        case RV_LA_DORA:
          if (ky.wall.empty()) {
            doras.push_back(a.tile);
          } else {
            dora_count++;
            if (pending_minkan > 0) pending_minkan--;
            sync_doras();
          }
          break;
        case RV_LA_ANGANG_ADDGANG:
          all(rinshan, false);
          flush_pending();
          if (a.meld_type == RV_MELD_ANKAN) {
            all(ippatsu, false), all(first_turn, false);
            last_kakan = last_babei = false;
            kakan_tile = -1;
            const int k34 = x.tile_raw_id;
            for (int i = 0; i < 4; i++) {
              auto it = std::find_if(hand[s].begin(), hand[s].end(), [&](uint8_t t) { return t / 4 == k34; });
              if (it != hand[s].end()) hand[s].erase(it);
            }
            M m{RV_MELD_ANKAN, 4, {(uint8_t)(k34 * 4), (uint8_t)(k34 * 4 + 1), (uint8_t)(k34 * 4 + 2), (uint8_t)(k34 * 4 + 3)}, -1, RV_NONE};
            melds[s].push_back(m);
            rinshan[s] = true;
            if (!ky.wall.empty()) dora_count++;
          } else {
            last_kakan = true;
            kakan_tile = a.tiles[0];
            rinshan[s] = true;
            bool upgraded = false;
            for (auto& m : melds[s])
              if (m.type == RV_MELD_PON && m.t[0] / 4 == a.tiles[0] / 4) {
                m.type = RV_MELD_KAKAN;
                if (m.n < 4) m.t[m.n++] = a.tiles[0];
                upgraded = true;
                break;
              }
            if (!upgraded) {
              M m{a.meld_type, a.n_tiles, {RV_NONE, RV_NONE, RV_NONE, RV_NONE}, -1, RV_NONE};
              for (int i = 0; i < a.n_tiles; i++) m.t[i] = a.tiles[i];
              melds[s].push_back(m);
            }
            match_and_remove(hand[s], a.tiles[0]);
            pending_minkan++;
          }
          sync_doras();
          break;
        case RV_LA_BABEI: {
          memcpy(ippatsu_before_babei, ippatsu, sizeof ippatsu);
          all(ippatsu, false), all(first_turn, false);
          last_babei = true;
          auto it = std::find_if(hand[s].begin(), hand[s].end(), [](uint8_t t) { return t / 4 == 30; });
          if (it != hand[s].end()) hand[s].erase(it);
          kita[s]++;
          rinshan[s] = true;
          break;
        }
        case RV_LA_HULE:
          for (int hi = 0; hi < a.n_hule; hi++) {
            const rv_hule& h = a.hules[hi];
            const int w = h.seat;
            if (w >= np) return err = "hule of a seat the kyoku does not have", false;
            const bool zimo = h.zimo;
            const bool chankan = !zimo && last_kakan && kakan_tile >= 0 && kakan_tile / 4 == h.hu_tile / 4;
            const bool ipp = !zimo && last_babei ? ippatsu_before_babei[w] : ippatsu[w];
            std::vector<uint8_t> tiles = hand[w];
            if (!zimo) tiles.push_back(h.hu_tile);
            std::vector<uint8_t> ura;
            if (liqi[w]) {
              if (h.n_li_doras != 0xFF) ura.assign(h.li_doras, h.li_doras + h.n_li_doras);
              else if (!ky.wall.empty()) ura = ura_from_wall();
              else ura.assign(ky.k.ura_doras, ky.k.ura_doras + ky.k.n_ura_doras);
            }
            if (tiles.size() > 14 || melds[w].size() > 4)
              return err = "a hand of more than 14 tiles or 4 melds at a hule", false;
            rv_win_context c;
            memset(&c, 0, sizeof c);
            rv_hand_query& q = c.query;
            memset(q.tiles, RV_NONE, sizeof q.tiles);
            memset(q.meld_tiles, RV_NONE, sizeof q.meld_tiles);
            q.n_tiles = (uint8_t)tiles.size();
            std::copy(tiles.begin(), tiles.end(), q.tiles);
            q.n_melds = (uint8_t)melds[w].size();
            for (size_t m = 0; m < melds[w].size(); m++) {
              q.meld_type[m] = melds[w][m].type;
              memcpy(q.meld_tiles[m], melds[w][m].t, 4);
              c.meld_from[m] = melds[w][m].from;
              c.meld_called[m] = melds[w][m].called;
            }
            q.win_tile = h.hu_tile;
            q.n_dora = (uint8_t)std::min<size_t>(doras.size(), 5);
            std::copy(doras.begin(), doras.begin() + q.n_dora, q.dora_ind);
            q.n_ura = (uint8_t)std::min<size_t>(ura.size(), 5);
            std::copy(ura.begin(), ura.begin() + q.n_ura, q.ura_ind);
            q.player_wind = (uint8_t)((w + np - ky.k.ju % np) % np);
            q.round_wind = (uint8_t)(ky.k.chang & 3);
            q.honba = 0;                                    // "Not tracked" (2008-2009)
            q.cond = (uint16_t)((zimo ? RV_C_TSUMO : 0) | (liqi[w] ? RV_C_RIICHI : 0) | (wliqi[w] ? RV_C_DOUBLE_RIICHI : 0) |
                                (ipp ? RV_C_IPPATSU : 0) | (left == 0 && zimo && !rinshan[w] ? RV_C_HAITEI : 0) |
                                (left == 0 && !zimo && !rinshan[w] ? RV_C_HOUTEI : 0) | (rinshan[w] ? RV_C_RINSHAN : 0) |
                                (chankan ? RV_C_CHANKAN : 0) | (first_turn[w] && zimo ? RV_C_TSUMO_FIRST_TURN : 0));
            q.sanma = 0;                                    // the iterator always builds the 4-player HandEvaluator (2035)
            q.kita_count = (uint8_t)kita[w];
            c.expected_yaku = h.fans;
            c.expected_han = h.count;
            c.expected_fu = h.fu;
            c.round = round;
            c.action = (int32_t)ai;
            c.seat = (uint8_t)w;
            out.push_back(c);
          }
          break;
        default: break;
      }
    }
    return true;
  }
};
}  // namespace

// run f(lo, hi) over [0, n) on a few host threads (the copies and the label pass of a large batch are memory-bound loops)
template <class F>
static void parallel_ranges(int64_t n, int64_t grain, F f) {
  int nt = (int)std::min<int64_t>(std::min<unsigned>(std::thread::hardware_concurrency(), 8u), (n + grain - 1) / std::max<int64_t>(grain, 1));
  if (nt <= 1) return f((int64_t)0, n);
  std::vector<std::thread> pool;
  const int64_t step = (n + nt - 1) / nt;
  for (int t = 1; t < nt; t++) pool.emplace_back(f, std::min(n, t * step), std::min(n, (t + 1) * step));
  f((int64_t)0, std::min(n, step));
  for (auto& t : pool) t.join();
}
extern "C" {
int rv_replay_from_text(const char* text, size_t len, uint32_t rule_bits, rv_replay** out) {
  if (!text || !out) return rv_internal_fail(RV_ERR_INVALID, "text / out is null");
  return parse_lines(text, len, rule_bits, out);
}
int rv_replay_from_jsonl(const char* path, uint32_t rule_bits, rv_replay** out) {
  if (!path || !out) return rv_internal_fail(RV_ERR_INVALID, "path / out is null");
  // gzopen reads plain files transparently and inflates those that start with the gzip magic bytes (0x1f 0x8b) —
  // detection by content, not by extension, as the reference does
  gzFile f = gzopen(path, "rb");
  if (!f) return rv_internal_fail(RV_ERR_INVALID, std::string("Failed to open file: ") + path);
  std::string text;
  char buf[1 << 16];
  int n;
  while ((n = gzread(f, buf, sizeof buf)) > 0) text.append(buf, (size_t)n);
  const bool bad = n < 0;
  gzclose(f);
  if (bad) return rv_internal_fail(RV_ERR_INVALID, std::string("Read error: ") + path);
  return parse_lines(text.data(), text.size(), rule_bits, out);
}
int rv_replay_from_mjsoul_text(const char* text, size_t len, uint32_t rule_bits, rv_replay** out) {
  if (!text || !out) return rv_internal_fail(RV_ERR_INVALID, "text / out is null");
  return parse_paifu(text, len, rule_bits, out);
}
int rv_replay_from_mjsoul_json(const char* path, uint32_t rule_bits, rv_replay** out) {
  if (!path || !out) return rv_internal_fail(RV_ERR_INVALID, "path / out is null");
  gzFile f = gzopen(path, "rb");                           // the reference expects gzip; gzopen also reads a plain file
  if (!f) return rv_internal_fail(RV_ERR_INVALID, std::string("Failed to open file: ") + path);
  std::string text;
  char buf[1 << 16];
  int n;
  while ((n = gzread(f, buf, sizeof buf)) > 0) text.append(buf, (size_t)n);
  const bool bad = n < 0;
  gzclose(f);
  if (bad) return rv_internal_fail(RV_ERR_INVALID, std::string("Failed to decompress: ") + path);
  return parse_paifu(text.data(), text.size(), rule_bits, out);
}
// ---- bulk loading for a data loader: many files parsed by a pool of host threads into ONE replay, then flat arrays ----------
int rv_replay_from_files(const char* const* paths, int n_paths, int format, uint32_t rule_bits, int threads, rv_replay** out,
                         int* n_failed) {
  if (!paths || n_paths < 0 || !out || (format != 0 && format != 1))
    return rv_internal_fail(RV_ERR_INVALID, "rv_replay_from_files: bad arguments");
  std::vector<rv_replay*> parts((size_t)n_paths, nullptr);
  std::atomic<int> next{0}, failed{0};
  auto work = [&]() {
    for (int i; (i = next.fetch_add(1)) < n_paths;) {
      rv_replay* r = nullptr;
      const int rc = !paths[i] ? RV_ERR_INVALID
                   : format == 0 ? rv_replay_from_jsonl(paths[i], rule_bits, &r) : rv_replay_from_mjsoul_json(paths[i], rule_bits, &r);
      if (rc == RV_OK) parts[(size_t)i] = r;
      else failed.fetch_add(1);                            // a file that does not parse is skipped, as the reference's datasets do
    }
  };
  int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  nt = std::max(1, std::min(nt, std::max(1, n_paths)));
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; t++) pool.emplace_back(work);
  work();
  for (auto& t : pool) t.join();
  std::unique_ptr<rv_replay> all(new rv_replay);
  for (rv_replay* r : parts) {
    if (!r) continue;
    for (auto& k : r->rounds) all->rounds.push_back(std::move(k));
    delete r;
  }
  if (n_failed) *n_failed = failed.load();
  *out = all.release();
  return RV_OK;
}
int rv_replay_totals(const rv_replay* r, int np, int64_t* n_rounds, int64_t* n_actions) {
  if (!r || (np != 3 && np != 4)) return rv_internal_fail(RV_ERR_INVALID, "rv_replay_totals: bad arguments");
  int64_t nr = 0, na = 0;
  for (auto& k : r->rounds)
    if (k.k.np == np) nr++, na += (int64_t)k.actions.size();
  if (n_rounds) *n_rounds = nr;
  if (n_actions) *n_actions = na;
  return RV_OK;
}
int rv_replay_flatten(const rv_replay* r, int np, rv_log_kyoku* kyokus, rv_log_action* actions, int64_t* first, int32_t* round_index) {
  if (!r || (np != 3 && np != 4) || !kyokus || !actions || !first)
    return rv_internal_fail(RV_ERR_INVALID, "rv_replay_flatten: bad arguments");
  std::vector<const Kyoku*> sel;
  int64_t a = 0;
  first[0] = 0;
  for (size_t k = 0; k < r->rounds.size(); k++) {
    const Kyoku& ky = r->rounds[k];
    if (ky.k.np != np) continue;
    if (round_index) round_index[sel.size()] = (int32_t)k;
    sel.push_back(&ky);
    a += (int64_t)ky.actions.size();
    first[sel.size()] = a;
  }
  parallel_ranges((int64_t)sel.size(), 256, [&](int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; i++) {
      kyokus[i] = sel[(size_t)i]->k;
      const auto& av = sel[(size_t)i]->actions;
      if (!av.empty()) memcpy(actions + first[i], av.data(), av.size() * sizeof(rv_log_action));
    }
  });
  return RV_OK;
}
int rv_replay_own_turn_labels(const rv_log_action* actions, int64_t n, int np, int16_t* seat, int16_t* action_id) {
  if (n < 0 || (n > 0 && (!actions || !seat || !action_id)) || (np != 3 && np != 4))
    return rv_internal_fail(RV_ERR_INVALID, "rv_replay_own_turn_labels: bad arguments");
  // Action::encode (action.rs:158-227): discard kind, 37 riichi, 42 + kind ankan / kakan, 79 tsumo; sanma (action_3p.rs) over
  // the 27 kinds it has: kind 0 -> 0, kinds 8.. -> kind - 7; 27 riichi, 29 + compact kan, 56 tsumo, 59 kita
  auto compact = [](int kind) { return kind == 0 ? 0 : kind >= 8 ? kind - 7 : -1; };
  parallel_ranges(n, 1 << 16, [&](int64_t lo, int64_t hi) {
  for (int64_t i = lo; i < hi; i++) {
    const rv_log_action& a = actions[i];
    int s = a.seat, id = -1;
    if (a.type == RV_LA_DISCARD) {
      if (a.flags & 1) id = np == 3 ? 27 : 37;
      else if (a.tile < 136) id = np == 3 ? compact(a.tile / 4) : a.tile / 4;
    } else if (a.type == RV_LA_ANGANG_ADDGANG && a.n_tiles > 0 && a.tiles[0] < 136) {
      const int c = np == 3 ? compact(a.tiles[0] / 4) : a.tiles[0] / 4;
      id = c < 0 ? -1 : (np == 3 ? 29 : 42) + c;
    } else if (a.type == RV_LA_HULE && a.n_hule > 0 && a.hules[0].zimo) {
      s = a.hules[0].seat;
      id = np == 3 ? 56 : 79;
    } else if (a.type == RV_LA_BABEI && np == 3) {
      id = 59;
    }
    seat[i] = (int16_t)(id >= 0 ? s : -1);
    action_id[i] = (int16_t)id;
  }
  });
  return RV_OK;
}
int rv_replay_free(rv_replay* r) {
  delete r;
  return RV_OK;
}
int rv_replay_num_rounds(const rv_replay* r) { return r ? (int)r->rounds.size() : 0; }
int rv_replay_kyoku(const rv_replay* r, int round, rv_log_kyoku* out) {
  if (!r || !out || round < 0 || round >= (int)r->rounds.size()) return rv_internal_fail(RV_ERR_INVALID, "round out of range");
  *out = r->rounds[round].k;
  return RV_OK;
}
int rv_replay_actions(const rv_replay* r, int round, rv_log_action* out, int cap, int* n_out) {
  if (!r || round < 0 || round >= (int)r->rounds.size()) return rv_internal_fail(RV_ERR_INVALID, "round out of range");
  const auto& a = r->rounds[round].actions;
  if (n_out) *n_out = (int)a.size();
  if (out)
    for (int i = 0; i < cap && i < (int)a.size(); i++) out[i] = a[i];
  return RV_OK;
}
int rv_replay_sizeof(int which) { return which == 0 ? (int)sizeof(rv_win_context) : which == 1 ? (int)sizeof(rv_log_action_aux) : -1; }
int rv_replay_actions_aux(const rv_replay* r, int round, rv_log_action_aux* out, int cap, int* n_out) {
  if (!r || round < 0 || round >= (int)r->rounds.size() || cap < 0 || (cap > 0 && !out))
    return rv_internal_fail(RV_ERR_INVALID, "rv_replay_actions_aux: bad arguments");
  const Kyoku& k = r->rounds[round];
  static const ActAux none;
  for (int i = 0; i < (int)k.actions.size() && i < cap; i++) out[i] = i < (int)k.aux.size() ? k.aux[i] : none;
  if (n_out) *n_out = (int)k.actions.size();
  return RV_OK;
}
int rv_replay_paishan(const rv_replay* r, int round, char* out, int cap, int* n_out) {
  if (!r || round < 0 || round >= (int)r->rounds.size() || !n_out || cap < 0 || (cap > 0 && !out))
    return rv_internal_fail(RV_ERR_INVALID, "rv_replay_paishan: bad arguments");
  const Kyoku& k = r->rounds[round];
  *n_out = k.has_paishan ? (int)k.paishan.size() : -1;
  if (k.has_paishan && cap > 0) memcpy(out, k.paishan.data(), std::min<size_t>(cap, k.paishan.size()));
  return RV_OK;
}
int rv_replay_win_contexts(const rv_replay* r, int round, rv_win_context* out, int cap, int* n_out) {
  if (!r || round >= (int)r->rounds.size() || cap < 0 || (cap > 0 && !out))
    return rv_internal_fail(RV_ERR_INVALID, "rv_replay_win_contexts: bad arguments");
  std::vector<rv_win_context> ctxs;
  std::string err;
  const int lo = round < 0 ? 0 : round, hi = round < 0 ? (int)r->rounds.size() : round + 1;
  for (int i = lo; i < hi; i++) {
    WinWalk w(r->rounds[i]);
    if (!w.run(i, ctxs, err)) return rv_internal_fail(RV_ERR_INVALID, "rv_replay_win_contexts: round " + std::to_string(i) + ": " + err);
  }
  for (size_t i = 0; i < ctxs.size() && (int)i < cap; i++) out[i] = ctxs[i];
  if (n_out) *n_out = (int)ctxs.size();
  return RV_OK;
}
}  // extern "C"
