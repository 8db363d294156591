// Host-side rendering of the binary event stream as MJAI JSON.
// Replaces the event builders scattered through state/mod.rs (e.g. 1785-1820, 1380-1387,
// 2021-2046) plus _push_mjai_event's per-player masking (state/mod.rs:2094-2148) and
// parser::tid_to_mjai (parser.rs:301-334).  serde_json without `preserve_order` emits
// object keys alphabetically and without spaces; this file reproduces that byte layout.
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/riichienv_b200.h"

static std::string tid_to_mjai(int tid) {
  if (tid == 16) return "5mr";
  if (tid == 52) return "5pr";
  if (tid == 88) return "5sr";
  if (tid < 108) {
    static const char suit[3] = {'m', 'p', 's'};
    char b[3] = {(char)('1' + (tid % 36) / 4), suit[tid / 36], 0};
    return b;
  }
  static const char* honors[7] = {"E", "S", "W", "N", "P", "F", "C"};
  int n = (tid - 108) / 4;
  return n < 7 ? honors[n] : "?";
}
static std::string q(const std::string& s) { return "\"" + s + "\""; }
static std::string ilist(const int32_t* v, int n) {
  std::string s = "[";
  for (int i = 0; i < n; i++) s += (i ? "," : "") + std::to_string(v[i]);
  return s + "]";
}
static const char* reason_str(int code, std::string& tmp) {
  switch (code) {
    case RV_RK_EXHAUSTIVE: return "exhaustive_draw";
    case RV_RK_NAGASHI: return "nagashimangan";
    case RV_RK_KYUSHU: return "kyushu_kyuhai";
    case RV_RK_SUFUURENTA: return "sufuurenta";
    case RV_RK_SUUKANSANSEN: return "suukansansen";
    case RV_RK_SUUCHA_RIICHI: return "suucha_riichi";
    case RV_RK_SANCHAHO: return "sanchaho";
  }
  tmp = "Error: Illegal Action by Player " + std::to_string(code - RV_RK_ILLEGAL_BASE);
  return tmp.c_str();
}

extern "C" int rv_event_to_json(const uint32_t* w, uint32_t n_words, int viewer, char* out, uint32_t out_cap) {
  if (!w || n_words == 0) return RV_ERR_INVALID;
  int type = w[0] & 0xFF, nw = (w[0] >> 8) & 0xFF, a = (w[0] >> 16) & 0xFF, b = (w[0] >> 24) & 0xFF;
  if (nw == 0 || (uint32_t)nw > n_words) return RV_ERR_INVALID;
  // the record length every event type is written with (ev_push callers in game.cuh): a stream that disagrees is refused,
  // not read past its end
  switch (type) {
    case RV_EV_START_KYOKU: if (nw != 19 && nw != 15) return RV_ERR_INVALID; break;
    case RV_EV_PON: case RV_EV_CHI: case RV_EV_DAIMINKAN: case RV_EV_ANKAN: case RV_EV_KAKAN: if (nw < 2) return RV_ERR_INVALID; break;
    case RV_EV_HORA: if (nw != 10 && nw != 9) return RV_ERR_INVALID; break;
    case RV_EV_RYUKYOKU: if (nw != 5 && nw != 4) return RV_ERR_INVALID; break;
    default: break;
  }
  std::string s, tmp;
  auto cons = [&](uint32_t word, int from, int to) {
    std::string r = "[";
    bool first = true;
    for (int k = from; k < to; k++) {
      int t = (word >> (8 * k)) & 0xFF;
      if (t == RV_NONE) continue;
      r += (first ? "" : ",") + q(tid_to_mjai(t));
      first = false;
    }
    return r + "]";
  };
  switch (type) {
    case RV_EV_START_GAME: s = "{\"type\":\"start_game\"}"; break;
    case RV_EV_END_KYOKU: s = "{\"type\":\"end_kyoku\"}"; break;
    case RV_EV_END_GAME: s = "{\"type\":\"end_game\"}"; break;
    case RV_EV_START_KYOKU: {
      static const char* winds[4] = {"E", "S", "W", "N"};
      int honba = w[1] & 0xFF, dora = (w[1] >> 8) & 0xFF, kyotaku = (w[1] >> 16) & 0xFFFF;
      int np = nw == 19 ? 4 : 3;   // 4P: 19 words, 3P: 15 words
      int32_t sc[4] = {0, 0, 0, 0};
      for (int p = 0; p < np; p++) sc[p] = (int32_t)w[2 + p];
      const uint8_t* th = (const uint8_t*)&w[2 + np];
      std::string tehais = "[";
      for (int p = 0; p < np; p++) {
        tehais += p ? ",[" : "[";
        bool first = true;
        for (int k = 0; k < 13; k++) {
          int t = th[p * 13 + k];
          if (t == RV_NONE) continue;
          tehais += (first ? "" : ",") + q((viewer >= 0 && viewer != p) ? "?" : tid_to_mjai(t));
          first = false;
        }
        tehais += "]";
      }
      tehais += "]";
      s = "{\"bakaze\":" + q(winds[a & 3]) + ",\"dora_marker\":" + q(tid_to_mjai(dora)) + ",\"honba\":" + std::to_string(honba) +
          ",\"kyoku\":" + std::to_string(b + 1) + ",\"kyotaku\":" + std::to_string(kyotaku) + ",\"oya\":" + std::to_string(b) +
          ",\"scores\":" + ilist(sc, np) + ",\"tehais\":" + tehais + ",\"type\":\"start_kyoku\"}";
      break;
    }
    case RV_EV_TSUMO:
      s = "{\"actor\":" + std::to_string(a) + ",\"pai\":" + q((viewer >= 0 && viewer != a) ? "?" : tid_to_mjai(b)) + ",\"type\":\"tsumo\"}";
      break;
    case RV_EV_DAHAI:
    case RV_EV_DAHAI_TSUMOGIRI:
      s = "{\"actor\":" + std::to_string(a) + ",\"pai\":" + q(tid_to_mjai(b)) + ",\"tsumogiri\":" +
          (type == RV_EV_DAHAI_TSUMOGIRI ? "true" : "false") + ",\"type\":\"dahai\"}";
      break;
    case RV_EV_REACH: s = "{\"actor\":" + std::to_string(a) + ",\"type\":\"reach\"}"; break;
    case RV_EV_REACH_ACCEPTED: s = "{\"actor\":" + std::to_string(a) + ",\"type\":\"reach_accepted\"}"; break;
    case RV_EV_PON:
    case RV_EV_CHI:
    case RV_EV_DAIMINKAN: {
      const char* nm = type == RV_EV_PON ? "pon" : type == RV_EV_CHI ? "chi" : "daiminkan";
      s = "{\"actor\":" + std::to_string(a) + ",\"consumed\":" + cons(w[1], 1, 4) + ",\"pai\":" + q(tid_to_mjai(b)) +
          ",\"target\":" + std::to_string(w[1] & 0xFF) + ",\"type\":" + q(nm) + "}";
      break;
    }
    case RV_EV_ANKAN:
    case RV_EV_KAKAN:
      s = "{\"actor\":" + std::to_string(a) + ",\"consumed\":" + cons(w[1], 0, 4) + ",\"pai\":" + q(tid_to_mjai(b)) +
          ",\"type\":" + q(type == RV_EV_ANKAN ? "ankan" : "kakan") + "}";
      break;
    case RV_EV_DORA: s = "{\"dora_marker\":" + q(tid_to_mjai(b)) + ",\"type\":\"dora\"}"; break;
    case RV_EV_KITA: s = "{\"actor\":" + std::to_string(a) + ",\"pai\":" + q(tid_to_mjai(b)) + ",\"type\":\"kita\"}"; break;
    case RV_EV_HORA: {
      bool tsumo = w[1] & 1;
      int n_ura = (w[1] >> 8) & 0xFF;
      const uint8_t* ub = (const uint8_t*)&w[2];
      int np = nw - 6;   // 4P: 10 words, 3P: 9 words
      int32_t d[4] = {0, 0, 0, 0};
      for (int p = 0; p < np && p < 4; p++) d[p] = (int32_t)w[4 + p];
      std::string ura = "[";
      for (int k = 0; k < n_ura && k < 5; k++) ura += (k ? "," : "") + q(tid_to_mjai(ub[k]));
      ura += "]";
      s = "{\"actor\":" + std::to_string(a) + ",\"deltas\":" + ilist(d, np) + ",\"target\":" + std::to_string(b) +
          (tsumo ? ",\"tsumo\":true" : "") + ",\"type\":\"hora\",\"ura_markers\":" + ura + "}";
      break;
    }
    case RV_EV_RYUKYOKU: {
      int np = nw - 1;
      int32_t d[4] = {0, 0, 0, 0};
      for (int p = 0; p < np && p < 4; p++) d[p] = (int32_t)w[1 + p];
      s = "{\"deltas\":" + ilist(d, np) + ",\"reason\":" + q(reason_str(a, tmp)) + ",\"type\":\"ryukyoku\"}";
      break;
    }
    default:
      return RV_ERR_INVALID;
  }
  if (out && out_cap) {
    size_t n = s.size() < out_cap - 1 ? s.size() : out_cap - 1;
    memcpy(out, s.data(), n);
    out[n] = 0;
  }
  return nw;
}
