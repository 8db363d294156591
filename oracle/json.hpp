// ORACLE — TEST INFRASTRUCTURE ONLY.
// MJAI text log of the oracle, written at event time the way the reference writes it: every event is a JSON object built
// key by key (state/mod.rs:158-162, 441-480, 571-581, 867-884, 1115-1131, 1195-1224, 1365-1368, 1498-1517, 1556-1563,
// 1674, 1785-1819, 1835-1838, 1957-1963, 2031-2036, 2075-2080; sanma state_3p/sanma.rs:47-54), serialised by serde_json
// whose Map (no `preserve_order` feature, Cargo.lock:849-859) is a BTreeMap — keys come out sorted, no spaces — and pushed
// to the global log and to four per-seat logs with other seats' start hands and draws masked (state/mod.rs:2094-2148).
// This is a second, independent renderer: the product renders its binary event words on the host (csrc/json.cpp);
// tests compare the two texts.  Tile names: parser.rs:301-334.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace orc {

inline std::string mjai_tile(int tid) {   // parser.rs:301-334
  switch (tid) {
    case 16: return "5mr";
    case 52: return "5pr";
    case 88: return "5sr";
  }
  const int kind = tid / 4;
  if (kind < 27) return std::string(1, (char)('1' + kind % 9)) + "mps"[kind / 9];
  static const char* z[7] = {"E", "S", "W", "N", "P", "F", "C"};
  return kind < 34 ? z[kind - 27] : "?";
}

struct JsonObj {   // serde_json::Map<String, Value> as a BTreeMap; values are kept serialised
  std::map<std::string, std::string> kv;
  static std::string quoted(const std::string& s) {
    std::string o = "\"";
    for (char c : s) {
      if (c == '"' || c == '\\') o += '\\';
      o += c;
    }
    return o + "\"";
  }
  JsonObj& str(const char* k, const std::string& v) { kv[k] = quoted(v); return *this; }
  JsonObj& num(const char* k, long long v) { kv[k] = std::to_string(v); return *this; }
  JsonObj& boolean(const char* k, bool v) { kv[k] = v ? "true" : "false"; return *this; }
  template <class It>
  JsonObj& nums(const char* k, It b, It e) {
    std::string s = "[";
    for (It i = b; i != e; ++i) s += (i == b ? "" : ",") + std::to_string((long long)*i);
    kv[k] = s + "]";
    return *this;
  }
  JsonObj& strs(const char* k, const std::vector<std::string>& v) {
    kv[k] = list(v);
    return *this;
  }
  static std::string list(const std::vector<std::string>& v) {
    std::string s = "[";
    for (size_t i = 0; i < v.size(); i++) s += (i ? "," : "") + quoted(v[i]);
    return s + "]";
  }
  std::string dump() const {
    std::string s = "{";
    bool first = true;
    for (auto& e : kv) {
      s += (first ? "" : ",") + quoted(e.first) + ":" + e.second;
      first = false;
    }
    return s + "}";
  }
};

struct MjaiLog {
  std::vector<std::string> all;
  std::vector<std::string> seat[4];
  int np = 4;
  void clear() {
    all.clear();
    for (auto& s : seat) s.clear();
  }
  // _push_mjai_event (state/mod.rs:2094-2148).  `tehais`: the start hands of a start_kyoku event (masked per viewer);
  // `draw_actor` >= 0: a tsumo event whose tile only that seat sees.
  void push(JsonObj ev, const std::vector<std::vector<std::string>>* tehais = nullptr, int draw_actor = -1) {
    if (tehais) {
      std::string t = "[";
      for (size_t i = 0; i < tehais->size(); i++) t += (i ? "," : "") + JsonObj::list((*tehais)[i]);
      ev.kv["tehais"] = t + "]";
    }
    all.push_back(ev.dump());
    for (int pid = 0; pid < np; pid++) {
      JsonObj m = ev;
      if (tehais) {
        std::string t = "[";
        for (size_t i = 0; i < tehais->size(); i++) {
          std::vector<std::string> h = (*tehais)[i];
          if ((int)i != pid)
            for (auto& x : h) x = "?";
          t += (i ? "," : "") + JsonObj::list(h);
        }
        m.kv["tehais"] = t + "]";
      } else if (draw_actor >= 0 && draw_actor != pid) {
        m.str("pai", "?");
      }
      seat[pid].push_back(m.dump());
    }
  }
};

}  // namespace orc
