// json_config.cc -- config.json -> RenderConfig, the surface of LoadJSONConfig (main.cc:98-205).
//
// Same keys, same type rules, unknown keys ignored:
//   obj_filename, eson_filename, magicavoxel_filename, material_filename   string
//   scene_scale                                                             number
//   scene_fit, plane                                                        boolean
//   eye, up, lookat                                                         array of exactly 3 (non-numbers read as 0)
//   resolution                                                              array of exactly 2 -> width, height (truncated)
//   num_passes                                                              number
//   num_photons                                                             number, OVERWRITES num_passes (main.cc:192-195)
// `fov` is not a key of the reference (always 45, render.h:34).  Keys this implementation adds -- all
// optional, all defaulting to the reference's behaviour -- are listed in mallie_api.h (RenderConfig).
// The reference parses with parson (deps/parson); this is a small recursive-descent JSON reader with
// the same acceptance for config files: a top-level object is required, duplicate keys are an error.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "mallie_api.h"

namespace {

struct JValue;
typedef std::shared_ptr<JValue> JRef;

struct JValue {
  enum Kind { kNull, kBool, kNumber, kString, kArray, kObject } kind = kNull;
  bool b = false;
  double num = 0.0;
  std::string str;
  std::vector<JRef> items;
  std::vector<std::pair<std::string, JRef>> members;

  const JValue *get(const char *name) const {
    if (kind != kObject) return nullptr;
    for (const auto &m : members)
      if (m.first == name) return m.second.get();
    return nullptr;
  }
};

class JParser {
public:
  explicit JParser(const std::string &text) : s_(text), i_(0) {}
  JRef parse_document() {
    JRef v = value(0);
    if (!v) return nullptr;
    skip();
    return v;
  }

private:
  const std::string &s_;
  size_t i_;

  void skip() {
    while (i_ < s_.size() && (s_[i_] == ' ' || s_[i_] == '\t' || s_[i_] == '\n' || s_[i_] == '\r')) i_++;
  }
  bool literal(const char *w) {
    const size_t n = strlen(w);
    if (s_.compare(i_, n, w) != 0) return false;
    i_ += n;
    return true;
  }
  static void append_utf8(std::string &out, unsigned cp) {
    if (cp < 0x80) {
      out += (char)cp;
    } else if (cp < 0x800) {
      out += (char)(0xC0 | (cp >> 6));
      out += (char)(0x80 | (cp & 0x3F));
    } else {
      out += (char)(0xE0 | (cp >> 12));
      out += (char)(0x80 | ((cp >> 6) & 0x3F));
      out += (char)(0x80 | (cp & 0x3F));
    }
  }
  bool string(std::string &out) {
    if (i_ >= s_.size() || s_[i_] != '"') return false;
    i_++;
    while (i_ < s_.size() && s_[i_] != '"') {
      char c = s_[i_++];
      if (c == '\\') {
        if (i_ >= s_.size()) return false;
        const char e = s_[i_++];
        switch (e) {
        case '"': out += '"'; break;
        case '\\': out += '\\'; break;
        case '/': out += '/'; break;
        case 'b': out += '\b'; break;
        case 'f': out += '\f'; break;
        case 'n': out += '\n'; break;
        case 'r': out += '\r'; break;
        case 't': out += '\t'; break;
        case 'u': {
          if (i_ + 4 > s_.size()) return false;
          unsigned cp = 0;
          for (int k = 0; k < 4; k++) {
            const char h = s_[i_++];
            cp <<= 4;
            if (h >= '0' && h <= '9') cp |= (unsigned)(h - '0');
            else if (h >= 'a' && h <= 'f') cp |= (unsigned)(h - 'a' + 10);
            else if (h >= 'A' && h <= 'F') cp |= (unsigned)(h - 'A' + 10);
            else return false;
          }
          append_utf8(out, cp);
          break;
        }
        default: return false;
        }
      } else {
        out += c;
      }
    }
    if (i_ >= s_.size()) return false;
    i_++; // closing quote
    return true;
  }
  JRef value(int depth) {
    if (depth > 64) return nullptr;
    skip();
    if (i_ >= s_.size()) return nullptr;
    JRef v = std::make_shared<JValue>();
    const char c = s_[i_];
    if (c == '{') {
      i_++;
      v->kind = JValue::kObject;
      skip();
      if (i_ < s_.size() && s_[i_] == '}') {
        i_++;
        return v;
      }
      for (;;) {
        skip();
        std::string key;
        if (!string(key)) return nullptr;
        skip();
        if (i_ >= s_.size() || s_[i_] != ':') return nullptr;
        i_++;
        JRef item = value(depth + 1);
        if (!item) return nullptr;
        if (v->get(key.c_str())) return nullptr; // duplicate key
        v->members.emplace_back(key, item);
        skip();
        if (i_ >= s_.size()) return nullptr;
        if (s_[i_] == ',') {
          i_++;
          continue;
        }
        if (s_[i_] == '}') {
          i_++;
          return v;
        }
        return nullptr;
      }
    }
    if (c == '[') {
      i_++;
      v->kind = JValue::kArray;
      skip();
      if (i_ < s_.size() && s_[i_] == ']') {
        i_++;
        return v;
      }
      for (;;) {
        JRef item = value(depth + 1);
        if (!item) return nullptr;
        v->items.push_back(item);
        skip();
        if (i_ >= s_.size()) return nullptr;
        if (s_[i_] == ',') {
          i_++;
          continue;
        }
        if (s_[i_] == ']') {
          i_++;
          return v;
        }
        return nullptr;
      }
    }
    if (c == '"') {
      v->kind = JValue::kString;
      return string(v->str) ? v : nullptr;
    }
    if (c == 't') {
      v->kind = JValue::kBool, v->b = true;
      return literal("true") ? v : nullptr;
    }
    if (c == 'f') {
      v->kind = JValue::kBool, v->b = false;
      return literal("false") ? v : nullptr;
    }
    if (c == 'n') return literal("null") ? v : nullptr;
    if (c == '-' || (c >= '0' && c <= '9')) {
      char *end = nullptr;
      v->kind = JValue::kNumber;
      v->num = strtod(s_.c_str() + i_, &end);
      if (end == s_.c_str() + i_) return nullptr;
      i_ = (size_t)(end - s_.c_str());
      return v;
    }
    return nullptr;
  }
};

// `~` / `~/...` -> $HOME (the reference expands paths with wordexp, filepath_util.cc); other paths as is.
std::string expand_path(const std::string &p) {
  if (!p.empty() && p[0] == '~' && (p.size() == 1 || p[1] == '/')) {
    const char *home = getenv("HOME");
    if (home) return std::string(home) + p.substr(1);
  }
  return p;
}

inline double number_or_zero(const JValue &arr, size_t i) {
  return (i < arr.items.size() && arr.items[i]->kind == JValue::kNumber) ? arr.items[i]->num : 0.0;
}

void read_vec3(const JValue &root, const char *key, double out[3]) {
  const JValue *a = root.get(key);
  if (a && a->kind == JValue::kArray && a->items.size() == 3)
    for (int k = 0; k < 3; k++) out[k] = number_or_zero(*a, (size_t)k);
}

void read_string(const JValue &root, const char *key, std::string &out) {
  const JValue *v = root.get(key);
  if (v && v->kind == JValue::kString) out = expand_path(v->str);
}

template <class T> void read_number(const JValue &root, const char *key, T &out) {
  const JValue *v = root.get(key);
  if (v && v->kind == JValue::kNumber) out = (T)v->num;
}

void read_bool(const JValue &root, const char *key, bool &out) {
  const JValue *v = root.get(key);
  if (v && v->kind == JValue::kBool) out = v->b;
}

} // namespace

namespace mallie {

bool LoadJSONConfigFromString(RenderConfig &config, const std::string &text) {
  JParser parser(text);
  const JRef root = parser.parse_document();
  if (!root || root->kind != JValue::kObject) return false;
  const JValue &o = *root;

  read_string(o, "obj_filename", config.obj_filename);
  read_string(o, "eson_filename", config.eson_filename);
  read_string(o, "magicavoxel_filename", config.magicavoxel_filename);
  read_string(o, "material_filename", config.material_filename);
  read_number(o, "scene_scale", config.scene_scale);
  read_bool(o, "scene_fit", config.scene_fit);
  read_vec3(o, "eye", config.eye);
  read_vec3(o, "up", config.up);
  read_vec3(o, "lookat", config.lookat);
  const JValue *res = o.get("resolution");
  if (res && res->kind == JValue::kArray && res->items.size() == 2) {
    config.width = (int)number_or_zero(*res, 0);
    config.height = (int)number_or_zero(*res, 1);
  }
  read_number(o, "num_passes", config.num_passes);
  read_number(o, "num_photons", config.num_passes); // sic: main.cc:192-195
  read_bool(o, "plane", config.plane);

  // ---- additions (absent => reference behaviour) ----
  read_number(o, "max_path_length", config.max_path_length);
  read_vec3(o, "light", config.light);
  read_number(o, "device", config.device);
  read_number(o, "gpus", config.num_gpus);
  const JValue *sh = o.get("shader");
  if (sh && sh->kind == JValue::kString) {
    if (sh->str == "pathtrace") config.shader = MB200_SHADER_PATHTRACE;
    else if (sh->str == "primary_shadow") config.shader = MB200_SHADER_PRIMARY_SHADOW;
    else if (sh->str == "primary") config.shader = MB200_SHADER_PRIMARY_ONLY;
  }
  return true;
}

bool LoadJSONConfig(RenderConfig &config, const std::string &filename) {
  std::ifstream is(filename.c_str());
  if (!is) {
    std::cerr << "File not found: " << filename << std::endl;
    return false;
  }
  std::stringstream ss;
  ss << is.rdbuf();
  return LoadJSONConfigFromString(config, ss.str());
}

} // namespace mallie
