// json_min.hpp -- a small JSON reader, enough for the reference's config schema
// (config/MPC_API.json: numbers, nested arrays, strings; parsed by nlohmann/json in the
// reference, src/ModelPredictiveControlAPI.cpp:12-13).  Throws std::runtime_error on malformed
// input, mirroring the reference's "throws on a bad config" behaviour (cpp:13,437-480).
#pragma once
#include <cctype>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace smpc {

struct JsonValue {
  enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
  double num = 0.0;
  bool boolean = false;
  std::string str;
  std::vector<JsonValue> arr;
  std::map<std::string, JsonValue> obj;

  bool has(const std::string &k) const { return kind == Object && obj.count(k); }
  const JsonValue &at(const std::string &k) const {
    auto it = obj.find(k);
    if (kind != Object || it == obj.end()) throw std::runtime_error("missing key '" + k + "'");
    return it->second;
  }
  // flattens a number / vector / matrix into row-major doubles; reports its shape
  void flatten(std::vector<double> &out, int &rows, int &cols) const {
    out.clear();
    if (kind == Number) { out.push_back(num); rows = cols = 1; return; }
    if (kind != Array || arr.empty()) throw std::runtime_error("expected a number or a non-empty array");
    if (arr[0].kind == Array) {
      rows = (int)arr.size(); cols = (int)arr[0].arr.size();
      for (const auto &r : arr) {
        if (r.kind != Array || (int)r.arr.size() != cols) throw std::runtime_error("inconsistent matrix row length");
        for (const auto &v : r.arr) { if (v.kind != Number) throw std::runtime_error("non-numeric matrix entry"); out.push_back(v.num); }
      }
    } else {
      rows = 1; cols = (int)arr.size();
      for (const auto &v : arr) { if (v.kind != Number) throw std::runtime_error("non-numeric vector entry"); out.push_back(v.num); }
    }
  }
};

class JsonParser {
 public:
  explicit JsonParser(const std::string &s) : s_(s) {}
  JsonValue parse() { JsonValue v = value(); ws(); if (i_ != s_.size()) fail("trailing characters"); return v; }

 private:
  const std::string &s_;
  size_t i_ = 0;
  [[noreturn]] void fail(const std::string &m) const { throw std::runtime_error("JSON: " + m + " at offset " + std::to_string(i_)); }
  void ws() { while (i_ < s_.size() && std::isspace((unsigned char)s_[i_])) ++i_; }
  char peek() { ws(); if (i_ >= s_.size()) fail("unexpected end"); return s_[i_]; }
  void expect(char c) { if (peek() != c) fail(std::string("expected '") + c + "'"); ++i_; }
  JsonValue value() {
    char c = peek();
    JsonValue v;
    if (c == '{') {
      v.kind = JsonValue::Object; ++i_;
      if (peek() == '}') { ++i_; return v; }
      for (;;) {
        if (peek() != '"') fail("expected a string key");
        std::string k = string();
        expect(':');
        v.obj[k] = value();
        if (peek() == ',') { ++i_; continue; }
        expect('}'); return v;
      }
    }
    if (c == '[') {
      v.kind = JsonValue::Array; ++i_;
      if (peek() == ']') { ++i_; return v; }
      for (;;) {
        v.arr.push_back(value());
        if (peek() == ',') { ++i_; continue; }
        expect(']'); return v;
      }
    }
    if (c == '"') { v.kind = JsonValue::String; v.str = string(); return v; }
    if (s_.compare(i_, 4, "true") == 0) { i_ += 4; v.kind = JsonValue::Bool; v.boolean = true; return v; }
    if (s_.compare(i_, 5, "false") == 0) { i_ += 5; v.kind = JsonValue::Bool; return v; }
    if (s_.compare(i_, 4, "null") == 0) { i_ += 4; return v; }
    const char *b = s_.c_str() + i_; char *e = nullptr;
    double d = std::strtod(b, &e);
    if (e == b) fail("unexpected character");
    i_ += (size_t)(e - b);
    v.kind = JsonValue::Number; v.num = d; return v;
  }
  std::string string() {
    expect('"');
    std::string out;
    while (i_ < s_.size() && s_[i_] != '"') {
      if (s_[i_] == '\\' && i_ + 1 < s_.size()) {
        char e = s_[i_ + 1];
        out += (e == 'n' ? '\n' : e == 't' ? '\t' : e);
        i_ += 2;
      } else out += s_[i_++];
    }
    if (i_ >= s_.size()) fail("unterminated string");
    ++i_;
    return out;
  }
};

}  // namespace smpc
