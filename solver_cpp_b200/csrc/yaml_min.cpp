// Minimal YAML reader (see yaml_min.h).  Line-oriented recursive descent for block structure, a small character parser
// for flow collections ([..], {..}), which may span lines.
#include "yaml_min.h"

#include <cctype>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace {

struct Line { int indent; std::string text; int no; };

[[noreturn]] void fail(int line, const std::string& what) {
  throw std::runtime_error("YAML line " + std::to_string(line) + ": " + what);
}

std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && std::isspace((unsigned char)s[a])) a++;
  while (b > a && std::isspace((unsigned char)s[b - 1])) b--;
  return s.substr(a, b - a);
}

// removes a trailing comment (a '#' at the line start or after white space, outside quotes)
std::string strip_comment(const std::string& s) {
  char q = 0;
  for (size_t i = 0; i < s.size(); i++) {
    const char ch = s[i];
    if (q) { if (ch == q) q = 0; continue; }
    if (ch == '"' || ch == '\'') { q = ch; continue; }
    if (ch == '#' && (i == 0 || std::isspace((unsigned char)s[i - 1]))) return s.substr(0, i);
  }
  return s;
}

std::string unquote(const std::string& s) {
  if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\''))) return s.substr(1, s.size() - 2);
  return s;
}

std::shared_ptr<YamlNode> make_scalar(const std::string& raw) {
  auto n = std::make_shared<YamlNode>();
  const std::string t = trim(raw);
  if (t.empty() || t == "~" || t == "null") { n->kind = YamlNode::Null; return n; }
  n->kind = YamlNode::Scalar;
  n->scalar = unquote(t);
  return n;
}

// position of the ':' that ends a mapping key (followed by white space or the end), outside quotes and brackets; npos if none
size_t key_colon(const std::string& s) {
  char q = 0; int depth = 0;
  for (size_t i = 0; i < s.size(); i++) {
    const char ch = s[i];
    if (q) { if (ch == q) q = 0; continue; }
    if (ch == '"' || ch == '\'') { q = ch; continue; }
    if (ch == '[' || ch == '{') depth++;
    if (ch == ']' || ch == '}') depth--;
    if (ch == ':' && depth == 0 && (i + 1 == s.size() || std::isspace((unsigned char)s[i + 1]))) return i;
  }
  return std::string::npos;
}

int bracket_balance(const std::string& s) {
  char q = 0; int depth = 0;
  for (char ch : s) {
    if (q) { if (ch == q) q = 0; continue; }
    if (ch == '"' || ch == '\'') { q = ch; continue; }
    if (ch == '[' || ch == '{') depth++;
    if (ch == ']' || ch == '}') depth--;
  }
  return depth;
}

// ---- flow collections
struct Flow {
  const std::string& s; size_t p; int line;
  void ws() { while (p < s.size() && std::isspace((unsigned char)s[p])) p++; }
  std::shared_ptr<YamlNode> value() {
    ws();
    if (p >= s.size()) fail(line, "unexpected end of flow collection");
    if (s[p] == '[') return seq();
    if (s[p] == '{') return map();
    const size_t b = p;
    char q = 0;
    while (p < s.size()) {
      const char ch = s[p];
      if (q) { if (ch == q) q = 0; p++; continue; }
      if (ch == '"' || ch == '\'') { q = ch; p++; continue; }
      if (ch == ',' || ch == ']' || ch == '}') break;
      p++;
    }
    return make_scalar(s.substr(b, p - b));
  }
  std::shared_ptr<YamlNode> seq() {
    auto n = std::make_shared<YamlNode>(); n->kind = YamlNode::Seq;
    p++;                                                      // '['
    for (;;) {
      ws();
      if (p >= s.size()) fail(line, "unterminated [");
      if (s[p] == ']') { p++; break; }
      n->seq.push_back(value());
      ws();
      if (p < s.size() && s[p] == ',') { p++; continue; }
      if (p < s.size() && s[p] == ']') { p++; break; }
      fail(line, "expected , or ] in flow sequence");
    }
    return n;
  }
  std::shared_ptr<YamlNode> map() {
    auto n = std::make_shared<YamlNode>(); n->kind = YamlNode::Map;
    p++;                                                      // '{'
    for (;;) {
      ws();
      if (p >= s.size()) fail(line, "unterminated {");
      if (s[p] == '}') { p++; break; }
      const size_t b = p;
      while (p < s.size() && s[p] != ':' && s[p] != '}' ) p++;
      if (p >= s.size() || s[p] != ':') fail(line, "expected key: value in flow mapping");
      const std::string key = unquote(trim(s.substr(b, p - b)));
      p++;
      n->map.emplace_back(key, value());
      ws();
      if (p < s.size() && s[p] == ',') { p++; continue; }
      if (p < s.size() && s[p] == '}') { p++; break; }
      fail(line, "expected , or } in flow mapping");
    }
    return n;
  }
};

std::shared_ptr<YamlNode> inline_value(const std::string& text, int line) {
  const std::string t = trim(text);
  if (!t.empty() && (t[0] == '[' || t[0] == '{')) {
    Flow f{t, 0, line};
    auto n = f.value();
    f.ws();
    if (f.p != t.size()) fail(line, "trailing characters after flow collection");
    return n;
  }
  return make_scalar(t);
}

// ---- block structure
struct Parser {
  std::vector<Line> L;
  size_t i = 0;

  // the value text starting on line i (already positioned after "key:" / "- "), extended over following lines while a flow
  // collection is open
  std::string gather_flow(std::string first) {
    while (bracket_balance(first) > 0) {
      if (i >= L.size()) fail(L.back().no, "unterminated flow collection");
      first += " " + L[i].text;
      i++;
    }
    return first;
  }

  // a plain scalar continues over following lines that are indented deeper than its parent (folded with one blank each)
  std::string gather_plain(std::string first, int parent_indent) {
    if (!first.empty() && (first[0] == '[' || first[0] == '{')) return gather_flow(first);
    while (i < L.size() && L[i].indent > parent_indent) {
      if (key_colon(L[i].text) != std::string::npos) fail(L[i].no, "unexpected indentation (a mapping entry below a scalar value)");
      first += " " + L[i].text; i++;
    }
    return first;
  }

  std::shared_ptr<YamlNode> block(int indent) {
    if (i >= L.size() || L[i].indent < indent) { auto n = std::make_shared<YamlNode>(); return n; }
    const int ind = L[i].indent;
    if (L[i].text[0] == '-' && (L[i].text.size() == 1 || L[i].text[1] == ' ')) return sequence(ind);
    return mapping(ind);
  }

  std::shared_ptr<YamlNode> sequence(int ind) {
    auto n = std::make_shared<YamlNode>(); n->kind = YamlNode::Seq;
    while (i < L.size() && L[i].indent == ind && L[i].text[0] == '-' && (L[i].text.size() == 1 || L[i].text[1] == ' ')) {
      const int no = L[i].no;
      std::string rest = L[i].text.size() > 1 ? L[i].text.substr(2) : "";
      size_t lead = 0;
      while (lead < rest.size() && rest[lead] == ' ') lead++;
      rest = rest.substr(lead);
      if (rest.empty()) { i++; n->seq.push_back(block(ind + 1)); continue; }
      if (rest[0] == '-' && (rest.size() == 1 || rest[1] == ' ')) {
        // "- - item": a nested block sequence whose first entry sits on the dash line (how PyYAML's block style writes a
        // matrix); re-read the line as that sequence
        L[i].indent = ind + 2 + (int)lead;
        L[i].text = rest;
        n->seq.push_back(sequence(L[i].indent));
        continue;
      }
      const bool flow = rest[0] == '[' || rest[0] == '{';
      if (!flow && key_colon(rest) != std::string::npos) {
        // "- key: value": a mapping whose first entry sits on the dash line; re-read the line as that entry
        L[i].indent = ind + 2 + (int)lead;
        L[i].text = rest;
        n->seq.push_back(mapping(L[i].indent));
        continue;
      }
      i++;
      n->seq.push_back(inline_value(gather_plain(rest, ind), no));
    }
    return n;
  }

  std::shared_ptr<YamlNode> mapping(int ind) {
    auto n = std::make_shared<YamlNode>(); n->kind = YamlNode::Map;
    while (i < L.size() && L[i].indent == ind) {
      const std::string& t = L[i].text;
      const int no = L[i].no;
      if (t[0] == '-' && (t.size() == 1 || t[1] == ' ')) break;             // a sequence of the parent at this indent
      const size_t c = key_colon(t);
      if (c == std::string::npos) fail(no, "expected `key: value`, got `" + t + "`");
      const std::string key = unquote(trim(t.substr(0, c)));
      const std::string rest = trim(t.substr(c + 1));
      i++;
      if (!rest.empty()) { n->map.emplace_back(key, inline_value(gather_plain(rest, ind), no)); continue; }
      if (i < L.size() && L[i].indent > ind) n->map.emplace_back(key, block(L[i].indent));
      else if (i < L.size() && L[i].indent == ind && L[i].text[0] == '-' && (L[i].text.size() == 1 || L[i].text[1] == ' '))
        n->map.emplace_back(key, sequence(ind));                            // "key:" followed by "- item" at the same indent
      else n->map.emplace_back(key, std::make_shared<YamlNode>());
    }
    if (i < L.size() && L[i].indent > ind) fail(L[i].no, "unexpected indentation");
    return n;
  }
};

}  // namespace

std::shared_ptr<YamlNode> yaml_parse(const std::string& text) {
  Parser P;
  std::istringstream in(text);
  std::string raw;
  int no = 0;
  while (std::getline(in, raw)) {
    no++;
    if (!raw.empty() && raw.back() == '\r') raw.pop_back();
    std::string s = strip_comment(raw);
    size_t ind = 0;
    while (ind < s.size() && s[ind] == ' ') ind++;
    if (ind < s.size() && s[ind] == '\t') fail(no, "tab in indentation");
    const std::string t = trim(s);
    if (t.empty() || t == "---" || t == "...") continue;
    if (t[0] == '%') continue;                                              // directive
    P.L.push_back({(int)ind, t, no});
  }
  if (P.L.empty()) return std::make_shared<YamlNode>();
  auto root = P.block(P.L[0].indent);
  if (P.i != P.L.size()) fail(P.L[P.i].no, "could not parse `" + P.L[P.i].text + "`");
  return root;
}

std::shared_ptr<YamlNode> yaml_parse_file(const std::string& path) {
  std::ifstream f(path);
  if (!f.good()) throw std::runtime_error("cannot open " + path);
  std::stringstream ss;
  ss << f.rdbuf();
  return yaml_parse(ss.str());
}

const YamlNode* YamlNode::get(const std::string& key) const {
  if (kind != Map) return nullptr;
  for (const auto& kv : map) if (kv.first == key) return kv.second.get();
  return nullptr;
}
double YamlNode::as_double() const {
  if (kind != Scalar) throw std::runtime_error("YAML: expected a number");
  char* end = nullptr;
  const double v = std::strtod(scalar.c_str(), &end);
  if (end == scalar.c_str() || *end != '\0') throw std::runtime_error("YAML: not a number: " + scalar);
  return v;
}
long YamlNode::as_int() const {
  const double v = as_double();
  if (v != (double)(long)v) throw std::runtime_error("YAML: not an integer: " + scalar);
  return (long)v;
}
bool YamlNode::as_bool() const {
  if (kind != Scalar) throw std::runtime_error("YAML: expected a boolean");
  std::string t;
  for (char ch : scalar) t += (char)std::tolower((unsigned char)ch);
  if (t == "true" || t == "yes" || t == "on" || t == "1") return true;
  if (t == "false" || t == "no" || t == "off" || t == "0") return false;
  throw std::runtime_error("YAML: not a boolean: " + scalar);
}
std::vector<double> YamlNode::as_doubles() const {
  std::vector<double> out;
  if (kind == Seq) { for (const auto& e : seq) out.push_back(e->as_double()); }
  else out.push_back(as_double());
  return out;
}
std::vector<int> YamlNode::as_ints() const {
  std::vector<int> out;
  if (kind == Seq) { for (const auto& e : seq) out.push_back((int)e->as_int()); }
  else out.push_back((int)as_int());
  return out;
}
