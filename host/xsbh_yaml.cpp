// xsbh_yaml.cpp -- reader for the YAML subset of exaStamp decks (see xsbh_yaml.h)
#include "xsbh_yaml.h"

#include <cctype>
#include <cstdlib>
#include <fstream>
#include <sstream>

namespace xsbh {

// ---------------------------------------------------------------------------------------------- Node
const Node* Node::find(const std::string& key) const {
  if (kind != Map) return nullptr;
  for (auto& kv : map) if (kv.first == key) return &kv.second;
  return nullptr;
}
Node* Node::find(const std::string& key) { return const_cast<Node*>(static_cast<const Node*>(this)->find(key)); }
const Node& Node::operator[](const std::string& key) const {
  const Node* n = find(key);
  if (!n) throw YamlError("key '" + key + "' not found");
  return *n;
}
const Node& Node::operator[](size_t i) const {
  if (kind != Seq || i >= seq.size()) throw YamlError("sequence index out of range");
  return seq[i];
}
Node& Node::set(const std::string& key, Node v) {
  if (kind == Null) kind = Map;
  if (kind != Map) throw YamlError("set('" + key + "') on a non-map node");
  for (auto& kv : map) if (kv.first == key) { kv.second = std::move(v); return kv.second; }
  map.emplace_back(key, std::move(v));
  return map.back().second;
}
const std::string& Node::as_string() const {
  if (kind != Scalar) throw YamlError("scalar expected");
  return scalar;
}
static std::string lower(std::string s) { for (auto& c : s) c = (char)std::tolower((unsigned char)c); return s; }
bool Node::as_bool() const {
  std::string s = lower(as_string());
  if (s == "true" || s == "yes" || s == "on" || s == "1") return true;
  if (s == "false" || s == "no" || s == "off" || s == "0") return false;
  throw YamlError("boolean expected, got '" + scalar + "'");
}
long long Node::as_int() const {
  const std::string& s = as_string();
  char* e = nullptr;
  long long v = std::strtoll(s.c_str(), &e, 10);
  if (e == s.c_str() || *e) {
    // 1e3-style integers
    double d = std::strtod(s.c_str(), &e);
    if (e == s.c_str() || *e || d != (double)(long long)d) throw YamlError("integer expected, got '" + s + "'");
    return (long long)d;
  }
  return v;
}
double Node::as_double() const {
  const std::string& s = as_string();
  char* e = nullptr;
  double v = std::strtod(s.c_str(), &e);
  if (e == s.c_str() || *e) throw YamlError("number expected, got '" + s + "'");
  return v;
}
std::string Node::dump(int indent) const {
  std::string pad(indent, ' '), out;
  switch (kind) {
    case Null: return "~";
    case Scalar: return "\"" + scalar + "\"";
    case Seq:
      out = "[";
      for (size_t i = 0; i < seq.size(); ++i) out += (i ? ", " : "") + seq[i].dump(indent);
      return out + "]";
    case Map:
      out = "{";
      for (size_t i = 0; i < map.size(); ++i) out += (i ? ", " : "") + map[i].first + ": " + map[i].second.dump(indent);
      return out + "}";
  }
  return out;
}

void merge_into(Node& base, const Node& over) {
  if (base.is_map() && over.is_map()) {
    for (auto& kv : over.map) {
      Node* b = base.find(kv.first);
      if (b && b->is_map() && kv.second.is_map()) merge_into(*b, kv.second);
      else base.set(kv.first, kv.second);
    }
  } else {
    base = over;
  }
}

// ---------------------------------------------------------------------------------------------- parser
namespace {

struct Line { int indent; std::string text; int no; };

std::string rtrim(std::string s) { while (!s.empty() && std::isspace((unsigned char)s.back())) s.pop_back(); return s; }
std::string ltrim(const std::string& s) { size_t i = 0; while (i < s.size() && std::isspace((unsigned char)s[i])) ++i; return s.substr(i); }
std::string trim(const std::string& s) { return rtrim(ltrim(s)); }

std::string strip_comment(const std::string& s) {
  char q = 0;
  for (size_t i = 0; i < s.size(); ++i) {
    char c = s[i];
    if (q) { if (c == q) q = 0; continue; }
    if (c == '"' || c == '\'') { q = c; continue; }
    if (c == '#' && (i == 0 || std::isspace((unsigned char)s[i - 1]))) return s.substr(0, i);
  }
  return s;
}

std::string unquote(const std::string& s) {
  if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\''))) {
    std::string in = s.substr(1, s.size() - 2), out;
    if (s.front() == '\'') return in;
    for (size_t i = 0; i < in.size(); ++i) {
      if (in[i] == '\\' && i + 1 < in.size()) {
        char c = in[++i];
        out += c == 'n' ? '\n' : c == 't' ? '\t' : c;
      } else out += in[i];
    }
    return out;
  }
  return s;
}

// index of the ':' that separates a block-map key from its value, or npos
size_t find_key_colon(const std::string& t) {
  if (t.empty() || t[0] == '{' || t[0] == '[' || t[0] == '&' || t[0] == '*' || t[0] == '|' || t[0] == '>') return std::string::npos;
  char q = 0;
  int depth = 0;
  for (size_t i = 0; i < t.size(); ++i) {
    char c = t[i];
    if (q) { if (c == q) q = 0; continue; }
    if (c == '"' || c == '\'') { q = c; continue; }
    if (c == '{' || c == '[') ++depth;
    else if (c == '}' || c == ']') --depth;
    else if (c == ':' && depth == 0 && (i + 1 == t.size() || std::isspace((unsigned char)t[i + 1]))) return i;
  }
  return std::string::npos;
}

bool is_seq_item(const std::string& t) { return !t.empty() && t[0] == '-' && (t.size() == 1 || std::isspace((unsigned char)t[1])); }

int bracket_balance(const std::string& t) {
  char q = 0; int d = 0;
  for (char c : t) {
    if (q) { if (c == q) q = 0; continue; }
    if (c == '"' || c == '\'') q = c;
    else if (c == '{' || c == '[') ++d;
    else if (c == '}' || c == ']') --d;
  }
  return d;
}

class Parser {
public:
  explicit Parser(const std::string& text) {
    std::istringstream is(text);
    std::string raw; int no = 0;
    while (std::getline(is, raw)) {
      ++no;
      for (auto& c : raw) if (c == '\t') c = ' ';
      std::string s = rtrim(strip_comment(raw));
      size_t ind = 0; while (ind < s.size() && s[ind] == ' ') ++ind;
      if (ind == s.size()) continue;
      std::string t = s.substr(ind);
      if (t == "---" || t == "...") continue;
      lines_.push_back({(int)ind, t, no});
    }
  }
  Node parse() {
    Node n = block(0);
    if (pos_ < lines_.size()) fail("unexpected content (bad indentation?)");
    return n;
  }

private:
  std::vector<Line> lines_;
  size_t pos_ = 0;
  std::map<std::string, Node> anchors_;

  [[noreturn]] void fail(const std::string& msg) {
    int no = pos_ < lines_.size() ? lines_[pos_].no : (lines_.empty() ? 0 : lines_.back().no);
    throw YamlError("yaml line " + std::to_string(no) + ": " + msg);
  }

  Node block(int min_indent) {
    if (pos_ >= lines_.size() || lines_[pos_].indent < min_indent) return Node();
    const Line& L = lines_[pos_];
    if (is_seq_item(L.text)) return sequence(L.indent);
    if (find_key_colon(L.text) != std::string::npos) return mapping(L.indent);
    std::string t = L.text;
    ++pos_;
    return value_after(t, L.indent - 1, false);
  }

  Node mapping(int indent) {
    Node m = Node::make_map();
    while (pos_ < lines_.size() && lines_[pos_].indent == indent && !is_seq_item(lines_[pos_].text)) {
      std::string t = lines_[pos_].text;
      size_t c = find_key_colon(t);
      if (c == std::string::npos) fail("'key: value' expected, got '" + t + "'");
      std::string key = unquote(trim(t.substr(0, c)));
      std::string rest = trim(t.substr(c + 1));
      ++pos_;
      Node v = value_after(rest, indent, true);
      if (key == "<<") {
        auto merge_one = [&](const Node& src) {
          if (!src.is_map()) fail("merge key '<<' needs a map");
          for (auto& kv : src.map) if (!m.has(kv.first)) m.set(kv.first, kv.second);
        };
        if (v.is_seq()) for (auto& s : v.seq) merge_one(s); else merge_one(v);
      } else {
        m.set(key, std::move(v));
      }
    }
    if (pos_ < lines_.size() && lines_[pos_].indent > indent) fail("unexpected indentation");
    return m;
  }

  Node sequence(int indent) {
    Node s = Node::make_seq();
    while (pos_ < lines_.size() && lines_[pos_].indent == indent && is_seq_item(lines_[pos_].text)) {
      std::string t = lines_[pos_].text;
      size_t off = 1; while (off < t.size() && t[off] == ' ') ++off;
      std::string rest = t.substr(off);
      if (rest.empty()) {
        ++pos_;
        s.seq.push_back(pos_ < lines_.size() && lines_[pos_].indent > indent ? block(lines_[pos_].indent) : Node());
      } else if (find_key_colon(rest) != std::string::npos || is_seq_item(rest)) {
        // "- key: value" starts a map (or nested sequence) whose indentation is the column of `key`
        lines_[pos_] = Line{indent + (int)off, rest, lines_[pos_].no};
        s.seq.push_back(block(indent + (int)off));
      } else {
        ++pos_;
        s.seq.push_back(value_after(rest, indent, false));
      }
    }
    return s;
  }

  // value that follows "key:" or "- " on the same line (possibly empty => nested block on the next lines)
  Node value_after(std::string rest, int parent_indent, bool from_map) {
    std::string anchor;
    if (!rest.empty() && rest[0] == '&') {
      size_t e = 1; while (e < rest.size() && !std::isspace((unsigned char)rest[e])) ++e;
      anchor = rest.substr(1, e - 1);
      rest = trim(rest.substr(e));
    }
    Node v;
    if (rest.empty()) {
      if (pos_ < lines_.size() && lines_[pos_].indent > parent_indent) v = block(lines_[pos_].indent);
      else if (from_map && pos_ < lines_.size() && lines_[pos_].indent == parent_indent && is_seq_item(lines_[pos_].text)) v = sequence(parent_indent);
    } else if (rest[0] == '*') {
      auto it = anchors_.find(trim(rest.substr(1)));
      if (it == anchors_.end()) fail("unknown alias '" + rest + "'");
      v = it->second;
    } else if (rest[0] == '|' || rest[0] == '>') {
      std::string acc; char sep = rest[0] == '|' ? '\n' : ' ';
      while (pos_ < lines_.size() && lines_[pos_].indent > parent_indent) { if (!acc.empty()) acc += sep; acc += lines_[pos_].text; ++pos_; }
      v = Node::make_scalar(acc);
    } else if (rest[0] == '{' || rest[0] == '[') {
      while (bracket_balance(rest) > 0) {
        if (pos_ >= lines_.size()) fail("unterminated flow collection");
        rest += " " + lines_[pos_].text; ++pos_;
      }
      size_t i = 0;
      v = flow(rest, i);
      while (i < rest.size() && std::isspace((unsigned char)rest[i])) ++i;
      if (i != rest.size()) fail("trailing characters after flow collection: '" + rest.substr(i) + "'");
    } else {
      v = Node::make_scalar(unquote(rest));
      if (rest == "~" || rest == "null") v = Node();
    }
    if (!anchor.empty()) anchors_[anchor] = v;
    return v;
  }

  static void skip_ws(const std::string& s, size_t& i) { while (i < s.size() && std::isspace((unsigned char)s[i])) ++i; }

  std::string flow_scalar(const std::string& s, size_t& i, bool is_key) {
    skip_ws(s, i);
    if (i < s.size() && (s[i] == '"' || s[i] == '\'')) {
      char q = s[i]; size_t b = i++;
      while (i < s.size() && s[i] != q) { if (q == '"' && s[i] == '\\') ++i; ++i; }
      if (i >= s.size()) fail("unterminated quoted string");
      ++i;
      return unquote(s.substr(b, i - b));
    }
    size_t b = i;
    while (i < s.size()) {
      char c = s[i];
      if (c == ',' || c == '}' || c == ']') break;
      if (is_key && c == ':' && (i + 1 == s.size() || std::isspace((unsigned char)s[i + 1]) || s[i + 1] == '{' || s[i + 1] == '[')) break;
      ++i;
    }
    return trim(s.substr(b, i - b));
  }

  Node flow(const std::string& s, size_t& i) {
    skip_ws(s, i);
    if (i >= s.size()) return Node();
    if (s[i] == '{') {
      ++i;
      Node m = Node::make_map();
      for (;;) {
        skip_ws(s, i);
        if (i >= s.size()) fail("unterminated flow map");
        if (s[i] == '}') { ++i; break; }
        if (s[i] == ',') { ++i; continue; }
        std::string key = flow_scalar(s, i, true);
        skip_ws(s, i);
        Node v;
        if (i < s.size() && s[i] == ':') { ++i; v = flow(s, i); }
        m.set(key, std::move(v));
      }
      return m;
    }
    if (s[i] == '[') {
      ++i;
      Node q = Node::make_seq();
      for (;;) {
        skip_ws(s, i);
        if (i >= s.size()) fail("unterminated flow sequence");
        if (s[i] == ']') { ++i; break; }
        if (s[i] == ',') { ++i; continue; }
        q.seq.push_back(flow(s, i));
      }
      return q;
    }
    bool quoted = s[i] == '"' || s[i] == '\'';
    std::string t = flow_scalar(s, i, false);
    if (!quoted && !t.empty() && t[0] == '*') {
      auto it = anchors_.find(t.substr(1));
      if (it == anchors_.end()) fail("unknown alias '" + t + "'");
      return it->second;
    }
    if (!quoted && (t == "~" || t == "null" || t.empty())) return Node();
    return Node::make_scalar(t);
  }
};

std::string dir_of(const std::string& p) { size_t s = p.find_last_of('/'); return s == std::string::npos ? std::string(".") : p.substr(0, s); }
bool file_exists(const std::string& p) { std::ifstream f(p); return f.good(); }

}  // namespace

Node parse_yaml(const std::string& text) { return Parser(text).parse(); }

Node load_yaml_file(const std::string& path, const std::vector<std::string>& search_dirs) {
  std::ifstream f(path);
  if (!f) throw YamlError("cannot open '" + path + "'");
  std::stringstream ss; ss << f.rdbuf();
  Node doc;
  try { doc = parse_yaml(ss.str()); } catch (const YamlError& e) { throw YamlError(path + ": " + e.what()); }
  if (doc.is_null()) doc = Node::make_map();
  const Node* inc = doc.find("includes");
  if (!inc) return doc;
  Node base = Node::make_map();
  std::vector<std::string> names;
  if (inc->is_seq()) for (auto& n : inc->seq) names.push_back(n.as_string()); else if (inc->is_scalar()) names.push_back(inc->as_string());
  for (auto& n : names) {
    std::vector<std::string> cand{dir_of(path) + "/" + n};
    if (!n.empty() && n[0] == '/') cand.insert(cand.begin(), n);
    for (auto& d : search_dirs) cand.push_back(d + "/" + n);
    std::string found;
    for (auto& c : cand) if (file_exists(c)) { found = c; break; }
    if (found.empty()) throw YamlError(path + ": include '" + n + "' not found");
    merge_into(base, load_yaml_file(found, search_dirs));
  }
  Node own = Node::make_map();
  for (auto& kv : doc.map) if (kv.first != "includes") own.set(kv.first, kv.second);
  merge_into(base, own);
  return base;
}

}  // namespace xsbh
