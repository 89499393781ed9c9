// xsbh_yaml.h -- the YAML subset exaStamp input decks (*.msp / *.yaml) are written in.
//
// The reference parses decks with yaml-cpp through onika; neither is present in this
// image, so the host operator layer carries its own reader for exactly what the decks under
// data/regression_new/potentials/ use: block maps and sequences, flow maps / sequences (possibly spanning
// lines), plain / quoted scalars, comments, anchors (&name), aliases (*name), merge keys (<<), `includes:`.
#pragma once
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace xsbh {

struct YamlError : std::runtime_error { using std::runtime_error::runtime_error; };

class Node {
public:
  enum Kind { Null, Scalar, Seq, Map };
  Kind kind = Null;
  std::string scalar;                                  // Kind::Scalar
  std::vector<Node> seq;                               // Kind::Seq
  std::vector<std::pair<std::string, Node>> map;       // Kind::Map, insertion order kept (operator order is API)

  Node() = default;
  static Node make_scalar(std::string s) { Node n; n.kind = Scalar; n.scalar = std::move(s); return n; }
  static Node make_map() { Node n; n.kind = Map; return n; }
  static Node make_seq() { Node n; n.kind = Seq; return n; }

  bool is_null() const { return kind == Null; }
  bool is_scalar() const { return kind == Scalar; }
  bool is_seq() const { return kind == Seq; }
  bool is_map() const { return kind == Map; }
  size_t size() const { return kind == Seq ? seq.size() : kind == Map ? map.size() : 0; }

  const Node* find(const std::string& key) const;      // nullptr when absent or not a map
  Node* find(const std::string& key);
  const Node& operator[](const std::string& key) const;   // throws YamlError when absent
  const Node& operator[](size_t i) const;
  Node& set(const std::string& key, Node v);               // insert or replace
  bool has(const std::string& key) const { return find(key) != nullptr; }

  // scalar conversions (throw YamlError with the offending text)
  const std::string& as_string() const;
  bool as_bool() const;
  long long as_int() const;
  double as_double() const;                                // plain number, no unit

  std::string dump(int indent = 0) const;                  // debug / round-trip tests
};

Node parse_yaml(const std::string& text);
// loads `path`, then resolves the top-level `includes:` list (paths relative to the including file, searched also in
// `search_dirs`): included documents are merged first, the including file overrides key by key (maps merge
// recursively, everything else is replaced) -- onika's deck layering, data/config/main-config.msp:1-20.
Node load_yaml_file(const std::string& path, const std::vector<std::string>& search_dirs = {});
void merge_into(Node& base, const Node& over);

}  // namespace xsbh
