#pragma once
#include "../xsref_common.h"
#include <onika/physics/units.h>
#include <sstream>
namespace YAML {
  // scalar-or-map node, just enough for the convert<> specialisations of the potential headers
  struct Node {
    bool is_map = false; std::string scalar; std::map<std::string, Node> kids;
    bool IsMap() const { return is_map; }
    bool IsScalar() const { return !is_map; }
    bool IsSequence() const { return false; }
    explicit operator bool () const { return is_map || !scalar.empty(); }
    Node operator [] (const std::string& k) const { auto it = kids.find(k); return it == kids.end() ? Node{} : it->second; }
    template<class T> T as() const;
  };
  template<class T> struct convert;
  template<class T> inline T Node::as() const {
    if constexpr ( std::is_same_v<T,std::string> ) { return scalar; }
    else if constexpr ( std::is_same_v<T,onika::physics::Quantity> ) { return onika::physics::Quantity{ std::stod(scalar) }; }
    else if constexpr ( std::is_arithmetic_v<T> ) { std::istringstream iss(scalar); T v{}; iss >> v; return v; }
    else { T v{}; convert<T>::decode(*this, v); return v; }
  }
}
