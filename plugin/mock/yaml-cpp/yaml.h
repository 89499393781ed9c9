// mock of the yaml-cpp subset the shim touches
#pragma once
#include <string>
#include <vector>
namespace YAML {
struct Node {
  bool IsDefined() const { return true; }
  bool IsMap() const { return true; }
  bool IsSequence() const { return true; }
  bool IsScalar() const { return true; }
  size_t size() const { return 0; }
  Node operator[](const std::string&) const { return Node(); }
  Node operator[](const char*) const { return Node(); }
  Node operator[](size_t) const { return Node(); }
  template<class T> T as() const { return T(); }
  template<class T> T as(const T& fallback) const { return fallback; }
  const Node* begin() const { return nullptr; }
  const Node* end() const { return nullptr; }
};
}
