// mock of onika/scg/operator.h + operator_slot.h: OperatorNode, ADD_SLOT, slot access (*slot, slot->, has_value())
#pragma once
#include <functional>
#include <memory>
#include <string>
#include <vector>
#include <iostream>
namespace onika {
namespace parallel { struct ParallelExecutionContext { int gpu_device_index() const { return 0; } void* gpu_stream() const { return nullptr; } }; }
namespace scg {
struct DocString { const char* s; };
enum SlotDirection { INPUT, OUTPUT, INPUT_OUTPUT, PRIVATE };
struct RequiredTag {}; struct OptionalTag {};
static constexpr RequiredTag REQUIRED{}; static constexpr OptionalTag OPTIONAL{};
template<class T> struct OperatorSlot {
  std::shared_ptr<T> v = std::make_shared<T>();
  bool present = true;
  OperatorSlot() = default;
  template<class... A> OperatorSlot(SlotDirection, A&&...) {}
  T& operator*() const { return *v; }
  T* operator->() const { return v.get(); }
  bool has_value() const { return present; }
  T* get_pointer() const { return v.get(); }
};
class OperatorNode {
public:
  // in onika the slot keywords are visible inside every operator class
  static constexpr SlotDirection INPUT = SlotDirection::INPUT, OUTPUT = SlotDirection::OUTPUT, INPUT_OUTPUT = SlotDirection::INPUT_OUTPUT, PRIVATE = SlotDirection::PRIVATE;
  static constexpr RequiredTag REQUIRED{}; static constexpr OptionalTag OPTIONAL{};
  virtual ~OperatorNode() = default;
  virtual void execute() = 0;
  virtual std::string documentation() const { return ""; }
  onika::parallel::ParallelExecutionContext* parallel_execution_context() const { return nullptr; }
  const std::string& name() const { static std::string n; return n; }
};
class OperatorNodeFactory {
public:
  using Creator = std::function<std::shared_ptr<OperatorNode>()>;
  static OperatorNodeFactory* instance() { static OperatorNodeFactory f; return &f; }
  void register_factory(const std::string&, Creator) {}
};
} // scg
struct FatalStream { template<class T> FatalStream& operator<<(const T&) { return *this; } FatalStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; } };
inline FatalStream fatal_error() { return FatalStream(); }
inline std::ostream& lout_stream() { return std::cout; }
} // onika
#define ADD_SLOT(T, name, ...) ::onika::scg::OperatorSlot< T > name { __VA_ARGS__ }
#define ONIKA_AUTORUN_INIT(name) static void xsb_mock_autorun_##name(); static const int xsb_mock_autorun_flag_##name = (xsb_mock_autorun_##name(), 0); static void xsb_mock_autorun_##name()
