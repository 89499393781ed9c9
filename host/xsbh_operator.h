// xsbh_operator.h -- host operator layer above the C ABI (include/xsb200.h).
//
// exaStamp force fields are onika::scg::OperatorNode subclasses with ADD_SLOT members, registered under a YAML name by
// OperatorNodeFactory::register_factory (src/potential/pair_potential_template/pair_potential_impl.hxx:510-513) and
// scheduled from a YAML deck (data/config/main-config.msp).  onika is not available here, so this layer restates the
// part of that contract the short-range force path needs -- same operator names, same slot names / defaults / error
// behaviour, same deck layering -- and forwards every compute to libxsb200.so.  No particle arithmetic happens on the
// host: an operator without a GPU fails.
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "../include/xsb200.h"
#include "xsbh_units.h"
#include "xsbh_yaml.h"

namespace xsbh {

struct OperatorError : std::runtime_error { using std::runtime_error::runtime_error; };

struct Species { std::string name; double mass = 1.0; double z = 0.0; double charge = 0.0; };

// thermodynamic_state values (src/thermo_state/simulation_thermodynamic_state.cpp:81-230), internal units
struct ThermoState {
  uint64_t natoms = 0;
  double kinetic = 0, potential = 0, mass = 0, volume = 0;
  double momentum[3] = {0, 0, 0};
  double ke_tensor[3] = {0, 0, 0};      // sum m v_a^2 (diagonal)
  double virial_diag[3] = {0, 0, 0};    // sum of per-atom virial diagonals (when the potentials produced them)
  double temperature() const;           // 2 Ek / (3 N kB)
  double total() const { return kinetic + potential; }
};

// what the graph's shared slots hold (grid, domain, species, rcut_max, chunk_neighbors, ...), one per rank
struct Simulation {
  xsb_ctx* ctx = nullptr;
  int device = 0, rank = 0, nranks = 1;
  int rank_dims[3] = {1, 1, 1}, rank_coord[3] = {0, 0, 0};
  bool cuda_required = true;            // false only for `--dry-run` graph resolution (no compute operator may run)
  // domain (exanb::Domain): bounds in grid space, xform, periodicity; grid: cells of this rank's brick + ghost layers
  double bounds_min[3] = {0, 0, 0}, bounds_max[3] = {0, 0, 0};
  double cell_size = 0.0;
  int grid_dims[3] = {0, 0, 0};         // global own cells
  int ghost_layers = 1;
  bool periodic[3] = {true, true, true};
  double xform[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  bool domain_ready = false, grid_ready = false, scheme_ready = false, neighbors_ready = false;
  // species / type map
  std::vector<Species> species;
  int species_index(const std::string& name) const;      // -1 when unknown
  // host-side staging of particles before the first xsb_particles_assign (lattice, readers, noise)
  std::vector<double> hx, hy, hz, hvx, hvy, hvz; std::vector<uint8_t> htype; std::vector<uint64_t> hid;
  bool staged_dirty = false;
  // global slots (config_globals.msp)
  double dt = 1.0e-3, rcut_inc = 1.0, rcut_max = 0.0, ghost_dist_max = 0.0, nbh_dist = 0.0, max_displ = 0.0;
  double physical_time = 0.0;
  long long timestep = 0, max_iteration = 0;
  long long thermo_screen_frequency = 10;
  bool trigger_thermo_state = true;     // energy step: force operators accumulate ep (and virial)
  bool compute_virial = false;          // grid flavor carries field::virial
  bool mixed_precision = false;         // xsb extension: XSB_FLAG_MIXED for the pair operators and eam_alloy_force
  std::string eam_alloy_loaded;         // setfl file whose tables the context holds (xsb_eam_alloy_set)
  std::map<std::string, bool> flags;    // trigger_move_particles, md_loop_continue, ...
  std::map<std::string, Node> shared_slots;   // graph-level named values connected to slots by `rebind`
  ThermoState thermo;
  bool preinit = false;                 // compute_force executed on the empty grid to collect rcut_max (main-config.msp:52-74)
  int verbosity = 1;
  std::vector<std::string> search_dirs; // data file lookup (onika data_file_path)
  std::string data_path(const std::string& file) const;
  std::vector<std::string> trace;       // names of the operators executed, in order (tests, --trace)
  bool tracing = false;

  void check(int status, const char* what) const;   // throws OperatorError with xsb_last_error
  void flush_staged();                               // staged host particles -> grid (xsb_grid_set + xsb_particles_assign)
  xsb_domain_desc domain_desc() const;
  ~Simulation();
};

// one node of the graph; `slots` is the YAML map given to this instance merged over the operator's top-level defaults
class Operator {
public:
  virtual ~Operator() = default;
  virtual void execute(Simulation& sim) = 0;
  std::string name;           // YAML operator name
  Node slots;                 // resolved slot values (map) or Null
  std::set<std::string> injected;   // slots that arrived through a batch-level `rebind` (ignored when not declared)
  // slot access with the reference's semantics: REQUIRED slots throw, optional ones fall back to the default
  const Node& required(const std::string& slot) const;
  const Node* optional(const std::string& slot) const { return slots.find(slot); }
  double quantity_slot(const std::string& slot, double dflt) const { return quantity_or(slots.find(slot), dflt); }
  bool bool_slot(const std::string& slot, bool dflt) const { const Node* n = slots.find(slot); return n && !n->is_null() ? n->as_bool() : dflt; }
  long long int_slot(const std::string& slot, long long dflt) const { const Node* n = slots.find(slot); return n && !n->is_null() ? n->as_int() : dflt; }
  std::string string_slot(const std::string& slot, const std::string& dflt) const { const Node* n = slots.find(slot); return n && n->is_scalar() ? n->as_string() : dflt; }
  // rejects slot names the operator does not declare (onika fails on unknown slots at graph build)
  void check_slots(const std::set<std::string>& declared) const;
};

using OperatorCreator = std::function<std::unique_ptr<Operator>()>;

class OperatorFactory {
public:
  static OperatorFactory& instance();
  void register_factory(const std::string& name, OperatorCreator c);
  bool has(const std::string& name) const { return creators_.count(name) != 0; }
  std::unique_ptr<Operator> make(const std::string& name) const;
  std::vector<std::string> names() const;
private:
  std::map<std::string, OperatorCreator> creators_;
};

struct OperatorRegistrar { OperatorRegistrar(const char* name, OperatorCreator c) { OperatorFactory::instance().register_factory(name, std::move(c)); } };
#define XSBH_REGISTER_OPERATOR(yaml_name, cls) \
  static ::xsbh::OperatorRegistrar xsbh_reg_##cls(yaml_name, []() { return std::unique_ptr<::xsbh::Operator>(new cls()); })

// batch / conditional / loop node (onika "batch" operators: body, condition, loop, rebind)
class Batch : public Operator {
public:
  std::vector<std::unique_ptr<Operator>> body;
  std::string condition;      // flag name, optionally prefixed by "not "
  bool loop = false;
  void execute(Simulation& sim) override;
};

// deck = built-in defaults (host/xsbh_default_config.cpp) <- includes <- user file; see load_deck
Node load_deck(const std::string& path, const std::vector<std::string>& search_dirs);
Node default_config();
// resolves `name` in the deck into an executable node (aliases, batches, operator instances with slot defaults)
std::unique_ptr<Operator> build_graph(const Node& deck, const std::string& name, Simulation& sim);
// reads `global:` and `configuration:` into the simulation
void apply_globals(const Node& deck, Simulation& sim);
// names of every operator a resolved graph would execute, depth first (used by tests and --dry-run)
void list_graph(const Operator& op, std::vector<std::string>& out);

}  // namespace xsbh
