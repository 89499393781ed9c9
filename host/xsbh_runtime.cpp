// xsbh_runtime.cpp -- operator factory, batch nodes, deck layering and graph resolution (see xsbh_operator.h)
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <fstream>

#include "xsbh_operator.h"

namespace xsbh {

// ------------------------------------------------------------------------------------------------ Simulation
double ThermoState::temperature() const {
  if (natoms == 0) return 0.0;
  // T = 2 (Ek - p^2 / 2M) / (3 N kB), per axis then averaged (simulation_thermodynamic_state.cpp:202, print op)
  double t = 0.0;
  for (int a = 0; a < 3; ++a) t += 2.0 * (0.5 * ke_tensor[a] - 0.5 * momentum[a] * momentum[a] / (mass > 0 ? mass : 1.0));
  return t / (3.0 * double(natoms) * kBoltzmann);
}

int Simulation::species_index(const std::string& name) const {
  for (size_t i = 0; i < species.size(); ++i) if (species[i].name == name) return (int)i;
  return -1;
}

std::string Simulation::data_path(const std::string& file) const {
  auto exists = [](const std::string& p) { std::ifstream f(p); return f.good(); };
  if (exists(file)) return file;
  for (auto& d : search_dirs) if (exists(d + "/" + file)) return d + "/" + file;
  return file;
}

void Simulation::check(int status, const char* what) const {
  if (status == XSB_OK) return;
  std::string msg = std::string(what) + " failed (" + std::to_string(status) + ")";
  if (ctx) msg += std::string(": ") + xsb_last_error(ctx);
  throw OperatorError(msg);
}

Simulation::~Simulation() { if (ctx) xsb_destroy(ctx); }

xsb_domain_desc Simulation::domain_desc() const {
  xsb_domain_desc d{};
  for (int a = 0; a < 3; ++a) {
    d.global_cells[a] = grid_dims[a]; d.periodic[a] = periodic[a]; d.rank_dims[a] = rank_dims[a]; d.rank_coord[a] = rank_coord[a];
    d.box[a] = grid_dims[a] * cell_size;
  }
  return d;
}

static int block_start(int coord, int cells, int parts) { return int((long long)coord * cells / parts); }

void Simulation::flush_staged() {
  if (!staged_dirty) return;
  if (!domain_ready) throw OperatorError("particles were created before the `domain` operator ran");
  if (!ctx) throw OperatorError("no GPU context: hw_device_init (init_cuda) must run before particles are placed on the grid");
  xsb_grid_desc g{};
  // ghost thickness: enough cell layers to cover nbh_dist measured through the narrowest physical cell extent
  double min_scale = 1e300;
  for (int a = 0; a < 3; ++a) {
    double col = std::sqrt(xform[a] * xform[a] + xform[3 + a] * xform[3 + a] + xform[6 + a] * xform[6 + a]);
    min_scale = std::min(min_scale, col);
  }
  double need = nbh_dist > 0 ? nbh_dist : rcut_max + rcut_inc;
  // operators that evaluate the embedding of ghost atoms themselves ask for a deeper halo (ghost_dist_max = 2 rcut:
  // eam_potential.cu:109-112, eam_potential_multimat.cu:116-120)
  if (ghost_dist_max > 0) need = std::max(need, ghost_dist_max + rcut_inc);
  ghost_layers = std::max(1, (int)std::ceil(need / (cell_size * min_scale) - 1e-12));
  for (int a = 0; a < 3; ++a) {
    int b0 = block_start(rank_coord[a], grid_dims[a], rank_dims[a]), b1 = block_start(rank_coord[a] + 1, grid_dims[a], rank_dims[a]);
    g.dims[a] = (b1 - b0) + 2 * ghost_layers;
    g.origin[a] = bounds_min[a] + (b0 - ghost_layers) * cell_size;
  }
  g.ghost_layers = ghost_layers;
  g.cell_size = cell_size;
  bool ident = true;
  for (int i = 0; i < 9; ++i) { g.xform[i] = xform[i]; if (xform[i] != (i % 4 == 0 ? 1.0 : 0.0)) ident = false; }
  g.xform_is_identity = ident;
  check(xsb_grid_set(ctx, &g), "xsb_grid_set");
  const uint64_t n = hx.size();
  check(xsb_particles_assign(ctx, n, hx.data(), hy.data(), hz.data(), hvx.empty() ? nullptr : hvx.data(), hvy.empty() ? nullptr : hvy.data(),
                             hvz.empty() ? nullptr : hvz.data(), htype.empty() ? nullptr : htype.data(), hid.empty() ? nullptr : hid.data()),
        "xsb_particles_assign");
  check(xsb_sync(ctx), "xsb_sync");
  hx.clear(); hy.clear(); hz.clear(); hvx.clear(); hvy.clear(); hvz.clear(); htype.clear(); hid.clear();
  hx.shrink_to_fit(); hy.shrink_to_fit(); hz.shrink_to_fit();
  staged_dirty = false; grid_ready = true; scheme_ready = false; neighbors_ready = false;
}

// ------------------------------------------------------------------------------------------------ Operator
const Node& Operator::required(const std::string& slot) const {
  const Node* n = slots.find(slot);
  if (!n || n->is_null()) throw OperatorError("operator '" + name + "': required slot '" + slot + "' is not set");
  return *n;
}

void Operator::check_slots(const std::set<std::string>& declared) const {
  if (!slots.is_map()) return;
  // slots every onika operator accepts
  static const std::set<std::string> common = {"profiling", "verbose", "name", "rebind", "gpu", "omp_num_threads", "debug", "log_level"};
  for (auto& kv : slots.map)
    if (!declared.count(kv.first) && !common.count(kv.first) && !injected.count(kv.first))
      throw OperatorError("operator '" + name + "' has no slot named '" + kv.first + "'");
}

OperatorFactory& OperatorFactory::instance() { static OperatorFactory f; return f; }
void OperatorFactory::register_factory(const std::string& name, OperatorCreator c) { creators_[name] = std::move(c); }
std::unique_ptr<Operator> OperatorFactory::make(const std::string& name) const {
  auto it = creators_.find(name);
  if (it == creators_.end()) throw OperatorError("no operator registered under the name '" + name + "'");
  auto op = it->second();
  op->name = name;
  return op;
}
std::vector<std::string> OperatorFactory::names() const {
  std::vector<std::string> v;
  for (auto& kv : creators_) v.push_back(kv.first);
  return v;
}

void Batch::execute(Simulation& sim) {
  auto cond_ok = [&]() {
    if (condition.empty()) return true;
    bool neg = condition.rfind("not ", 0) == 0;
    std::string flag = neg ? condition.substr(4) : condition;
    auto it = sim.flags.find(flag);
    bool v = it != sim.flags.end() && it->second;
    return neg ? !v : v;
  };
  if (loop) {
    while (cond_ok()) for (auto& op : body) op->execute(sim);
  } else if (cond_ok()) {
    for (auto& op : body) op->execute(sim);
  }
}

// operators of the reference graphs that belong to its control plane (AMR sub-grids, load balancing, locks, memory
// compaction, logging ...) and have no counterpart on this path: accepted and skipped so unmodified decks resolve
static const std::set<std::string>& passive_names() {
  static const std::set<std::string> s = {
      "nop", "print_logo_banner", "print_version_info", "message", "mpi_comm_world", "update_ghost_config", "finalize_cuda", "make_empty_grid",
      "grid_flavor", "grid_flavor_full", "grid_flavor_multimat", "grid_flavor_minimal", "grid_flavor_multimat_mechanics", "init_parameters",
      "generate_default_species", "particle_regions", "particles_regions", "init_prolog", "init_epilog", "init_rcb_grid", "grid_post_processing",
      "grid_memory_compact", "reduce_species_after_read", "print_domain", "performance_adviser", "memory_stats", "rebuild_amr", "amr_grid_pairs",
      "resize_particle_locks", "extend_domain", "load_balance", "trigger_load_balance", "load_balancing_if_triggered", "load_balance_auto_tune_start",
      "load_balance_auto_tune_end", "loadbalance_log_helper", "lb_event_counter", "profile_ghost_comm_scheme", "trigger_restart", "trigger_analysis",
      "trigger_snapshot", "write_restart_if_triggered", "perform_analysis_if_triggered", "write_snapshot_if_triggered", "write_final_restart",
      "write_restart", "default_thermostate_file", "thermostate_file_if_triggered", "trigger_thermostate_file", "nose_hoover_additional_step",
      "md_loop_prolog", "md_loop_epilog", "simulation_epilog_extra", "check_values", "grid_stats", "chunk_neighbors_stats",
      "final_dump", "species", "input_data"};
  return s;
}

namespace {

class NopOperator : public Operator {
public:
  void execute(Simulation& sim) override { if (sim.tracing) sim.trace.push_back(name); }
};

struct Resolver {
  const Node& deck;
  Simulation& sim;

  // YAML slots for operator `opname` = its top-level defaults overlaid by the instance's own map, with rebinds applied
  Node merged_slots(const std::string& opname, const Node* inst, const std::map<std::string, std::string>& rebind, std::set<std::string>& injected) {
    Node s = Node::make_map();
    const Node* top = deck.find(opname);
    if (top && top->is_map() && !top->has("body")) merge_into(s, *top);
    if (inst && inst->is_map()) merge_into(s, *inst);
    // rebind { slot: shared_name }: the slot is connected to a graph-level named value.  An instance that carries the
    // slot publishes it (eam_alloy_init: parameters -> eam_alloy_parameters); later instances without it read it back.
    for (auto& rb : rebind) {
      if (const Node* own = s.find(rb.first)) { sim.shared_slots[rb.second] = *own; continue; }
      if (const Node* v = deck.find(rb.second)) { s.set(rb.first, *v); injected.insert(rb.first); continue; }
      auto it = sim.shared_slots.find(rb.second);
      if (it != sim.shared_slots.end()) { s.set(rb.first, it->second); injected.insert(rb.first); }
    }
    return s;
  }

  std::unique_ptr<Operator> batch_from(const std::string& name, const Node& body, const Node* meta, std::map<std::string, std::string> rebind, int depth) {
    auto b = std::make_unique<Batch>();
    b->name = name;
    if (meta) {
      if (const Node* c = meta->find("condition")) b->condition = c->as_string();
      if (const Node* l = meta->find("loop")) b->loop = l->as_bool();
      if (const Node* r = meta->find("rebind")) if (r->is_map()) for (auto& kv : r->map) if (kv.second.is_scalar()) rebind[kv.first] = kv.second.as_string();
    }
    if (!body.is_seq()) throw OperatorError("batch '" + name + "': body must be a list");
    for (auto& item : body.seq) {
      if (item.is_scalar()) b->body.push_back(resolve(item.as_string(), nullptr, rebind, depth + 1));
      else if (item.is_map() && item.map.size() == 1) b->body.push_back(resolve(item.map[0].first, &item.map[0].second, rebind, depth + 1));
      else throw OperatorError("batch '" + name + "': every entry must be an operator name or a single-key map");
    }
    return b;
  }

  std::unique_ptr<Operator> resolve(const std::string& name, const Node* inst, const std::map<std::string, std::string>& rebind, int depth) {
    if (depth > 64) throw OperatorError("operator graph recursion too deep at '" + name + "' (alias cycle?)");
    // inline batch: "- helper: { rebind: ..., body: [...] }"
    if (inst && inst->is_map() && inst->has("body")) return batch_from(name, (*inst)["body"], inst, rebind, depth);
    const Node* top = deck.find(name);
    if (OperatorFactory::instance().has(name) && !(top && (top->is_scalar() || top->is_seq()))) {
      auto op = OperatorFactory::instance().make(name);
      op->slots = merged_slots(name, inst, rebind, op->injected);
      return op;
    }
    if (top) {
      if (top->is_scalar()) return resolve(top->as_string(), inst, rebind, depth + 1);
      if (top->is_seq()) return batch_from(name, *top, nullptr, rebind, depth);
      if (top->is_map() && top->has("body")) return batch_from(name, (*top)["body"], top, rebind, depth);
    }
    if (passive_names().count(name)) { auto op = std::make_unique<NopOperator>(); op->name = name; return op; }
    throw OperatorError("unknown operator '" + name + "' (not registered, not defined in the deck)");
  }
};

}  // namespace

std::unique_ptr<Operator> build_graph(const Node& deck, const std::string& name, Simulation& sim) {
  Resolver r{deck, sim};
  return r.resolve(name, nullptr, {}, 0);
}

void list_graph(const Operator& op, std::vector<std::string>& out) {
  if (auto* b = dynamic_cast<const Batch*>(&op)) { for (auto& c : b->body) list_graph(*c, out); }
  else out.push_back(op.name);
}

void apply_globals(const Node& deck, Simulation& sim) {
  const Node* g = deck.find("global");
  if (g && g->is_map()) {
    sim.dt = quantity_or(g->find("dt"), sim.dt);
    sim.rcut_inc = quantity_or(g->find("rcut_inc"), sim.rcut_inc);
    if (const Node* n = g->find("max_iteration")) sim.max_iteration = n->as_int();
    if (const Node* n = g->find("simulation_end_iteration")) sim.max_iteration = n->as_int();      // older deck vocabulary
    if (const Node* n = g->find("simulation_log_frequency")) sim.thermo_screen_frequency = n->as_int();
    if (const Node* n = g->find("timestep")) sim.timestep = n->as_int();
    if (const Node* n = g->find("simulation_thermostate_screen_frequency")) sim.thermo_screen_frequency = n->as_int();
    if (const Node* n = g->find("enable_mixed_precision")) sim.mixed_precision = n->as_bool();     // xsb extension
    if (const Node* n = g->find("compute_virial")) sim.compute_virial = n->as_bool();
  }
  if (const Node* gf = deck.find("grid_flavor")) if (gf->is_scalar() && gf->as_string().find("mechanics") != std::string::npos) sim.compute_virial = true;
}

Node load_deck(const std::string& path, const std::vector<std::string>& search_dirs) {
  Node deck = default_config();
  Node user = load_yaml_file(path, search_dirs);
  merge_into(deck, user);
  return deck;
}

}  // namespace xsbh
