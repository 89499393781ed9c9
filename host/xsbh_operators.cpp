// xsbh_operators.cpp -- the operators of the short-range force path under their exaStamp YAML names.  Each class cites
// the reference operator whose slots / defaults / failure behaviour it keeps; every compute goes to the C ABI.
#include <unistd.h>

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <thread>

#include "xsbh_operator.h"
#include "xsbh_readers.h"

namespace xsbh {
namespace {

#define TRACE(sim) do { if ((sim).tracing) (sim).trace.push_back(name); } while (0)

void need_gpu(const Simulation& sim, const std::string& op) {
  if (!sim.ctx) throw OperatorError("operator '" + op + "' needs the GPU context: init_cuda has not run or no sm_100 device is usable (there is no CPU fallback)");
}

// counter-based generator: value k of particle `id` under `seed` is a pure function of (seed, id, k), so lattices
// with noise do not depend on the rank decomposition.  (The reference's generators live in exaNBody: its streams
// cannot be reproduced bit for bit -- decks that need identical positions use read_xyz_file_with_xform.)
inline uint64_t mix64(uint64_t z) { z += 0x9e3779b97f4a7c15ull; z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }
inline double uniform01(uint64_t seed, uint64_t id, uint64_t k) { return ((mix64(mix64(seed ^ (id * 0x2545f4914f6cdd1dull)) + k) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
inline void gauss3(uint64_t seed, uint64_t id, double out[3]) {
  double g[4];
  for (int p = 0; p < 2; ++p) {
    double u1 = uniform01(seed, id, 2 * p), u2 = uniform01(seed, id, 2 * p + 1);
    double r = std::sqrt(-2.0 * std::log(u1));
    g[2 * p] = r * std::cos(2.0 * M_PI * u2); g[2 * p + 1] = r * std::sin(2.0 * M_PI * u2);
  }
  out[0] = g[0]; out[1] = g[1]; out[2] = g[2];
}

std::vector<double> masses(const Simulation& sim) {
  std::vector<double> m;
  for (auto& s : sim.species) m.push_back(s.mass);
  if (m.empty()) m.push_back(1.0);
  return m;
}

int force_flags(const Simulation& sim, bool ghost) {
  int f = 0;
  if (ghost) f |= XSB_FLAG_GHOST;
  if (sim.trigger_thermo_state) f |= XSB_FLAG_ENERGY;
  if (sim.compute_virial && sim.trigger_thermo_state) f |= XSB_FLAG_VIRIAL;
  return f;
}

uint32_t field_bit(const std::string& f) {
  static const std::map<std::string, int> m = {{"rx", XSB_F_RX}, {"ry", XSB_F_RY}, {"rz", XSB_F_RZ}, {"fx", XSB_F_FX}, {"fy", XSB_F_FY}, {"fz", XSB_F_FZ},
                                               {"ep", XSB_F_EP}, {"vx", XSB_F_VX}, {"vy", XSB_F_VY}, {"vz", XSB_F_VZ}, {"virial", XSB_F_VIRIAL},
                                               {"rho_dEmb", XSB_F_RHO_DEMB}, {"type", XSB_F_TYPE}, {"id", XSB_F_ID}};
  auto it = m.find(f);
  if (it == m.end()) throw OperatorError("unknown particle field '" + f + "'");
  return 1u << it->second;
}

// ================================================================================================ hardware
// init_cuda (main-config.msp:78-86): one context per rank.  Ranks come from the launcher environment
// (RANK / WORLD_SIZE / LOCAL_RANK, as set by `xsb200-run --gpus N` or torch.distributed.run); the NCCL id travels
// through a file because there is no MPI on this path.
class InitCuda : public Operator {
public:
  void execute(Simulation& sim) override {
    TRACE(sim);
    check_slots({"enable_cuda", "single_gpu", "rotate_gpu", "smem_bksize", "device"});
    if (!bool_slot("enable_cuda", true)) throw OperatorError("init_cuda: enable_cuda=false is not supported, this build has no CPU path");
    if (sim.ctx) return;
    const char* e;
    sim.rank = (e = std::getenv("RANK")) ? std::atoi(e) : 0;
    sim.nranks = (e = std::getenv("WORLD_SIZE")) ? std::atoi(e) : 1;
    sim.device = (int)int_slot("device", (e = std::getenv("LOCAL_RANK")) ? std::atoi(e) : 0);
    int rc = xsb_create(sim.device, &sim.ctx);
    if (rc != XSB_OK) {
      std::string msg = sim.ctx ? xsb_last_error(sim.ctx) : "xsb_create failed";
      if (sim.ctx) { xsb_destroy(sim.ctx); sim.ctx = nullptr; }
      if (sim.cuda_required) throw OperatorError("init_cuda: " + msg);
      return;
    }
    static const int dims[9][3] = {{1, 1, 1}, {1, 1, 1}, {2, 1, 1}, {0, 0, 0}, {2, 2, 1}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {2, 2, 2}};
    if (sim.nranks > 8 || dims[sim.nranks][0] == 0) throw OperatorError("init_cuda: 1, 2, 4 or 8 ranks per node are supported");
    for (int a = 0; a < 3; ++a) sim.rank_dims[a] = dims[sim.nranks][a];
    sim.rank_coord[0] = sim.rank % sim.rank_dims[0];
    sim.rank_coord[1] = (sim.rank / sim.rank_dims[0]) % sim.rank_dims[1];
    sim.rank_coord[2] = sim.rank / (sim.rank_dims[0] * sim.rank_dims[1]);
    if (sim.nranks > 1) {
      const char* port = std::getenv("MASTER_PORT");
      std::string path = std::string("/tmp/xsb200_nccl_id_") + (port ? port : "0") + "_" + std::to_string((long)getppid());
      if (const char* f = std::getenv("XSB_NCCL_ID_FILE")) path = f;
      char id[128];
      if (sim.rank == 0) {
        sim.check(xsb_comm_unique_id(id), "xsb_comm_unique_id");
        std::string tmp = path + ".tmp";
        { std::ofstream o(tmp, std::ios::binary); o.write(id, 128); }
        std::rename(tmp.c_str(), path.c_str());
      } else {
        for (int tries = 0;; ++tries) {
          std::ifstream i(path, std::ios::binary);
          if (i.read(id, 128)) break;
          if (tries > 600) throw OperatorError("init_cuda: timed out waiting for the NCCL id file " + path);
          std::this_thread::sleep_for(std::chrono::milliseconds(100));
        }
      }
      sim.check(xsb_comm_init(sim.ctx, sim.nranks, sim.rank, id), "xsb_comm_init");
    }
  }
};
XSBH_REGISTER_OPERATOR("init_cuda", InitCuda);

// ================================================================================================ domain, species
// domain (ext exanb operator; slots as used by every deck: data/config/main-config.msp:204-208 and
// potentials/pair/lj/single_specy_nosym.msp:19-27)
class DomainOp : public Operator {
public:
  void execute(Simulation& sim) override {
    TRACE(sim);
    check_slots({"cell_size", "grid_dims", "bounds", "xform", "periodic", "expandable", "mirror"});
    if (const Node* p = optional("periodic")) for (int a = 0; a < 3 && a < (int)p->size(); ++a) sim.periodic[a] = (*p)[a].as_bool();
    if (const Node* x = optional("xform")) {
      if (!x->is_seq() || x->size() != 3) throw OperatorError("domain: xform must be a 3x3 matrix");
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) sim.xform[3 * i + j] = (*x)[i][j].as_double();
    }
    double cs = quantity_slot("cell_size", 0.0);
    int gd[3] = {0, 0, 0};
    if (const Node* g = optional("grid_dims")) for (int a = 0; a < 3 && a < (int)g->size(); ++a) gd[a] = (int)(*g)[a].as_int();
    const Node* b = optional("bounds");
    bool have_bounds = b && b->is_seq() && b->size() == 2;
    if (have_bounds) for (int a = 0; a < 3; ++a) { sim.bounds_min[a] = quantity((*b)[0][a]); sim.bounds_max[a] = quantity((*b)[1][a]); }
    else if (cs > 0 && gd[0] > 0) for (int a = 0; a < 3; ++a) { sim.bounds_min[a] = 0.0; sim.bounds_max[a] = gd[a] * cs; have_bounds = true; }
    if (!have_bounds) return;     // the top-level default `domain` node: bounds arrive later (setup_system / reader)
    finalize(sim, cs, gd);
  }
  // cubic cells of edge cell_size tile the grid-space bounds; when the requested box is not a multiple of a common
  // edge the difference is absorbed by a diagonal scaling of xform, which is how exanb::Domain represents it
  static void finalize(Simulation& sim, double cs, const int gd_in[3]) {
    double L[3]; int gd[3];
    for (int a = 0; a < 3; ++a) { L[a] = sim.bounds_max[a] - sim.bounds_min[a]; gd[a] = gd_in[a]; if (L[a] <= 0) throw OperatorError("domain: empty bounds"); }
    double want = cs > 0 ? cs : (sim.nbh_dist > 0 ? sim.nbh_dist : sim.rcut_max + sim.rcut_inc);
    if (want <= 0) throw OperatorError("domain: cell_size is 0 and no potential declared a cutoff (rcut_max = 0)");
    for (int a = 0; a < 3; ++a) if (gd[a] <= 0) gd[a] = std::max(1, cs > 0 ? (int)std::llround(L[a] / cs) : (int)std::floor(L[a] / want + 1e-9));
    double edge = L[0] / gd[0];
    bool cubic = true;
    for (int a = 0; a < 3; ++a) if (std::fabs(L[a] / gd[a] - edge) > 1e-9 * edge) cubic = false;
    if (!cubic) {
      for (int a = 0; a < 3; ++a) {
        double s = (L[a] / gd[a]) / edge;
        for (int r = 0; r < 3; ++r) sim.xform[3 * r + a] *= s;
        sim.bounds_min[a] /= s; sim.bounds_max[a] = sim.bounds_min[a] + gd[a] * edge;
      }
    }
    sim.cell_size = edge;
    for (int a = 0; a < 3; ++a) { sim.grid_dims[a] = gd[a]; if (gd[a] < sim.rank_dims[a]) throw OperatorError("domain: fewer cells than ranks along an axis"); }
    sim.domain_ready = true;
  }
};
XSBH_REGISTER_OPERATOR("domain", DomainOp);

void add_species(Simulation& sim, const std::string& name, const Node* props) {
  int i = sim.species_index(name);
  if (i < 0) { sim.species.push_back(Species{name}); i = (int)sim.species.size() - 1; }
  if (props && props->is_map()) {
    sim.species[i].mass = quantity_or(props->find("mass"), sim.species[i].mass);
    sim.species[i].z = quantity_or(props->find("z"), sim.species[i].z);
    sim.species[i].charge = quantity_or(props->find("charge"), sim.species[i].charge);
  }
}

// particle_types / particle_type_add_properties (setup_system entries of the decks)
class ParticleTypes : public Operator {
public:
  void execute(Simulation& sim) override {
    TRACE(sim);
    check_slots({"particle_type_map"});
    const Node& m = required("particle_type_map");
    if (!m.is_map()) throw OperatorError("particle_types: particle_type_map must be a map name -> id");
    std::vector<std::pair<long long, std::string>> order;
    for (auto& kv : m.map) order.emplace_back(kv.second.as_int(), kv.first);
    std::sort(order.begin(), order.end());
    for (size_t i = 0; i < order.size(); ++i) {
      if (order[i].first != (long long)i) throw OperatorError("particle_types: ids must be 0..n-1");
      add_species(sim, order[i].second, nullptr);
      if (sim.species_index(order[i].second) != (int)i) throw OperatorError("particle_types: '" + order[i].second + "' conflicts with the species list order");
    }
  }
};
XSBH_REGISTER_OPERATOR("particle_types", ParticleTypes);

class ParticleTypeAddProperties : public Operator {
public:
  void execute(Simulation& sim) override {
    TRACE(sim);
    if (slots.is_map()) for (auto& kv : slots.map) if (kv.second.is_map()) add_species(sim, kv.first, &kv.second);
  }
};
XSBH_REGISTER_OPERATOR("particle_type_add_properties", ParticleTypeAddProperties);

// ================================================================================================ particle sources
// lattice (ext exanb operator): slots structure / types / size as in the decks.  Fills the domain with whole unit
// cells; only the unit cells overlapping this rank's brick are generated.
class Lattice : public Operator {
public:
  void execute(Simulation& sim) override {
    TRACE(sim);
    check_slots({"structure", "types", "size", "repeats", "init_domain", "shift", "np", "positions", "region", "grid_cell_values", "noise"});
    std::string st = required("structure").as_string();
    for (auto& c : st) c = (char)std::toupper((unsigned char)c);
    std::vector<std::array<double, 3>> basis;
    if (st == "SC") basis = {{0.25, 0.25, 0.25}};
    else if (st == "BCC") basis = {{0.25, 0.25, 0.25}, {0.75, 0.75, 0.75}};
    else if (st == "FCC") basis = {{0.25, 0.25, 0.25}, {0.25, 0.75, 0.75}, {0.75, 0.25, 0.75}, {0.75, 0.75, 0.25}};
    else throw OperatorError("lattice: structure '" + st + "' is not supported (SC, BCC, FCC)");
    const Node& ty = required("types");
    if (!ty.is_seq() || ty.size() != basis.size()) throw OperatorError("lattice: `types` needs one entry per basis atom");
    std::vector<uint8_t> tid;
    for (auto& t : ty.seq) {
      int i = sim.species_index(t.as_string());
      if (i < 0) { add_species(sim, t.as_string(), nullptr); i = sim.species_index(t.as_string()); }
      tid.push_back((uint8_t)i);
    }
    const Node& sz = required("size");
    double a[3]; for (int k = 0; k < 3; ++k) a[k] = quantity(sz[k]);
    long long rep[3];
    if (const Node* r = optional("repeats")) {
      for (int k = 0; k < 3; ++k) rep[k] = (*r)[k].as_int();
      if (!sim.domain_ready) {   // init_domain: the lattice defines the box
        for (int k = 0; k < 3; ++k) { sim.bounds_min[k] = 0.0; sim.bounds_max[k] = rep[k] * a[k]; }
        int gd[3] = {0, 0, 0};
        DomainOp::finalize(sim, 0.0, gd);
      }
    } else {
      if (!sim.domain_ready) throw OperatorError("lattice: no domain bounds and no `repeats`");
    }
    // physical box edges (grid space x column scaling of xform); lattice coordinates are physical
    double Lp[3], sc[3];
    for (int k = 0; k < 3; ++k) {
      sc[k] = std::sqrt(sim.xform[k] * sim.xform[k] + sim.xform[3 + k] * sim.xform[3 + k] + sim.xform[6 + k] * sim.xform[6 + k]);
      Lp[k] = (sim.bounds_max[k] - sim.bounds_min[k]) * sc[k];
      if (!optional("repeats")) rep[k] = std::llround(Lp[k] / a[k]);
      if (rep[k] < 1) throw OperatorError("lattice: the domain is smaller than one unit cell");
    }
    // brick of this rank in unit-cell indices (a unit cell belongs to the rank that owns its atoms one by one)
    const size_t nb = basis.size();
    double lo[3], hi[3];
    for (int k = 0; k < 3; ++k) {
      int c0 = int((long long)sim.rank_coord[k] * sim.grid_dims[k] / sim.rank_dims[k]), c1 = int((long long)(sim.rank_coord[k] + 1) * sim.grid_dims[k] / sim.rank_dims[k]);
      lo[k] = sim.bounds_min[k] + c0 * sim.cell_size; hi[k] = sim.bounds_min[k] + c1 * sim.cell_size;
    }
    long long u0[3], u1[3];
    for (int k = 0; k < 3; ++k) {
      u0[k] = std::max(0LL, (long long)std::floor((lo[k] - sim.bounds_min[k]) * sc[k] / a[k]) - 1);
      u1[k] = std::min(rep[k], (long long)std::ceil((hi[k] - sim.bounds_min[k]) * sc[k] / a[k]) + 1);
    }
    for (long long i = u0[0]; i < u1[0]; ++i) for (long long j = u0[1]; j < u1[1]; ++j) for (long long k = u0[2]; k < u1[2]; ++k)
      for (size_t b = 0; b < nb; ++b) {
        // grid-space position (diagonal xform only matters through sc[])
        double g[3] = {sim.bounds_min[0] + (i + basis[b][0]) * a[0] / sc[0], sim.bounds_min[1] + (j + basis[b][1]) * a[1] / sc[1],
                       sim.bounds_min[2] + (k + basis[b][2]) * a[2] / sc[2]};
        if (g[0] < lo[0] || g[0] >= hi[0] || g[1] < lo[1] || g[1] >= hi[1] || g[2] < lo[2] || g[2] >= hi[2]) continue;
        sim.hx.push_back(g[0]); sim.hy.push_back(g[1]); sim.hz.push_back(g[2]);
        sim.htype.push_back(tid[b]);
        sim.hid.push_back(uint64_t(((i * rep[1] + j) * rep[2] + k) * (long long)nb + (long long)b));
      }
    sim.hvx.assign(sim.hx.size(), 0.0); sim.hvy.assign(sim.hx.size(), 0.0); sim.hvz.assign(sim.hx.size(), 0.0);
    sim.staged_dirty = true;
  }
};
XSBH_REGISTER_OPERATOR("lattice", Lattice);

// gaussian_noise_r / gaussian_noise_v (ext exanb operators; slots sigma [, seed])
class GaussianNoise : public Operator {
public:
  bool velocity;
  explicit GaussianNoise(bool v) : velocity(v) {}
  void execute(Simulation& sim) override {
    TRACE(sim);
    check_slots({"sigma", "seed", "dt", "deterministic_noise", "region", "ghost"});
    if (!sim.staged_dirty && sim.grid_ready) throw OperatorError(name + ": particles are already on the device; apply noise inside setup_system");
    double sigma = quantity_slot("sigma", 0.0);
    uint64_t seed = (uint64_t)int_slot("seed", velocity ? 1234567 : 7654321);
    double inv[3] = {1, 1, 1};
    if (!velocity) for (int k = 0; k < 3; ++k) inv[k] = 1.0 / std::sqrt(sim.xform[k] * sim.xform[k] + sim.xform[3 + k] * sim.xform[3 + k] + sim.xform[6 + k] * sim.xform[6 + k]);
    for (size_t p = 0; p < sim.hx.size(); ++p) {
      double g[3]; gauss3(seed, sim.hid[p], g);
      if (velocity) { sim.hvx[p] += sigma * g[0]; sim.hvy[p] += sigma * g[1]; sim.hvz[p] += sigma * g[2]; }
      else { sim.hx[p] += sigma * g[0] * inv[0]; sim.hy[p] += sigma * g[1] * inv[1]; sim.hz[p] += sigma * g[2] * inv[2]; }
    }
  }
};
static OperatorRegistrar reg_noise_r("gaussian_noise_r", []() { return std::unique_ptr<Operator>(new GaussianNoise(false)); });
static OperatorRegistrar reg_noise_v("gaussian_noise_v", []() { return std::unique_ptr<Operator>(new GaussianNoise(true)); });

// read_xyz_file_with_xform (src/io/read_xyz_file_with_xform.cpp:140-146): slots filename, bounds_mode, read_velocities.
// Line 2 of the file carries the cell: `Lattice="ax ay az bx by bz cx cy cz"` (extended xyz) or three lengths.
class ReadXyz : public Operator {
public:
  void execute(Simulation& sim) override {
    TRACE(sim);
    check_slots({"filename", "file", "bounds_mode", "read_velocities", "enlarge_bounds", "pbc_adjust_xform", "adjust_bounds_to_particles"});
    const Node* fn = optional("filename"); if (!fn) fn = optional("file");
    if (!fn) throw OperatorError(name + ": required slot 'filename' is not set");
    XyzData d = read_xyz(sim.data_path(fn->as_string()), bool_slot("read_velocities", false));
    // orthorhombic part goes to the bounds, the rest to xform (H = xform * diag(L))
    double L[3];
    for (int k = 0; k < 3; ++k) L[k] = std::sqrt(d.cell[k] * d.cell[k] + d.cell[3 + k] * d.cell[3 + k] + d.cell[6 + k] * d.cell[6 + k]);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) sim.xform[3 * r + c] = d.cell[3 * r + c] / L[c];
    for (int k = 0; k < 3; ++k) { sim.bounds_min[k] = 0.0; sim.bounds_max[k] = L[k]; }
    int gd[3] = {0, 0, 0};
    DomainOp::finalize(sim, 0.0, gd);
    // grid-space positions = xform^-1 * physical
    double X[9], Xi[9]; std::copy(sim.xform, sim.xform + 9, X);
    double det = X[0] * (X[4] * X[8] - X[5] * X[7]) - X[1] * (X[3] * X[8] - X[5] * X[6]) + X[2] * (X[3] * X[7] - X[4] * X[6]);
    Xi[0] = (X[4] * X[8] - X[5] * X[7]) / det; Xi[1] = -(X[1] * X[8] - X[2] * X[7]) / det; Xi[2] = (X[1] * X[5] - X[2] * X[4]) / det;
    Xi[3] = -(X[3] * X[8] - X[5] * X[6]) / det; Xi[4] = (X[0] * X[8] - X[2] * X[6]) / det; Xi[5] = -(X[0] * X[5] - X[2] * X[3]) / det;
    Xi[6] = (X[3] * X[7] - X[4] * X[6]) / det; Xi[7] = -(X[0] * X[7] - X[1] * X[6]) / det; Xi[8] = (X[0] * X[4] - X[1] * X[3]) / det;
    for (size_t p = 0; p < d.x.size(); ++p) {
      int t = sim.species_index(d.species[p]);
      if (t < 0) { add_species(sim, d.species[p], nullptr); t = sim.species_index(d.species[p]); }
      double g[3] = {Xi[0] * d.x[p] + Xi[1] * d.y[p] + Xi[2] * d.z[p], Xi[3] * d.x[p] + Xi[4] * d.y[p] + Xi[5] * d.z[p], Xi[6] * d.x[p] + Xi[7] * d.y[p] + Xi[8] * d.z[p]};
      bool mine = true;
      for (int k = 0; k < 3; ++k) {
        double Lg = sim.bounds_max[k] - sim.bounds_min[k];
        if (sim.periodic[k]) { g[k] -= std::floor((g[k] - sim.bounds_min[k]) / Lg) * Lg; if (g[k] >= sim.bounds_max[k]) g[k] = sim.bounds_min[k]; }
        int c = std::min(sim.grid_dims[k] - 1, std::max(0, (int)std::floor((g[k] - sim.bounds_min[k]) / sim.cell_size)));
        int c0 = int((long long)sim.rank_coord[k] * sim.grid_dims[k] / sim.rank_dims[k]), c1 = int((long long)(sim.rank_coord[k] + 1) * sim.grid_dims[k] / sim.rank_dims[k]);
        if (c < c0 || c >= c1) mine = false;
      }
      if (!mine) continue;
      sim.hx.push_back(g[0]); sim.hy.push_back(g[1]); sim.hz.push_back(g[2]);
      sim.hvx.push_back(d.vx.empty() ? 0.0 : d.vx[p]); sim.hvy.push_back(d.vy.empty() ? 0.0 : d.vy[p]); sim.hvz.push_back(d.vz.empty() ? 0.0 : d.vz[p]);
      sim.htype.push_back((uint8_t)t); sim.hid.push_back(p);
    }
    sim.staged_dirty = true;
  }
};
XSBH_REGISTER_OPERATOR("read_xyz_file_with_xform", ReadXyz);
static OperatorRegistrar reg_read_xyz("read_xyz_file", []() { return std::unique_ptr<Operator>(new ReadXyz()); });

// replicate_domain { repeat: [nx, ny, nz] } (ext exanb operator used by the snap decks): tiles the staged particles
class ReplicateDomain : public Operator {
public:
  void execute(Simulation& sim) override {
    TRACE(sim);
    check_slots({"repeat"});
    if (!sim.staged_dirty) throw OperatorError("replicate_domain: must follow a reader / lattice inside setup_system");
    if (sim.nranks > 1) throw OperatorError("replicate_domain: single-rank only (readers keep the particles of the local brick)");
    const Node& r = required("repeat");
    long long rep[3]; for (int k = 0; k < 3; ++k) rep[k] = r[k].as_int();
    const size_t n0 = sim.hx.size();
    double L[3]; for (int k = 0; k < 3; ++k) L[k] = sim.bounds_max[k] - sim.bounds_min[k];
    for (long long i = 0; i < rep[0]; ++i) for (long long j = 0; j < rep[1]; ++j) for (long long k = 0; k < rep[2]; ++k) {
      if (i == 0 && j == 0 && k == 0) continue;
      const uint64_t img = uint64_t((i * rep[1] + j) * rep[2] + k);
      for (size_t p = 0; p < n0; ++p) {
        sim.hx.push_back(sim.hx[p] + i * L[0]); sim.hy.push_back(sim.hy[p] + j * L[1]); sim.hz.push_back(sim.hz[p] + k * L[2]);
        sim.hvx.push_back(sim.hvx[p]); sim.hvy.push_back(sim.hvy[p]); sim.hvz.push_back(sim.hvz[p]);
        sim.htype.push_back(sim.htype[p]); sim.hid.push_back(sim.hid[p] + img * n0);
      }
    }
    for (int k = 0; k < 3; ++k) sim.bounds_max[k] = sim.bounds_min[k] + rep[k] * L[k];
    int gd[3] = {0, 0, 0};
    DomainOp::finalize(sim, 0.0, gd);
  }
};
XSBH_REGISTER_OPERATOR("replicate_domain", ReplicateDomain);

// staged host particles -> device grid.  Not a reference operator: upstream readers insert straight into the grid.
class PlaceParticles : public Operator {
public:
  void execute(Simulation& sim) override {
    TRACE(sim);
    if (sim.staged_dirty) { need_gpu(sim, name); sim.flush_staged(); }
  }
};
XSBH_REGISTER_OPERATOR("place_particles", PlaceParticles);

// ================================================================================================ neighbours, ghosts
// nbh_dist (ext; main-config.msp:66-74): nbh_dist_lab = rcut_max + rcut_inc, max_displ = rcut_inc / 2
class NbhDist : public Operator {
public:
  void execute(Simulation& sim) override {
    TRACE(sim);
    sim.preinit = false;
    sim.nbh_dist = sim.rcut_max + sim.rcut_inc;
    sim.max_displ = 0.5 * sim.rcut_inc;
    if (bool_slot("verbose", false) && sim.rank == 0 && sim.verbosity > 0)
      std::printf("rcut_max = %.6g ang, rcut_inc = %.6g ang, nbh_dist = %.6g ang, max_displ = %.6g ang\n", sim.rcut_max, sim.rcut_inc, sim.nbh_dist, sim.max_displ);
  }
};
XSBH_REGISTER_OPERATOR("nbh_dist", NbhDist);

// chunk_neighbors (config: data/config/config_move_particles.msp:54-61)
class ChunkNeighbors : public Operator {
public:
  void execute(Simulation& sim) override {
    TRACE(sim);
    check_slots({"config", "chunk_size", "nbh_dist_lab", "enable_cuda"});
    need_gpu(sim, name);
    xsb_chunk_neighbors_config c{1, 1, 1, 0, 1.05};
    if (const Node* cfg = optional("config")) {
      static const std::set<std::string> known = {"chunk_size", "build_particle_offset", "subcell_compaction", "free_scratch_memory", "scratch_mem_per_cell",
                                                  "stream_prealloc_factor", "random_access", "half_symmetric", "skip_ghosts", "dual_particle_offset"};
      for (auto& kv : cfg->map) if (!known.count(kv.first)) throw OperatorError("chunk_neighbors: unknown config entry '" + kv.first + "'");
      if (const Node* n = cfg->find("chunk_size")) c.chunk_size = (int)n->as_int();
      if (const Node* n = cfg->find("build_particle_offset")) c.build_particle_offset = n->as_bool();
      if (const Node* n = cfg->find("subcell_compaction")) c.subcell_compaction = n->as_bool();
      if (const Node* n = cfg->find("free_scratch_memory")) c.free_scratch_memory = n->as_bool();
      if (const Node* n = cfg->find("stream_prealloc_factor")) c.stream_prealloc_factor = n->as_double();
    }
    // chunk size must be a power of two (type_pair_rcut_neighbors.cpp:90-104)
    if (c.chunk_size < 1 || (c.chunk_size & (c.chunk_size - 1))) throw OperatorError("chunk_neighbors: chunk_size must be a power of two");
    sim.check(xsb_chunk_neighbors_build(sim.ctx, sim.nbh_dist, &c), "xsb_chunk_neighbors_build");
    sim.neighbors_ready = true;
  }
};
XSBH_REGISTER_OPERATOR("chunk_neighbors", ChunkNeighbors);

// move_particles + migrate_cell_particles (config_move_particles.msp:89-125)
class MoveParticles : public Operator {
public:
  void execute(Simulation& sim) override {
    TRACE(sim);
    need_gpu(sim, name);
    if (!sim.scheme_ready) return;          // fresh from xsb_particles_assign: already binned
    xsb_domain_desc d = sim.domain_desc();
    sim.check(xsb_particles_rebin(sim.ctx, &d), "xsb_particles_rebin");
    sim.scheme_ready = false; sim.neighbors_ready = false;
  }
};
XSBH_REGISTER_OPERATOR("move_particles", MoveParticles);
class MigrateCellParticles : public Operator {
public:
  void execute(Simulation& sim) override { TRACE(sim); }   // done by move_particles (xsb_particles_rebin migrates across ranks)
};
XSBH_REGISTER_OPERATOR("migrate_cell_particles", MigrateCellParticles);

class BackupR : public Operator {
public:
  void execute(Simulation& sim) override {
    TRACE(sim);
    need_gpu(sim, name);
    // the reference backs positions up before the ghosts exist; own-particle order does not change afterwards
    sim.flags["backup_pending"] = true;
  }
};
XSBH_REGISTER_OPERATOR("backup_r", BackupR);

class GhostCommScheme : public Operator {
public:
  void execute(Simulation& sim) override {
    TRACE(sim);
    need_gpu(sim, name);
    xsb_domain_desc d = sim.domain_desc();
    sim.check(xsb_ghost_comm_scheme(sim.ctx, &d), "xsb_ghost_comm_scheme");
    sim.scheme_ready = true;
    sim.flags["ghosts_fresh"] = true;
    if (sim.flags["backup_pending"]) { sim.check(xsb_backup_r(sim.ctx), "xsb_backup_r"); sim.flags["backup_pending"] = false; }
  }
};
XSBH_REGISTER_OPERATOR("ghost_comm_scheme", GhostCommScheme);

// ghost_update_r / ghost_update_all_no_fv / ghost_update_opt (src/mpi/update_ghosts.cu:30-47)
class GhostUpdate : public Operator {
public:
  uint32_t mask; bool all;
  GhostUpdate(uint32_t m, bool a) : mask(m), all(a) {}
  void execute(Simulation& sim) override {
    TRACE(sim);
    check_slots({"opt_fields", "gpu_buffer_pack", "async_buffer_pack", "staging_buffer", "serialize_pack_send", "wait_all", "mpi_tag", "device_side_buffer"});
    if (sim.preinit) return;      // compute_force run on the empty grid to publish rcut_max (main-config.msp:52-74)
    need_gpu(sim, name);
    if (!sim.scheme_ready) throw OperatorError(name + ": ghost_comm_scheme has not run");
    uint32_t m = mask;
    if (const Node* f = optional("opt_fields")) for (auto& n : f->seq) m |= field_bit(n.as_string());
    if (all && sim.flags["ghosts_fresh"]) { sim.flags["ghosts_fresh"] = false; return; }   // the scheme call already copied r, v, type, id
    if (m) sim.check(xsb_ghost_update(sim.ctx, m), "xsb_ghost_update");
  }
};
static const uint32_t kR = (1u << XSB_F_RX) | (1u << XSB_F_RY) | (1u << XSB_F_RZ);
static OperatorRegistrar reg_gur("ghost_update_r", []() { return std::unique_ptr<Operator>(new GhostUpdate(kR, false)); });
static OperatorRegistrar reg_gua("ghost_update_all_no_fv", []() { return std::unique_ptr<Operator>(new GhostUpdate(kR | (1u << XSB_F_VX) | (1u << XSB_F_VY) | (1u << XSB_F_VZ), true)); });
static OperatorRegistrar reg_gual("ghost_update_all", []() { return std::unique_ptr<Operator>(new GhostUpdate(kR | (1u << XSB_F_VX) | (1u << XSB_F_VY) | (1u << XSB_F_VZ), true)); });
static OperatorRegistrar reg_guo("ghost_update_opt", []() { return std::unique_ptr<Operator>(new GhostUpdate(0, false)); });
static OperatorRegistrar reg_gue("ghost_update_emb", []() { return std::unique_ptr<Operator>(new GhostUpdate(1u << XSB_F_RHO_DEMB, false)); });

// update_force_energy_from_ghost / update_opt_from_ghost (src/mpi/update_from_ghosts.cu:29-48)
class UpdateFromGhosts : public Operator {
public:
  uint32_t mask;
  explicit UpdateFromGhosts(uint32_t m) : mask(m) {}
  void execute(Simulation& sim) override {
    TRACE(sim);
    if (sim.preinit) return;
    need_gpu(sim, name);
    uint32_t m = mask;
    if (const Node* f = optional("opt_fields")) for (auto& n : f->seq) m |= field_bit(n.as_string());
    if (m) sim.check(xsb_ghost_reduce_add(sim.ctx, m), "xsb_ghost_reduce_add");
  }
};
static OperatorRegistrar reg_ufg("update_force_energy_from_ghost", []() {
  return std::unique_ptr<Operator>(new UpdateFromGhosts((1u << XSB_F_FX) | (1u << XSB_F_FY) | (1u << XSB_F_FZ) | (1u << XSB_F_EP)));
});
// update_virial_force_energy_from_ghost (src/mpi/update_from_ghosts.cu:43): + the 9 virial components
static OperatorRegistrar reg_uvfg("update_virial_force_energy_from_ghost", []() {
  return std::unique_ptr<Operator>(new UpdateFromGhosts((1u << XSB_F_FX) | (1u << XSB_F_FY) | (1u << XSB_F_FZ) | (1u << XSB_F_EP) | (1u << XSB_F_VIRIAL)));
});
static OperatorRegistrar reg_uog("update_opt_from_ghost", []() { return std::unique_ptr<Operator>(new UpdateFromGhosts(0)); });

// particle_displ_over (config_move_particles.msp:19-23): result = max displacement since backup_r > threshold
class ParticleDisplOver : public Operator {
public:
  void execute(Simulation& sim) override {
    TRACE(sim);
    need_gpu(sim, name);
    int over = 0; double dmax = 0.0;
    double thr = quantity_slot("threshold", sim.max_displ);
    if (const Node* t = optional("threshold")) if (t->is_scalar() && t->as_string() == "max_displ") thr = sim.max_displ;
    sim.check(xsb_particle_displ_over(sim.ctx, thr, &over, &dmax), "xsb_particle_displ_over");
    sim.flags["trigger_move_particles"] = over != 0;
    sim.flags["move_flag"] = over != 0;
  }
};
XSBH_REGISTER_OPERATOR("particle_displ_over", ParticleDisplOver);

// ================================================================================================ per-particle steps
class ZeroForceEnergy : public Operator {   // src/compute/zero_force_energy.cu:98-135, slot ghost
public:
  void execute(Simulation& sim) override {
    TRACE(sim);
    check_slots({"ghost"});
    if (sim.preinit) return;
    need_gpu(sim, name);
    sim.check(xsb_zero_force_energy(sim.ctx, bool_slot("ghost", false)), "xsb_zero_force_energy");
  }
};
XSBH_REGISTER_OPERATOR("zero_force_energy", ZeroForceEnergy);

class ForceToAccel : public Operator {      // src/compute/force_to_accel.cu:82-101
public:
  void execute(Simulation& sim) override {
    TRACE(sim);
    if (sim.preinit) return;
    need_gpu(sim, name);
    auto m = masses(sim);
    sim.check(xsb_force_to_accel(sim.ctx, (int)m.size(), m.data()), "xsb_force_to_accel");
  }
};
XSBH_REGISTER_OPERATOR("force_to_accel", ForceToAccel);

class PushFVR : public Operator {            // ext; config_numerical_schemes.msp:23-27, slots dt_scale, xform_mode
public:
  bool with_r;
  explicit PushFVR(bool r) : with_r(r) {}
  void execute(Simulation& sim) override {
    TRACE(sim);
    check_slots({"dt_scale", "xform_mode", "dt"});
    need_gpu(sim, name);
    double dt = sim.dt * quantity_slot("dt_scale", 1.0);
    if (with_r) sim.check(xsb_push_f_v_r(sim.ctx, dt), "xsb_push_f_v_r");
    else sim.check(xsb_push_f_v(sim.ctx, dt), "xsb_push_f_v");
  }
};
static OperatorRegistrar reg_pfvr("push_f_v_r", []() { return std::unique_ptr<Operator>(new PushFVR(true)); });
static OperatorRegistrar reg_pfv("push_f_v", []() { return std::unique_ptr<Operator>(new PushFVR(false)); });

// ================================================================================================ force operators
// ---- <pot>_compute_force / <pot>_multi_force for the pair potentials of the C ABI (xsb_pair_pot) --------------------
// slots: pair_potential_impl.hxx:104-122; parameter maps: lennard_jones.h:60-71 {epsilon, sigma}, zbl/potential.h:64-79
// {r1, rc} (+ the atomic numbers of the pair from `species`, USTAMP_POTENTIAL_PAIR_PARAMS_EXTRACT zbl/potential.h:312),
// exp6.h:44-57 {A, B, C, D}, buckingham.h:60-70 {A, Rho, C}.
struct PairPotDesc { int pot; std::vector<const char*> names; bool needs_z; };
static const PairPotDesc& pot_desc(int pot) {
  // yukawa.h:56-70 {A, kappa}; relax/potential.h:56-72 {r1, rc}; zero/potential.h:37-45 (no parameters)
  static const PairPotDesc d[7] = {{XSB_POT_LJ, {"epsilon", "sigma"}, false}, {XSB_POT_ZBL, {"r1", "rc"}, true},
                                   {XSB_POT_EXP6, {"A", "B", "C", "D"}, false}, {XSB_POT_BUCKINGHAM, {"A", "Rho", "C"}, false},
                                   {XSB_POT_YUKAWA, {"A", "kappa"}, false}, {XSB_POT_RELAX, {"r1", "rc"}, false}, {XSB_POT_ZERO, {}, false}};
  return d[pot];
}
// raw parameter vector in C-ABI order; absent entries of `common_parameters` default to 0 like the reference's structs
static std::vector<double> pot_params(const std::string& op, const PairPotDesc& d, const Node* map, bool required, double za, double zb) {
  std::vector<double> v;
  for (const char* n : d.names) {
    const Node* x = map && map->is_map() ? map->find(n) : nullptr;
    if (!x && required) throw OperatorError(op + ": parameter '" + n + "' is missing");
    v.push_back(x ? quantity(*x) : 0.0);
  }
  if (d.needs_z) { v.push_back(za); v.push_back(zb); }
  return v;
}

class PairComputeForce : public Operator {
public:
  int pot;
  explicit PairComputeForce(int p) : pot(p) {}
  void execute(Simulation& sim) override {
    TRACE(sim);
    check_slots({"parameters", "rcut", "rcut_max", "chunk_neighbors", "species", "type", "ghost", "grid", "domain", "compact_nbh_weight", "enable_pair_weights", "particle_locks"});
    const PairPotDesc& d = pot_desc(pot);
    const double rcut = quantity(required("rcut"));
    const Node& p = required("parameters");
    sim.rcut_max = std::max(sim.rcut_max, rcut);       // IN_OUT slot rcut_max (pair_potential_impl.hxx:131-140)
    int t = -1;
    if (optional("type")) {
      t = sim.species_index(required("type").as_string());
      if (t < 0 && !sim.preinit) throw OperatorError(name + ": unknown species '" + required("type").as_string() + "'");
    }
    const double z = sim.species.empty() ? 0.0 : sim.species[t >= 0 ? t : 0].z;     // single-material operator: the pair is (species, species)
    std::vector<double> prm = pot_params(name, d, &p, true, z, z);
    if (sim.preinit) return;
    need_gpu(sim, name);
    const int fl = force_flags(sim, bool_slot("ghost", false)) | (sim.mixed_precision ? XSB_FLAG_MIXED : 0);
    if (t >= 0) {
      // slot `type` restricts the operator to one species (pair_potential_impl.hxx:143-158): a multi table with one live pair
      const int nt = (int)std::max<size_t>(1, sim.species.size());
      const size_t w = prm.size() + 1;
      std::vector<double> rows(size_t(nt) * (nt + 1) / 2 * w, 0.0);
      size_t id = size_t(t) * (t + 1) / 2 + t;
      std::copy(prm.begin(), prm.end(), rows.begin() + id * w); rows[id * w + prm.size()] = rcut;
      sim.check(xsb_pair_multi_force(sim.ctx, pot, nt, rows.data(), (int)prm.size(), rcut, fl), "xsb_pair_multi_force");
      return;
    }
    sim.check(xsb_pair_force(sim.ctx, pot, prm.data(), (int)prm.size(), rcut, fl), "xsb_pair_force");
  }
};

// <pot>_multi_force (pair_potential_force_op_multiparam.h:249-284 YAML; table build pair_potential_impl.hxx:209-368)
class PairMultiForce : public Operator {
public:
  int pot;
  explicit PairMultiForce(int p) : pot(p) {}
  void execute(Simulation& sim) override {
    TRACE(sim);
    check_slots({"parameters", "common_parameters", "rcut", "rcut_max", "chunk_neighbors", "species", "ghost", "grid", "domain", "compact_nbh_weight", "enable_pair_weights"});
    const PairPotDesc& d = pot_desc(pot);
    const double rcut = quantity(required("rcut"));
    const Node& list = required("parameters");
    if (!list.is_seq()) throw OperatorError(name + ": `parameters` must be a list of { type_a, type_b, rcut, parameters }");
    double rmax = rcut;
    for (auto& e : list.seq) rmax = std::max(rmax, quantity_or(e.find("rcut"), rcut));
    sim.rcut_max = std::max(sim.rcut_max, rmax);
    if (sim.preinit) return;
    need_gpu(sim, name);
    const int nt = (int)sim.species.size();
    if (nt < 1) throw OperatorError(name + ": no species defined");
    const size_t np = size_t(nt) * (nt + 1) / 2, w = d.names.size() + (d.needs_z ? 2 : 0) + 1;
    std::vector<double> rows(np * w);
    // pairs without user parameters use common_parameters and the operator rcut (pair_potential_impl.hxx:295-321)
    for (int hi = 0; hi < nt; ++hi) for (int lo = 0; lo <= hi; ++lo) {
      std::vector<double> prm = pot_params(name, d, optional("common_parameters"), false, sim.species[lo].z, sim.species[hi].z);
      size_t id = size_t(hi) * (hi + 1) / 2 + lo;
      std::copy(prm.begin(), prm.end(), rows.begin() + id * w); rows[id * w + w - 1] = rcut;
    }
    for (auto& e : list.seq) {
      int a = sim.species_index(e["type_a"].as_string()), b = sim.species_index(e["type_b"].as_string());
      if (a < 0 || b < 0) throw OperatorError(name + ": unknown species in pair " + e["type_a"].as_string() + "/" + e["type_b"].as_string());
      int hi = std::max(a, b), lo = std::min(a, b);
      size_t id = size_t(hi) * (hi + 1) / 2 + lo;            // unique_pair_id (ext, symmetric triangular index)
      std::vector<double> prm = pot_params(name, d, &e["parameters"], true, sim.species[lo].z, sim.species[hi].z);
      std::copy(prm.begin(), prm.end(), rows.begin() + id * w); rows[id * w + w - 1] = quantity_or(e.find("rcut"), rcut);
    }
    sim.check(xsb_pair_multi_force(sim.ctx, pot, nt, rows.data(), (int)w - 1, rmax, force_flags(sim, bool_slot("ghost", false)) | (sim.mixed_precision ? XSB_FLAG_MIXED : 0)),
              "xsb_pair_multi_force");
  }
};
// <pot>_compute_force_symetric (pair_potential_singlemat_symetric.cpp:335-346): the reference walks half lists and scatters -f
// to the neighbour under particle locks, then folds ghost forces back (config_update_symmetric_forces.msp).  Here the same
// totals come from the full-list kernel (one writer per atom, nothing lands on ghosts), so the surrounding zero-ghost /
// update_force_energy_from_ghost nodes of those decks add zeros.
#define XSBH_PAIR_OPS(potname, POT) \
  static OperatorRegistrar reg_##potname##_cf(#potname "_compute_force", []() { return std::unique_ptr<Operator>(new PairComputeForce(POT)); }); \
  static OperatorRegistrar reg_##potname##_sy(#potname "_compute_force_symetric", []() { return std::unique_ptr<Operator>(new PairComputeForce(POT)); }); \
  static OperatorRegistrar reg_##potname##_mf(#potname "_multi_force", []() { return std::unique_ptr<Operator>(new PairMultiForce(POT)); });
XSBH_PAIR_OPS(lj, XSB_POT_LJ)
XSBH_PAIR_OPS(zbl, XSB_POT_ZBL)
XSBH_PAIR_OPS(exp6, XSB_POT_EXP6)
XSBH_PAIR_OPS(buckingham, XSB_POT_BUCKINGHAM)
XSBH_PAIR_OPS(yukawa, XSB_POT_YUKAWA)
XSBH_PAIR_OPS(relax, XSB_POT_RELAX)
XSBH_PAIR_OPS(zero, XSB_POT_ZERO)

// johnson_force / johnson_emb / johnson_force_reuse_emb / johnson_init (eam_potential.cu:92-100,178-193; johnson.h:176-204)
// The same operator class serves the other analytic single-species models of eam_potential_template: sutton_chen
// (sutton_chen.h:71-83 {c, epsilon, a0, n, m}) and vniitf (vniitf.h:139-157, 13 scalars); only the parameter names differ.
class JohnsonForce : public Operator {
public:
  int phases;     // bit0 emb, bit1 emb over ghosts, bit2 force
  int model;      // xsb_eam_model
  explicit JohnsonForce(int ph, int m = XSB_EAM_JOHNSON) : phases(ph), model(m) {}
  void execute(Simulation& sim) override {
    TRACE(sim);
    check_slots({"parameters", "rcut", "rcut_max", "ghost_dist_max", "chunk_neighbors", "grid", "domain", "eam_extra_fields"});
    const double rcut = quantity(required("rcut"));
    const Node& p = required("parameters");
    static const std::vector<const char*> all_names[3] = {
      {"re", "fe", "rhoe", "alpha", "beta", "A", "B", "kappa", "lambda", "Fn0", "Fn1", "Fn2", "Fn3", "F0", "F1", "F2", "F3", "Fo", "eta"},
      {"c", "epsilon", "a0", "n", "m"},
      {"rmax", "rmin", "rt0", "Ecoh", "E0", "beta", "A", "Z", "n", "alpha", "D", "eta", "mu"}};
    const std::vector<const char*>& names = all_names[model];
    double prm[19];
    for (size_t i = 0; i < names.size(); ++i) {
      const Node* v = p.find(names[i]);
      if (!v) throw OperatorError(name + ": parameter '" + names[i] + "' is missing");
      prm[i] = quantity(*v);
    }
    sim.rcut_max = std::max(sim.rcut_max, rcut);
    sim.ghost_dist_max = std::max(sim.ghost_dist_max, 2.0 * rcut);    // ComputeGhostEmb (eam_potential.cu:109-112)
    if (sim.preinit || phases == 0) return;
    need_gpu(sim, name);
    int fl = sim.compute_virial && sim.trigger_thermo_state ? XSB_FLAG_VIRIAL : 0;
    if (sim.mixed_precision) fl |= XSB_FLAG_MIXED;      // xsb extension (global `enable_mixed_precision`): FP32 rho(r), phi(r), tolerance 1e-5
    sim.check(xsb_eam_analytic_force(sim.ctx, model, prm, int(names.size()), rcut, phases, fl), "xsb_eam_analytic_force");
  }
};
static OperatorRegistrar reg_jf("johnson_force", []() { return std::unique_ptr<Operator>(new JohnsonForce(7)); });
static OperatorRegistrar reg_je("johnson_emb", []() { return std::unique_ptr<Operator>(new JohnsonForce(3)); });
static OperatorRegistrar reg_jr("johnson_force_reuse_emb", []() { return std::unique_ptr<Operator>(new JohnsonForce(4)); });
static OperatorRegistrar reg_ji("johnson_init", []() { return std::unique_ptr<Operator>(new JohnsonForce(0)); });
#define XSBH_EAM1_OPS(nm, MODEL) \
  static OperatorRegistrar reg_##nm##_f(#nm "_force", []() { return std::unique_ptr<Operator>(new JohnsonForce(7, MODEL)); }); \
  static OperatorRegistrar reg_##nm##_e(#nm "_emb", []() { return std::unique_ptr<Operator>(new JohnsonForce(3, MODEL)); }); \
  static OperatorRegistrar reg_##nm##_r(#nm "_force_reuse_emb", []() { return std::unique_ptr<Operator>(new JohnsonForce(4, MODEL)); }); \
  static OperatorRegistrar reg_##nm##_i(#nm "_init", []() { return std::unique_ptr<Operator>(new JohnsonForce(0, MODEL)); });
XSBH_EAM1_OPS(sutton_chen, XSB_EAM_SUTTON_CHEN)
XSBH_EAM1_OPS(vniitf, XSB_EAM_VNIITF)

// eam_alloy_force / eam_alloy_init (eam_potential_multimat.cu:88-109 slots; eam_alloy.cpp:66-84 parameters)
class EamAlloyForce : public Operator {
public:
  bool init_only;
  explicit EamAlloyForce(bool i) : init_only(i) {}
  void execute(Simulation& sim) override {
    TRACE(sim);
    check_slots({"species", "parameters", "types", "rcut", "rcut_max", "ghost_dist_max", "chunk_neighbors", "grid", "domain", "trigger_thermo_state", "compute_virial",
                 "eam_rho", "eam_rho2emb", "eam_ghost", "eam_force", "eam_symmetry", "particle_locks", "eam_extra_fields"});
    const double rcut = quantity(required("rcut"));
    const bool rho = bool_slot("eam_rho", true), r2e = bool_slot("eam_rho2emb", true), ghost = bool_slot("eam_ghost", true), force = bool_slot("eam_force", true);
    // eam_symmetry=true (half lists + locks upstream) is computed with the same full-list kernels: rho and forces of
    // owned atoms are complete without ghost contributions, so the update_*_from_ghost nodes of the _sym graphs add zeros
    sim.rcut_max = std::max(sim.rcut_max, rcut);
    if ((rho || force) && ghost) sim.ghost_dist_max = std::max(sim.ghost_dist_max, 2.0 * rcut);   // eam_potential_multimat.cu:116-120
    const Node& p = required("parameters");
    std::string file = p.is_scalar() ? p.as_string() : p["file"].as_string();
    if (sim.preinit) return;
    need_gpu(sim, name);
    std::string path = sim.data_path(file);
    // the tables live in the context: the instances of one graph (eam_alloy_init, eam_rho, eam_rho2emb, eam_force) share them;
    // uploading the same file again would also drop the per-pair cache the rho phase left for the force phase
    if (sim.eam_alloy_loaded != path) {
      xsb_eam_alloy_tables t{};
      char names[512];
      int rc = xsb_eam_alloy_read(path.c_str(), &t, names, sizeof(names));
      if (rc != XSB_OK) throw OperatorError(name + ": cannot read setfl file '" + path + "'");
      // type i <-> element i of the file (eam_alloy.h:95-98 maps types to elements in order)
      std::istringstream is(names); std::string el; int k = 0;
      while (is >> el) { if (k < (int)sim.species.size() && sim.species[k].name != el && sim.rank == 0 && sim.verbosity > 0)
                            std::fprintf(stderr, "warning: %s: species %d is '%s' but element %d of %s is '%s'\n", name.c_str(), k, sim.species[k].name.c_str(), k, file.c_str(), el.c_str());
                         ++k; }
      rc = xsb_eam_alloy_set(sim.ctx, &t);
      xsb_eam_alloy_free(&t);
      sim.check(rc, "xsb_eam_alloy_set");
      sim.eam_alloy_loaded = path;
    }
    if (init_only) return;
    int phases = (rho ? XSB_EAM_RHO : 0) | (r2e ? XSB_EAM_RHO2EMB : 0) | (ghost ? XSB_EAM_GHOST : 0) | (force ? XSB_EAM_FORCE : 0);
    bool eflag = optional("trigger_thermo_state") ? bool_slot("trigger_thermo_state", true) : sim.trigger_thermo_state;   // eam_potential_multimat.cu:125-149
    if (eflag) phases |= XSB_EAM_EFLAG;
    int fl = eflag && (bool_slot("compute_virial", false) || sim.compute_virial) ? XSB_FLAG_VIRIAL : 0;
    if (sim.mixed_precision) fl |= XSB_FLAG_MIXED;      // xsb extension (global `enable_mixed_precision`): FP32 spline + pair math, tolerance 1e-5
    sim.check(xsb_eam_alloy_force(sim.ctx, rcut, phases, fl), "xsb_eam_alloy_force");
  }
};
static OperatorRegistrar reg_eaf("eam_alloy_force", []() { return std::unique_ptr<Operator>(new EamAlloyForce(false)); });
static OperatorRegistrar reg_eai("eam_alloy_init", []() { return std::unique_ptr<Operator>(new EamAlloyForce(true)); });

// snap_force (snap/snap_force.cu:26-36; parameters { nt, param, coef } used at snaplmp.cpp:69,115-121)
class SnapForce : public Operator {
public:
  bool configured = false;
  double rcut = 0.0;
  bool fp32 = false;          // snap_force_fp32: the reference's SNAP_FP32_MATH plugin (snap/snap_force.cu:25-29) = XSB_FLAG_MIXED
  explicit SnapForce(bool f32 = false) : fp32(f32) {}
  void execute(Simulation& sim) override {
    TRACE(sim);
    check_slots({"parameters", "rcut_max", "chunk_neighbors", "ghost", "grid", "domain", "bispectrumchkfile", "conv_coef_units", "trigger_thermo_state", "species", "particle_locks"});
    const Node& p = required("parameters");
    if (!configured || sim.preinit) {
      SnapFiles sf = read_snap_files(sim.data_path(p["param"].as_string()), sim.data_path(p["coef"].as_string()));
      if (const Node* nt = p.find("nt")) if (nt->as_int() != (long long)sf.elements.size()) throw OperatorError(name + ": parameters.nt does not match the number of elements in the coefficient file");
      // coefficient blocks follow the species order of the simulation when names match, else file order
      std::vector<int> order(sf.elements.size());
      for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
      if (sim.species.size() >= sf.elements.size()) {
        bool all = true; std::vector<int> o2(sf.elements.size(), -1);
        for (size_t e = 0; e < sf.elements.size(); ++e) { int s = sim.species_index(sf.elements[e]); if (s < 0 || s >= (int)sf.elements.size()) all = false; else o2[s] = (int)e; }
        if (all) order = o2;
      }
      const int ncoef = xsb_snap_ncoeff(sf.twojmax);
      if (ncoef < 0) throw OperatorError(name + ": twojmax " + std::to_string(sf.twojmax) + " is outside 0..8");
      if (sf.ncoeff_all != ncoef + 1) throw OperatorError(name + ": coefficient count " + std::to_string(sf.ncoeff_all) + " does not match twojmax (linear SNAP expects " + std::to_string(ncoef + 1) + ")");
      std::vector<double> rad, wj, beta;
      const double conv = quantity_slot("conv_coef_units", kEv);      // coefficients are in eV (snap_force_op.h:77)
      for (int e : order) {
        rad.push_back(sf.radelem[e]); wj.push_back(sf.wjelem[e]);
        for (int k = 0; k <= ncoef; ++k) beta.push_back(sf.beta[size_t(e) * (ncoef + 1) + k] * conv);
      }
      xsb_snap_params sp{};
      sp.twojmax = sf.twojmax; sp.switchflag = sf.switchflag; sp.bzeroflag = sf.bzeroflag; sp.nelements = (int)order.size();
      sp.quadraticflag = sf.quadraticflag; sp.chemflag = sf.chemflag; sp.switchinnerflag = 0;
      sp.rfac0 = sf.rfac0; sp.rmin0 = sf.rmin0; sp.rcutfac = sf.rcutfac;
      sp.radelem = rad.data(); sp.wjelem = wj.data(); sp.beta = beta.data();
      double rmax = 0.0; for (double r : rad) rmax = std::max(rmax, r);
      rcut = 2.0 * rmax * sf.rcutfac;
      if (sim.ctx) { sim.check(xsb_snap_set(sim.ctx, &sp), "xsb_snap_set"); configured = true; }
    }
    sim.rcut_max = std::max(sim.rcut_max, rcut);
    if (sim.preinit) return;
    need_gpu(sim, name);
    sim.check(xsb_snap_force(sim.ctx, force_flags(sim, bool_slot("ghost", false)) | ((fp32 || sim.mixed_precision) ? XSB_FLAG_MIXED : 0)), "xsb_snap_force");
    int ovf = 0;
    sim.check(xsb_snap_overflow(sim.ctx, &ovf), "xsb_snap_overflow");
    if (ovf) throw OperatorError(name + ": an atom has more in-range neighbours than the kernel's capacity");
  }
};
XSBH_REGISTER_OPERATOR("snap_force", SnapForce);
// snaplmp_force (snaplmp.cpp:59-360) is the same operator computed through LAMMPS' sna.cpp upstream: same slots, same files
static OperatorRegistrar reg_snaplmp("snaplmp_force", []() { return std::unique_ptr<Operator>(new SnapForce()); });
// snap_force_fp32 (potentials/snap/multi_WBe_fp32.msp:31,60): the FP32-arithmetic build of the same operator
static OperatorRegistrar reg_snapf32("snap_force_fp32", []() { return std::unique_ptr<Operator>(new SnapForce(true)); });

// ================================================================================================ thermodynamic state, loop control
class TriggerThermoState : public Operator {     // config_thermostate.msp: screen frequency trigger
public:
  void execute(Simulation& sim) override {
    TRACE(sim);
    bool t = bool_slot("force", false);
    long long f = sim.thermo_screen_frequency;
    if (f > 0 && sim.timestep % f == 0) t = true;
    if (sim.timestep == sim.max_iteration) t = true;
    sim.trigger_thermo_state = t;
    sim.flags["trigger_thermo_state"] = t;
  }
};
XSBH_REGISTER_OPERATOR("trigger_thermo_state", TriggerThermoState);

class ThermodynamicStateOp : public Operator {   // src/thermo_state/simulation_thermodynamic_state.cpp:73-230
public:
  void execute(Simulation& sim) override {
    TRACE(sim);
    check_slots({"potential_energy_shift"});
    need_gpu(sim, name);
    auto m = masses(sim);
    double t[27];
    sim.check(xsb_thermo_state(sim.ctx, (int)m.size(), m.data(), t), "xsb_thermo_state");
    ThermoState& s = sim.thermo;
    for (int a = 0; a < 3; ++a) { s.virial_diag[a] = t[4 * a]; s.ke_tensor[a] = 2.0 * t[9 + 4 * a]; s.momentum[a] = t[18 + a]; }
    s.kinetic = t[21] + t[22] + t[23];
    s.potential = t[24] + quantity_slot("potential_energy_shift", 0.0);
    s.mass = t[25]; s.natoms = (uint64_t)t[26];
    const double* X = sim.xform;
    double det = X[0] * (X[4] * X[8] - X[5] * X[7]) - X[1] * (X[3] * X[8] - X[5] * X[6]) + X[2] * (X[3] * X[7] - X[4] * X[6]);
    s.volume = det;
    for (int a = 0; a < 3; ++a) s.volume *= sim.bounds_max[a] - sim.bounds_min[a];
  }
};
XSBH_REGISTER_OPERATOR("simulation_thermodynamic_state", ThermodynamicStateOp);

class PrintThermodynamicState : public Operator {   // src/io/print_thermodynamic_state.cpp (log_mode default columns)
public:
  void execute(Simulation& sim) override {
    TRACE(sim);
    if (sim.rank != 0 || sim.verbosity <= 0) return;
    const ThermoState& s = sim.thermo;
    if (bool_slot("print_header", false)) std::printf("%10s %14s %20s %20s %20s %14s %16s %s\n", "Step", "Time (ps)", "Tot. E. (eV)", "Kin. E. (eV)", "Pot. E. (eV)", "Temp. (K)", "Pressure (Pa)", "N");
    double p = 0.0;   // hydrostatic pressure = mean of (2 Ek_a (c.o.m. removed) + W_aa) / V  (simulation_thermodynamic_state.cpp:202-224)
    for (int a = 0; a < 3; ++a) p += ((s.ke_tensor[a] - s.momentum[a] * s.momentum[a] / (s.mass > 0 ? s.mass : 1.0)) + s.virial_diag[a]) / s.volume;
    p /= 3.0;
    const double pa = p * kInternalEnergyJ / 1e-30;
    std::printf("%10lld %14.6e %20.12e %20.12e %20.12e %14.6f %16.8e %llu\n", sim.timestep, sim.physical_time, s.total() / kEv, s.kinetic / kEv, s.potential / kEv,
                s.temperature(), pa, (unsigned long long)s.natoms);
    std::fflush(stdout);
  }
};
XSBH_REGISTER_OPERATOR("print_thermodynamic_state", PrintThermodynamicState);

class NextTimeStep : public Operator {
public:
  void execute(Simulation& sim) override { TRACE(sim); sim.timestep += 1; sim.physical_time += sim.dt; }
};
XSBH_REGISTER_OPERATOR("next_time_step", NextTimeStep);

class SimContinue : public Operator {             // main-config.msp:178-183
public:
  void execute(Simulation& sim) override { TRACE(sim); sim.flags["md_loop_continue"] = sim.timestep <= sim.max_iteration; }
};
XSBH_REGISTER_OPERATOR("sim_continue", SimContinue);

// dump_particles { file }: own particles (id, type, r, v, f, ep) in a flat binary file read by tests/test_host_decks.py.
// xsb extension standing in for check_values / write_dump_atoms of the reference control plane.
class DumpParticles : public Operator {
public:
  void execute(Simulation& sim) override {
    TRACE(sim);
    check_slots({"file", "filename"});
    need_gpu(sim, name);
    std::string file = string_slot("file", string_slot("filename", "particles.xsbdump"));
    if (sim.nranks > 1) file += "." + std::to_string(sim.rank);
    const uint64_t n = xsb_num_particles(sim.ctx), nc = xsb_num_cells(sim.ctx);
    std::vector<uint64_t> off(nc + 1), id(n); std::vector<uint8_t> ty(n);
    sim.check(xsb_cell_offsets_download(sim.ctx, off.data()), "xsb_cell_offsets_download");
    std::vector<std::vector<double>> f(11, std::vector<double>(n));
    const int fields[11] = {XSB_F_RX, XSB_F_RY, XSB_F_RZ, XSB_F_VX, XSB_F_VY, XSB_F_VZ, XSB_F_FX, XSB_F_FY, XSB_F_FZ, XSB_F_EP, XSB_F_RHO_DEMB};
    for (int k = 0; k < 11; ++k) sim.check(xsb_field_download(sim.ctx, fields[k], f[k].data()), "xsb_field_download");
    sim.check(xsb_field_download(sim.ctx, XSB_F_ID, id.data()), "xsb_field_download");
    sim.check(xsb_field_download(sim.ctx, XSB_F_TYPE, ty.data()), "xsb_field_download");
    // local grid dims: brick + ghost layers
    int dims[3];
    for (int a = 0; a < 3; ++a) dims[a] = int((long long)(sim.rank_coord[a] + 1) * sim.grid_dims[a] / sim.rank_dims[a]) - int((long long)sim.rank_coord[a] * sim.grid_dims[a] / sim.rank_dims[a]) + 2 * sim.ghost_layers;
    std::vector<uint64_t> own;
    for (uint64_t c = 0; c < nc; ++c) {
      int i = int(c % dims[0]), j = int((c / dims[0]) % dims[1]), k = int(c / (uint64_t(dims[0]) * dims[1])), gl = sim.ghost_layers;
      if (i < gl || i >= dims[0] - gl || j < gl || j >= dims[1] - gl || k < gl || k >= dims[2] - gl) continue;
      for (uint64_t p = off[c]; p < off[c + 1]; ++p) own.push_back(p);
    }
    std::ofstream o(file, std::ios::binary);
    if (!o) throw OperatorError(name + ": cannot write '" + file + "'");
    const char magic[8] = {'X', 'S', 'B', 'D', 'U', 'M', 'P', '1'};
    uint64_t hdr[2] = {own.size(), 11};
    double meta[16] = {sim.bounds_min[0], sim.bounds_min[1], sim.bounds_min[2], sim.bounds_max[0], sim.bounds_max[1], sim.bounds_max[2], sim.cell_size,
                       sim.xform[0], sim.xform[1], sim.xform[2], sim.xform[3], sim.xform[4], sim.xform[5], sim.xform[6], sim.xform[7], sim.xform[8]};
    o.write(magic, 8); o.write((const char*)hdr, sizeof(hdr)); o.write((const char*)meta, sizeof(meta));
    for (uint64_t p : own) o.write((const char*)&id[p], 8);
    for (uint64_t p : own) o.write((const char*)&ty[p], 1);
    for (int k = 0; k < 11; ++k) for (uint64_t p : own) o.write((const char*)&f[k][p], 8);
  }
};
XSBH_REGISTER_OPERATOR("dump_particles", DumpParticles);

}  // namespace
}  // namespace xsbh
