// xsb200-run -- runs an exaStamp-style YAML deck on the B200 path: deck -> operator graph -> libxsb200 C ABI.
// Stands where `onika-exec deck.msp` stands upstream, for the operators of the short-range force path.
//
//   xsb200-run deck.msp [--gpus N] [--set key.sub value]... [--data-dir DIR]... [--dry-run] [--trace] [--quiet]
//   xsb200-run --list-operators
//   xsb200-run --parse deck.msp            (prints the layered deck, no graph)
//   xsb200-run --quantity "0.0104 eV"      (prints the value in internal units)
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "xsbh_operator.h"

using namespace xsbh;

static void set_path(Node& deck, const std::string& path, const std::string& value) {
  Node* cur = &deck;
  size_t b = 0;
  for (;;) {
    size_t d = path.find('.', b);
    std::string key = path.substr(b, d == std::string::npos ? std::string::npos : d - b);
    if (d == std::string::npos) { cur->set(key, parse_yaml(value)); return; }
    Node* nxt = cur->find(key);
    if (!nxt || !nxt->is_map()) nxt = &cur->set(key, Node::make_map());
    cur = nxt; b = d + 1;
  }
}

static int run(const std::string& deck_path, const std::vector<std::pair<std::string, std::string>>& sets, std::vector<std::string> dirs, bool dry, bool trace, bool quiet) {
  size_t s = deck_path.find_last_of('/');
  dirs.insert(dirs.begin(), s == std::string::npos ? std::string(".") : deck_path.substr(0, s));
  Node deck = load_deck(deck_path, dirs);
  for (auto& kv : sets) set_path(deck, kv.first, kv.second);
  Simulation sim;
  sim.search_dirs = dirs;
  sim.tracing = trace || dry;
  sim.verbosity = quiet ? 0 : 1;
  sim.cuda_required = !dry;
  apply_globals(deck, sim);
  if (const Node* sp = deck.find("species")) if (sp->is_seq())
    for (auto& e : sp->seq) if (e.is_map() && e.map.size() == 1) {
      Species x; x.name = e.map[0].first;
      const Node& p = e.map[0].second;
      x.mass = quantity_or(p.find("mass"), 1.0); x.z = quantity_or(p.find("z"), 0.0); x.charge = quantity_or(p.find("charge"), 0.0);
      sim.species.push_back(x);
    }
  auto graph = build_graph(deck, "simulation", sim);
  if (dry) {
    std::vector<std::string> ops; list_graph(*graph, ops);
    std::printf("graph:");
    for (auto& o : ops) std::printf(" %s", o.c_str());
    std::printf("\n");
    // cutoff pre-pass only: force operators publish rcut_max without touching a grid (main-config.msp:52-74)
    sim.preinit = true;
    build_graph(deck, "preinit_rcut_max", sim)->execute(sim);
    std::printf("rcut_max %.17g rcut_inc %.17g nbh_dist %.17g max_displ %.17g dt %.17g species %zu\n", sim.rcut_max, sim.rcut_inc, sim.nbh_dist, sim.max_displ, sim.dt, sim.species.size());
    return 0;
  }
  sim.preinit = true;
  graph->execute(sim);
  if (trace && sim.rank == 0) { std::printf("trace:"); for (auto& o : sim.trace) std::printf(" %s", o.c_str()); std::printf("\n"); }
  if (sim.ctx) sim.check(xsb_sync(sim.ctx), "xsb_sync");
  if (trace && sim.ctx && sim.rank == 0) {
    // operator chains served in one pass (a Lennard-Jones operator behind eam_alloy_force, xsb200.h: xsb_chain_stats)
    uint64_t fused = 0; sim.check(xsb_chain_stats(sim.ctx, &fused), "xsb_chain_stats");
    std::printf("fused_pair_operators: %llu\n", (unsigned long long)fused);
  }
  return 0;
}

int main(int argc, char** argv) {
  std::string deck;
  std::vector<std::pair<std::string, std::string>> sets;
  std::vector<std::string> dirs;
  bool dry = false, trace = false, quiet = false, parse_only = false;
  int gpus = 1;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "--list-operators") { for (auto& n : OperatorFactory::instance().names()) std::printf("%s\n", n.c_str()); return 0; }
    else if (a == "--quantity" && i + 1 < argc) {
      try { std::printf("%.17g\n", quantity(std::string(argv[++i]))); return 0; } catch (const std::exception& e) { std::fprintf(stderr, "fatal: %s\n", e.what()); return 1; }
    }
    else if (a == "--dry-run") dry = true;
    else if (a == "--trace") trace = true;
    else if (a == "--quiet") quiet = true;
    else if (a == "--parse") parse_only = true;
    else if (a == "--gpus" && i + 1 < argc) gpus = std::atoi(argv[++i]);
    else if (a == "--data-dir" && i + 1 < argc) dirs.push_back(argv[++i]);
    else if (a == "--set" && i + 2 < argc) { sets.emplace_back(argv[i + 1], argv[i + 2]); i += 2; }
    else if (!a.empty() && a[0] == '-') { std::fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    else deck = a;
  }
  if (deck.empty()) { std::fprintf(stderr, "usage: xsb200-run deck.msp [--gpus N] [--set key value] [--data-dir DIR] [--dry-run] [--trace] [--quiet]\n"); return 2; }
  try {
    if (parse_only) { std::printf("%s\n", load_deck(deck, dirs).dump().c_str()); return 0; }
    if (gpus > 1 && !std::getenv("RANK")) {
      // one process per GPU; children find each other through the NCCL id file (init_cuda)
      std::string idfile = "/tmp/xsb200_nccl_id_" + std::to_string((long)getpid());
      std::remove(idfile.c_str());
      setenv("XSB_NCCL_ID_FILE", idfile.c_str(), 1);
      setenv("WORLD_SIZE", std::to_string(gpus).c_str(), 1);
      std::vector<pid_t> kids;
      for (int r = 0; r < gpus; ++r) {
        pid_t p = fork();
        if (p == 0) {
          setenv("RANK", std::to_string(r).c_str(), 1); setenv("LOCAL_RANK", std::to_string(r).c_str(), 1);
          int rc = 1;
          try { rc = run(deck, sets, dirs, dry, trace, quiet || r != 0); } catch (const std::exception& e) { std::fprintf(stderr, "[rank %d] fatal: %s\n", r, e.what()); }
          std::fflush(stdout);
          _exit(rc);
        }
        kids.push_back(p);
      }
      int bad = 0;
      for (pid_t p : kids) { int st = 0; waitpid(p, &st, 0); if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) bad = 1; }
      std::remove(idfile.c_str());
      return bad;
    }
    return run(deck, sets, dirs, dry, trace, quiet);
  } catch (const std::exception& e) {
    // the reference aborts through fatal_error(); same observable behaviour: message + non-zero exit
    std::fprintf(stderr, "fatal: %s\n", e.what());
    return 1;
  }
}
