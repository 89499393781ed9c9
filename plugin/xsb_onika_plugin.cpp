// xsb_onika_plugin.cpp -- the onika plugin that makes libxsb200.so a drop-in for exaStamp's short-range force operators.
//
// exaStamp exposes no C ABI: a force field is an onika::scg::OperatorNode subclass with ADD_SLOT members, registered under its
// YAML name by a static constructor (reference: src/potential/pair_potential_template/pair_potential_impl.hxx:510-513,
// src/potential/eam_potential_template/eam_potential_multimat.cu:65-109, src/potential/snap/snap_force.cu:32-35,
// src/mpi/update_ghosts.cu:30-47, src/mpi/update_from_ghosts.cu:29-48).  Every class below keeps the reference operator's
// name and slots and forwards to include/xsb200.h.  Build it where exaNBody is installed (plugin/CMakeLists.txt:
// find_package(exaNBody) + onika_add_plugin); in this repository it is type-checked against plugin/mock/ by
// tests/test_plugin_syntax.py, and the same operator layer WITHOUT onika is compiled and tested as host/xsbh_operators.cpp.
//
// Operator names: with -DXSB_REPLACE_REFERENCE_OPERATORS the plugin registers the reference's own names (load it INSTEAD of
// exaStampPotential / exaStampMPI ...); otherwise every name gets the prefix "xsb_" so both can live in one process for A/B runs.
#include <onika/scg/operator.h>
#include <onika/scg/operator_factory.h>
#include <onika/scg/operator_slot.h>
#include <onika/log.h>
#include <onika/math/basic_types.h>
#include <exanb/core/grid.h>
#include <exanb/core/domain.h>
#include <exanb/core/make_grid_variant_operator.h>
#include <exanb/particle_neighbors/chunk_neighbors.h>
#include <exaStamp/particle_species/particle_specie.h>
#include <yaml-cpp/yaml.h>

#include <xsb200.h>

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#ifdef XSB_REPLACE_REFERENCE_OPERATORS
#  define XSB_OPNAME(n) n
#else
#  define XSB_OPNAME(n) "xsb_" n
#endif

namespace xsbplugin
{
  using namespace exanb;
  using onika::scg::OperatorNode;
  using onika::scg::OperatorNodeFactory;
  using onika::scg::DocString;
  using exaStamp::ParticleSpecies;
  using StringVector = std::vector<std::string>;

  // 1 eV in exaStamp internal units (ang, Da, ps): EXASTAMP_CONST_QUANTITY(1 eV) of the reference (snap_force_op.h:77)
  static constexpr double XSB_EV = 1.602176634e-19 / (1.66053906660e-27 * 1.0e4);

  // ------------------------------------------------------------------------------------------------------------------
  // one context per rank, created by the first operator that runs (after init_cuda), destroyed at unload
  // ------------------------------------------------------------------------------------------------------------------
  struct Bridge
  {
    xsb_ctx* ctx = nullptr;
    const void* bound_grid = nullptr;     // grid object the flat SoA currently mirrors
    size_t bound_particles = 0;
    bool eam_loaded = false; std::string eam_file;
    bool snap_loaded = false; std::string snap_key;
    ~Bridge() { if( ctx ) xsb_destroy(ctx); }
  };
  static Bridge g_bridge;

  static xsb_ctx* context(const OperatorNode* op)
  {
    if( !g_bridge.ctx )
    {
      auto* pec = op->parallel_execution_context();
      const int dev = pec ? pec->gpu_device_index() : 0;
      if( xsb_create(dev, &g_bridge.ctx) != XSB_OK )
        onika::fatal_error() << "xsb_create: " << xsb_last_error(g_bridge.ctx) << std::endl;      // the reference's fatal_error()/abort behaviour
    }
    return g_bridge.ctx;
  }
# define XSB_CK(call) do { if( (call) != XSB_OK ) onika::fatal_error() << #call << ": " << xsb_last_error(g_bridge.ctx) << std::endl; } while(0)

  // ------------------------------------------------------------------------------------------------------------------
  // row a1: exanb::Grid (one SoA block per cell, unified memory) <-> the library's flat SoA sorted by cell.
  // grid.cell_particle_offset_data() IS the flat layout; per-cell blocks are copied device-to-device.
  // ------------------------------------------------------------------------------------------------------------------
  template<class GridT> static void bind_grid(xsb_ctx* c, GridT& grid, const Domain& domain, bool positions_only)
  {
    const size_t nc = grid.number_of_cells();
    if( g_bridge.bound_grid != &grid || g_bridge.bound_particles != grid.number_of_particles() ) positions_only = false;
    if( !positions_only )
    {
      xsb_grid_desc g{}; const IJK d = grid.dimension();
      g.dims[0] = int32_t(d.i); g.dims[1] = int32_t(d.j); g.dims[2] = int32_t(d.k);
      g.ghost_layers = int32_t(grid.ghost_layers()); g.cell_size = domain.cell_size();
      const Vec3d o = grid.cell_position(IJK{0, 0, 0}); g.origin[0] = o.x; g.origin[1] = o.y; g.origin[2] = o.z;
      const Mat3d& X = domain.xform();
      const double m[9] = { X.m11, X.m12, X.m13, X.m21, X.m22, X.m23, X.m31, X.m32, X.m33 };
      std::copy(m, m + 9, g.xform); g.xform_is_identity = domain.xform_is_identity() ? 1 : 0;
      XSB_CK(xsb_grid_set(c, &g));
      std::vector<uint64_t> off(nc + 1);
      const size_t* po = grid.cell_particle_offset_data();
      for(size_t i = 0; i <= nc; i++) off[i] = po[i];
      XSB_CK(xsb_particles_set_cells(c, off.data()));
      g_bridge.bound_grid = &grid; g_bridge.bound_particles = grid.number_of_particles();
    }
    // per-cell SoA -> flat SoA.  exaNBody allocates cells in unified memory, so the "host" pointers below are valid device
    // pointers too; xsb_field_upload of the gathered staging array is the portable path, a D2D gather kernel the fast one.
    const size_t n = grid.number_of_particles();
    std::vector<double> stage(n);
    const size_t* po = grid.cell_particle_offset_data();
    auto gather = [&](int field, auto fid)
    {
      for(size_t cell = 0; cell < nc; cell++)
      {
        const double* src = grid.cells()[cell][fid];
        std::copy(src, src + grid.cells()[cell].size(), stage.begin() + po[cell]);
      }
      XSB_CK(xsb_field_upload(c, field, stage.data()));
    };
    gather(XSB_F_RX, field::rx); gather(XSB_F_RY, field::ry); gather(XSB_F_RZ, field::rz);
    if( !positions_only )
    {
      std::vector<uint8_t> types(n);
      for(size_t cell = 0; cell < nc; cell++)
      {
        const uint8_t* src = grid.cells()[cell][field::type];
        std::copy(src, src + grid.cells()[cell].size(), types.begin() + po[cell]);
      }
      XSB_CK(xsb_field_upload(c, XSB_F_TYPE, types.data()));
    }
    XSB_CK(xsb_sync(c));
  }

  // flat f, ep (virial) -> per-cell blocks.  Operators ACCUMULATE like the reference functors: the flat arrays are zeroed
  // before each forwarded operator and their content is added to the grid's fields afterwards.
  template<class GridT> static void add_forces_to_grid(xsb_ctx* c, GridT& grid, bool with_virial)
  {
    const size_t nc = grid.number_of_cells(), n = grid.number_of_particles();
    std::vector<double> stage(n);
    const size_t* po = grid.cell_particle_offset_data();
    auto scatter_add = [&](int field, auto fid)
    {
      XSB_CK(xsb_field_download(c, field, stage.data()));
      for(size_t cell = 0; cell < nc; cell++)
      {
        double* dst = grid.cells()[cell][fid];
        const size_t np = grid.cells()[cell].size();
        for(size_t p = 0; p < np; p++) dst[p] += stage[po[cell] + p];
      }
    };
    scatter_add(XSB_F_FX, field::fx); scatter_add(XSB_F_FY, field::fy); scatter_add(XSB_F_FZ, field::fz); scatter_add(XSB_F_EP, field::ep);
    if( with_virial )
    {
      std::vector<double> v(9 * n);
      XSB_CK(xsb_field_download(c, XSB_F_VIRIAL, v.data()));
      for(size_t cell = 0; cell < nc; cell++)
      {
        double* dst = grid.cells()[cell][field::virial];      // Mat3d per particle = 9 doubles, row-major
        const size_t np = grid.cells()[cell].size();
        for(size_t p = 0; p < 9 * np; p++) dst[p] += v[9 * po[cell] + p];
      }
    }
  }

  // ------------------------------------------------------------------------------------------------------------------
  // chunk_neighbors (config: data/config/config_move_particles.msp:54-61)
  // ------------------------------------------------------------------------------------------------------------------
  template<class GridT>
  class XsbChunkNeighbors : public OperatorNode
  {
    ADD_SLOT( GridT                     , grid            , INPUT , REQUIRED );
    ADD_SLOT( Domain                    , domain          , INPUT , REQUIRED );
    ADD_SLOT( double                    , nbh_dist_lab    , INPUT , REQUIRED );
    ADD_SLOT( YAML::Node                , config          , INPUT , OPTIONAL , DocString{"chunk_size, build_particle_offset, subcell_compaction, free_scratch_memory, stream_prealloc_factor"} );
    ADD_SLOT( exanb::GridChunkNeighbors , chunk_neighbors , INPUT_OUTPUT , DocString{"left untouched: the library owns its list (export with xsb_chunk_neighbors_export for other consumers)"} );
  public:
    void execute() override final
    {
      if( grid->number_of_cells() == 0 ) return;
      xsb_ctx* c = context(this);
      bind_grid(c, *grid, *domain, false);
      xsb_chunk_neighbors_config cfg{ 1, 1, 1, 0, 1.05 };
      if( config.has_value() )
      {
        const YAML::Node& n = *config;
        if( n["chunk_size"].IsDefined() ) cfg.chunk_size = n["chunk_size"].template as<int>();
        if( n["build_particle_offset"].IsDefined() ) cfg.build_particle_offset = n["build_particle_offset"].template as<bool>() ? 1 : 0;
        if( n["free_scratch_memory"].IsDefined() ) cfg.free_scratch_memory = n["free_scratch_memory"].template as<bool>() ? 1 : 0;
        if( n["stream_prealloc_factor"].IsDefined() ) cfg.stream_prealloc_factor = n["stream_prealloc_factor"].template as<double>();
      }
      XSB_CK(xsb_chunk_neighbors_build(c, *nbh_dist_lab, &cfg));
    }
  };

  // ------------------------------------------------------------------------------------------------------------------
  // <pot>_compute_force (pair_potential_impl.hxx:39-500; slots :104-122) and <pot>_multi_force
  // ------------------------------------------------------------------------------------------------------------------
  static int pair_param_count(int pot) { return pot == XSB_POT_ZERO ? 0 : (pot == XSB_POT_LJ || pot == XSB_POT_YUKAWA || pot == XSB_POT_RELAX) ? 2 : pot == XSB_POT_BUCKINGHAM ? 3 : 4; }

  // parameters of one pair in the library's order; quantities arrive converted to internal units by onika's YAML layer
  static std::vector<double> pair_params(int pot, const YAML::Node& p, unsigned za, unsigned zb)
  {
    switch( pot )
    {
      case XSB_POT_LJ:         return { p["epsilon"].as<double>(), p["sigma"].as<double>() };                      // lennard_jones.h:60-71
      case XSB_POT_ZBL:        return { p["r1"].as<double>(), p["rc"].as<double>(), double(za), double(zb) };      // zbl/potential.h:36-57 (z from species)
      case XSB_POT_EXP6:       return { p["A"].as<double>(), p["B"].as<double>(), p["C"].as<double>(), p["D"].as<double>() };
      case XSB_POT_BUCKINGHAM: return { p["A"].as<double>(), p["Rho"].as<double>(), p["C"].as<double>() };
      case XSB_POT_YUKAWA:     return { p["A"].as<double>(), p["kappa"].as<double>() };                            // yukawa.h:56-70
      case XSB_POT_RELAX:      return { p["r1"].as<double>(), p["rc"].as<double>() };                              // relax/potential.h:56-72
      default:                 return {};                                                                          // zero
    }
  }

  template<class GridT, int POT, bool MULTI>
  class XsbPairForce : public OperatorNode
  {
    ADD_SLOT( YAML::Node                , parameters        , INPUT , REQUIRED );
    ADD_SLOT( YAML::Node                , common_parameters , INPUT , OPTIONAL );
    ADD_SLOT( double                    , rcut              , INPUT , REQUIRED );
    ADD_SLOT( double                    , rcut_max          , INPUT_OUTPUT , 0.0 );
    ADD_SLOT( exanb::GridChunkNeighbors , chunk_neighbors   , INPUT , OPTIONAL );
    ADD_SLOT( ParticleSpecies           , species           , INPUT , OPTIONAL );
    ADD_SLOT( std::string               , type              , INPUT , OPTIONAL );
    ADD_SLOT( bool                      , ghost             , INPUT , false );
    ADD_SLOT( bool                      , mixed_precision   , INPUT , false , DocString{"FP32 pair math, FP64 accumulation (tolerance 1e-5)"} );
    ADD_SLOT( GridT                     , grid              , INPUT_OUTPUT );
    ADD_SLOT( Domain                    , domain            , INPUT , REQUIRED );
  public:
    void execute() override final
    {
      *rcut_max = std::max(*rcut_max, *rcut);                       // pre-pass on the empty grid (pair_potential_impl.hxx:131-197)
      if( grid->number_of_cells() == 0 ) return;
      xsb_ctx* c = context(this);
      bind_grid(c, *grid, *domain, true);
      const bool vir = grid->has_allocated_field(field::virial);
      const int flags = XSB_FLAG_ENERGY | (*ghost ? XSB_FLAG_GHOST : 0) | (vir ? XSB_FLAG_VIRIAL : 0) | (*mixed_precision ? XSB_FLAG_MIXED : 0);
      XSB_CK(xsb_zero_force_energy(c, 1));
      if constexpr ( !MULTI )
      {
        unsigned z = 0;
        if( species.has_value() && !species->empty() ) z = (*species)[0].m_z;
        const std::vector<double> p = pair_params(POT, *parameters, z, z);
        XSB_CK(xsb_pair_force(c, POT, p.data(), int(p.size()), *rcut, flags));
      }
      else
      {
        // one row per unique_pair_id(type_a, type_b) = hi*(hi+1)/2+lo : {params..., rcut}; pairs without user parameters take
        // common_parameters and the operator's rcut (pair_potential_impl.hxx:295-321)
        const int nt = species.has_value() ? int(species->size()) : 1, np = pair_param_count(POT);
        auto type_index = [&](const std::string& s) { for(int i = 0; i < nt; i++) if( (*species)[size_t(i)].name() == s ) return i; return -1; };
        std::vector<double> rows(size_t(nt * (nt + 1) / 2) * size_t(np + 1), 0.0);
        for(int hi = 0; hi < nt; hi++) for(int lo = 0; lo <= hi; lo++)
        {
          std::vector<double> p(size_t(np), 0.0);
          if( common_parameters.has_value() ) p = pair_params(POT, *common_parameters, (*species)[size_t(lo)].m_z, (*species)[size_t(hi)].m_z);
          double* r = rows.data() + size_t(hi * (hi + 1) / 2 + lo) * size_t(np + 1);
          std::copy(p.begin(), p.end(), r); r[np] = *rcut;
        }
        for(size_t i = 0; i < parameters->size(); i++)
        {
          const YAML::Node e = (*parameters)[i];
          const int a = type_index(e["type_a"].as<std::string>()), b = type_index(e["type_b"].as<std::string>());
          if( a < 0 || b < 0 ) { onika::fatal_error() << "unknown species in pair parameters" << std::endl; continue; }
          const int hi = std::max(a, b), lo = std::min(a, b);
          const std::vector<double> p = pair_params(POT, e["parameters"], (*species)[size_t(lo)].m_z, (*species)[size_t(hi)].m_z);
          double* r = rows.data() + size_t(hi * (hi + 1) / 2 + lo) * size_t(np + 1);
          std::copy(p.begin(), p.end(), r); r[np] = e["rcut"].as<double>();
        }
        XSB_CK(xsb_pair_multi_force(c, POT, nt, rows.data(), np, *rcut, flags));
      }
      add_forces_to_grid(c, *grid, vir);
    }
  };
  template<class GridT> using XsbLjComputeForce   = XsbPairForce<GridT, XSB_POT_LJ, false>;
  template<class GridT> using XsbLjMultiForce     = XsbPairForce<GridT, XSB_POT_LJ, true>;
  template<class GridT> using XsbZblComputeForce  = XsbPairForce<GridT, XSB_POT_ZBL, false>;
  template<class GridT> using XsbZblMultiForce    = XsbPairForce<GridT, XSB_POT_ZBL, true>;
  template<class GridT> using XsbExp6ComputeForce = XsbPairForce<GridT, XSB_POT_EXP6, false>;
  template<class GridT> using XsbBuckComputeForce = XsbPairForce<GridT, XSB_POT_BUCKINGHAM, false>;
  template<class GridT> using XsbYukawaComputeForce = XsbPairForce<GridT, XSB_POT_YUKAWA, false>;
  template<class GridT> using XsbRelaxComputeForce  = XsbPairForce<GridT, XSB_POT_RELAX, false>;
  template<class GridT> using XsbZeroComputeForce   = XsbPairForce<GridT, XSB_POT_ZERO, false>;

  // ------------------------------------------------------------------------------------------------------------------
  // johnson_force / johnson_emb / johnson_force_reuse_emb (eam_potential.cu:69-176; 19 scalars johnson.h:176-204)
  // ------------------------------------------------------------------------------------------------------------------
  // MODEL = xsb_eam_model: the same operator template serves sutton_chen (sutton_chen.h:71-83) and vniitf (vniitf.h:139-157)
  template<class GridT, int PHASES, int MODEL = XSB_EAM_JOHNSON>
  class XsbJohnson : public OperatorNode
  {
    ADD_SLOT( YAML::Node , parameters     , INPUT , REQUIRED );
    ADD_SLOT( double     , rcut           , INPUT , REQUIRED );
    ADD_SLOT( double     , rcut_max       , INPUT_OUTPUT , 0.0 );
    ADD_SLOT( double     , ghost_dist_max , INPUT_OUTPUT , 0.0 );
    ADD_SLOT( exanb::GridChunkNeighbors , chunk_neighbors , INPUT , OPTIONAL );
    ADD_SLOT( GridT      , grid           , INPUT_OUTPUT );
    ADD_SLOT( Domain     , domain         , INPUT , REQUIRED );
    ADD_SLOT( bool       , mixed_precision, INPUT , false );
  public:
    void execute() override final
    {
      *rcut_max = std::max(*rcut_max, *rcut);
      *ghost_dist_max = std::max(*ghost_dist_max, 2.0 * (*rcut));      // single-pass EAM needs F'(rho) of ghosts: 2 rcut of ghost atoms
      if( grid->number_of_cells() == 0 ) return;
      static const std::vector<const char*> all_names[3] = {
        { "re", "fe", "rhoe", "alpha", "beta", "A", "B", "kappa", "lambda", "Fn0", "Fn1", "Fn2", "Fn3", "F0", "F1", "F2", "F3", "Fo", "eta" },
        { "c", "epsilon", "a0", "n", "m" },
        { "rmax", "rmin", "rt0", "Ecoh", "E0", "beta", "A", "Z", "n", "alpha", "D", "eta", "mu" } };
      const std::vector<const char*>& names = all_names[MODEL];
      double p[19]; for(size_t i = 0; i < names.size(); i++) p[i] = (*parameters)[names[i]].template as<double>();
      xsb_ctx* c = context(this);
      bind_grid(c, *grid, *domain, true);
      const bool vir = grid->has_allocated_field(field::virial);
      XSB_CK(xsb_zero_force_energy(c, 1));
      XSB_CK(xsb_eam_analytic_force(c, MODEL, p, int(names.size()), *rcut, PHASES, (vir ? XSB_FLAG_VIRIAL : 0) | (*mixed_precision ? XSB_FLAG_MIXED : 0)));
      add_forces_to_grid(c, *grid, vir);
    }
  };
  template<class GridT> using XsbJohnsonForce   = XsbJohnson<GridT, 1 | 2 | 4>;
  template<class GridT> using XsbJohnsonEmb     = XsbJohnson<GridT, 1 | 2>;
  template<class GridT> using XsbJohnsonReuse   = XsbJohnson<GridT, 4>;
  template<class GridT> using XsbSuttonChenForce = XsbJohnson<GridT, 1 | 2 | 4, XSB_EAM_SUTTON_CHEN>;
  template<class GridT> using XsbSuttonChenEmb   = XsbJohnson<GridT, 1 | 2, XSB_EAM_SUTTON_CHEN>;
  template<class GridT> using XsbSuttonChenReuse = XsbJohnson<GridT, 4, XSB_EAM_SUTTON_CHEN>;
  template<class GridT> using XsbVniitfForce     = XsbJohnson<GridT, 1 | 2 | 4, XSB_EAM_VNIITF>;
  template<class GridT> using XsbVniitfEmb       = XsbJohnson<GridT, 1 | 2, XSB_EAM_VNIITF>;
  template<class GridT> using XsbVniitfReuse     = XsbJohnson<GridT, 4, XSB_EAM_VNIITF>;

  // ------------------------------------------------------------------------------------------------------------------
  // eam_alloy_init / eam_alloy_force (eam_potential_multimat.cu:58-259; slots :88-109)
  // ------------------------------------------------------------------------------------------------------------------
  class XsbEamAlloyInit : public OperatorNode
  {
    ADD_SLOT( YAML::Node , parameters , INPUT_OUTPUT , REQUIRED , DocString{"{file: <setfl>} ; shared with eam_alloy_force through rebind"} );
  public:
    void execute() override final
    {
      const std::string file = (*parameters)["file"].IsDefined() ? (*parameters)["file"].as<std::string>() : parameters->as<std::string>();
      if( g_bridge.eam_loaded && g_bridge.eam_file == file ) return;
      xsb_eam_alloy_tables t{}; char names[256];
      if( xsb_eam_alloy_read(file.c_str(), &t, names, sizeof(names)) != XSB_OK ) { onika::fatal_error() << "cannot read setfl file " << file << std::endl; return; }
      XSB_CK(xsb_eam_alloy_set(context(this), &t));
      xsb_eam_alloy_free(&t);
      g_bridge.eam_loaded = true; g_bridge.eam_file = file;
    }
  };

  template<class GridT>
  class XsbEamAlloyForce : public OperatorNode
  {
    ADD_SLOT( ParticleSpecies , species              , INPUT , REQUIRED );
    ADD_SLOT( YAML::Node      , parameters           , INPUT_OUTPUT , REQUIRED );
    ADD_SLOT( StringVector    , types                , INPUT , StringVector{} );
    ADD_SLOT( double          , rcut                 , INPUT , REQUIRED );
    ADD_SLOT( double          , rcut_max             , INPUT_OUTPUT , 0.0 );
    ADD_SLOT( double          , ghost_dist_max       , INPUT_OUTPUT , 0.0 );
    ADD_SLOT( exanb::GridChunkNeighbors , chunk_neighbors , INPUT , OPTIONAL );
    ADD_SLOT( GridT           , grid                 , INPUT_OUTPUT );
    ADD_SLOT( Domain          , domain               , INPUT , REQUIRED );
    ADD_SLOT( bool            , trigger_thermo_state , INPUT , OPTIONAL );
    ADD_SLOT( bool            , compute_virial       , INPUT , false );
    ADD_SLOT( bool            , eam_rho              , INPUT , true );
    ADD_SLOT( bool            , eam_rho2emb          , INPUT , true );
    ADD_SLOT( bool            , eam_ghost            , INPUT , true );
    ADD_SLOT( bool            , eam_force            , INPUT , true );
    ADD_SLOT( bool            , eam_symmetry         , INPUT , false , DocString{"accepted; computed with full lists (Newton-off): identical forces on own atoms, ghost contributions are zero"} );
    ADD_SLOT( bool            , mixed_precision      , INPUT , false );
    ADD_SLOT( exanb::GridParticleLocks , particle_locks , INPUT_OUTPUT , OPTIONAL );
  public:
    void execute() override final
    {
      *rcut_max = std::max(*rcut_max, *rcut);
      *ghost_dist_max = std::max(*ghost_dist_max, (*eam_ghost) ? 2.0 * (*rcut) : *rcut);
      if( grid->number_of_cells() == 0 ) return;
      xsb_ctx* c = context(this);
      if( !g_bridge.eam_loaded ) { onika::fatal_error() << "eam_alloy_force: eam_alloy_init has not run" << std::endl; return; }
      bind_grid(c, *grid, *domain, true);
      const bool eflag = trigger_thermo_state.has_value() ? *trigger_thermo_state : true;      // eam_potential_multimat.cu:125-149
      const bool vir = eflag && *compute_virial;
      int phases = (*eam_rho ? XSB_EAM_RHO : 0) | (*eam_rho2emb ? XSB_EAM_RHO2EMB : 0) | (*eam_ghost ? XSB_EAM_GHOST : 0) | (*eam_force ? XSB_EAM_FORCE : 0) | (eflag ? XSB_EAM_EFLAG : 0);
      if( *eam_force ) XSB_CK(xsb_zero_force_energy(c, 1));
      XSB_CK(xsb_eam_alloy_force(c, *rcut, phases, (vir ? XSB_FLAG_VIRIAL : 0) | (*mixed_precision ? XSB_FLAG_MIXED : 0)));
      if( *eam_force || (*eam_rho2emb && eflag) ) add_forces_to_grid(c, *grid, vir);
    }
  };

  // ------------------------------------------------------------------------------------------------------------------
  // snap_force (snap_force.cu:26-36 -> ext md::SnapForceGeneric; parameters {nt, param, coef} = LAMMPS files)
  // ------------------------------------------------------------------------------------------------------------------
  // implemented in plugin/xsb_snap_files.cpp of a full build; here: the reader contract
  struct SnapFiles { int twojmax = 0, switchflag = 1, bzeroflag = 1; double rcutfac = 0, rfac0 = 0.99363, rmin0 = 0; std::vector<double> radelem, wjelem, beta; };
  SnapFiles read_snap_files(const std::string& param, const std::string& coef);      // same grammar as host/xsbh_readers.cpp

  // FP32 = true: snap_force_fp32, the reference's SNAP_FP32_MATH plugin (snap/snap_force.cu:25-29, multi_WBe_fp32.msp:31,60)
  template<class GridT, bool FP32 = false>
  class XsbSnapForceT : public OperatorNode
  {
    ADD_SLOT( YAML::Node , parameters           , INPUT , REQUIRED , DocString{"{nt, param, coef}"} );
    ADD_SLOT( double     , rcut_max             , INPUT_OUTPUT , 0.0 );
    ADD_SLOT( exanb::GridChunkNeighbors , chunk_neighbors , INPUT , OPTIONAL );
    ADD_SLOT( bool       , ghost                , INPUT , false );
    ADD_SLOT( bool       , trigger_thermo_state , INPUT , OPTIONAL );
    ADD_SLOT( GridT      , grid                 , INPUT_OUTPUT );
    ADD_SLOT( Domain     , domain               , INPUT , REQUIRED );
  public:
    void execute() override final
    {
      xsb_ctx* c = context(this);
      const std::string key = (*parameters)["param"].as<std::string>() + "|" + (*parameters)["coef"].as<std::string>();
      if( !g_bridge.snap_loaded || g_bridge.snap_key != key )
      {
        SnapFiles f = read_snap_files((*parameters)["param"].as<std::string>(), (*parameters)["coef"].as<std::string>());
        for(double& b : f.beta) b *= XSB_EV;                                   // coefficients in eV -> internal (snap_force_op.h:77)
        xsb_snap_params p{};
        p.twojmax = f.twojmax; p.switchflag = f.switchflag; p.bzeroflag = f.bzeroflag; p.nelements = int(f.radelem.size());
        p.rfac0 = f.rfac0; p.rmin0 = f.rmin0; p.rcutfac = f.rcutfac;
        p.radelem = f.radelem.data(); p.wjelem = f.wjelem.data(); p.beta = f.beta.data();
        XSB_CK(xsb_snap_set(c, &p));
        g_bridge.snap_loaded = true; g_bridge.snap_key = key;
      }
      *rcut_max = std::max(*rcut_max, xsb_snap_rcut_max(c));
      if( grid->number_of_cells() == 0 ) return;
      bind_grid(c, *grid, *domain, true);
      const bool eflag = trigger_thermo_state.has_value() ? *trigger_thermo_state : true;
      const bool vir = eflag && grid->has_allocated_field(field::virial);
      XSB_CK(xsb_zero_force_energy(c, 1));
      // f_i += f_ij, f_j -= f_ij: contributions to ghost neighbours land on the ghost copies, the deck's
      // update_force_energy_from_ghost folds them back (config_update_symmetric_forces.msp)
      XSB_CK(xsb_snap_force(c, (*ghost ? XSB_FLAG_GHOST : 0) | (eflag ? XSB_FLAG_ENERGY : 0) | (vir ? XSB_FLAG_VIRIAL : 0) | (FP32 ? XSB_FLAG_MIXED : 0)));
      add_forces_to_grid(c, *grid, vir);
    }
  };
  template<class GridT> using XsbSnapForce     = XsbSnapForceT<GridT, false>;
  template<class GridT> using XsbSnapForceFp32 = XsbSnapForceT<GridT, true>;

  // ------------------------------------------------------------------------------------------------------------------
  // ghost operators (src/mpi/update_ghosts.cu:30-47, update_from_ghosts.cu:29-48) on the library's own exchange
  // (NCCL P2P; xsb_comm_init with an id broadcast over MPI, xsb_ghost_comm_scheme with the brick decomposition)
  // ------------------------------------------------------------------------------------------------------------------
  static uint32_t field_bit(const std::string& s)
  {
    static const char* n[] = { "rx", "ry", "rz", "fx", "fy", "fz", "ep", "vx", "vy", "vz", "virial", "rho_dEmb", "type", "id" };
    for(int i = 0; i < XSB_F_COUNT_; i++) if( s == n[i] ) return 1u << i;
    onika::fatal_error() << "unknown field '" << s << "' in opt_fields" << std::endl;
    return 0;
  }

  template<uint32_t MASK, bool REVERSE>
  class XsbGhostOp : public OperatorNode
  {
    ADD_SLOT( StringVector , opt_fields , INPUT , StringVector{} );
  public:
    void execute() override final
    {
      if( !g_bridge.ctx ) return;                                   // nothing bound yet (pre-pass on the empty grid)
      uint32_t m = MASK; for(const std::string& s : *opt_fields) m |= field_bit(s);
      if( !m ) return;
      if( REVERSE ) XSB_CK(xsb_ghost_reduce_add(g_bridge.ctx, m)); else XSB_CK(xsb_ghost_update(g_bridge.ctx, m));
    }
  };
  static constexpr uint32_t R_BITS = (1u << XSB_F_RX) | (1u << XSB_F_RY) | (1u << XSB_F_RZ);
  static constexpr uint32_t FE_BITS = (1u << XSB_F_FX) | (1u << XSB_F_FY) | (1u << XSB_F_FZ) | (1u << XSB_F_EP);
  using XsbGhostUpdateR = XsbGhostOp<R_BITS, false>;
  using XsbGhostUpdateAllNoFV = XsbGhostOp<R_BITS | (1u << XSB_F_VX) | (1u << XSB_F_VY) | (1u << XSB_F_VZ) | (1u << XSB_F_TYPE) | (1u << XSB_F_ID), false>;
  using XsbGhostUpdateOpt = XsbGhostOp<0u, false>;
  using XsbUpdateForceEnergyFromGhost = XsbGhostOp<FE_BITS, true>;
  using XsbUpdateVirialForceEnergyFromGhost = XsbGhostOp<FE_BITS | (1u << XSB_F_VIRIAL), true>;
  using XsbUpdateOptFromGhost = XsbGhostOp<0u, true>;

  // zero_force_energy (src/compute/zero_force_energy.cu:98-135) for decks that keep the whole step on the device
  class XsbZeroForceEnergy : public OperatorNode
  {
    ADD_SLOT( bool , ghost , INPUT , false );
  public:
    void execute() override final { if( g_bridge.ctx ) XSB_CK(xsb_zero_force_energy(g_bridge.ctx, *ghost ? 1 : 0)); }
  };

  // === register factories ===
  ONIKA_AUTORUN_INIT(xsb200_plugin)
  {
    auto* F = OperatorNodeFactory::instance();
    F->register_factory( XSB_OPNAME("chunk_neighbors")           , make_grid_variant_operator< XsbChunkNeighbors > );
    F->register_factory( XSB_OPNAME("lj_compute_force")          , make_grid_variant_operator< XsbLjComputeForce > );
    F->register_factory( XSB_OPNAME("lj_compute_force_symetric") , make_grid_variant_operator< XsbLjComputeForce > );
    F->register_factory( XSB_OPNAME("lj_multi_force")            , make_grid_variant_operator< XsbLjMultiForce > );
    F->register_factory( XSB_OPNAME("zbl_compute_force")         , make_grid_variant_operator< XsbZblComputeForce > );
    F->register_factory( XSB_OPNAME("zbl_multi_force")           , make_grid_variant_operator< XsbZblMultiForce > );
    F->register_factory( XSB_OPNAME("exp6_compute_force")        , make_grid_variant_operator< XsbExp6ComputeForce > );
    F->register_factory( XSB_OPNAME("buckingham_compute_force")  , make_grid_variant_operator< XsbBuckComputeForce > );
    F->register_factory( XSB_OPNAME("zbl_compute_force_symetric")        , make_grid_variant_operator< XsbZblComputeForce > );
    F->register_factory( XSB_OPNAME("exp6_compute_force_symetric")       , make_grid_variant_operator< XsbExp6ComputeForce > );
    F->register_factory( XSB_OPNAME("buckingham_compute_force_symetric") , make_grid_variant_operator< XsbBuckComputeForce > );
    F->register_factory( XSB_OPNAME("yukawa_compute_force_symetric")     , make_grid_variant_operator< XsbYukawaComputeForce > );
    F->register_factory( XSB_OPNAME("relax_compute_force_symetric")      , make_grid_variant_operator< XsbRelaxComputeForce > );
    F->register_factory( XSB_OPNAME("zero_compute_force_symetric")       , make_grid_variant_operator< XsbZeroComputeForce > );
    F->register_factory( XSB_OPNAME("yukawa_compute_force")      , make_grid_variant_operator< XsbYukawaComputeForce > );
    F->register_factory( XSB_OPNAME("relax_compute_force")       , make_grid_variant_operator< XsbRelaxComputeForce > );
    F->register_factory( XSB_OPNAME("zero_compute_force")        , make_grid_variant_operator< XsbZeroComputeForce > );
    F->register_factory( XSB_OPNAME("sutton_chen_force")         , make_grid_variant_operator< XsbSuttonChenForce > );
    F->register_factory( XSB_OPNAME("sutton_chen_emb")           , make_grid_variant_operator< XsbSuttonChenEmb > );
    F->register_factory( XSB_OPNAME("sutton_chen_force_reuse_emb"), make_grid_variant_operator< XsbSuttonChenReuse > );
    F->register_factory( XSB_OPNAME("vniitf_force")              , make_grid_variant_operator< XsbVniitfForce > );
    F->register_factory( XSB_OPNAME("vniitf_emb")                , make_grid_variant_operator< XsbVniitfEmb > );
    F->register_factory( XSB_OPNAME("vniitf_force_reuse_emb")    , make_grid_variant_operator< XsbVniitfReuse > );
    F->register_factory( XSB_OPNAME("johnson_force")             , make_grid_variant_operator< XsbJohnsonForce > );
    F->register_factory( XSB_OPNAME("johnson_emb")               , make_grid_variant_operator< XsbJohnsonEmb > );
    F->register_factory( XSB_OPNAME("johnson_force_reuse_emb")   , make_grid_variant_operator< XsbJohnsonReuse > );
    F->register_factory( XSB_OPNAME("eam_alloy_init")            , make_simple_operator< XsbEamAlloyInit > );
    F->register_factory( XSB_OPNAME("eam_alloy_force")           , make_grid_variant_operator< XsbEamAlloyForce > );
    F->register_factory( XSB_OPNAME("snap_force")                , make_grid_variant_operator< XsbSnapForce > );
    F->register_factory( XSB_OPNAME("snap_force_fp32")           , make_grid_variant_operator< XsbSnapForceFp32 > );
    F->register_factory( XSB_OPNAME("ghost_update_r")            , make_simple_operator< XsbGhostUpdateR > );
    F->register_factory( XSB_OPNAME("ghost_update_all_no_fv")    , make_simple_operator< XsbGhostUpdateAllNoFV > );
    F->register_factory( XSB_OPNAME("ghost_update_opt")          , make_simple_operator< XsbGhostUpdateOpt > );
    F->register_factory( XSB_OPNAME("update_force_energy_from_ghost")        , make_simple_operator< XsbUpdateForceEnergyFromGhost > );
    F->register_factory( XSB_OPNAME("update_virial_force_energy_from_ghost") , make_simple_operator< XsbUpdateVirialForceEnergyFromGhost > );
    F->register_factory( XSB_OPNAME("update_opt_from_ghost")     , make_simple_operator< XsbUpdateOptFromGhost > );
    F->register_factory( XSB_OPNAME("zero_force_energy")         , make_simple_operator< XsbZeroForceEnergy > );
  }

} // namespace xsbplugin
