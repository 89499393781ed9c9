// TEST INFRASTRUCTURE ONLY -- CPU oracle.  Imported solely by tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference leg.  The product path (exastamp_b200/) never links,
// loads or calls anything in this directory.
//
// xs_oracle.cpp : CPU restatement (C++17 + OpenMP) of the exaStamp short-range force hot path:
//   * chunk_neighbors list build emitting the reference's per-cell uint16 stream
//     (format: src/rigidmol/compute_pair_rigidmol.h:154-234 ; config data/config/config_move_particles.msp:54-61)
//   * compute_cell_particle_pairs traversal (src/rigidmol/compute_pair_rigidmol.h:87-316, cell loop :354-376)
//   * pair / EAM functors (files cited at each function)
// The traversal + builder live in the un-vendored exaNBody v2.0.1 (docs/notes/BUILD.txt:4): they are
// restated from the in-tree decoder/callers.  PARITY PINNING: the arithmetic is pinned against the
// reference's own headers (oracle/_ref, tests/test_oracle_math.py); the neighbour-list *set* is pinned
// against a brute-force O(N^2) search (the semantics of ext verify_chunk_neighbors); the *encoding
// constants* (cell-index bit layout, offset-table layout) are "parity unpinned" (ext, not visible).
#include "orc_math.h"
#include <algorithm>
#include <cassert>
#include <cstring>
#include <omp.h>

using namespace orc;

extern "C" {

struct orc_grid_t
{
  int32_t dims[3];        // cells per axis, ghost layers included
  int32_t ghost_layers;
  double  cell_size;      // grid-space cell edge
  double  origin[3];      // grid-space corner of cell (0,0,0)
  double  xform[9];       // row-major; physical = xform * grid-space  (Domain::xform)
  int32_t xform_is_identity;
  int32_t pad_;
};

}

namespace
{

struct Vec3 { double x, y, z; };

inline Vec3 xform_apply(const orc_grid_t& g, Vec3 d)
{
  if( g.xform_is_identity ) return d;
  const double* m = g.xform;
  return Vec3{ m[0]*d.x + m[1]*d.y + m[2]*d.z, m[3]*d.x + m[4]*d.y + m[5]*d.z, m[6]*d.x + m[7]*d.y + m[8]*d.z };
}

inline void inverse3(const double* m, double* inv)
{
  const double det = m[0]*(m[4]*m[8]-m[5]*m[7]) - m[1]*(m[3]*m[8]-m[5]*m[6]) + m[2]*(m[3]*m[7]-m[4]*m[6]);
  const double id = 1.0 / det;
  inv[0] =  (m[4]*m[8]-m[5]*m[7])*id; inv[1] = -(m[1]*m[8]-m[2]*m[7])*id; inv[2] =  (m[1]*m[5]-m[2]*m[4])*id;
  inv[3] = -(m[3]*m[8]-m[5]*m[6])*id; inv[4] =  (m[0]*m[8]-m[2]*m[6])*id; inv[5] = -(m[0]*m[5]-m[2]*m[3])*id;
  inv[6] =  (m[3]*m[7]-m[4]*m[6])*id; inv[7] = -(m[0]*m[7]-m[1]*m[6])*id; inv[8] =  (m[0]*m[4]-m[1]*m[3])*id;
}

// number of cell layers to scan along each grid axis so that every grid-space displacement whose
// physical length is < dist is covered: extent_i = dist * || row_i(xform^-1) ||
inline void search_range(const orc_grid_t& g, double dist, int R[3])
{
  double inv[9] = {1,0,0,0,1,0,0,0,1};
  if( !g.xform_is_identity ) inverse3(g.xform, inv);
  for(int a = 0; a < 3; a++)
  {
    const double n = std::sqrt(inv[3*a]*inv[3*a] + inv[3*a+1]*inv[3*a+1] + inv[3*a+2]*inv[3*a+2]);
    R[a] = int(std::ceil(dist * n / g.cell_size));
    if( R[a] < 1 ) R[a] = 1;
    if( R[a] > 15 ) R[a] = 15; // encodable range of the 5-bit relative cell index
  }
}

// relative cell index <-> uint16.  exanb::encode_cell_index / decode_cell_index are external (used at
// compute_pair_rigidmol.h:205); 5 bits per axis biased by 16 is our documented stand-in.
inline uint16_t encode_cell_index(int ri, int rj, int rk) { return uint16_t((ri + 16) | ((rj + 16) << 5) | ((rk + 16) << 10)); }
inline void decode_cell_index(uint16_t e, int& ri, int& rj, int& rk) { ri = int(e & 31) - 16; rj = int((e >> 5) & 31) - 16; rk = int((e >> 10) & 31) - 16; }

struct Nbh
{
  int chunk_size = 1;
  bool has_offsets = true;
  std::vector< std::vector<uint16_t> > streams; // one per cell
};

// default neighbour filter: d2>0 && d2<rcut2 (in-tree filters: src/molecule/extramolecular_neighbors.cpp:56-68,
// src/particle_species/type_pair_rcut_neighbors.cpp:57-64)
inline bool nbh_filter(double d2, double rcut2) { return d2 > 0.0 && d2 < rcut2; }

} // namespace

extern "C" {

// ---- chunk_neighbors -------------------------------------------------------------------------
void* orc_nbh_build(const orc_grid_t* g, const uint64_t* cell_off, const double* rx, const double* ry, const double* rz,
                    double nbh_dist_lab, int chunk_size, int build_particle_offset)
{
  Nbh* nb = new Nbh;
  nb->chunk_size = chunk_size;
  nb->has_offsets = build_particle_offset != 0;
  const int nx = g->dims[0], ny = g->dims[1], nz = g->dims[2];
  const size_t ncells = size_t(nx) * ny * nz;
  nb->streams.resize(ncells);
  int R[3]; search_range(*g, nbh_dist_lab, R);
  const double d2max = nbh_dist_lab * nbh_dist_lab;
  const int CS = chunk_size;

# pragma omp parallel for schedule(dynamic)
  for(size_t cell_a = 0; cell_a < ncells; cell_a++)
  {
    const int ia = int(cell_a % nx), ja = int((cell_a / nx) % ny), ka = int(cell_a / (size_t(nx) * ny));
    const size_t na = cell_off[cell_a + 1] - cell_off[cell_a];
    std::vector<uint16_t>& st = nb->streams[cell_a];
    std::vector<uint32_t> poff(na, 0);
    std::vector<uint16_t> body;
    for(size_t p_a = 0; p_a < na; p_a++)
    {
      poff[p_a] = uint32_t(body.size());
      const size_t ga = cell_off[cell_a] + p_a;
      const size_t groups_pos = body.size();
      body.push_back(0);
      unsigned cell_groups = 0;
      for(int rk = -R[2]; rk <= R[2]; rk++) for(int rj = -R[1]; rj <= R[1]; rj++) for(int ri = -R[0]; ri <= R[0]; ri++)
      {
        const int ib = ia + ri, jb = ja + rj, kb = ka + rk;
        if( ib < 0 || ib >= nx || jb < 0 || jb >= ny || kb < 0 || kb >= nz ) continue;
        const size_t cell_b = size_t(ib) + size_t(nx) * (size_t(jb) + size_t(ny) * kb);
        const size_t nbp = cell_off[cell_b + 1] - cell_off[cell_b];
        size_t nchunks_pos = 0; unsigned nchunks = 0; long last_chunk = -1;
        for(size_t p_b = 0; p_b < nbp; p_b++)
        {
          if( cell_b == cell_a && p_b == p_a ) continue;
          const size_t gb = cell_off[cell_b] + p_b;
          Vec3 dr = xform_apply(*g, Vec3{ rx[gb] - rx[ga], ry[gb] - ry[ga], rz[gb] - rz[ga] });
          const double d2 = dr.x*dr.x + dr.y*dr.y + dr.z*dr.z;
          if( !nbh_filter(d2, d2max) ) continue;
          const long chunk = long(p_b / CS);
          if( chunk == last_chunk ) continue;
          if( nchunks == 0 )
          {
            body.push_back(encode_cell_index(ri, rj, rk));
            nchunks_pos = body.size();
            body.push_back(0);
            ++cell_groups;
          }
          body.push_back(uint16_t(chunk));
          ++nchunks;
          last_chunk = chunk;
        }
        if( nchunks ) body[nchunks_pos] = uint16_t(nchunks);
      }
      body[groups_pos] = uint16_t(cell_groups);
    }
    st.clear();
    if( nb->has_offsets )
    {
      st.resize(2 * na);
      for(size_t p = 0; p < na; p++)
      {
        const uint32_t o = uint32_t(2 * na) + poff[p];
        st[2*p] = uint16_t(o & 0xFFFFu); st[2*p+1] = uint16_t(o >> 16);
      }
    }
    st.insert(st.end(), body.begin(), body.end());
  }
  return nb;
}

void orc_nbh_free(void* h) { delete static_cast<Nbh*>(h); }

uint64_t orc_nbh_total_size(void* h)
{
  uint64_t s = 0; for(const auto& v : static_cast<Nbh*>(h)->streams) s += v.size(); return s;
}

// concatenated export: stream_off[ncells+1] (uint16 units) + data
void orc_nbh_export(void* h, uint64_t* stream_off, uint16_t* data)
{
  Nbh* nb = static_cast<Nbh*>(h);
  uint64_t o = 0;
  for(size_t c = 0; c < nb->streams.size(); c++)
  {
    stream_off[c] = o;
    if( data ) std::memcpy(data + o, nb->streams[c].data(), nb->streams[c].size() * sizeof(uint16_t));
    o += nb->streams[c].size();
  }
  stream_off[nb->streams.size()] = o;
}

// import a stream built elsewhere (e.g. the CUDA library's export) so the reference traversal can run on it
void* orc_nbh_import(uint64_t ncells, const uint64_t* stream_off, const uint16_t* data, int chunk_size, int has_offsets)
{
  Nbh* nb = new Nbh; nb->chunk_size = chunk_size; nb->has_offsets = has_offsets != 0;
  nb->streams.resize(ncells);
  for(uint64_t c = 0; c < ncells; c++) nb->streams[c].assign(data + stream_off[c], data + stream_off[c+1]);
  return nb;
}

} // extern "C"

namespace
{

// Walk the neighbours of every particle of cell_a in stream order, exactly like the in-tree decoder
// (compute_pair_rigidmol.h:154-234): f(p_a, start/stop) and per candidate f(cell_b, p_b, global_b).
template<class StartF, class NbhF, class StopF>
inline void walk_cell(const orc_grid_t& g, const uint64_t* cell_off, const Nbh& nb, size_t cell_a, StartF start, NbhF nbh, StopF stop)
{
  const int nx = g.dims[0], ny = g.dims[1];
  const int ia = int(cell_a % nx), ja = int((cell_a / nx) % ny), ka = int(cell_a / (size_t(nx) * ny));
  const size_t na = cell_off[cell_a + 1] - cell_off[cell_a];
  if( na == 0 ) return;
  const uint16_t* stream = nb.streams[cell_a].data();
  if( nb.has_offsets ) stream += 2 * na;   // chunknbh_stream_info(): skip per-particle offset table
  const unsigned CS = unsigned(nb.chunk_size);
  for(size_t p_a = 0; p_a < na; p_a++)
  {
    start(p_a);
    const unsigned cell_groups = *(stream++);
    for(unsigned cg = 0; cg < cell_groups; cg++)
    {
      int ri, rj, rk; decode_cell_index(*(stream++), ri, rj, rk);
      const size_t cell_b = size_t(ia + ri) + size_t(nx) * (size_t(ja + rj) + size_t(ny) * (ka + rk));
      const size_t nbp = cell_off[cell_b + 1] - cell_off[cell_b];
      const unsigned nchunks = *(stream++);
      for(unsigned c = 0; c < nchunks; c++)
      {
        const unsigned chunk_start = unsigned(*(stream++)) * CS;
        for(unsigned i = 0; i < CS; i++)
        {
          const size_t p_b = chunk_start + i;
          if( p_b < nbp && (cell_b != cell_a || p_b != p_a) ) nbh(p_a, cell_b, p_b);
        }
      }
    }
    stop(p_a);
  }
}

inline bool is_ghost_cell(const orc_grid_t& g, size_t cell)
{
  const int nx = g.dims[0], ny = g.dims[1], nz = g.dims[2], gl = g.ghost_layers;
  const int i = int(cell % nx), j = int((cell / nx) % ny), k = int(cell / (size_t(nx) * ny));
  return i < gl || i >= nx - gl || j < gl || j >= ny - gl || k < gl || k >= nz - gl;
}

// ComputePairBuffer2 stand-in (ext): drx,dry,drz,d2 + neighbour global index
struct PairBuf
{
  std::vector<double> drx, dry, drz, d2; std::vector<size_t> gb; size_t count = 0;
  void clear() { count = 0; drx.clear(); dry.clear(); drz.clear(); d2.clear(); gb.clear(); }
  void push(Vec3 dr, double dd, size_t g) { drx.push_back(dr.x); dry.push_back(dr.y); drz.push_back(dr.z); d2.push_back(dd); gb.push_back(g); ++count; }
};

struct Particles { const uint64_t* cell_off; const double *rx, *ry, *rz; const uint8_t* type; };

// generic traversal: for each central atom, fill the buffer with in-range pairs (d2 <= rcut2, inclusive:
// compute_pair_rigidmol.h:262) then call op(ga, buf).  OpenMP over cells, schedule(dynamic) (:354-376).
template<class Op>
inline void compute_cell_particle_pairs(const orc_grid_t& g, const Particles& P, const Nbh& nb, double rcut, bool ghost, Op op)
{
  const size_t ncells = size_t(g.dims[0]) * g.dims[1] * g.dims[2];
  const double rcut2 = rcut * rcut;
# pragma omp parallel
  {
    PairBuf buf;
#   pragma omp for schedule(dynamic)
    for(size_t cell_a = 0; cell_a < ncells; cell_a++)
    {
      if( !ghost && is_ghost_cell(g, cell_a) ) continue;
      walk_cell(g, P.cell_off, nb, cell_a,
        [&](size_t) { buf.clear(); },
        [&](size_t p_a, size_t cell_b, size_t p_b)
        {
          const size_t ga = P.cell_off[cell_a] + p_a, gb = P.cell_off[cell_b] + p_b;
          Vec3 dr = xform_apply(g, Vec3{ P.rx[gb] - P.rx[ga], P.ry[gb] - P.ry[ga], P.rz[gb] - P.rz[ga] });
          const double d2 = dr.x*dr.x + dr.y*dr.y + dr.z*dr.z;
          if( d2 <= rcut2 ) buf.push(dr, d2, gb);
        },
        [&](size_t p_a) { op(P.cell_off[cell_a] + p_a, buf); } );
    }
  }
}

// vir += -0.5 * tensor(fe,dr) ; Mat3d row-major m11..m33, tensor(a,b)_ij = a_i*b_j (ext)
inline void vir_add(double* v, double fx, double fy, double fz, double dx, double dy, double dz)
{
  v[0] += fx*dx*-0.5; v[1] += fx*dy*-0.5; v[2] += fx*dz*-0.5;
  v[3] += fy*dx*-0.5; v[4] += fy*dy*-0.5; v[5] += fy*dz*-0.5;
  v[6] += fz*dx*-0.5; v[7] += fz*dy*-0.5; v[8] += fz*dz*-0.5;
}

} // namespace

extern "C" {

// decode streams into a flat list (CSR): counts per particle (global index), then indices; returns total.
// pass idx=nullptr to only count.  Used to compare list *content* with the CUDA library.
uint64_t orc_nbh_decode(const orc_grid_t* g, const uint64_t* cell_off, void* h, uint32_t* counts, uint64_t* idx_off, uint32_t* idx)
{
  const Nbh& nb = *static_cast<Nbh*>(h);
  const size_t ncells = size_t(g->dims[0]) * g->dims[1] * g->dims[2];
  const size_t N = cell_off[ncells];
  for(size_t i = 0; i < N; i++) counts[i] = 0;
  for(size_t c = 0; c < ncells; c++)
    walk_cell(*g, cell_off, nb, c, [](size_t){}, [&](size_t p_a, size_t, size_t) { ++counts[cell_off[c] + p_a]; }, [](size_t){});
  uint64_t tot = 0;
  for(size_t i = 0; i < N; i++) { idx_off[i] = tot; tot += counts[i]; }
  idx_off[N] = tot;
  if( idx )
  {
    for(size_t c = 0; c < ncells; c++)
    {
      uint64_t w = 0;
      walk_cell(*g, cell_off, nb, c, [&](size_t p_a){ w = idx_off[cell_off[c] + p_a]; },
                [&](size_t, size_t cell_b, size_t p_b) { idx[w++] = uint32_t(cell_off[cell_b] + p_b); }, [](size_t){});
    }
  }
  return tot;
}

// brute-force O(N^2) neighbour counts with the same filter (set semantics of ext verify_chunk_neighbors)
void orc_nbh_bruteforce_counts(const orc_grid_t* g, uint64_t N, const double* rx, const double* ry, const double* rz, double nbh_dist_lab, uint32_t* counts)
{
  const double d2max = nbh_dist_lab * nbh_dist_lab;
# pragma omp parallel for schedule(static)
  for(uint64_t a = 0; a < N; a++)
  {
    uint32_t c = 0;
    for(uint64_t b = 0; b < N; b++)
    {
      if( a == b ) continue;
      Vec3 dr = xform_apply(*g, Vec3{ rx[b] - rx[a], ry[b] - ry[a], rz[b] - rz[a] });
      if( nbh_filter(dr.x*dr.x + dr.y*dr.y + dr.z*dr.z, d2max) ) ++c;
    }
    counts[a] = c;
  }
}

// one evaluation of a pair potential and its energy cutoff (pins the restatement against the reference headers)
void orc_pair_eval(int pot, const double* params, double r, double* e, double* de) { double a = 0.0, b = 0.0; PairPot(pot, params).compute(r, a, b); *e = a; *de = b; }
double orc_pair_ecut(int pot, const double* params, double rcut) { return PairPot(pot, params).energy_cutoff(rcut); }

// ---- <pot>_compute_force, single species, buffered protocol ------------------------------------
// ForceOp body: src/potential/pair_potential_template/force_op_impl2.hxx:22-78 ; operator slots and
// ecut: pair_potential_impl.hxx:373-377,488-498.  pot: 0 = lj.  vir may be null (9 doubles/atom, AoS Mat3d).
void orc_pair_force(const orc_grid_t* g, const uint64_t* cell_off, const double* rx, const double* ry, const double* rz,
                    void* nbh, int pot, const double* params, double rcut, int ghost,
                    double* fx, double* fy, double* fz, double* ep, double* vir)
{
  const Particles P{ cell_off, rx, ry, rz, nullptr };
  const PairPot p(pot, params);
  const double ecut = p.energy_cutoff(rcut);
  compute_cell_particle_pairs(*g, P, *static_cast<Nbh*>(nbh), rcut, ghost != 0, [&](size_t ga, const PairBuf& tab)
  {
    double _ep = 0., _fx = 0., _fy = 0., _fz = 0.; double _vir[9] = {0,0,0,0,0,0,0,0,0};
    const double weight = 1.0;
    for(size_t i = 0; i < tab.count; i++)
    {
      const double r = std::sqrt(tab.d2[i]);
      double e = 0.0, de = 0.0;
      p.compute(r, e, de);
      e *= weight; de *= weight;
      e -= ecut * weight;
      de /= r;
      const double fe_x = de * tab.drx[i], fe_y = de * tab.dry[i], fe_z = de * tab.drz[i];
      _fx += fe_x; _fy += fe_y; _fz += fe_z;
      _ep += .5 * e;
      vir_add(_vir, fe_x, fe_y, fe_z, tab.drx[i], tab.dry[i], tab.drz[i]);
    }
    if( ep ) ep[ga] += _ep;
    fx[ga] += _fx; fy[ga] += _fy; fz[ga] += _fz;
    if( vir ) for(int k = 0; k < 9; k++) vir[9*ga + k] += _vir[k];
  });
}

// ---- <pot>_multi_force, buffer-less protocol with per type-pair parameters ------------------------
// PairMultiForceOp: pair_potential_force_op_multiparam.h:76-223 (overload with ep: :78-118 / :121-150).
// pair_params[pair_id] = {params..., rcut, ecut} (params: PairPot::nparams(pot) scalars); traversal radius = rcut_max
// (pair_potential_impl.hxx:196,478).
void orc_pair_multi_force(const orc_grid_t* g, const uint64_t* cell_off, const double* rx, const double* ry, const double* rz, const uint8_t* type,
                          void* nbh, int pot, int n_pair_params, const double* pair_params, double rcut_max, int ghost,
                          double* fx, double* fy, double* fz, double* ep, double* vir)
{
  const int np = PairPot::nparams(pot), stride = np + 2;
  std::vector<PairPot> pots(n_pair_params);
  for(int i = 0; i < n_pair_params; i++) pots[i] = PairPot(pot, pair_params + size_t(stride) * i);
  const Particles P{ cell_off, rx, ry, rz, type };
  compute_cell_particle_pairs(*g, P, *static_cast<Nbh*>(nbh), rcut_max, ghost != 0, [&](size_t ga, const PairBuf& tab)
  {
    // buffer-less: the reference accumulates straight into the particle fields, pair by pair in list order
    const unsigned type_a = type[ga];
    for(size_t i = 0; i < tab.count; i++)
    {
      const double r = std::sqrt(tab.d2[i]);
      const unsigned type_b = type[tab.gb[i]];
      const unsigned pid = unique_pair_id(type_a, type_b);
      const double* pp = pair_params + size_t(stride) * pid + np;     // {rcut, ecut}
      if( r <= pp[0] )
      {
        double e = 0.0, de = 0.0;
        pots[pid].compute(r, e, de);
        const double weight = 1.0;
        if( vir && ep ) { e *= weight; de *= weight; e -= pp[1] * weight; de /= r; }   // :106-109
        else            { e -= pp[1]; de /= r; e *= weight; de *= weight; }               // :142-145
        const double fe_x = de * tab.drx[i], fe_y = de * tab.dry[i], fe_z = de * tab.drz[i];
        fx[ga] += fe_x; fy[ga] += fe_y; fz[ga] += fe_z;
        if( ep ) ep[ga] += .5 * e;
        if( vir ) vir_add(vir + 9*ga, fe_x, fe_y, fe_z, tab.drx[i], tab.dry[i], tab.drz[i]);
      }
    }
  });
}

// ---- single-species analytic EAM (johnson_force): two buffered passes -----------------------------
// operator: src/potential/eam_potential_template/eam_potential.cu:105-174 ; functors
// eam_force_op_singlemat.h:39-84 (EmbOp) and :86-171 (ForceOp).  flags: bit0 = compute emb pass,
// bit1 = emb pass over ghost cells too (ComputeGhostEmb), bit2 = force pass.
// model: 0 johnson, 1 sutton_chen, 2 vniitf -- the operator template is the same, only the three functions differ
// (<name>/potential.h: USTAMP_POTENTIAL_EAM_RHO / _PHI / _EMB)
void orc_eam_analytic(const orc_grid_t* g, const uint64_t* cell_off, const double* rx, const double* ry, const double* rz,
                      void* nbh, int model, const double* params, double rcut, int flags,
                      double* fx, double* fy, double* fz, double* ep, double* vir, double* rho_dEmb);
void orc_eam_johnson(const orc_grid_t* g, const uint64_t* cell_off, const double* rx, const double* ry, const double* rz,
                     void* nbh, const double* params19, double rcut, int flags,
                     double* fx, double* fy, double* fz, double* ep, double* vir, double* rho_dEmb)
{ orc_eam_analytic(g, cell_off, rx, ry, rz, nbh, 0, params19, rcut, flags, fx, fy, fz, ep, vir, rho_dEmb); }
void orc_eam_analytic(const orc_grid_t* g, const uint64_t* cell_off, const double* rx, const double* ry, const double* rz,
                      void* nbh, int model, const double* params, double rcut, int flags,
                      double* fx, double* fy, double* fz, double* ep, double* vir, double* rho_dEmb)
{
  const Particles P{ cell_off, rx, ry, rz, nullptr };
  const EamAnalytic p(model, params);
  const Nbh& nb = *static_cast<Nbh*>(nbh);
  if( flags & 1 )
  {
    const size_t N = cell_off[size_t(g->dims[0]) * g->dims[1] * g->dims[2]];
    for(size_t i = 0; i < N; i++) rho_dEmb[i] = 0.0;      // m_rho_emb.clear(); resize() (eam_potential.cu:126-127)
    compute_cell_particle_pairs(*g, P, nb, rcut, (flags & 2) != 0, [&](size_t ga, const PairBuf& tab)
    {
      if( tab.count == 0 ) return;                        // ext traversal only calls the functor for non-empty buffers
      double particle_rho = 0.;
      for(size_t i = 0; i < tab.count; i++)
      {
        const double r = std::sqrt(tab.d2[i]);
        double Rho = 0., dRho = 0.;
        p.rho(r, Rho, dRho);
        particle_rho += Rho;
      }
      double Emb = 0., dEmb = 0.;
      p.fEmbed(particle_rho, Emb, dEmb);
      ep[ga] += Emb;
      rho_dEmb[ga] = dEmb;
    });
  }
  if( flags & 4 )
  {
    compute_cell_particle_pairs(*g, P, nb, rcut, false, [&](size_t ga, const PairBuf& tab)
    {
      if( tab.count == 0 ) return;
      const double dEmb = rho_dEmb[ga];
      double _ep = 0., _fx = 0., _fy = 0., _fz = 0.; double _vir[9] = {0,0,0,0,0,0,0,0,0};
      for(size_t i = 0; i < tab.count; i++)
      {
        const double r = std::sqrt(tab.d2[i]);
        double Rho = 0., dRho = 0., Phi = 0., dPhi = 0.;
        p.rho(r, Rho, dRho);
        p.phi(r, Phi, dPhi);
        const double de = ( dRho * ( dEmb + rho_dEmb[tab.gb[i]] ) + dPhi ) / r;
        const double fe_x = de * tab.drx[i], fe_y = de * tab.dry[i], fe_z = de * tab.drz[i];
        _fx += fe_x; _fy += fe_y; _fz += fe_z;
        _ep += .5 * Phi;
        vir_add(_vir, fe_x, fe_y, fe_z, tab.drx[i], tab.dry[i], tab.drz[i]);
      }
      ep[ga] += _ep; fx[ga] += _fx; fy[ga] += _fy; fz[ga] += _fz;
      if( vir ) for(int k = 0; k < 9; k++) vir[9*ga + k] += _vir[k];
    });
  }
}

// ---- eam_alloy_force (multi-species, tabulated): rho -> rho2emb -> force ---------------------------
// operator: src/potential/eam_potential_template/eam_potential_multimat.cu:113-257 ; functors
// eam_force_op_multimat.h:107-343 (Newton-off variants).  phase flags as the operator slots:
// bit0 eam_rho, bit1 eam_rho2emb, bit2 eam_ghost, bit3 eam_force, bit4 eflag (trigger_thermo_state),
// bit5 compute_virial.
void* orc_eam_alloy_load(const char* path) { EamAlloy* e = new EamAlloy; if( !eam_alloy_read(path, *e) ) { delete e; return nullptr; } return e; }
void  orc_eam_alloy_free(void* h) { delete static_cast<EamAlloy*>(h); }
void  orc_eam_alloy_info(void* h, int* nelements, int* nr, int* nrho, double* rdr, double* rdrho, double* rc, double* rhomax)
{
  const EamAlloy& e = *static_cast<EamAlloy*>(h);
  *nelements = e.nelements; *nr = e.nr; *nrho = e.nrho; *rdr = e.rdr; *rdrho = e.rdrho; *rc = e.rc; *rhomax = e.rhomax;
}
const double* orc_eam_alloy_table(void* h, int which) // 0 frho, 1 rhor, 2 z2r
{
  const EamAlloy& e = *static_cast<EamAlloy*>(h);
  return which == 0 ? e.frho.data() : which == 1 ? e.rhor.data() : e.z2r.data();
}
double orc_eam_alloy_eval(void* h, int what, double x, int ti, int tj, double fpi, double fpj, double* out2)
{
  const EamAlloy& e = *static_cast<EamAlloy*>(h);
  if( what == 0 ) return eam_alloy_rho_noderiv(e, x, ti, tj);
  if( what == 1 ) { double phi = 0, fp = 0; eam_alloy_fEmbed(e, x, phi, fp, ti); *out2 = fp; return phi; }
  double phi = 0; const double fpair = eam_alloy_mm_force(e, phi, x, fpi, fpj, ti, tj); *out2 = phi; return fpair;
}

void orc_eam_alloy(const orc_grid_t* g, const uint64_t* cell_off, const double* rx, const double* ry, const double* rz, const uint8_t* type,
                   void* nbh, void* eam_h, double rcut, int flags,
                   double* fx, double* fy, double* fz, double* ep, double* vir, double* rho_dEmb)
{
  const Particles P{ cell_off, rx, ry, rz, type };
  const EamAlloy& eam = *static_cast<EamAlloy*>(eam_h);
  const Nbh& nb = *static_cast<Nbh*>(nbh);
  const size_t ncells = size_t(g->dims[0]) * g->dims[1] * g->dims[2];
  const size_t N = cell_off[ncells];
  const bool eam_rho = flags & 1, eam_rho2emb = flags & 2, eam_ghost = flags & 4, eam_force = flags & 8;
  // need_virial = log_energy && compute_virial, then log_energy |= need_virial (eam_potential_multimat.cu:134-149):
  // compute_virial never changes the outcome -- the virial field is written whenever eflag is set (:262-266).
  const bool eflag = (flags & 16) != 0;
  if( eam_rho )
  {
    for(size_t i = 0; i < N; i++) rho_dEmb[i] = 0.0;    // parallel_memset (eam_potential_multimat.cu:188)
    compute_cell_particle_pairs(*g, P, nb, rcut, eam_ghost, [&](size_t ga, const PairBuf& tab)
    {
      double rho = 0.0;                                   // ComputePairParticleContextStart
      const int type_a = type[ga];
      for(size_t i = 0; i < tab.count; i++)
      {
        const int type_b = type[tab.gb[i]];
        const double r = std::sqrt(tab.d2[i]);
        rho += eam_alloy_rho_noderiv(eam, r, type_b, type_a);   // eam_force_op_multimat.h:156
      }
      rho_dEmb[ga] += rho;                                // ContextStop :139
    });
  }
  if( eam_rho2emb )
  {
    // compute_cell_particles(grid, eam_ghost, Rho2EmbOp) (eam_potential_multimat.cu:206-211)
#   pragma omp parallel for schedule(static)
    for(size_t c = 0; c < ncells; c++)
    {
      if( !eam_ghost && is_ghost_cell(*g, c) ) continue;
      for(size_t ga = cell_off[c]; ga < cell_off[c+1]; ga++)
      {
        double emb = 0., dEmb = 0.;
        eam_alloy_fEmbed(eam, rho_dEmb[ga], emb, dEmb, type[ga]);
        rho_dEmb[ga] = dEmb;
        if( eflag ) ep[ga] += emb;                        // Rho2EmbOp overload with ep (:176-187) only when log_energy
      }
    }
  }
  if( eam_force )
  {
    compute_cell_particle_pairs(*g, P, nb, rcut, false, [&](size_t ga, const PairBuf& tab)
    {
      const double fpi = rho_dEmb[ga];
      double f[3] = {0,0,0}, e = 0.0; double v[9] = {0,0,0,0,0,0,0,0,0};
      const int type_a = type[ga];
      for(size_t i = 0; i < tab.count; i++)
      {
        const int type_b = type[tab.gb[i]];
        const double r = std::sqrt(tab.d2[i]);
        double phi = 0.;
        const double fpair = eam_alloy_mm_force(eam, phi, r, fpi, rho_dEmb[tab.gb[i]], type_a, type_b);
        const double fe_x = tab.drx[i] * fpair, fe_y = tab.dry[i] * fpair, fe_z = tab.drz[i] * fpair;
        f[0] += fe_x; f[1] += fe_y; f[2] += fe_z;
        if( eflag ) { e += .5 * phi; vir_add(v, fe_x, fe_y, fe_z, tab.drx[i], tab.dry[i], tab.drz[i]); }
      }
      fx[ga] += f[0]; fy[ga] += f[1]; fz[ga] += f[2];
      if( eflag ) { ep[ga] += e; if( vir ) for(int k = 0; k < 9; k++) vir[9*ga + k] += v[k]; }
    });
  }
}

// ---- snap_force: traversal + the call sequence of src/potential/snaplmp/snap_force_op.h:177-337 ----------------------
// (filter rsq < cutsq_ij && rsq > 1e-20 :191 ; compute_ui :214 ; compute_yi :246 ; per neighbour duidrj + deidrj :248-257 ;
//  f_i += fij, f_j -= fij :277-294 ; virial -fij (x) rij on the centre :267-268 ; energy e0 + beta.B :297-337).
// The per-atom arithmetic lives in snap_oracle.cpp.  flags: bit0 ghost, bit1 energy, bit2 virial.
void orc_snap_atom(void* h, int n, const double* dx, const double* dy, const double* dz, const int* elem_j, int elem_i, double* B, double* dB, double* energy, double* dedr);
void orc_snap_cut(void* h, int elem_i, int elem_j, double* rcut);
void orc_snap_force(const orc_grid_t* g, const uint64_t* cell_off, const double* rx, const double* ry, const double* rz, const uint8_t* type,
                    void* nbh, void* snap, double rcut_max, int flags, double* fx, double* fy, double* fz, double* ep, double* vir)
{
  const Particles P{ cell_off, rx, ry, rz, type };
  const bool ghost = flags & 1, eflag = flags & 2, vflag = flags & 4;
  // f_j -= fij crosses cell boundaries: serial over cells keeps the oracle race-free and deterministic
  const size_t ncells = size_t(g->dims[0]) * g->dims[1] * g->dims[2];
  const Nbh& nb = *static_cast<Nbh*>(nbh);
  const double rcut2 = rcut_max * rcut_max;
  std::vector<double> dx, dy, dz, dedr; std::vector<int> ej; std::vector<size_t> gj;
  for(size_t cell_a = 0; cell_a < ncells; cell_a++)
  {
    if( !ghost && is_ghost_cell(*g, cell_a) ) continue;
    walk_cell(*g, cell_off, nb, cell_a,
      [&](size_t) { dx.clear(); dy.clear(); dz.clear(); ej.clear(); gj.clear(); },
      [&](size_t p_a, size_t cell_b, size_t p_b)
      {
        const size_t ga = cell_off[cell_a] + p_a, gb = cell_off[cell_b] + p_b;
        Vec3 dr = xform_apply(*g, Vec3{ rx[gb] - rx[ga], ry[gb] - ry[ga], rz[gb] - rz[ga] });
        const double d2 = dr.x*dr.x + dr.y*dr.y + dr.z*dr.z;
        if( d2 > rcut2 ) return;
        double rc = 0.0; orc_snap_cut(snap, type ? type[ga] : 0, type ? type[gb] : 0, &rc);
        if( d2 < rc * rc && d2 > 1e-20 ) { dx.push_back(dr.x); dy.push_back(dr.y); dz.push_back(dr.z); ej.push_back(type ? type[gb] : 0); gj.push_back(gb); }
      },
      [&](size_t p_a)
      {
        const size_t ga = cell_off[cell_a] + p_a;
        const int n = int(dx.size());
        dedr.assign(3 * size_t(n) + 3, 0.0);
        double e = 0.0;
        orc_snap_atom(snap, n, dx.data(), dy.data(), dz.data(), ej.data(), type ? type[ga] : 0, nullptr, nullptr, &e, dedr.data());
        for(int i = 0; i < n; i++)
        {
          const double* f = &dedr[3 * i];
          fx[ga] += f[0]; fy[ga] += f[1]; fz[ga] += f[2];
          fx[gj[i]] -= f[0]; fy[gj[i]] -= f[1]; fz[gj[i]] -= f[2];
          if( vflag && vir )
          {
            double* v = vir + 9 * ga;
            v[0] -= f[0]*dx[i]; v[1] -= f[0]*dy[i]; v[2] -= f[0]*dz[i];
            v[3] -= f[1]*dx[i]; v[4] -= f[1]*dy[i]; v[5] -= f[1]*dz[i];
            v[6] -= f[2]*dx[i]; v[7] -= f[2]*dy[i]; v[8] -= f[2]*dz[i];
          }
        }
        if( eflag && ep ) ep[ga] += e;
      });
  }
}

// scalar entry points used to pin the restated math against oracle/_ref and tests/golden/ref_math.json
void orc_lj_eval(double epsilon, double sigma, double r, double* e, double* de) { lj_compute_energy(LJParams{epsilon, sigma}, r, *e, *de); }
void orc_johnson_eval(const double* params19, int what, double x, double* f, double* df)
{
  JohnsonParams p; std::memcpy(&p, params19, sizeof(p));
  if( what == 0 ) johnson_phi(p, x, *f, *df);
  else if( what == 1 ) johnson_rho(p, x, *f, *df);
  else johnson_fEmbed(p, x, *f, *df);
}

void orc_eam_analytic_eval(int model, const double* params, int what, double x, double* f, double* df)
{
  const EamAnalytic p(model, params);
  if( what == 0 ) p.phi(x, *f, *df); else if( what == 1 ) p.rho(x, *f, *df); else p.fEmbed(x, *f, *df);
}

int orc_num_threads() { return omp_get_max_threads(); }
// bench.py's CPU arm sets the team size explicitly: torchrun exports OMP_NUM_THREADS=1 to every rank
void orc_set_num_threads(int n) { if( n > 0 ) omp_set_num_threads(n); }
double orc_ev_internal() { return EV_INTERNAL; }

} // extern "C"
