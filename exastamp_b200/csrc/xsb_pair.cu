// xsb_pair.cu -- <pot>_compute_force and <pot>_multi_force.
// Reference functors: src/potential/pair_potential_template/pair_potential_force_op_singlemat.h:45-218 with body
// force_op_impl2.hxx:22-78 (single species) and pair_potential_force_op_multiparam.h:76-223 (per type pair).
// Potential math: src/potential/pair_potentials/lennard_jones/include/.../lennard_jones.h:40-50.
#include "xsb_tilepass.cuh"
#include "xsb_pairpot.cuh"

namespace xsb
{

template<int TPA, bool XFORM, bool MULTI, bool VIRIAL, class real>
__global__ void __launch_bounds__(256) pair_force_kernel(ParticleView P, XForm X, LJMulti prm, double rcut2_max,
                                                          double* __restrict__ fx, double* __restrict__ fy, double* __restrict__ fz,
                                                          double* __restrict__ ep, double* __restrict__ vir)
{
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned g = t / TPA, sub = t % TPA;
  const bool valid = g < P.n_atoms;
  unsigned a = 0; unsigned long long e0 = 0, e1 = 0;
  double xa = 0, ya = 0, za = 0; unsigned ta = 0;
  if( valid )
  {
    a = P.atoms ? P.atoms[g] : g;
    e0 = P.nbh_off[a]; e1 = P.nbh_off[a + 1];
    xa = P.rx[a]; ya = P.ry[a]; za = P.rz[a];
    if( MULTI ) ta = P.type[a];
  }
  double sfx = 0, sfy = 0, sfz = 0, sep = 0;
  double v0 = 0, v1 = 0, v2 = 0, v3 = 0, v4 = 0, v5 = 0, v6 = 0, v7 = 0, v8 = 0;
  for(unsigned long long e = e0 + sub; e < e1; e += TPA)
  {
    const unsigned b = P.nbh_idx[e];
    double dx = P.rx[b] - xa, dy = P.ry[b] - ya, dz = P.rz[b] - za;
    apply_xform<XFORM>(X, dx, dy, dz);
    const double d2 = dx * dx + dy * dy + dz * dz;
    if( d2 <= rcut2_max )
    {
      const LJPair& pp = MULTI ? prm.pp[unique_pair_id(ta, P.type[b])] : prm.pp[0];
      if( !MULTI || d2 <= pp.rcut2 )
      {
        real e_, de_r;
        lj_eval<real>(pp, real(d2), e_, de_r);
        const double fex = double(de_r) * dx, fey = double(de_r) * dy, fez = double(de_r) * dz;
        sfx += fex; sfy += fey; sfz += fez; sep += 0.5 * double(e_);
        if( VIRIAL )
        {
          v0 -= 0.5 * fex * dx; v1 -= 0.5 * fex * dy; v2 -= 0.5 * fex * dz;
          v3 -= 0.5 * fey * dx; v4 -= 0.5 * fey * dy; v5 -= 0.5 * fey * dz;
          v6 -= 0.5 * fez * dx; v7 -= 0.5 * fez * dy; v8 -= 0.5 * fez * dz;
        }
      }
    }
  }
  sfx = group_sum<TPA>(sfx); sfy = group_sum<TPA>(sfy); sfz = group_sum<TPA>(sfz); sep = group_sum<TPA>(sep);
  if( VIRIAL )
  {
    v0 = group_sum<TPA>(v0); v1 = group_sum<TPA>(v1); v2 = group_sum<TPA>(v2);
    v3 = group_sum<TPA>(v3); v4 = group_sum<TPA>(v4); v5 = group_sum<TPA>(v5);
    v6 = group_sum<TPA>(v6); v7 = group_sum<TPA>(v7); v8 = group_sum<TPA>(v8);
  }
  if( valid && sub == 0 )
  {
    fx[a] += sfx; fy[a] += sfy; fz[a] += sfz;
    if( ep ) ep[a] += sep;
    if( VIRIAL )
    {
      double* v = vir + 9ull * a;
      v[0] += v0; v[1] += v1; v[2] += v2; v[3] += v3; v[4] += v4; v[5] += v5; v[6] += v6; v[7] += v7; v[8] += v8;
    }
  }
}

// the same functor for the persistent tile kernel (xsb_tilepass.cuh)
template<bool MULTI, bool VIRIAL, class real>
struct LJTileOp
{
  static constexpr bool HAS_W = false, TYPES = MULTI, D2_ONLY = false;
  double rcut2;
  LJMulti prm;
  double *fx, *fy, *fz, *ep, *vir;
  __host__ __device__ size_t table_bytes() const { return 0; }
  __device__ __forceinline__ void load_tables(unsigned char*, int) const {}
  struct Acc { double fx, fy, fz, ep; unsigned ta; Vir9 v; };
  __device__ __forceinline__ void init(Acc& A) const { A.fx = A.fy = A.fz = A.ep = 0.0; A.ta = 0; if( VIRIAL ) A.v.zero(); }
  __device__ __forceinline__ void start(Acc& A, unsigned, unsigned sa, const StageBuf<HAS_W, TYPES>& B, const unsigned char*) const { if( MULTI ) A.ta = B.t[sa]; }
  __device__ __forceinline__ void pair(Acc& A, double dx, double dy, double dz, double d2, unsigned j, const StageBuf<HAS_W, TYPES>& B, const unsigned char*) const
  {
    const LJPair& pp = MULTI ? prm.pp[unique_pair_id(A.ta, B.t[j])] : prm.pp[0];
    if( !MULTI || d2 <= pp.rcut2 )
    {
      real e_, de_r;
      lj_eval<real>(pp, real(d2), e_, de_r);
      const double fex = double(de_r) * dx, fey = double(de_r) * dy, fez = double(de_r) * dz;
      A.fx += fex; A.fy += fey; A.fz += fez; A.ep += 0.5 * double(e_);
      if( VIRIAL ) A.v.add(fex, fey, fez, dx, dy, dz);
    }
  }
  template<int TPA> __device__ __forceinline__ void finish(Acc& A, unsigned a, bool valid, unsigned sub) const
  {
    A.fx = group_sum<TPA>(A.fx); A.fy = group_sum<TPA>(A.fy); A.fz = group_sum<TPA>(A.fz); A.ep = group_sum<TPA>(A.ep);
    if( VIRIAL ) A.v.template reduce<TPA>();
    if( valid && sub == 0 )
    {
      red_add(fx + a, A.fx); red_add(fy + a, A.fy); red_add(fz + a, A.fz);
      if( ep ) red_add(ep + a, A.ep);
      if( VIRIAL ) A.v.store_add(vir, a);
    }
  }
};

static int pair_nparams(int pot)
{
  switch( pot )
  {
    case XSB_POT_LJ: case XSB_POT_YUKAWA: case XSB_POT_RELAX: return 2;
    case XSB_POT_BUCKINGHAM: return 3;
    case XSB_POT_ZBL: case XSB_POT_EXP6: return 4;
    case XSB_POT_ZERO: return 0;
    default: return -1;
  }
}

// host-side zbl helpers: the r-independent part of zbl_compute_energy (zbl/potential.h:180-246)
static double zbl_e_h(double r, const double* d, double zze) { return zze * (0.02817 * exp(-d[0] * r) + 0.28022 * exp(-d[1] * r) + 0.50986 * exp(-d[2] * r) + 0.18175 * exp(-d[3] * r)) / r; }
static double zbl_de_h(double r, const double* d, double zze)
{
  const double e1 = exp(-d[0] * r), e2 = exp(-d[1] * r), e3 = exp(-d[2] * r), e4 = exp(-d[3] * r), rinv = 1.0 / r;
  const double sum = 0.02817 * e1 + 0.28022 * e2 + 0.50986 * e3 + 0.18175 * e4;
  const double sum_p = -0.02817 * d[0] * e1 - 0.28022 * d[1] * e2 - 0.50986 * d[2] * e3 - 0.18175 * d[3] * e4;
  return zze * (sum_p - sum * rinv) * rinv;
}
static double zbl_d2e_h(double r, const double* d, double zze)
{
  const double e1 = exp(-d[0] * r), e2 = exp(-d[1] * r), e3 = exp(-d[2] * r), e4 = exp(-d[3] * r), rinv = 1.0 / r;
  const double sum = 0.02817 * e1 + 0.28022 * e2 + 0.50986 * e3 + 0.18175 * e4;
  const double sum_p = 0.02817 * e1 * d[0] + 0.28022 * e2 * d[1] + 0.50986 * e3 * d[2] + 0.18175 * e4 * d[3];
  const double sum_pp = 0.02817 * e1 * d[0] * d[0] + 0.28022 * e2 * d[1] * d[1] + 0.50986 * e3 * d[2] * d[2] + 0.18175 * e4 * d[3] * d[3];
  return zze * (sum_pp + 2.0 * sum_p * rinv + 2.0 * sum * rinv * rinv) * rinv;
}

// coefficients of potential `pot` for one type pair from its raw parameters (reference order) and the operator cutoff
static LJPair make_pair(int pot, const double* prm, double rcut)
{
  LJPair p{}; p.pot = pot; p.rcut2 = rcut * rcut; p.ecut = 0.0; p.rc2_pot = 0.0;
  if( pot == XSB_POT_LJ ) { p.k[0] = 4.0 * prm[0]; p.k[1] = 24.0 * prm[0]; p.k[2] = prm[1] * prm[1]; }
  else if( pot == XSB_POT_ZBL )
  {
    const double r1 = prm[0], rc = prm[1], za = prm[2], zb = prm[3];
    const double ainv = (pow(za, 0.23) + pow(zb, 0.23)) / 0.46850;
    p.k[0] = 0.20162 * ainv; p.k[1] = 0.40290 * ainv; p.k[2] = 0.94229 * ainv; p.k[3] = 3.19980 * ainv;
    p.k[4] = za * zb * 14.399645;
    const double tc = rc - r1, fc = zbl_e_h(rc, p.k, p.k[4]), fcp = zbl_de_h(rc, p.k, p.k[4]), fcpp = zbl_d2e_h(rc, p.k, p.k[4]);
    const double swa = (-3.0 * fcp + tc * fcpp) / (tc * tc), swb = (2.0 * fcp - tc * fcpp) / (tc * tc * tc);
    p.k[5] = swa; p.k[6] = swb; p.k[7] = swa / 3.0; p.k[8] = swb / 4.0; p.k[9] = -fc + (tc / 2.0) * fcp - (tc * tc / 12.0) * fcpp;
    p.k[10] = r1; p.rc2_pot = rc * rc;
  }
  else { for(int i = 0; i < pair_nparams(pot); i++) p.k[i] = prm[i]; }
  if( rcut > 0.0 )   // energy_cutoff(): e(rcut) through the same evaluation (pair_potential_impl.hxx:488-498)
  {
    double e = 0.0, de = 0.0;
    lj_eval<double>(p, rcut * rcut, e, de);
    p.ecut = e;     // lj_eval subtracted ecut = 0
  }
  return p;
}

// A deferred EAM force phase (xsb_ctx::pending_eam) takes this operator along when its pairs are a subset of the ones that
// pass visits and both accumulate the same fields; otherwise the EAM phase is launched first and the operator runs on its own.
template<bool MULTI>
static int join_pending_eam(xsb_ctx* ctx, const LJMulti& prm, int n_types, double rcut_max, int flags, bool* absorbed)
{
  *absorbed = false;
  if( !ctx->pending_eam.active ) return XSB_OK;
  const xsb_ctx::PendingEamForce pe = ctx->pending_eam;
  const bool eflag = pe.phases & XSB_EAM_EFLAG, virial = eflag && (pe.flags & XSB_FLAG_VIRIAL);
  const bool fits = prm.pp[0].pot == XSB_POT_LJ && !(flags & XSB_FLAG_GHOST) && bool(flags & XSB_FLAG_ENERGY) == eflag && bool(flags & XSB_FLAG_VIRIAL) == virial &&
                    bool(flags & XSB_FLAG_MIXED) == bool(pe.flags & XSB_FLAG_MIXED) && rcut_max <= pe.rcut &&
                    !(MULTI && n_types > 1 && ctx->eam.nelements < 2);      // a single-element pass stages no type bytes
  if( !fits ) return xsb_internal_flush_pending(ctx);
  ctx->pending_eam.active = false;
  LJMulti all = prm;
  if( !MULTI ) for(int i = 1; i < 16; i++) all.pp[i] = prm.pp[0];          // one parameter set for every type pair
  const int rc = xsb_internal_eam_force_phase(ctx, pe.rcut, pe.phases, pe.flags, &all);
  if( rc ) return rc;
  ctx->fused_chains++; *absorbed = true;
  return XSB_OK;
}

template<bool MULTI>
static int launch_pair(xsb_ctx* ctx, const LJMulti& prm, double rcut_max, int flags)
{
  XSB_REQUIRE(ctx, ctx->nbh_built, XSB_ERR_STATE, "chunk_neighbors must be built before a force operator");
  XSB_REQUIRE(ctx, rcut_max <= ctx->nbh_dist, XSB_ERR_INVALID, "rcut exceeds the neighbour-list distance nbh_dist_lab");
  XSB_CUDA(ctx, cudaSetDevice(ctx->device));
  const bool ghost = flags & XSB_FLAG_GHOST, virial = flags & XSB_FLAG_VIRIAL, mixed = flags & XSB_FLAG_MIXED;
  if( virial ) { int rc = xsb_internal_ensure_virial(ctx); if( rc ) return rc; }
  ParticleView P{ ctx->f64[XSB_F_RX].p, ctx->f64[XSB_F_RY].p, ctx->f64[XSB_F_RZ].p, ctx->type.p, ctx->nbh_off.p, ctx->nbh_idx.p,
                  ghost ? nullptr : ctx->own_atoms.p, unsigned(ghost ? ctx->n : ctx->n_own) };
  if( P.n_atoms == 0 ) return XSB_OK;
  if( ctx->tile_ok )
  {
    double *tfx = ctx->f64[XSB_F_FX].p, *tfy = ctx->f64[XSB_F_FY].p, *tfz = ctx->f64[XSB_F_FZ].p;
    double *tep = (flags & XSB_FLAG_ENERGY) ? ctx->f64[XSB_F_EP].p : nullptr, *tvir = virial ? ctx->f64[XSB_F_VIRIAL].p : nullptr;
    int rc;
    ctx->prof_begin(XSB_PROF_PAIR);
    // behind an EAM operator of the same step (compute_force: [eam_alloy_force, lj_multi_force], configs[4]) the in-range sub-list
    // that operator left is a superset of this potential's pairs when its cut-off is the larger one: walk it instead of the
    // full list (the distance test against this operator's own cut-off stays)
    const int lmode = ctx->sub_covers(rcut_max, ghost) ? LIST_SUB : LIST_FULL;
#   define XSB_LJ_TILE(VIR, REAL, TPA_, NT_) { LJTileOp<MULTI, VIR, REAL> op{ rcut_max * rcut_max, prm, tfx, tfy, tfz, tep, tvir }; rc = launch_tile_pass<TPA_, NT_>(ctx, ghost, op, nullptr, lmode); }
    if( mixed ) { if( virial ) XSB_LJ_TILE(true, float, 8, 512) else XSB_LJ_TILE(false, float, 16, 1024) }
    else        { if( virial ) XSB_LJ_TILE(true, double, 8, 512) else if( ctx->exp_tpa == 8 ) XSB_LJ_TILE(false, double, 8, 1024) else XSB_LJ_TILE(false, double, 16, 1024) }
#   undef XSB_LJ_TILE
    ctx->prof_end(XSB_PROF_PAIR);
    return rc;
  }
  const XForm X = make_xform(ctx->grid);
  constexpr int TPA = 8; const int block = 256;
  const unsigned grid = groups_grid<TPA>(P.n_atoms, block);
  double *fx = ctx->f64[XSB_F_FX].p, *fy = ctx->f64[XSB_F_FY].p, *fz = ctx->f64[XSB_F_FZ].p;
  double *ep = (flags & XSB_FLAG_ENERGY) ? ctx->f64[XSB_F_EP].p : nullptr, *vir = virial ? ctx->f64[XSB_F_VIRIAL].p : nullptr;
  const double rc2 = rcut_max * rcut_max;
  const bool xf = !ctx->grid.xform_is_identity;
# define XSB_PAIR_GO(XF, VIR, REAL) pair_force_kernel<TPA, XF, MULTI, VIR, REAL><<<grid, block, 0, ctx->stream>>>(P, X, prm, rc2, fx, fy, fz, ep, vir)
  ctx->prof_begin(XSB_PROF_PAIR);
  if( mixed ) { if( xf ) { if( virial ) XSB_PAIR_GO(true, true, float); else XSB_PAIR_GO(true, false, float); }
                else     { if( virial ) XSB_PAIR_GO(false, true, float); else XSB_PAIR_GO(false, false, float); } }
  else        { if( xf ) { if( virial ) XSB_PAIR_GO(true, true, double); else XSB_PAIR_GO(true, false, double); }
                else     { if( virial ) XSB_PAIR_GO(false, true, double); else XSB_PAIR_GO(false, false, double); } }
# undef XSB_PAIR_GO
  ctx->prof_end(XSB_PROF_PAIR);
  XSB_LAUNCH_CHECK(ctx);
  return XSB_OK;
}

} // namespace xsb

using namespace xsb;

extern "C" {

int xsb_pair_force(xsb_ctx* ctx, int pot, const double* params, int nparams, double rcut, int flags)
{
  XSB_ENTER_KEEP(ctx);
  XSB_REQUIRE(ctx, pair_nparams(pot) >= 0, XSB_ERR_UNSUPPORTED, "pair potential not implemented (lj, zbl, exp6, buckingham, yukawa, relax, zero)");
  XSB_REQUIRE(ctx, (params != nullptr || pair_nparams(pot) == 0) && nparams == pair_nparams(pot), XSB_ERR_INVALID,
              "wrong parameter count: lj {epsilon, sigma}, zbl {r1, rc, z_a, z_b}, exp6 {A, B, C, D}, buckingham {A, Rho, C}, yukawa {A, kappa}, relax {r1, rc}, zero {}");
  XSB_REQUIRE(ctx, rcut > 0.0, XSB_ERR_INVALID, "rcut must be > 0");
  LJMulti prm; prm.pp[0] = make_pair(pot, params, rcut);
  bool absorbed; const int jrc = join_pending_eam<false>(ctx, prm, 1, rcut, flags, &absorbed);
  if( jrc || absorbed ) return jrc;
  return launch_pair<false>(ctx, prm, rcut, flags);
}

int xsb_pair_multi_force(xsb_ctx* ctx, int pot, int n_types, const double* pair_params, int nparams, double rcut_max, int flags)
{
  XSB_ENTER_KEEP(ctx);
  XSB_REQUIRE(ctx, pair_nparams(pot) >= 0, XSB_ERR_UNSUPPORTED, "pair potential not implemented (lj, zbl, exp6, buckingham, yukawa, relax, zero)");
  XSB_REQUIRE(ctx, pair_params != nullptr && nparams == pair_nparams(pot), XSB_ERR_INVALID, "rows of {params..., rcut} expected");
  const int npairs = n_types * (n_types + 1) / 2;
  XSB_REQUIRE(ctx, n_types >= 1 && npairs <= 16, XSB_ERR_INVALID, "too many type pairs (MAX_TYPE_PAIR_IDS = 16)");
  LJMulti prm; double rmax = 0.0;
  for(int i = 0; i < npairs; i++)
  {
    const double* row = pair_params + size_t(nparams + 1) * i;
    prm.pp[i] = make_pair(pot, row, row[nparams]);
    if( row[nparams] > rmax ) rmax = row[nparams];
  }
  XSB_REQUIRE(ctx, rcut_max >= rmax, XSB_ERR_INVALID, "rcut_max is smaller than a pair rcut");
  bool absorbed; const int jrc = join_pending_eam<true>(ctx, prm, n_types, rcut_max, flags, &absorbed);
  if( jrc || absorbed ) return jrc;
  return launch_pair<true>(ctx, prm, rcut_max, flags);
}

// how many pair operators were evaluated inside the force pass of the eam_alloy_force operator in front of them
int xsb_chain_stats(xsb_ctx* ctx, uint64_t* fused_pair_operators)
{
  XSB_ENTER(ctx);
  if( fused_pair_operators ) *fused_pair_operators = ctx->fused_chains;
  return XSB_OK;
}

} // extern "C"
