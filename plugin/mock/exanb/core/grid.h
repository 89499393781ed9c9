// mock of exanb/core/grid.h + domain.h: the accessors listed in SURVEY.md 8a row a1
#pragma once
#include <cstddef>
#include <cstdint>
#include <onika/math/basic_types.h>
namespace exanb {
using onika::math::Vec3d; using onika::math::Mat3d; using onika::math::IJK;
namespace field {
struct _rx {}; struct _ry {}; struct _rz {}; struct _fx {}; struct _fy {}; struct _fz {}; struct _ep {}; struct _vx {}; struct _vy {}; struct _vz {};
struct _type {}; struct _id {}; struct _virial {};
static constexpr _rx rx{}; static constexpr _ry ry{}; static constexpr _rz rz{}; static constexpr _fx fx{}; static constexpr _fy fy{}; static constexpr _fz fz{};
static constexpr _ep ep{}; static constexpr _vx vx{}; static constexpr _vy vy{}; static constexpr _vz vz{}; static constexpr _type type{}; static constexpr _id id{};
static constexpr _virial virial{};
}
struct CellParticlesMock {
  size_t size() const { return 0; }
  template<class F> double* operator[](F) const { return nullptr; }
  const uint8_t* operator[](field::_type) const { return nullptr; }
  const uint64_t* operator[](field::_id) const { return nullptr; }
};
template<class... Fields> struct GridMock {
  using CellParticles = CellParticlesMock;
  IJK dimension() const { return IJK(); }
  long ghost_layers() const { return 1; }
  size_t number_of_cells() const { return 0; }
  size_t number_of_particles() const { return 0; }
  const CellParticles* cells() const { return nullptr; }
  CellParticles* cells() { return nullptr; }
  const size_t* cell_particle_offset_data() const { return nullptr; }
  Vec3d cell_position(const IJK&) const { return Vec3d(); }
  template<class F> bool has_allocated_field(F) const { return false; }
};
struct Domain {
  double cell_size() const { return 1.0; }
  const Mat3d& xform() const { static Mat3d m; return m; }
  bool xform_is_identity() const { return true; }
  bool periodic_boundary_x() const { return true; }
  bool periodic_boundary_y() const { return true; }
  bool periodic_boundary_z() const { return true; }
  IJK grid_dimension() const { return IJK(); }
};
struct GridChunkNeighbors {};
struct GridParticleLocks {};
}
