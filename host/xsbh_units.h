// xsbh_units.h -- "0.0104 eV", "3.4 ang", "2.522E-20 J", "50. m/s": quantities with units as the decks write them,
// converted to exaStamp's internal unit system (angstrom, Da, ps, e, K, particle, cd, rad:
// include/exaStamp/unit_system.h:28-36).  Restates onika::physics::Quantity::convert (ext, onika/physics/units.h)
// for the unit names that occur under data/ (length m nm ang, mass kg g Da, time s ps fs, charge C e-, temperature K,
// amount mol particle, angle rad degree, energy J eV kcal cal, plus products / quotients / integer powers).
#pragma once
#include <string>

#include "xsbh_yaml.h"

namespace xsbh {

struct UnitError : std::runtime_error { using std::runtime_error::runtime_error; };

// multiplicative factor that converts one `unit_expr` (e.g. "kcal/mol/ang^2") to internal units
double unit_factor(const std::string& unit_expr);
// "8.0 ang" -> 8.0 ; "1.0e-3 ps" -> 1e-3 ; "300" -> 300 (no unit: already internal, like the reference)
double quantity(const std::string& text);
double quantity(const Node& n);
inline double quantity_or(const Node* n, double dflt) { return n && !n->is_null() ? quantity(*n) : dflt; }

constexpr double kAvogadro = 6.02214076e23;
constexpr double kElementaryCharge = 1.602176634e-19;   // C
constexpr double kDalton = 1.66053906660e-27;           // kg
constexpr double kBoltzmannSI = 1.380649e-23;           // J/K
// 1 internal energy unit = Da ang^2 / ps^2 = 1.66053906660e-23 J
constexpr double kInternalEnergyJ = kDalton * 1.0e-20 / 1.0e-24;
constexpr double kEv = kElementaryCharge / kInternalEnergyJ;   // 1 eV in internal units (= EXASTAMP_CONST_QUANTITY(1 eV))
constexpr double kBoltzmann = kBoltzmannSI / kInternalEnergyJ; // internal energy per K

}  // namespace xsbh
