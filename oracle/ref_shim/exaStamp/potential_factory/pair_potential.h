#pragma once
#include "../../xsref_common.h"
namespace exanb {}
namespace exaStamp {
  // only the type name is needed by the pair math headers (parameter is unused by lj / zbl takes z)
  struct PairPotentialAtom { double m_mass = 0, m_charge = 0; unsigned m_z = 0; };
  struct PairPotentialMinimalParameters { PairPotentialAtom m_atom_a, m_atom_b; };
}
