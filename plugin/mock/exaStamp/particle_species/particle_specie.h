// mock of exaStamp/particle_species/particle_specie.h (src/particle_species/include/...:53-101)
#pragma once
#include <string>
#include <vector>
namespace exaStamp {
struct ParticleSpecie { double m_mass = 0.; unsigned int m_z = 0; std::string name() const { return ""; } };
using ParticleSpecies = std::vector<ParticleSpecie>;
}
