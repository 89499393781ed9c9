#include "xsbh_readers.h"

#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>

#include "xsbh_operator.h"

namespace xsbh {

static bool next_data_line(std::istream& in, std::string& line) {
  while (std::getline(in, line)) {
    size_t h = line.find('#');
    if (h != std::string::npos) line = line.substr(0, h);
    size_t b = line.find_first_not_of(" \t\r\n");
    if (b == std::string::npos) continue;
    line = line.substr(b);
    return true;
  }
  return false;
}

SnapFiles read_snap_files(const std::string& param_path, const std::string& coef_path) {
  SnapFiles s;
  std::ifstream pf(param_path);
  if (!pf) throw OperatorError("cannot open SNAP parameter file '" + param_path + "'");
  std::string line;
  bool have_rcut = false, have_twoj = false;
  while (next_data_line(pf, line)) {
    std::istringstream is(line);
    std::string key, val;
    if (!(is >> key >> val)) throw OperatorError(param_path + ": malformed line '" + line + "'");
    if (key == "rcutfac") { s.rcutfac = std::atof(val.c_str()); have_rcut = true; }
    else if (key == "twojmax") { s.twojmax = std::atoi(val.c_str()); have_twoj = true; }
    else if (key == "rfac0") s.rfac0 = std::atof(val.c_str());
    else if (key == "rmin0") s.rmin0 = std::atof(val.c_str());
    else if (key == "switchflag") s.switchflag = std::atoi(val.c_str());
    else if (key == "bzeroflag") s.bzeroflag = std::atoi(val.c_str());
    else if (key == "quadraticflag") s.quadraticflag = std::atoi(val.c_str());
    else if (key == "chemflag") s.chemflag = std::atoi(val.c_str());
    else if (key == "bnormflag" || key == "wselfallflag" || key == "switchinnerflag" || key == "diagonalstyle") {
      if (std::atoi(val.c_str()) != 0 && key != "diagonalstyle") throw OperatorError(param_path + ": " + key + " != 0 is not supported");
    } else throw OperatorError(param_path + ": unknown keyword '" + key + "'");
  }
  if (!have_rcut || !have_twoj) throw OperatorError(param_path + ": rcutfac and twojmax are required");
  std::ifstream cf(coef_path);
  if (!cf) throw OperatorError("cannot open SNAP coefficient file '" + coef_path + "'");
  if (!next_data_line(cf, line)) throw OperatorError(coef_path + ": empty file");
  int nel = 0;
  { std::istringstream is(line); if (!(is >> nel >> s.ncoeff_all) || nel < 1 || s.ncoeff_all < 1) throw OperatorError(coef_path + ": bad header '" + line + "'"); }
  for (int e = 0; e < nel; ++e) {
    if (!next_data_line(cf, line)) throw OperatorError(coef_path + ": missing element block");
    std::istringstream is(line);
    std::string el; double rad, w;
    if (!(is >> el >> rad >> w)) throw OperatorError(coef_path + ": bad element line '" + line + "'");
    s.elements.push_back(el); s.radelem.push_back(rad); s.wjelem.push_back(w);
    for (int k = 0; k < s.ncoeff_all; ++k) {
      if (!next_data_line(cf, line)) throw OperatorError(coef_path + ": missing coefficient");
      s.beta.push_back(std::atof(line.c_str()));
    }
  }
  return s;
}

XyzData read_xyz(const std::string& path, bool read_velocities) {
  std::ifstream in(path);
  if (!in) throw OperatorError("cannot open xyz file '" + path + "'");
  XyzData d;
  std::string line;
  if (!std::getline(in, line)) throw OperatorError(path + ": empty file");
  const long n = std::atol(line.c_str());
  if (n < 0) throw OperatorError(path + ": bad atom count");
  if (!std::getline(in, line)) throw OperatorError(path + ": missing comment line");
  size_t lp = line.find("Lattice=\"");
  if (lp != std::string::npos) {
    std::istringstream is(line.substr(lp + 9));
    double v[9];
    for (int k = 0; k < 9; ++k) if (!(is >> v[k])) throw OperatorError(path + ": bad Lattice=\"...\" entry");
    // extended xyz lists the vectors a, b, c; store them as columns
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) d.cell[3 * r + c] = v[3 * c + r];
  } else {
    std::istringstream is(line);
    double L[3];
    if (!(is >> L[0] >> L[1] >> L[2])) throw OperatorError(path + ": line 2 must hold the cell (Lattice=\"...\" or three lengths)");
    for (int k = 0; k < 9; ++k) d.cell[k] = 0.0;
    d.cell[0] = L[0]; d.cell[4] = L[1]; d.cell[8] = L[2];
  }
  for (long i = 0; i < n; ++i) {
    if (!std::getline(in, line)) throw OperatorError(path + ": file ends before atom " + std::to_string(i));
    std::istringstream is(line);
    std::string sp; double x, y, z;
    if (!(is >> sp >> x >> y >> z)) throw OperatorError(path + ": bad atom line '" + line + "'");
    d.species.push_back(sp); d.x.push_back(x); d.y.push_back(y); d.z.push_back(z);
    if (read_velocities) {
      double vx, vy, vz;
      if (!(is >> vx >> vy >> vz)) throw OperatorError(path + ": velocities requested but missing on line '" + line + "'");
      d.vx.push_back(vx); d.vy.push_back(vy); d.vz.push_back(vz);
    }
  }
  return d;
}

}  // namespace xsbh
