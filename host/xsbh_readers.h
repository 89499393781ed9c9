// xsbh_readers.h -- data files of the force path that the C ABI does not read itself (setfl is read by
// xsb_eam_alloy_read): LAMMPS .snapparam / .snapcoeff (reference reader: src/potential/snaplegacy/lib/
// snap_read_lammps.cpp:25-93 -- note it pre-multiplies coefficients by 1e-4 e/amu, we keep eV and convert in the
// operator) and (extended) xyz (src/io/read_xyz_file_with_xform.cpp).
#pragma once
#include <string>
#include <vector>

namespace xsbh {

struct SnapFiles {
  int twojmax = 0, switchflag = 1, bzeroflag = 1, quadraticflag = 0, chemflag = 0;
  double rcutfac = 0.0, rfac0 = 0.99363, rmin0 = 0.0;
  int ncoeff_all = 0;                       // coefficients per element including beta0
  std::vector<std::string> elements;
  std::vector<double> radelem, wjelem;
  std::vector<double> beta;                 // [nelements][ncoeff_all], eV
};
SnapFiles read_snap_files(const std::string& param_path, const std::string& coef_path);

struct XyzData {
  double cell[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};     // columns = cell vectors a, b, c (row-major storage)
  std::vector<std::string> species;
  std::vector<double> x, y, z, vx, vy, vz;
};
XyzData read_xyz(const std::string& path, bool read_velocities);

}  // namespace xsbh
