// symmetries.h — point-group symmetry lists for the reconstruction path.
//
// Stands in for xmippCore's SymList as used at reconstruct_fourier.cpp:272-286:
// readSymmetryFile(name) accepts a point-group name (c1, cN, cNv, cNh, sN, dN, dNv, dNh, t, td, th,
// o, oh, i, i1..i4, i1h..i4h; src/xmipp/applications/tests/function_tests/test_symmetries_main.cpp:31-51)
// or a symmetry description file (lines "rot_axis <fold> <x> <y> <z>", "mirror_plane <x> <y> <z>",
// "inversion"); trueSymsNo()/getMatrices then enumerate every group element except the identity.
#pragma once
#include <array>
#include <string>
#include <vector>

namespace rfhost {

using Mat3 = std::array<double, 9>;   // row-major

// all non-identity elements of the group; throws std::runtime_error for an unknown name / bad file
std::vector<Mat3> symmetryMatrices(const std::string& nameOrFile);

}  // namespace rfhost
