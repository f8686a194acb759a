// FCIDUMP and binary-RDM file I/O at the edges of the CI path: the formats MACIS reads its
// integrals from and writes its results to (external/macis/include/macis/util/fcidump.hpp:26-155,
// src/macis/fcidump.cxx), so fixtures and results interchange with MACIS / pymacis. Same function
// names, argument meaning and error behaviour (std::runtime_error); arrays are column-major:
// T[p + q LDT], V[p + q LDV + r LDV^2 + s LDV^3] = (pq|rs).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace qdk_b200::io {

struct FCIDumpHeader {  // fcidump.hpp:26-32
  uint32_t norb = 0;
  uint32_t nelec = 0;
  int32_t ms2 = 0;
  int32_t isym = 1;
  std::vector<int32_t> orbsym;
};

FCIDumpHeader fcidump_read_header(const std::string& fname);
uint32_t read_fcidump_norb(const std::string& fname);
double read_fcidump_core(const std::string& fname);
void read_fcidump_1body(const std::string& fname, double* T, size_t LDT);
void read_fcidump_2body(const std::string& fname, double* V, size_t LDV);
// one pass over the file (fcidump.hpp:103-117)
void read_fcidump_all(const std::string& fname, double* T, size_t LDT, double* V, size_t LDV, double& E_core);
// every element with |value| >= threshold, two-body first, then one-body, then the core energy
// (fcidump.cxx:440-485: "integral p q r s" lines, 1-based indices)
void write_fcidump(const std::string& fname, const FCIDumpHeader& header, const double* T, size_t LDT,
                   const double* V, size_t LDV, double E_core, double threshold = 1e-15);

// binary RDM files (int32 norb, norb^2 doubles, norb^4 doubles; fcidump.cxx:487-560)
void read_rdms_binary(const std::string& fname, size_t norb, double* ORDM, size_t LDD1, double* TRDM, size_t LDD2);
void write_rdms_binary(const std::string& fname, size_t norb, const double* ORDM, size_t LDD1, const double* TRDM,
                       size_t LDD2);

// ---- text wavefunction files (external/macis/include/macis/wavefunction_io.hpp:22-116):
//   <nstates> <norb> <nalpha> <nbeta>
//   <coefficient, scientific, 16 digits, width 30> <canonical string: one of 0 u d 2 per orbital>
// Determinants are (alpha, beta) occupation words, bit p = orbital p, as everywhere on this path.
std::string to_canonical_string(uint64_t alpha, uint64_t beta, size_t norb);      // sd_operations.hpp:478-498
std::pair<uint64_t, uint64_t> from_canonical_string(const std::string& str);      // sd_operations.hpp:508-525
struct WavefunctionFile {
  size_t nstates = 0, norb = 0, nalpha = 0, nbeta = 0;  // header (informational, like the reference)
  std::vector<uint64_t> alpha, beta;
  std::vector<double> coeffs;
};
WavefunctionFile read_wavefunction(const std::string& fname);
// throws "Invalid Wave Function Dimensions" on a size mismatch; writes nothing for an empty list
void write_wavefunction(const std::string& fname, size_t norb, const std::vector<uint64_t>& alpha,
                        const std::vector<uint64_t>& beta, const std::vector<double>& coeffs);

}  // namespace qdk_b200::io
