// FCIDUMP / binary RDM I/O (see qdk_b200/fcidump.hpp). Reader semantics follow
// external/macis/src/macis/fcidump.cxx: the header runs from "&FCI" to "&END"; the first
// five-token line after it fixes the layout ("integral p q r s" or "p q r s integral"); indices
// are 1-based; (0,0,0,0) is the core energy, lines with all four indices non-zero are two-body
// integrals stored under the eight permutations (pq|rs) = (pq|sr) = ... = (sr|qp), everything else
// is a one-body element stored symmetrically.
#include "qdk_b200/fcidump.hpp"

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <regex>
#include <sstream>
#include <stdexcept>

namespace qdk_b200::io {
namespace {

std::string slurp(const std::string& fname) {
  std::ifstream file(fname, std::ios::binary);
  if (!file.is_open()) throw std::runtime_error("Could not open file: " + fname);
  std::ostringstream ss;
  ss << file.rdbuf();
  return ss.str();
}

bool looks_float(const std::string& tok) {  // is_float of the reference: a letter (exponent) or a point
  for (char c : tok)
    if (std::isalpha(static_cast<unsigned char>(c)) || c == '.') return true;
  return false;
}

enum class Layout { Unknown, IntegralFirst, IndicesFirst };

Layout detect_layout(const std::string& line) {
  std::istringstream ss(line);
  std::vector<std::string> tok;
  for (std::string t; ss >> t;) tok.push_back(t);
  if (tok.size() != 5) return Layout::Unknown;
  const bool first = looks_float(tok[0]), last = looks_float(tok[4]);
  if (first && !last) return Layout::IntegralFirst;
  if (!first && last) return Layout::IndicesFirst;
  // both or neither: non-negative integers in the first four places mean indices first
  for (int i = 0; i < 4; ++i) {
    char* end = nullptr;
    const long v = std::strtol(tok[i].c_str(), &end, 10);
    if (end == tok[i].c_str() || v < 0) return Layout::IntegralFirst;
  }
  return Layout::IndicesFirst;
}

// calls f(p, q, r, s, value) for every data line after the header
void for_each_entry(const std::string& fname, const std::function<bool(int, int, int, int, double)>& f) {
  const std::string content = slurp(fname);
  const char* ptr = content.c_str();
  const char* end = ptr + content.size();
  bool header_passed = false;
  Layout layout = Layout::Unknown;
  std::string line;
  while (ptr < end) {
    const char* le = static_cast<const char*>(std::memchr(ptr, '\n', size_t(end - ptr)));
    if (!le) le = end;
    line.assign(ptr, le);
    ptr = le + 1;
    if (!header_passed) {
      if (line.find("&END") != std::string::npos) header_passed = true;
      continue;
    }
    if (layout == Layout::Unknown) {
      layout = detect_layout(line);
      if (layout == Layout::Unknown) continue;
    }
    int p, q, r, s;
    double v;
    const int parsed = layout == Layout::IntegralFirst ? std::sscanf(line.c_str(), "%lf %d %d %d %d", &v, &p, &q, &r, &s)
                                                       : std::sscanf(line.c_str(), "%d %d %d %d %lf", &p, &q, &r, &s, &v);
    if (parsed != 5) continue;
    if (!f(p, q, r, s, v)) return;
  }
}

inline bool is_core(int p, int q, int r, int s) { return !(p || q || r || s); }
inline bool is_two_body(int p, int q, int r, int s) { return p && q && r && s; }

void check_index(int p, uint32_t norb) {
  if (p < 1 || uint32_t(p) > norb) throw std::runtime_error("FCIDUMP orbital index out of range");
}

void store_two_body(double* V, size_t L, int p, int q, int r, int s, double v) {
  const size_t L2 = L * L, L3 = L2 * L;
  const size_t P = size_t(p - 1), Q = size_t(q - 1), R = size_t(r - 1), S = size_t(s - 1);
  V[P + Q * L + R * L2 + S * L3] = v;
  V[P + Q * L + S * L2 + R * L3] = v;
  V[Q + P * L + R * L2 + S * L3] = v;
  V[Q + P * L + S * L2 + R * L3] = v;
  V[R + S * L + P * L2 + Q * L3] = v;
  V[S + R * L + P * L2 + Q * L3] = v;
  V[R + S * L + Q * L2 + P * L3] = v;
  V[S + R * L + Q * L2 + P * L3] = v;
}

}  // namespace

FCIDumpHeader fcidump_read_header(const std::string& fname) {
  std::ifstream file(fname);
  std::string line, text;
  bool in_header = false;
  while (std::getline(file, line)) {
    if (line.find("&FCI") != std::string::npos) in_header = true;
    if (in_header) {
      text += line + " ";
      if (line.find("&END") != std::string::npos) break;
    }
  }
  if (text.empty()) throw std::runtime_error("No FCIDUMP header found");
  FCIDumpHeader h;
  std::smatch m;
  if (std::regex_search(text, m, std::regex(R"(NORB\s*=\s*(\d+))"))) h.norb = uint32_t(std::stoul(m[1].str()));
  if (std::regex_search(text, m, std::regex(R"(NELEC\s*=\s*(\d+))"))) h.nelec = uint32_t(std::stoul(m[1].str()));
  if (std::regex_search(text, m, std::regex(R"(MS2\s*=\s*(-?\d+))"))) h.ms2 = std::stoi(m[1].str());
  if (std::regex_search(text, m, std::regex(R"(ISYM\s*=\s*(\d+))"))) h.isym = std::stoi(m[1].str());
  if (std::regex_search(text, m, std::regex(R"(ORBSYM\s*=\s*([\d,\s]+))"))) {
    const std::string list = m[1].str();
    const std::regex num(R"(\d+)");
    for (std::sregex_iterator it(list.begin(), list.end(), num), e; it != e; ++it) h.orbsym.push_back(std::stoi(it->str()));
  }
  return h;
}

uint32_t read_fcidump_norb(const std::string& fname) {
  const FCIDumpHeader h = fcidump_read_header(fname);
  if (h.norb == 0) throw std::runtime_error("NORB not found or is zero in FCIDUMP header");
  return h.norb;
}

double read_fcidump_core(const std::string& fname) {
  double core = 0.0;
  for_each_entry(fname, [&](int p, int q, int r, int s, double v) {
    if (is_core(p, q, r, s)) { core = v; return false; }
    return true;
  });
  return core;
}

void read_fcidump_1body(const std::string& fname, double* T, size_t LDT) {
  const uint32_t norb = read_fcidump_norb(fname);
  if (LDT < norb) throw std::runtime_error("T is of improper dimension");
  for_each_entry(fname, [&](int p, int q, int r, int s, double v) {
    if (!is_core(p, q, r, s) && !is_two_body(p, q, r, s)) {
      check_index(p, norb); check_index(q, norb);
      T[size_t(p - 1) + size_t(q - 1) * LDT] = v;
      T[size_t(q - 1) + size_t(p - 1) * LDT] = v;
    }
    return true;
  });
}

void read_fcidump_2body(const std::string& fname, double* V, size_t LDV) {
  const uint32_t norb = read_fcidump_norb(fname);
  if (LDV < norb) throw std::runtime_error("V is of improper dimension");
  for_each_entry(fname, [&](int p, int q, int r, int s, double v) {
    if (is_two_body(p, q, r, s)) {
      check_index(p, norb); check_index(q, norb); check_index(r, norb); check_index(s, norb);
      store_two_body(V, LDV, p, q, r, s, v);
    }
    return true;
  });
}

void read_fcidump_all(const std::string& fname, double* T, size_t LDT, double* V, size_t LDV, double& E_core) {
  const uint32_t norb = read_fcidump_norb(fname);
  if (LDT < norb) throw std::runtime_error("T is of improper dimension");
  if (LDV < norb) throw std::runtime_error("V is of improper dimension");
  E_core = 0.0;
  bool core_seen = false;
  for (size_t j = 0; j < norb; ++j)  // unlike the single-purpose readers this one zeroes its outputs
    for (size_t i = 0; i < norb; ++i) T[i + j * LDT] = 0.0;
  {
    const size_t L2 = LDV * LDV, L3 = L2 * LDV;
    for (size_t l = 0; l < norb; ++l)
      for (size_t k = 0; k < norb; ++k)
        for (size_t j = 0; j < norb; ++j)
          for (size_t i = 0; i < norb; ++i) V[i + j * LDV + k * L2 + l * L3] = 0.0;
  }
  for_each_entry(fname, [&](int p, int q, int r, int s, double v) {
    if (is_core(p, q, r, s)) {
      if (!core_seen) { E_core = v; core_seen = true; }
    } else if (is_two_body(p, q, r, s)) {
      check_index(p, norb); check_index(q, norb); check_index(r, norb); check_index(s, norb);
      store_two_body(V, LDV, p, q, r, s, v);
    } else {
      check_index(p, norb); check_index(q, norb);
      T[size_t(p - 1) + size_t(q - 1) * LDT] = v;
      T[size_t(q - 1) + size_t(p - 1) * LDT] = v;
    }
    return true;
  });
}

void write_fcidump(const std::string& fname, const FCIDumpHeader& header, const double* T, size_t LDT, const double* V,
                   size_t LDV, double E_core, double threshold) {
  std::FILE* fh = std::fopen(fname.c_str(), "w");
  if (!fh) throw std::runtime_error("Could not open file: " + fname);
  std::fprintf(fh, "&FCI NORB=%u,NELEC=%u,MS2=%d,\n  ISYM=%d,\n", header.norb, header.nelec, header.ms2, header.isym);
  if (!header.orbsym.empty()) {
    std::fprintf(fh, "  ORBSYM=");
    for (size_t i = 0; i < header.orbsym.size(); ++i) std::fprintf(fh, i ? ",%d" : "%d", header.orbsym[i]);
    std::fprintf(fh, "\n");
  }
  std::fprintf(fh, "&END\n");
  const size_t n = header.norb, L2 = LDV * LDV, L3 = L2 * LDV;
  auto put = [&](double v, size_t p, size_t q, size_t r, size_t s) {
    std::fprintf(fh, "%25.14e %8zu %8zu %8zu %8zu\n", v, p, q, r, s);
  };
  for (size_t i = 0; i < n; ++i)
    for (size_t j = 0; j < n; ++j)
      for (size_t k = 0; k < n; ++k)
        for (size_t l = 0; l < n; ++l) {
          const double v = V[i + j * LDV + k * L2 + l * L3];
          if (std::abs(v) < threshold) continue;
          put(v, i + 1, j + 1, k + 1, l + 1);
        }
  for (size_t i = 0; i < n; ++i)
    for (size_t j = 0; j < n; ++j) {
      const double v = T[i + j * LDT];
      if (std::abs(v) < threshold) continue;
      put(v, i + 1, j + 1, 0, 0);
    }
  put(E_core, 0, 0, 0, 0);
  std::fclose(fh);
}

void read_rdms_binary(const std::string& fname, size_t norb, double* ORDM, size_t LDD1, double* TRDM, size_t LDD2) {
  std::ifstream in(fname, std::ios::binary);
  if (!in) throw std::runtime_error(fname + " not available");
  int32_t n_read = 0;
  in.read(reinterpret_cast<char*>(&n_read), sizeof(int32_t));
  if (size_t(n_read) != norb)
    throw std::runtime_error("NORB in RDM file doesn't match " + std::to_string(norb) + " " + std::to_string(n_read));
  const size_t n2 = norb * norb, n4 = n2 * n2;
  std::vector<double> raw(n4);
  in.read(reinterpret_cast<char*>(raw.data()), std::streamsize(n2 * sizeof(double)));
  for (size_t i = 0; i < norb; ++i)
    for (size_t j = 0; j < norb; ++j) ORDM[i + j * LDD1] = raw[i + j * norb];
  in.read(reinterpret_cast<char*>(raw.data()), std::streamsize(n4 * sizeof(double)));
  if (!in) throw std::runtime_error(fname + " is truncated");
  const size_t L2 = LDD2 * LDD2, L3 = L2 * LDD2;
  for (size_t i = 0; i < norb; ++i)
    for (size_t j = 0; j < norb; ++j)
      for (size_t k = 0; k < norb; ++k)
        for (size_t l = 0; l < norb; ++l)
          TRDM[i + j * LDD2 + k * L2 + l * L3] = raw[i + j * norb + k * n2 + l * n2 * norb];
}

void write_rdms_binary(const std::string& fname, size_t norb, const double* ORDM, size_t LDD1, const double* TRDM,
                       size_t LDD2) {
  std::ofstream out(fname, std::ios::binary);
  if (!out) throw std::runtime_error("Could not open file: " + fname);
  const int32_t n32 = int32_t(norb);
  out.write(reinterpret_cast<const char*>(&n32), sizeof(int32_t));
  const size_t n2 = norb * norb;
  std::vector<double> raw(n2 * n2);
  for (size_t i = 0; i < norb; ++i)
    for (size_t j = 0; j < norb; ++j) raw[i + j * norb] = ORDM[i + j * LDD1];
  out.write(reinterpret_cast<const char*>(raw.data()), std::streamsize(n2 * sizeof(double)));
  const size_t L2 = LDD2 * LDD2, L3 = L2 * LDD2;
  for (size_t i = 0; i < norb; ++i)
    for (size_t j = 0; j < norb; ++j)
      for (size_t k = 0; k < norb; ++k)
        for (size_t l = 0; l < norb; ++l)
          raw[i + j * norb + k * n2 + l * n2 * norb] = TRDM[i + j * LDD2 + k * L2 + l * L3];
  out.write(reinterpret_cast<const char*>(raw.data()), std::streamsize(n2 * n2 * sizeof(double)));
}

std::string to_canonical_string(uint64_t alpha, uint64_t beta, size_t norb) {
  std::string out;
  out.reserve(norb);
  for (size_t i = 0; i < norb && i < 64; ++i) {
    const bool a = (alpha >> i) & 1u, b = (beta >> i) & 1u;
    out.push_back(a && b ? '2' : a ? 'u' : b ? 'd' : '0');
  }
  return out;
}

std::pair<uint64_t, uint64_t> from_canonical_string(const std::string& str) {
  uint64_t alpha = 0, beta = 0;
  for (size_t i = 0; i < str.size() && i < 64; ++i) {
    if (str[i] == '2') { alpha |= uint64_t(1) << i; beta |= uint64_t(1) << i; }
    else if (str[i] == 'u') alpha |= uint64_t(1) << i;
    else if (str[i] == 'd') beta |= uint64_t(1) << i;
  }
  return {alpha, beta};
}

WavefunctionFile read_wavefunction(const std::string& fname) {
  std::ifstream file(fname);
  if (!file.is_open()) throw std::runtime_error("Could not open file: " + fname);
  WavefunctionFile w;
  std::string line;
  if (std::getline(file, line)) {
    std::istringstream ss(line);
    ss >> w.nstates >> w.norb >> w.nalpha >> w.nbeta;
  }
  w.alpha.reserve(w.nstates);
  w.beta.reserve(w.nstates);
  w.coeffs.reserve(w.nstates);
  while (std::getline(file, line)) {
    std::istringstream ss(line);
    std::string c, d;
    if (!(ss >> c >> d)) continue;
    const auto ab = from_canonical_string(d);
    w.alpha.push_back(ab.first);
    w.beta.push_back(ab.second);
    w.coeffs.push_back(std::stod(c));
  }
  return w;
}

void write_wavefunction(const std::string& fname, size_t norb, const std::vector<uint64_t>& alpha,
                        const std::vector<uint64_t>& beta, const std::vector<double>& coeffs) {
  if (alpha.size() != coeffs.size() || beta.size() != coeffs.size())
    throw std::runtime_error("Invalid Wave Function Dimensions");
  if (coeffs.empty()) return;
  std::FILE* fh = std::fopen(fname.c_str(), "w");
  if (!fh) throw std::runtime_error("Could not open file: " + fname);
  std::fprintf(fh, "%zu %zu %d %d\n", coeffs.size(), norb, __builtin_popcountll(alpha[0]), __builtin_popcountll(beta[0]));
  for (size_t i = 0; i < coeffs.size(); ++i)
    std::fprintf(fh, "%30.16e %s \n", coeffs[i], to_canonical_string(alpha[i], beta[i], norb).c_str());
  std::fclose(fh);
}

}  // namespace qdk_b200::io
