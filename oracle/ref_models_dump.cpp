// TEST INFRASTRUCTURE.  Dumps, as JSON, what the REFERENCE'S OWN model code computes when it is compiled unmodified from
// /root/reference/src/{golden_site,haagerup_site,haagerup_q_site,f_data}.cpp against tests/itensor_stub (oracle/Makefile):
//   * every site operator the three site classes define (op(name) -> d x d matrix, M[i][j] = Op.elt(s(i+1), s'(j+1))),
//   * the `mpo += ...` term lists Hamiltonian() emits for a set of (boundary condition, L, U, couplings),
//   * state() labels, the Z3 sector charges of the haagerupq site index, F-symbols and RhoDefectCell of both categories.
// Nothing here is product code; the output is committed as tests/golden/reference_models.json (the reference is not on the GPU box).
#include <cstdio>
#include <string>
#include <vector>
#include "golden_site.h"
#include "haagerup_q_site.h"
#include "haagerup_site.h"

using namespace itensor;

static void put_c(Cplx v) { std::printf("[%.17g,%.17g]", v.real(), v.imag()); }

template <class Sites>
static void dump_ops(Sites& sites, const std::vector<std::string>& names) {
  std::printf("\"ops\":{");
  bool first = true;
  for (auto& nm : names) {
    ITensor O = sites.op(nm, 1);
    Index s = sites(1), sP = prime(s);
    int d = s.dim();
    std::printf("%s\"%s\":[", first ? "" : ",", nm.c_str());
    first = false;
    for (int i = 1; i <= d; ++i) {
      std::printf("%s[", i > 1 ? "," : "");
      for (int j = 1; j <= d; ++j) { if (j > 1) std::printf(","); put_c(O.eltC(s(i), sP(j))); }
      std::printf("]");
    }
    std::printf("]");
  }
  std::printf("}");
}

struct HCase { const char* bc; int n; double U; std::vector<double> cp; };

template <class Sites>
static void dump_terms(const std::vector<HCase>& cases) {
  std::printf("\"hamiltonians\":[");
  bool first = true;
  for (auto& c : cases) {
    Sites sites(c.n, {"ConserveQNs=", true});
    MPO H = sites.Hamiltonian(c.bc, c.n, c.U, c.cp);
    std::printf("%s{\"bc\":\"%s\",\"n\":%d,\"U\":%.17g,\"couplings\":[", first ? "" : ",", c.bc, c.n, c.U);
    first = false;
    for (size_t k = 0; k < c.cp.size(); ++k) std::printf("%s%.17g", k ? "," : "", c.cp[k]);
    std::printf("],\"terms\":[");
    bool f2 = true;
    for (auto& t : H.stub_terms()) {
      std::printf("%s{\"c\":", f2 ? "" : ",");
      f2 = false;
      put_c(t.coef);
      std::printf(",\"ops\":[");
      for (size_t k = 0; k < t.ops.size(); ++k) std::printf("%s[\"%s\",%d]", k ? "," : "", t.ops[k].op.c_str(), t.ops[k].i);
      std::printf("]}");
    }
    std::printf("]}");
  }
  std::printf("]");
}

template <class Sites>
static void dump_states(Sites& sites, const std::vector<std::string>& labels) {
  std::printf("\"states\":{");
  for (size_t k = 0; k < labels.size(); ++k) std::printf("%s\"%s\":%d", k ? "," : "", labels[k].c_str(), sites(1, labels[k]).val);
  std::printf("}");
}

static void dump_fdata(FData& f) {
  std::printf("\"nu\":%d,\"rk\":%d,\"qd\":[", f.nu_, f.rk_);
  for (int i = 1; i <= f.rk_; ++i) std::printf("%s%.17g", i > 1 ? "," : "", f.QD(i));
  std::printf("],\"fsymbols\":[");
  bool first = true;
  int r = f.rk_;
  for (int i = 1; i <= r; ++i) for (int j = 1; j <= r; ++j) for (int k = 1; k <= r; ++k)
    for (int l = 1; l <= r; ++l) for (int m = 1; m <= r; ++m) for (int n = 1; n <= r; ++n) {
      double v = f.FSymbol(i, j, k, l, m, n);
      if (v != 0) { std::printf("%s[%d,%d,%d,%d,%d,%d,%.17g]", first ? "" : ",", i, j, k, l, m, n, v); first = false; }
    }
  std::printf("],\"rho_defect_cell\":[");
  Index s1(r, "Site,a"), s2(r, "Site,b");
  ITensor G = f.RhoDefectCell(s1, s2);
  Index s1P = prime(s1), s2P = prime(s2);
  first = true;
  for (int i1 = 1; i1 <= r; ++i1) for (int j1 = 1; j1 <= r; ++j1) for (int i2 = 1; i2 <= r; ++i2) for (int j2 = 1; j2 <= r; ++j2) {
    double v = G.elt(s1(i1), s1P(j1), s2(i2), s2P(j2));
    if (v != 0) { std::printf("%s[%d,%d,%d,%d,%.17g]", first ? "" : ",", i1, j1, i2, j2, v); first = false; }
  }
  std::printf("]");
}

int main() {
  std::vector<HCase> g_cases, h_cases;
  for (const char* bc : {"p", "o", "s", "d", "sp"}) {
    g_cases.push_back({bc, 6, 2.0, {-1.0}});
    g_cases.push_back({bc, 8, 2.0, {1.0}});
    h_cases.push_back({bc, 4, 2.0, {-0.5, 0.25, 1.5}});       // small enough to compare dense matrices
    h_cases.push_back({bc, 6, 2.0, {-0.5, 0.25, 1.5}});
  }
  for (auto cp : std::vector<std::vector<double>>{{-1, 0, 0}, {0, -1, 0}, {0, 0, -1}}) { h_cases.push_back({"p", 6, 2.0, cp}); h_cases.push_back({"s", 6, 2.0, cp}); }
  g_cases.push_back({"o", 6, 2.0, {1.0}});
  g_cases.push_back({"p", 8, 2.0, {-1.0}});
  g_cases.push_back({"p", 24, 2.0, {-1.0}});                  // BASELINE configs[2]
  h_cases.push_back({"p", 8, 2.0, {-1, 0, 0}});               // BASELINE configs[4]
  h_cases.push_back({"o", 12, 2.0, {0, -1, 0}});              // BASELINE configs[3]

  std::printf("{\"source\":\"/root/reference/src/{golden_site,haagerup_site,haagerup_q_site,f_data}.cpp compiled unmodified against tests/itensor_stub\",\n");
  {
    Golden sites(6);
    std::printf("\"golden\":{\"d\":%d,", sites(1).dim());
    dump_ops(sites, {"id", "n1", "nt", "FF"});
    std::printf(",");
    dump_states(sites, {"0", "1", "2"});
    std::printf(",");
    dump_terms<Golden>(g_cases);
    std::printf("},\n");
  }
  {
    Haagerup sites(6);
    std::vector<std::string> names = {"id", "n1", "na", "nb", "nr", "nar", "nbr"};
    // every FF-type operator name the Hamiltonian uses is discovered from the term lists below; dump those that op() knows
    std::set<std::string> seen(names.begin(), names.end());
    for (auto& c : h_cases) {
      Haagerup s2(c.n);
      MPO H = s2.Hamiltonian(c.bc, c.n, c.U, c.cp);
      for (auto& t : H.stub_terms()) for (auto& o : t.ops) if (seen.insert(o.op).second) names.push_back(o.op);
    }
    std::printf("\"haagerup\":{\"d\":%d,", sites(1).dim());
    dump_ops(sites, names);
    std::printf(",");
    dump_states(sites, {"0", "1", "2", "r"});
    std::printf(",");
    dump_terms<Haagerup>(h_cases);
    std::printf("},\n");
  }
  {
    HaagerupQ sites(6, {"ConserveQNs=", true});
    std::vector<std::string> names = {"id", "m1", "m2", "m3", "m4", "m5", "m6", "q1", "q2", "q3"};
    for (const char* f : {"qr", "qar", "qbr"}) for (int i = 1; i <= 3; ++i) for (int j = 1; j <= 3; ++j) names.push_back(std::string(f) + std::to_string(i) + std::to_string(j));
    std::set<std::string> seen(names.begin(), names.end());
    for (auto& c : h_cases) {
      HaagerupQ s2(c.n, {"ConserveQNs=", true});
      MPO H = s2.Hamiltonian(c.bc, c.n, c.U, c.cp);
      for (auto& t : H.stub_terms()) for (auto& o : t.ops) if (seen.insert(o.op).second) names.push_back(o.op);
    }
    Index s = sites(1);
    std::printf("\"haagerupq\":{\"d\":%d,\"sector_charges\":[", s.dim());
    for (int b = 1; b <= s.nblock(); ++b) for (int k = 0; k < s.blocksize(b); ++k) std::printf("%s%d", (b > 1 || k) ? "," : "", s.qn(b).val());
    std::printf("],\"qn_mod\":%d,\"arrow\":%d,", s.qn(1).mod(), (int)s.dir());
    dump_ops(sites, names);
    std::printf(",");
    dump_states(sites, {"0", "1", "2"});
    std::printf(",");
    dump_terms<HaagerupQ>(h_cases);
    std::printf("},\n");
  }
  {
    GoldenFData g;
    HaagerupFData h;
    std::printf("\"golden_fdata\":{");
    dump_fdata(g);
    std::printf("},\n\"haagerup_fdata\":{");
    dump_fdata(h);
    std::printf("}\n");
  }
  std::printf("}\n");
  return 0;
}
