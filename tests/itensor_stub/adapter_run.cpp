// TEST INFRASTRUCTURE.  Executes include/chain_b200_itensor.h for real: builds an MPO and a product MPS as (stub) ITensor objects from
// a text file written by tests/test_adapter.py, calls `dmrg(H, psis, psi0, sweeps, {"Quiet",true,"Weight=",1000.0})` -- the reference's
// call of src/dmrg.h:544, redirected by the adapter's macro -- twice (ground state, then first excited state with the ground state as
// a previous state), and prints energies / link dimensions / norms.  Linked against tests/hostemu/libcb200_hostemu.so in the GPU-less
// container and against chain_b200/libchain_b200.so on the GPU box.  The stub has no QN storage, so this covers the dense models.
#include <cstdio>
#include <fstream>
#include <iostream>
#include "chain_b200_itensor.h"

using namespace itensor;

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  std::ifstream in(argv[1]);
  int N, d, nsweep;
  in >> N >> d >> nsweep;
  std::vector<int> maxdim(nsweep);
  std::vector<double> cutoff(nsweep);
  for (int i = 0; i < nsweep; ++i) in >> maxdim[i] >> cutoff[i];
  std::vector<int> k(N + 1);
  for (int j = 0; j <= N; ++j) in >> k[j];
  std::vector<Index> s(N + 1), a(N + 1), l(N + 1);
  for (int j = 1; j <= N; ++j) s[j] = Index(d, "Site,n=" + std::to_string(j));
  for (int j = 1; j < N; ++j) { a[j] = Index(k[j], "Link,H,l=" + std::to_string(j)); l[j] = Index(1, "Link,l=" + std::to_string(j)); }
  MPO H(N);
  for (int j = 1; j <= N; ++j) {                      // file order: W[a, s_out, s_in, a'] column-major (the C ABI's MPO layout)
    std::vector<Index> is;
    if (j > 1) is.push_back(a[j - 1]);
    is.push_back(prime(s[j]));
    is.push_back(s[j]);
    if (j < N) is.push_back(a[j]);
    ITensor W(is);
    for (int ar = 1; ar <= k[j]; ++ar) for (int si = 1; si <= d; ++si) for (int so = 1; so <= d; ++so) for (int al = 1; al <= k[j - 1]; ++al) {
      double v;
      in >> v;
      if (v == 0.0) continue;
      if (j == 1) W.set(prime(s[j])(so), s[j](si), a[j](ar), v);
      else if (j == N) W.set(a[j - 1](al), prime(s[j])(so), s[j](si), v);
      else W.set(a[j - 1](al), prime(s[j])(so), s[j](si), a[j](ar), v);
    }
    H.ref(j) = W;
  }
  MPS psi0(N);
  for (int j = 1; j <= N; ++j) {
    std::vector<Index> is;
    if (j > 1) is.push_back(l[j - 1]);
    is.push_back(s[j]);
    if (j < N) is.push_back(l[j]);
    ITensor A(is);
    for (int si = 1; si <= d; ++si) {
      double v;
      in >> v;
      if (j == 1) A.set(s[j](si), l[j](1), v);
      else if (j == N) A.set(l[j - 1](1), s[j](si), v);
      else A.set(l[j - 1](1), s[j](si), l[j](1), v);
    }
    psi0.ref(j) = A;
  }
  Sweeps sweeps(nsweep);
  for (int i = 1; i <= nsweep; ++i) { sweeps.setmaxdim(i, maxdim[i - 1]); sweeps.setcutoff(i, cutoff[i - 1]); sweeps.setniter(i, 2); sweeps.setnoise(i, 0.0); }
  Real en_;
  MPS psi_;
  std::vector<MPS> done;
  try {
    for (int state = 0; state < 2; ++state) {
      std::tie(en_, psi_) = dmrg(H, done, psi0, sweeps, {"Quiet", true, "Weight=", 1000.0});      // src/dmrg.h:544, verbatim shape
      std::printf("energy %.15g\n", en_);
      std::printf("links");
      for (int j = 1; j < N; ++j) std::printf(" %d", dim(linkIndex(psi_, j)));
      std::printf("\nnorm1 %.15g\n", norm(psi_(1)));
      done.push_back(psi_);
    }
  } catch (std::exception const& e) {
    std::printf("exception %s\n", e.what());
    return 1;
  }
  return 0;
}
