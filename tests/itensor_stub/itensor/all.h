// TEST INFRASTRUCTURE -- a miniature stand-in for "itensor/all.h" (ITensor C++ v3), written from scratch.
//
// ITensor is an un-vendored, un-pinned dependency of the reference (/root/reference/README.md:3-10) and is absent
// from this image, so nothing of the reference compiles as shipped.  This header declares -- and, for the dense
// pieces the reference's MODEL files use, implements -- just the part of the ITensor v3 API that
//   (1) /root/reference/src/{golden_site,haagerup_site,haagerup_q_site,f_data}.cpp touch, so that those files can be
//       compiled UNMODIFIED from where they lie (oracle/Makefile -> oracle/_ref/ref_models) and their operator tables,
//       F-symbols and `mpo += ...` term lists dumped: the only reference-held data the oracle can be pinned to;
//   (2) include/chain_b200_itensor.h (the drop-in adapter for `itensor::dmrg`, src/dmrg.h:544,563,596) touches, so that
//       `g++ -fsyntax-only` type-checks the adapter in CI (tests/test_adapter.py).
// It is NOT ITensor: no QN block storage, no SVD, no DMRG.  AutoMPO only records its terms; toMPO() hands them on.
#pragma once
#include <algorithm>
#include <array>
#include <cctype>
#include <cmath>
#include <complex>
#include <cstdio>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

namespace itensor {

using Real = double;
using Cplx = std::complex<double>;
using std::string;
using std::vector;
using std::sqrt;
using std::sin;
using std::cos;
using std::conj;

constexpr Real Pi = 3.14159265358979323846264338327950288;

inline Cplx operator"" _i(unsigned long long x) { return Cplx(0.0, (double)x); }
inline Cplx operator"" _i(long double x) { return Cplx(0.0, (double)x); }
inline Cplx operator*(int a, Cplx const& b) { return Cplx((double)a, 0.0) * b; }
inline Cplx operator*(Cplx const& a, int b) { return a * Cplx((double)b, 0.0); }
inline Cplx operator+(int a, Cplx const& b) { return Cplx((double)a, 0.0) + b; }
inline Cplx operator+(Cplx const& a, int b) { return a + Cplx((double)b, 0.0); }
inline Cplx operator-(int a, Cplx const& b) { return Cplx((double)a, 0.0) - b; }
inline Cplx operator-(Cplx const& a, int b) { return a - Cplx((double)b, 0.0); }

class ITError : public std::runtime_error {
 public:
  explicit ITError(string const& m = "") : std::runtime_error(m) {}
};
inline void Error(string const& m) { throw ITError(m); }

template <typename... A>
void print(A&&... a) { (std::cout << ... << a); }
template <typename... A>
void println(A&&... a) { (std::cout << ... << a); std::cout << std::endl; }
template <typename... A>
void printfln(const char*, A&&...) {}

// ---------------------------------------------------------------------------------------------------------- Args
class Args {
  std::map<string, double> num_;
  std::map<string, string> str_;
  static string key(string k) { while (!k.empty() && k.back() == '=') k.pop_back(); return k; }
  void add(const char* k, bool v) { num_[key(k)] = v ? 1.0 : 0.0; }
  void add(const char* k, int v) { num_[key(k)] = v; }
  void add(const char* k, long v) { num_[key(k)] = (double)v; }
  void add(const char* k, double v) { num_[key(k)] = v; }
  void add(const char* k, const char* v) { str_[key(k)] = v; }
  void add(const char* k, string const& v) { str_[key(k)] = v; }
  void fill() {}
  template <typename T, typename... R>
  void fill(const char* k, T const& v, R const&... r) { add(k, v); fill(r...); }
  template <typename... R>
  void fill(Args const& o, R const&... r) { for (auto& p : o.num_) num_[p.first] = p.second; for (auto& p : o.str_) str_[p.first] = p.second; fill(r...); }

 public:
  Args() {}
  template <typename T, typename... R>
  Args(const char* k, T const& v, R const&... r) { fill(k, v, r...); }
  template <typename... R>
  Args(Args const& o, const char* k, R const&... r) : num_(o.num_), str_(o.str_) { fill(k, r...); }
  static Args& global() { static Args g; return g; }
  bool defined(string const& k) const { return num_.count(key(k)) || str_.count(key(k)); }
  bool getBool(string const& k, bool d) const { auto it = num_.find(key(k)); return it == num_.end() ? d : it->second != 0.0; }
  bool getBool(string const& k) const { return getBool(k, false); }
  int getInt(string const& k, int d) const { auto it = num_.find(key(k)); return it == num_.end() ? d : (int)it->second; }
  Real getReal(string const& k, Real d) const { auto it = num_.find(key(k)); return it == num_.end() ? d : it->second; }
  string getString(string const& k, string const& d) const { auto it = str_.find(key(k)); return it == str_.end() ? d : it->second; }
  template <typename T>
  void add(string const& k, T const& v) { add(k.c_str(), v); }
};

// ------------------------------------------------------------------------------------------------------- QN, Index
enum Arrow { In = -1, Out = 1 };
inline Arrow operator-(Arrow a) { return a == In ? Out : In; }

struct QNum {
  string name;
  int val = 0, mod = 1;
  QNum() {}
  QNum(const char* n, int v, int m = 1) : name(n), val(v), mod(m) {}
  QNum(int v) : name(""), val(v), mod(1) {}
};
class QN {
  QNum q_;

 public:
  QN() {}
  QN(QNum q) : q_(std::move(q)) {}
  int val(string const& = "") const { return q_.val; }
  int val(const char*) const { return q_.val; }
  int mod(string const& = "") const { return q_.mod; }
  string const& name() const { return q_.name; }
};

class TagSet {
  string t_;

 public:
  TagSet() {}
  TagSet(const char* t) : t_(t) {}
  TagSet(string t) : t_(std::move(t)) {}
  string const& str() const { return t_; }
};

struct IndexVal;
class Index {
  long id_ = 0;
  int dim_ = 0, plev_ = 0;
  Arrow dir_ = Out;
  TagSet tags_;
  vector<std::pair<QN, int>> sect_;
  static long next_id() { static long c = 0; return ++c; }
  void take() {}
  template <typename... R>
  void take(QN const& q, int n, R const&... r) { sect_.emplace_back(q, n); dim_ += n; take(r...); }
  template <typename... R>
  void take(Arrow a, R const&... r) { dir_ = a; take(r...); }
  template <typename... R>
  void take(TagSet const& t, R const&... r) { tags_ = t; take(r...); }

 public:
  using qnstorage = vector<std::pair<QN, long>>;
  Index() {}
  Index(long dim, TagSet const& t = TagSet()) : id_(next_id()), dim_((int)dim), tags_(t) {}
  Index(int dim, TagSet const& t = TagSet()) : id_(next_id()), dim_(dim), tags_(t) {}
  Index(qnstorage&& q, Arrow a, TagSet const& t = TagSet()) : id_(next_id()), dir_(a), tags_(t) { for (auto& p : q) { sect_.emplace_back(p.first, (int)p.second); dim_ += (int)p.second; } }
  template <typename... R>
  Index(QN const& q, int n, R const&... r) : id_(next_id()) { take(q, n, r...); }
  explicit operator bool() const { return id_ != 0; }
  long id() const { return id_; }
  int dim() const { return dim_; }
  int primeLevel() const { return plev_; }
  Arrow dir() const { return dir_; }
  TagSet const& tags() const { return tags_; }
  int nblock() const { return (int)sect_.size(); }
  QN const& qn(int b) const { return sect_.at(b - 1).first; }          // 1-based, as in ITensor
  int blocksize(int b) const { return sect_.at(b - 1).second; }
  bool hasQNs() const { return !sect_.empty(); }
  Index& prime(int inc = 1) { plev_ += inc; return *this; }
  Index& setPrime(int p) { plev_ = p; return *this; }
  Index& dag() { dir_ = -dir_; return *this; }
  IndexVal operator()(int v) const;
  friend bool operator==(Index const& a, Index const& b) { return a.id_ == b.id_ && a.plev_ == b.plev_; }
  friend bool operator!=(Index const& a, Index const& b) { return !(a == b); }
};
inline int dim(Index const& i) { return i.dim(); }
inline Arrow dir(Index const& i) { return i.dir(); }
inline bool hasQNs(Index const& i) { return i.hasQNs(); }
inline int nblock(Index const& i) { return i.nblock(); }
inline QN const& qn(Index const& i, int b) { return i.qn(b); }
inline int blocksize(Index const& i, int b) { return i.blocksize(b); }
inline Index prime(Index i, int inc = 1) { i.prime(inc); return i; }
inline Index dag(Index i) { i.dag(); return i; }
inline Index removeQNs(Index const& i) { Index r(i.dim(), i.tags()); return r; }

struct IndexVal {
  Index index;
  int val = 0;
  IndexVal() {}
  IndexVal(Index i, int v) : index(std::move(i)), val(v) {}
};
inline IndexVal Index::operator()(int v) const {
  if (v < 1 || v > dim_) throw ITError("IndexVal out of range");
  return IndexVal(*this, v);
}

// ------------------------------------------------------------------------------------------------------- ITensor
// Dense, complex storage, column-major over inds() (first index fastest) -- the layout of ITensor's Dense<T> storage.
class ITensor {
  vector<Index> is_;
  vector<Cplx> d_;
  bool real_ = true;
  void gather(vector<Index>&) {}
  template <typename... R>
  void gather(vector<Index>& v, Index const& i, R const&... r) { v.push_back(i); gather(v, r...); }
  size_t offset(vector<IndexVal> const& iv) const {
    if (iv.size() != is_.size()) throw ITError("wrong number of IndexVals");
    size_t off = 0, stride = 1;
    for (auto const& I : is_) {
      bool found = false;
      for (auto const& v : iv)
        if (v.index == I) { off += stride * (size_t)(v.val - 1); found = true; break; }
      if (!found) throw ITError("IndexVal does not match any index of the ITensor");
      stride *= (size_t)I.dim();
    }
    return off;
  }
  template <typename V>
  void set_impl(vector<IndexVal>& iv, V const& v) { Cplx c(v); if (c.imag() != 0.0) real_ = false; d_[offset(iv)] = c; }
  template <typename... R>
  void set_impl(vector<IndexVal>& iv, IndexVal const& a, R const&... r) { iv.push_back(a); set_impl(iv, r...); }
  void elt_impl(vector<IndexVal>&) const {}
  template <typename... R>
  void elt_impl(vector<IndexVal>& iv, IndexVal const& a, R const&... r) const { iv.push_back(a); elt_impl(iv, r...); }

 public:
  ITensor() {}
  explicit ITensor(vector<Index> is) : is_(std::move(is)) { size_t n = 1; for (auto& i : is_) n *= (size_t)i.dim(); d_.assign(n, Cplx(0)); }
  template <typename... R>
  explicit ITensor(Index const& i, R const&... r) { gather(is_, i, r...); size_t n = 1; for (auto& j : is_) n *= (size_t)j.dim(); d_.assign(n, Cplx(0)); }
  explicit operator bool() const { return !is_.empty() || !d_.empty(); }
  vector<Index> const& inds() const { return is_; }
  int order() const { return (int)is_.size(); }
  Index const& index(int j) const { return is_.at(j - 1); }
  bool isComplex() const { return !real_; }
  vector<Cplx>& data() { return d_; }
  vector<Cplx> const& data() const { return d_; }
  void markComplex() { real_ = false; }
  template <typename... R>
  void set(IndexVal const& a, R const&... r) { vector<IndexVal> iv; set_impl(iv, a, r...); }
  template <typename... R>
  Cplx eltC(IndexVal const& a, R const&... r) const { vector<IndexVal> iv; elt_impl(iv, a, r...); return d_[offset(iv)]; }
  template <typename... R>
  Real elt(IndexVal const& a, R const&... r) const { return eltC(a, r...).real(); }
  template <typename F>
  void visit(F&& f) const { for (auto const& v : d_) { if (real_) f(v.real()); else f(v); } }
  template <typename F>
  ITensor& generate(F&& f) { for (auto& v : d_) { v = Cplx(f()); if (v.imag() != 0.0) real_ = false; } return *this; }
  ITensor& operator*=(ITensor const& o) { *this = (*this) * o; return *this; }
  ITensor& operator*=(Real x) { for (auto& v : d_) v *= x; return *this; }
  ITensor& operator/=(Real x) { for (auto& v : d_) v /= x; return *this; }
  // contraction over the indices the two tensors share
  friend ITensor operator*(ITensor const& A, ITensor const& B) {
    vector<int> a_free, b_free;
    vector<std::pair<int, int>> shared;
    for (int i = 0; i < (int)A.is_.size(); ++i) {
      int m = -1;
      for (int j = 0; j < (int)B.is_.size(); ++j) if (A.is_[i] == B.is_[j]) m = j;
      if (m < 0) a_free.push_back(i); else shared.emplace_back(i, m);
    }
    for (int j = 0; j < (int)B.is_.size(); ++j) {
      bool s = false;
      for (auto& p : shared) if (p.second == j) s = true;
      if (!s) b_free.push_back(j);
    }
    vector<Index> ci;
    for (int i : a_free) ci.push_back(A.is_[i]);
    for (int j : b_free) ci.push_back(B.is_[j]);
    ITensor C{ci};
    if (ci.empty()) C.d_.assign(1, Cplx(0));
    C.real_ = A.real_ && B.real_;
    auto strides = [](vector<Index> const& is) { vector<size_t> s(is.size()); size_t x = 1; for (size_t k = 0; k < is.size(); ++k) { s[k] = x; x *= (size_t)is[k].dim(); } return s; };
    auto sa = strides(A.is_), sb = strides(B.is_), sc = strides(ci);
    size_t nsh = 1;
    for (auto& p : shared) nsh *= (size_t)A.is_[p.first].dim();
    for (size_t c = 0; c < C.d_.size(); ++c) {
      size_t oa = 0, ob = 0, rem = c;
      for (size_t k = 0; k < ci.size(); ++k) {
        size_t v = rem % (size_t)ci[k].dim(); rem /= (size_t)ci[k].dim();
        if (k < a_free.size()) oa += v * sa[a_free[k]]; else ob += v * sb[b_free[k - a_free.size()]];
      }
      Cplx acc(0);
      for (size_t s = 0; s < nsh; ++s) {
        size_t xa = oa, xb = ob, r2 = s;
        for (auto& p : shared) { size_t v = r2 % (size_t)A.is_[p.first].dim(); r2 /= (size_t)A.is_[p.first].dim(); xa += v * sa[p.first]; xb += v * sb[p.second]; }
        acc += A.d_[xa] * B.d_[xb];
      }
      C.d_[c] = acc;
    }
    return C;
  }
};
inline ITensor delta(Index const& a, Index const& b) { ITensor T(a, b); for (int i = 1; i <= std::min(a.dim(), b.dim()); ++i) T.set(a(i), b(i), 1.0); return T; }
inline ITensor toDense(ITensor T) { return T; }
inline ITensor setElt(IndexVal const& iv) { ITensor T(iv.index); T.set(iv, 1.0); return T; }
inline QN div(ITensor const&) { return QN(); }     // stub: dense tensors only
inline Index findIndex(ITensor const& T, TagSet const& ts) {
  // "Tag,plev" syntax of ITensor v3: the tag list may end in a prime level
  string t = ts.str();
  int plev = -1;
  auto c = t.rfind(',');
  if (c != string::npos && c + 1 < t.size() && std::isdigit((unsigned char)t[c + 1])) { plev = std::atoi(t.c_str() + c + 1); t = t.substr(0, c); }
  for (auto& i : T.inds()) if (i.tags().str().find(t) != string::npos && (plev < 0 || i.primeLevel() == plev)) return i;
  return Index();
}
template <typename... R>
ITensor permute(ITensor const& T, Index const& i, R const&... r) {
  vector<Index> o{i, r...};
  ITensor P{o};
  if (T.isComplex()) P.markComplex();
  vector<int> map(o.size());
  for (size_t k = 0; k < o.size(); ++k) { map[k] = -1; for (size_t m = 0; m < T.inds().size(); ++m) if (T.inds()[m] == o[k]) map[k] = (int)m; if (map[k] < 0) throw ITError("permute: index not found"); }
  vector<size_t> st(T.inds().size());
  size_t x = 1;
  for (size_t m = 0; m < st.size(); ++m) { st[m] = x; x *= (size_t)T.inds()[m].dim(); }
  for (size_t c = 0; c < P.data().size(); ++c) {
    size_t rem = c, off = 0;
    for (size_t k = 0; k < o.size(); ++k) { size_t v = rem % (size_t)o[k].dim(); rem /= (size_t)o[k].dim(); off += v * st[map[k]]; }
    P.data()[c] = T.data()[off];
  }
  return P;
}
inline ITensor removeQNs(ITensor T) { return T; }
inline ITensor dag(ITensor T) { for (auto& v : T.data()) v = std::conj(v); return T; }
inline bool isComplex(ITensor const& T) { return T.isComplex(); }
inline int order(ITensor const& T) { return T.order(); }
inline vector<Index> const& inds(ITensor const& T) { return T.inds(); }
inline Real norm(ITensor const& T) { Real s = 0; for (auto& v : T.data()) s += std::norm(v); return std::sqrt(s); }
inline bool hasIndex(ITensor const& T, Index const& I) { for (auto& i : T.inds()) if (i == I) return true; return false; }
inline Index commonIndex(ITensor const& A, ITensor const& B) { for (auto& i : A.inds()) if (hasIndex(B, i)) return i; return Index(); }

// ------------------------------------------------------------------------------------------------- site sets, AutoMPO
class SiteSet {
 protected:
  struct Model {
    virtual ~Model() {}
    virtual int N() const = 0;
    virtual Index index(int j) const = 0;
    virtual ITensor op(string const& name, int j, Args const& a) const = 0;
    virtual IndexVal state(int j, string const& s) const = 0;
  };
  std::shared_ptr<Model> m_;

 public:
  SiteSet() {}
  int length() const { return m_ ? m_->N() : 0; }
  int N() const { return length(); }
  Index operator()(int j) const { return m_->index(j); }
  IndexVal operator()(int j, string const& s) const { return m_->state(j, s); }
  ITensor op(string const& name, int j, Args const& a = Args::global()) const { return m_->op(name, j, a); }
  explicit operator bool() const { return (bool)m_; }
};
inline int length(SiteSet const& s) { return s.length(); }

template <class SiteType>
class BasicSiteSet : public SiteSet {
  struct M : Model {
    mutable vector<SiteType> s;
    int N() const override { return (int)s.size(); }
    Index index(int j) const override { return s.at(j - 1).index(); }
    ITensor op(string const& name, int j, Args const& a) const override { return s.at(j - 1).op(name, a); }
    IndexVal state(int j, string const& st) const override { return s.at(j - 1).state(st); }
  };

 public:
  BasicSiteSet() {}
  BasicSiteSet(int N, Args const& args = Args::global()) {
    auto m = std::make_shared<M>();
    for (int j = 1; j <= N; ++j) m->s.emplace_back(Args(args, "SiteNumber=", j));
    m_ = m;
  }
};

struct SiteTerm { string op; int i = 0; };
struct HTerm { Cplx coef = 1.0; vector<SiteTerm> ops; };

class AutoMPO {
  SiteSet sites_;
  vector<HTerm> terms_;

 public:
  class Accumulator {
    AutoMPO* p_;
    HTerm t_;
    string pending_;
    bool have_op_ = false;

   public:
    Accumulator(AutoMPO* p, Cplx c) : p_(p) { t_.coef = c; }
    Accumulator(AutoMPO* p, const char* op) : p_(p), pending_(op), have_op_(true) {}
    Accumulator(Accumulator&& o) : p_(o.p_), t_(std::move(o.t_)), pending_(std::move(o.pending_)), have_op_(o.have_op_) { o.p_ = nullptr; }
    ~Accumulator() { if (p_) p_->terms_.push_back(t_); }
    Accumulator& operator,(const char* op) { pending_ = op; have_op_ = true; return *this; }
    Accumulator& operator,(string const& op) { pending_ = op; have_op_ = true; return *this; }
    Accumulator& operator,(int v) {
      if (have_op_) { t_.ops.push_back(SiteTerm{pending_, v}); have_op_ = false; } else t_.coef *= (double)v;
      return *this;
    }
    Accumulator& operator,(Real v) { t_.coef *= v; return *this; }
    Accumulator& operator,(Cplx v) { t_.coef *= v; return *this; }
  };
  AutoMPO() {}
  explicit AutoMPO(SiteSet const& s) : sites_(s) {}
  SiteSet const& sites() const { return sites_; }
  vector<HTerm> const& terms() const { return terms_; }
  Accumulator operator+=(Real c) { return Accumulator(this, Cplx(c)); }
  Accumulator operator+=(int c) { return Accumulator(this, Cplx((double)c)); }
  Accumulator operator+=(Cplx c) { return Accumulator(this, c); }
  Accumulator operator+=(const char* op) { return Accumulator(this, op); }
};

// ------------------------------------------------------------------------------------------------------- MPS / MPO
// Containers only (1-based site access, ITensor-style free functions); enough for the adapter's marshalling code to type-check.
class MPS {
 protected:
  vector<ITensor> A_;

 public:
  MPS() {}
  explicit MPS(int N) : A_(N) {}
  explicit MPS(SiteSet const& s) : A_(s.length()) {}
  int length() const { return (int)A_.size(); }
  explicit operator bool() const { return !A_.empty(); }
  ITensor const& operator()(int j) const { return A_.at(j - 1); }
  ITensor& ref(int j) { return A_.at(j - 1); }
  void set(int j, ITensor T) { A_.at(j - 1) = std::move(T); }
  MPS& position(int, Args const& = Args::global()) { return *this; }
  MPS& orthogonalize(Args const& = Args::global()) { return *this; }
  void leftLim(int) {}
  void rightLim(int) {}
  Real normalize() { return 1.0; }
};
class MPO : public MPS {
  vector<HTerm> terms_;      // stub only: what AutoMPO recorded (a real MPO holds the compressed site tensors)
  SiteSet sites_;

 public:
  MPO() {}
  explicit MPO(int N) : MPS(N) {}
  explicit MPO(SiteSet const& s) : MPS(s), sites_(s) {}
  vector<HTerm> const& stub_terms() const { return terms_; }
  SiteSet const& stub_sites() const { return sites_; }
  friend MPO toMPO(AutoMPO const& a, Args const&);
};
inline MPO toMPO(AutoMPO const& a, Args const& = Args::global()) { MPO H(a.sites()); H.terms_ = a.terms(); return H; }
inline int length(MPS const& p) { return p.length(); }
inline Index siteIndex(MPS const& psi, int j) {
  for (auto& i : psi(j).inds()) if (i.tags().str().find("Site") != string::npos && i.primeLevel() == 0) return i;
  return Index();
}
inline Index linkIndex(MPS const& psi, int j) { return (j < 1 || j >= psi.length()) ? Index() : commonIndex(psi(j), psi(j + 1)); }
inline Index leftLinkIndex(MPS const& psi, int j) { return linkIndex(psi, j - 1); }
inline Index rightLinkIndex(MPS const& psi, int j) { return linkIndex(psi, j); }
inline bool hasQNs(MPS const& psi) { return psi.length() > 0 && psi(1).inds().size() > 0 && psi(1).inds()[0].hasQNs(); }
inline int orthoCenter(MPS const&) { return 1; }
inline bool isComplex(MPS const& psi) { for (int j = 1; j <= psi.length(); ++j) if (psi(j).isComplex()) return true; return false; }

class Sweeps {
  int n_ = 0;
  vector<int> maxdim_, mindim_, niter_;
  vector<Real> cutoff_, noise_;

 public:
  Sweeps() {}
  explicit Sweeps(int n) : n_(n), maxdim_(n, 1), mindim_(n, 1), niter_(n, 2), cutoff_(n, 0.0), noise_(n, 0.0) {}
  int nsweep() const { return n_; }
  int maxdim(int sw) const { return maxdim_.at(sw - 1); }
  int mindim(int sw) const { return mindim_.at(sw - 1); }
  int niter(int sw) const { return niter_.at(sw - 1); }
  Real cutoff(int sw) const { return cutoff_.at(sw - 1); }
  Real noise(int sw) const { return noise_.at(sw - 1); }
  void setmaxdim(int sw, int v) { maxdim_.at(sw - 1) = v; }
  void setmindim(int sw, int v) { mindim_.at(sw - 1) = v; }
  void setniter(int sw, int v) { niter_.at(sw - 1) = v; }
  void setcutoff(int sw, Real v) { cutoff_.at(sw - 1) = v; }
  void setnoise(int sw, Real v) { noise_.at(sw - 1) = v; }
};

}  // namespace itensor
