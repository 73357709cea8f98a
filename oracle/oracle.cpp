// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle_math.hpp header).
//
// CPU restatement of QUGaR's implicit-domain classification and cut-cell quadrature drivers on
// top of oracle_quad.hpp, exported through a small C interface for ctypes (tests/, bench.py's
// cpu_baseline leg).  Each function cites the reference code it follows.
#include <string>
#include "oracle_quad.hpp"
#include "oracle_reparam.hpp"

#include <array>
#include <cstring>
#include <map>
#include <memory>
#include <thread>

namespace orc {

const real g_gauss_table[] = {
#include "gauss_table.inc"
};

// ImmersedCellStatus (unfitted_domain_binary_part.hpp:34-39)
enum CellStatus : unsigned char { CELL_CUT = 0, CELL_FULL = 1, CELL_EMPTY = 2, CELL_UNKNOWN = 3 };
// temporary status incl. full_with_unf_bdry (impl_unfitted_domain.cpp:163-180)
enum CellTmp { TMP_CUT, TMP_FULL, TMP_EMPTY, TMP_FULL_UNF };
// ImmersedFacetStatus (unfitted_domain.hpp:36-48), same order
enum FacetStatus : unsigned char {
  F_CUT = 0,
  F_FULL,
  F_EMPTY,
  F_CUT_UNF_BDRY,
  F_CUT_EXT_BDRY,
  F_CUT_UNF_BDRY_EXT_BDRY,
  F_FULL_UNF_BDRY,
  F_FULL_EXT_BDRY,
  F_UNF_BDRY,
  F_EXT_BDRY,
  F_UNF_BDRY_EXT_BDRY
};

inline bool fs_is_full(int s) { return s == F_FULL; }
inline bool fs_is_empty(int s) { return s == F_EMPTY || s == F_FULL_EXT_BDRY || s == F_EXT_BDRY; }
inline bool fs_is_cut(int s)
{
  return s == F_CUT || s == F_CUT_UNF_BDRY || s == F_CUT_EXT_BDRY || s == F_CUT_UNF_BDRY_EXT_BDRY;
}
inline bool fs_is_full_unf(int s) { return s == F_FULL_UNF_BDRY; }
inline bool fs_has_unf(int s)
{
  return s == F_CUT_UNF_BDRY || s == F_CUT_UNF_BDRY_EXT_BDRY || s == F_FULL_UNF_BDRY || s == F_UNF_BDRY
         || s == F_UNF_BDRY_EXT_BDRY;
}
inline bool fs_has_ext(int s)
{
  return s == F_CUT_EXT_BDRY || s == F_CUT_UNF_BDRY_EXT_BDRY || s == F_FULL_EXT_BDRY || s == F_EXT_BDRY
         || s == F_UNF_BDRY_EXT_BDRY;
}

struct Rule
{
  int pdim = 0;  // coordinates per point
  bool has_normals = false;
  std::vector<int> n_per;
  std::vector<real> pts, wts, nrm;
};

template<int N> struct Domain
{
  Func phi;
  std::vector<real> breaks[N];
  int64_t n[N];
  std::vector<unsigned char> status;                            // per cell
  std::map<int64_t, std::array<unsigned char, 2 * N>> facets;   // facets_status_
  std::vector<unsigned char> full_unf;                          // 1 if full_with_unf_bdry (kept record)
  Stats stats;

  int64_t num_cells() const
  {
    int64_t t = 1;
    for (int d = 0; d < N; ++d) t *= n[d];
    return t;
  }
  void tid(int64_t id, int64_t *t) const
  {
    for (int d = 0; d < N; ++d) {
      t[d] = id % n[d];
      id /= n[d];
    }
  }
  Box<N> cell_box(int64_t id) const
  {
    int64_t t[N];
    tid(id, t);
    Box<N> b;
    for (int d = 0; d < N; ++d) {
      b.lo[d] = breaks[d][t[d]];
      b.hi[d] = breaks[d][t[d] + 1];
    }
    return b;
  }
  bool is_exterior_facet(int64_t id, int lf) const
  {
    int64_t t[N];
    tid(id, t);
    const int cd = lf / 2, side = lf % 2;
    return t[cd] == (side == 0 ? 0 : n[cd] - 1);
  }

  // --- create_compute_sign_function (impl_unfitted_domain.cpp:136-159) ---
  int box_sign(const Box<N> &b)
  {
    Interval<N> xint[N];
    for (int d = 0; d < N; ++d) {
      const real mid = 0.5 * (b.lo[d] + b.hi[d]);
      xint[d] = Interval<N>::variable(mid, d);
      Interval<N>::delta(d) = 0.5 * (b.hi[d] - b.lo[d]);
    }
    const Interval<N> res = eval<N, Interval<N>>(phi, xint);
    if (res.uniformSign()) return res.alpha < 0.0 ? -1 : 1;
    return 0;
  }

  static real box_volume(const Box<N> &b)
  {
    real v = 1.0;
    for (int d = 0; d < N; ++d) v *= (b.hi[d] - b.lo[d]);
    return v;
  }

  // --- classify_cell_general + classify_cell_from_quad (impl_unfitted_domain.cpp:163-190) ---
  CellTmp classify_cell(const Box<N> &b)
  {
    QuadRule<N> q;
    quad_gen<N>(phi, b, -1, 0, 1, q, &stats);
    const real vol = q.sumWeights();
    if (std::fabs(vol - 0.0) <= NEAR_EPS) return TMP_EMPTY;
    if (std::fabs(vol - box_volume(b)) <= NEAR_EPS) return TMP_FULL_UNF;
    return TMP_CUT;
  }

  // --- classify_facet_general + classify_facet_from_quad + classify_facet_full_unfitted_boundary
  //     (impl_unfitted_domain.cpp:237-349) ---
  int classify_facet(const Box<N> &b, int lf, CellTmp cell_status)
  {
    const int cd = lf / 2, side = lf % 2;
    QuadRule<N> q;
    quad_gen<N>(phi, b, cd, side, 1, q, &stats);
    if (q.size() == 0) return cell_status == TMP_FULL_UNF ? F_FULL_UNF_BDRY : F_EMPTY;
    bool ext = false, unf = false, cut = false;
    for (size_t i = 0; i < q.size(); ++i) {
      const real *x = &q.x[N * i];
      if (std::fabs(eval<N, real>(phi, x)) <= NEAR_EPS) {
        real g[N];
        grad<N, real>(phi, x, g);
        real s = g[0] * g[0];
        for (int d = 1; d < N; ++d) s += g[d] * g[d];
        const real nc = g[cd] / std::sqrt(s);
        if (!(std::fabs(std::fabs(nc) - 1.0) <= NEAR_EPS)) continue;
        if (side == 0 ? (nc < -NEAR_EPS) : (NEAR_EPS < nc))
          unf = true;
        else
          ext = true;
      } else {
        cut = true;
      }
    }
    if (cell_status == TMP_FULL_UNF && !cut) return F_FULL_UNF_BDRY;
    static const unsigned char values[8] = { F_EMPTY, F_EXT_BDRY, F_UNF_BDRY, F_UNF_BDRY_EXT_BDRY, F_CUT,
      F_CUT_EXT_BDRY, F_CUT_UNF_BDRY, F_CUT_UNF_BDRY_EXT_BDRY };
    const int st = values[(ext ? 1 : 0) + (unf ? 2 : 0) + (cut ? 4 : 0)];
    const real facet_vol = q.sumWeights();
    const int n_pts = (int)q.size();
    // prod(remove_component(extent, const_dir))
    real dom_vol = 1.0;
    for (int d = 0; d < N; ++d)
      if (d != cd) dom_vol *= (b.hi[d] - b.lo[d]);
    const real max_val = std::fabs(facet_vol) < std::fabs(dom_vol) ? std::fabs(dom_vol) : std::fabs(facet_vol);
    const bool is_full = std::fabs(facet_vol - dom_vol) <= (NEAR_EPS + (NEAR_EPS * 1.0e3) * max_val);
    if (is_full) {
      if ((st == F_UNF_BDRY || st == F_EXT_BDRY) && n_pts == 1) return st == F_UNF_BDRY ? F_FULL_UNF_BDRY : F_FULL_EXT_BDRY;
      if (st == F_CUT) return F_FULL;
    }
    return st;
  }

  std::vector<int64_t> undetermined;  // single cells whose box sign is undetermined (in tree order)

  // --- classify_undetermined_sign_cell (impl_unfitted_domain.cpp:485-513) ---
  // out: where the facet record goes (per-thread list when classifying in parallel)
  void classify_undetermined(int64_t id, std::vector<std::pair<int64_t, std::array<unsigned char, 2 * N>>> *out = nullptr)
  {
    const Box<N> b = cell_box(id);
    CellTmp cs = classify_cell(b);
    if (cs == TMP_CUT || cs == TMP_FULL_UNF) {
      std::array<unsigned char, 2 * N> fs;
      bool all_full = true;
      for (int lf = 0; lf < 2 * N; ++lf) {
        fs[lf] = (unsigned char)classify_facet(b, lf, cs);
        if (fs[lf] != F_FULL) all_full = false;
      }
      if (cs == TMP_FULL_UNF && all_full)
        cs = TMP_FULL;
      else if (out)
        out->emplace_back(id, fs);
      else
        facets.emplace(id, fs);
    }
    status[id] = cs == TMP_CUT ? CELL_CUT : (cs == TMP_EMPTY ? CELL_EMPTY : CELL_FULL);
  }

  // --- create_decomposition (impl_unfitted_domain.cpp:441-469), TensorIndexRangeTP::split
  //     (tensor_index_tp.cpp:258-273), BoundBox::extend (bbox.cpp:136-147) ---
  void decompose(const int64_t *lo, const int64_t *hi)
  {
    Box<N> b;
    for (int d = 0; d < N; ++d) {
      b.lo[d] = breaks[d][lo[d]];
      b.hi[d] = breaks[d][hi[d]];
      b.lo[d] -= EPS * 1000;
      b.hi[d] += EPS * 1000;
    }
    const int sign = box_sign(b);
    int64_t cnt = 1;
    for (int d = 0; d < N; ++d) cnt *= hi[d] - lo[d];
    if (sign == 0) {
      if (cnt == 1) {
        int64_t id = 0, stride = 1;
        for (int d = 0; d < N; ++d) {
          id += lo[d] * stride;
          stride *= n[d];
        }
        undetermined.push_back(id);
      } else {
        int sd = 0;
        for (int d = 1; d < N; ++d)
          if (hi[d] - lo[d] > hi[sd] - lo[sd]) sd = d;
        const int64_t mid = (lo[sd] + hi[sd]) / 2;
        int64_t h0[N], l1[N];
        for (int d = 0; d < N; ++d) {
          h0[d] = hi[d];
          l1[d] = lo[d];
        }
        h0[sd] = mid;
        l1[sd] = mid;
        decompose(lo, h0);
        decompose(l1, hi);
      }
    } else {
      const unsigned char st = sign > 0 ? CELL_EMPTY : CELL_FULL;
      fill(lo, hi, st);
    }
  }
  void fill(const int64_t *lo, const int64_t *hi, unsigned char st)
  {
    if (N == 2) {
      for (int64_t j = lo[1]; j < hi[1]; ++j)
        for (int64_t i = lo[0]; i < hi[0]; ++i) status[i + n[0] * j] = st;
    } else {
      for (int64_t k = lo[N - 1]; k < hi[N - 1]; ++k)
        for (int64_t j = lo[1]; j < hi[1]; ++j)
          for (int64_t i = lo[0]; i < hi[0]; ++i) status[i + n[0] * (j + n[1] * k)] = st;
    }
  }

  // The reference classifies each undetermined cell as the recursion reaches it; cells are independent, so
  // the "all-core" baseline classifies them on nthreads threads after the (cheap) tree walk.
  void build(int nthreads = 1)
  {
    status.assign((size_t)num_cells(), CELL_UNKNOWN);
    int64_t lo[N], hi[N];
    for (int d = 0; d < N; ++d) {
      lo[d] = 0;
      hi[d] = n[d];
    }
    undetermined.clear();
    decompose(lo, hi);
    const int64_t nu = (int64_t)undetermined.size();
    if (nthreads < 1) nthreads = 1;
    if ((int64_t)nthreads > nu) nthreads = nu > 0 ? (int)nu : 1;
    if (nthreads == 1) {
      for (int64_t i = 0; i < nu; ++i) classify_undetermined(undetermined[(size_t)i]);
    } else {
      std::vector<std::vector<std::pair<int64_t, std::array<unsigned char, 2 * N>>>> recs((size_t)nthreads);
      std::vector<std::thread> th;
      for (int t = 0; t < nthreads; ++t)
        th.emplace_back([&, t]() {
          const int64_t b = nu * t / nthreads, e = nu * (t + 1) / nthreads;
          for (int64_t i = b; i < e; ++i) classify_undetermined(undetermined[(size_t)i], &recs[(size_t)t]);
        });
      for (auto &t : th) t.join();
      for (auto &r : recs)
        for (auto &kv : r) facets.emplace(kv.first, kv.second);
    }
    undetermined.clear();
    undetermined.shrink_to_fit();
  }

  // status predicates with facets_status_ semantics (unfitted_domain.cpp:354-453)
  const std::array<unsigned char, 2 * N> *rec(int64_t id) const
  {
    auto it = facets.find(id);
    return it == facets.end() ? nullptr : &it->second;
  }
  bool is_cut_cell(int64_t id) const { return rec(id) && status[id] == CELL_CUT; }
  bool is_full_cell(int64_t id) const { return status[id] == CELL_FULL; }
  bool is_full_with_unf(int64_t id) const { return rec(id) && status[id] != CELL_CUT; }
  bool is_cell_with_unf_bdry(int64_t id) const { return is_cut_cell(id) || is_full_with_unf(id); }
  bool facet_has_unf(int64_t id, int lf) const
  {
    auto r = rec(id);
    return r ? fs_has_unf((*r)[lf]) : false;
  }
  bool facet_has_ext(int64_t id, int lf) const
  {
    auto r = rec(id);
    return r ? fs_has_ext((*r)[lf]) : false;
  }
  bool facet_is_cut(int64_t id, int lf) const
  {
    auto r = rec(id);
    return r ? fs_is_cut((*r)[lf]) : false;
  }
  bool facet_is_full_unf(int64_t id, int lf) const
  {
    auto r = rec(id);
    return r ? fs_is_full_unf((*r)[lf]) : false;
  }
  bool facet_is_full(int64_t id, int lf) const
  {
    auto r = rec(id);
    return r ? fs_is_full((*r)[lf]) : is_full_cell(id);
  }

  // ----------------------------------------------------------------------------------------
  // Quadrature drivers
  // ----------------------------------------------------------------------------------------

  // CutCellsQuadWrapper::evalIntegrand (impl_quadrature.cpp:65-69)
  struct CellSink : Target<N>
  {
    Rule &r;
    Box<N> b;
    real vol;
    CellSink(Rule &r_, const Box<N> &b_) : r(r_), b(b_), vol(box_volume(b_)) {}
    void evalIntegrand(const real *x, real w) override
    {
      for (int d = 0; d < N; ++d) r.pts.push_back((x[d] - b.lo[d]) / (b.hi[d] - b.lo[d]));
      r.wts.push_back(w / vol);
    }
  };
  // CutUnfBoundsQuadWrapper::evalIntegrand (impl_quadrature.cpp:87-111)
  struct SurfSink : Target<N>
  {
    Rule &r;
    const Func &phi;
    Box<N> b;
    real vol;
    SurfSink(Rule &r_, const Func &phi_, const Box<N> &b_) : r(r_), phi(phi_), b(b_), vol(box_volume(b_)) {}
    void evalIntegrand(const real *x, real w) override
    {
      for (int d = 0; d < N; ++d) r.pts.push_back((x[d] - b.lo[d]) / (b.hi[d] - b.lo[d]));
      real g[N];
      grad<N, real>(phi, x, g);
      real s = g[0] * g[0];
      for (int d = 1; d < N; ++d) s += g[d] * g[d];
      const real nn = std::sqrt(s);
      for (int d = 0; d < N; ++d) g[d] /= nn;
      for (int d = 0; d < N; ++d) g[d] *= (b.hi[d] - b.lo[d]);
      real s2 = g[0] * g[0];
      for (int d = 1; d < N; ++d) s2 += g[d] * g[d];
      const real n2 = std::sqrt(s2);
      for (int d = 0; d < N; ++d) g[d] /= n2;
      r.wts.push_back(w * n2 / vol);
      for (int d = 0; d < N; ++d) r.nrm.push_back(g[d]);
    }
  };
  // CutIsoBoundsQuadWrapper::evalIntegrand (impl_quadrature.cpp:130-135)
  struct FacetSink : Target<N>
  {
    Rule &r;
    Box<N> b;
    int cd;
    real fvol;
    FacetSink(Rule &r_, const Box<N> &b_, int cd_) : r(r_), b(b_), cd(cd_)
    {
      fvol = 1.0;
      for (int d = 0; d < N; ++d)
        if (d != cd) fvol *= (b.hi[d] - b.lo[d]);
    }
    void evalIntegrand(const real *x, real w) override
    {
      for (int d = 0; d < N; ++d)
        if (d != cd) r.pts.push_back((x[d] - b.lo[d]) / (b.hi[d] - b.lo[d]));
      r.wts.push_back(w / fvol);
    }
  };

  // Quadrature<dim>::create_Gauss_01 / create_tp_quadrature (quadrature.cpp:199-277): dir 0 fastest
  static void append_gauss_01(int dim, int n, Rule &r)
  {
    int tot = 1;
    for (int d = 0; d < dim; ++d) tot *= n;
    for (int t = 0; t < tot; ++t) {
      int idx = t;
      real w = 1.0;
      real xs[3];
      int is[3];
      for (int d = 0; d < dim; ++d) {
        is[d] = idx % n;
        idx /= n;
        xs[d] = gauss_x(n, is[d]);
      }
      // weights: w = w0 * w1 (* w2) in the reference's loop order
      if (dim == 1) w = gauss_w(n, is[0]);
      if (dim == 2) w = gauss_w(n, is[0]) * gauss_w(n, is[1]);
      if (dim == 3) w = gauss_w(n, is[0]) * gauss_w(n, is[1]) * gauss_w(n, is[2]);
      for (int d = 0; d < dim; ++d) r.pts.push_back(xs[d]);
      r.wts.push_back(w);
    }
  }

  // create_cell_quadrature (impl_quadrature.cpp:493-537)
  void cell_quadrature(int64_t id, int n, Rule &r, Stats *st)
  {
    if (is_cut_cell(id)) {
      const Box<N> b = cell_box(id);
      const size_t n0 = r.wts.size();
      CellSink sink(r, b);
      quad_gen<N>(phi, b, -1, 0, n, sink, st);
      r.n_per.push_back((int)(r.wts.size() - n0));
    } else if (is_full_cell(id)) {
      const size_t n0 = r.wts.size();
      append_gauss_01(N, n, r);
      r.n_per.push_back((int)(r.wts.size() - n0));
    } else {
      r.n_per.push_back(0);
    }
  }

  // purge_facet_points, cell version (impl_quadrature.cpp:238-288)
  void purge_cell_points(int64_t id, int lf, size_t first, Rule &r)
  {
    if (!facet_has_unf(id, lf) && !facet_has_ext(id, lf)) return;
    const real tol = 1000.0 * EPS;
    const int cd = lf / 2, side = lf % 2;
    const Box<N> b = cell_box(id);
    const real fc = side == 0 ? 0.0 : 1.0;
    std::vector<size_t> erase;
    for (size_t i = first; i < r.wts.size(); ++i) {
      const real *pt = &r.pts[N * i];
      if (!(std::fabs(pt[cd] - fc) <= tol)) continue;
      const real nc = r.nrm[N * i + cd];
      if ((1.0 - std::fabs(nc)) > tol) continue;  // tol.smaller_than(|n|, 1)
      real x[N];
      for (int d = 0; d < N; ++d) x[d] = pt[d] * (b.hi[d] - b.lo[d]) + b.lo[d];
      if (std::fabs(eval<N, real>(phi, x)) <= tol) erase.push_back(i);
    }
    for (size_t k = erase.size(); k-- > 0;) {
      const size_t i = erase[k];
      r.pts.erase(r.pts.begin() + N * i, r.pts.begin() + N * (i + 1));
      r.nrm.erase(r.nrm.begin() + N * i, r.nrm.begin() + N * (i + 1));
      r.wts.erase(r.wts.begin() + i);
    }
  }

  // purge_facet_points, facet version (impl_quadrature.cpp:291-347)
  void purge_facet_points(int64_t id, int lf, bool purge_unf, bool purge_unf_ext, bool purge_cut, size_t first, Rule &r)
  {
    const bool has_unf = facet_has_unf(id, lf);
    purge_unf = purge_unf && has_unf;
    purge_unf_ext = purge_unf_ext && has_unf;
    purge_cut = purge_cut && facet_is_cut(id, lf);
    if (!purge_unf && !purge_unf_ext && !purge_cut) return;
    const real tol = 1000.0 * EPS;
    const int cd = lf / 2, side = lf % 2;
    const Box<N> b = cell_box(id);
    const real fc = side == 0 ? 0.0 : 1.0;
    std::vector<size_t> erase;
    for (size_t i = first; i < r.wts.size(); ++i) {
      const real *pt = &r.pts[(N - 1) * i];
      real p01[N], x[N];
      for (int d = 0, k = 0; d < N; ++d) p01[d] = (d == cd) ? fc : pt[k++];
      for (int d = 0; d < N; ++d) x[d] = p01[d] * (b.hi[d] - b.lo[d]) + b.lo[d];
      if (std::fabs(eval<N, real>(phi, x)) <= tol) {
        real g[N];
        grad<N, real>(phi, x, g);
        const bool outside = side == 0 ? (g[cd] < -tol) : (tol < g[cd]);
        if ((outside && purge_unf) || (!outside && purge_unf_ext)) erase.push_back(i);
      } else if (purge_cut) {
        erase.push_back(i);
      }
    }
    for (size_t k = erase.size(); k-- > 0;) {
      const size_t i = erase[k];
      r.pts.erase(r.pts.begin() + (N - 1) * i, r.pts.begin() + (N - 1) * (i + 1));
      r.wts.erase(r.wts.begin() + i);
    }
  }

  // create_facet_quadrature (impl_quadrature.cpp:595-649); appends one n_per entry
  void facet_quadrature(int64_t id, int lf, int n, bool remove_unf, bool remove_cut, Rule &r, Stats *st)
  {
    const bool full_unf = facet_is_full_unf(id, lf);
    if (full_unf || facet_is_full(id, lf)) {
      if (full_unf && remove_unf) {
        r.n_per.push_back(0);
        return;
      }
      const size_t n0 = r.wts.size();
      append_gauss_01(N - 1, n, r);
      r.n_per.push_back((int)(r.wts.size() - n0));
    } else if (facet_is_cut(id, lf) || facet_has_unf(id, lf)) {
      const Box<N> b = cell_box(id);
      const int cd = lf / 2, side = lf % 2;
      const size_t n0 = r.wts.size();
      FacetSink sink(r, b, cd);
      quad_gen<N>(phi, b, cd, side, n, sink, st);
      purge_facet_points(id, lf, remove_unf, true, remove_cut, n0, r);
      r.n_per.push_back((int)(r.wts.size() - n0));
    } else {
      r.n_per.push_back(0);
    }
  }

  // create_cell_unfitted_bound_quadrature (impl_quadrature.cpp:540-592) + add_unf_bdry_quad (:349-389)
  void cell_unf_bound_quadrature(int64_t id, int n, bool include_facet, bool exclude_ext, Rule &r, Stats *st)
  {
    if (!is_cell_with_unf_bdry(id)) {
      r.n_per.push_back(0);
      return;
    }
    const size_t n0 = r.wts.size();
    if (is_cut_cell(id)) {
      const Box<N> b = cell_box(id);
      SurfSink sink(r, phi, b);
      quad_gen<N>(phi, b, N, 0, n, sink, st);
      for (int lf = 0; lf < 2 * N; ++lf)
        if (facet_has_unf(id, lf)) purge_cell_points(id, lf, n0, r);
    }
    if (include_facet) {
      for (int lf = 0; lf < 2 * N; ++lf) {
        if (!facet_has_unf(id, lf)) continue;
        if (exclude_ext && is_exterior_facet(id, lf)) continue;
        Rule fr;
        fr.pdim = N - 1;
        facet_quadrature(id, lf, n, false, true, fr, st);
        const int cd = lf / 2, side = lf % 2;
        const real fc = side == 0 ? 0.0 : 1.0;
        for (size_t i = 0; i < fr.wts.size(); ++i) {
          for (int d = 0, k = 0; d < N; ++d) r.pts.push_back(d == cd ? fc : fr.pts[(N - 1) * i + k++]);
          r.wts.push_back(fr.wts[i]);
          for (int d = 0; d < N; ++d) r.nrm.push_back(d == cd ? (side == 0 ? -1.0 : 1.0) : 0.0);
        }
      }
    }
    r.n_per.push_back((int)(r.wts.size() - n0));
  }

  // create_facets_quadrature_{interior,exterior}_integral lambdas (cut_quadrature.cpp:178-262)
  void facets_quadrature_entry(int64_t id, int lf, int n, bool exterior, Rule &r, Stats *st)
  {
    const bool ext_facet = is_exterior_facet(id, lf);
    if (!exterior) {
      if (!ext_facet)
        facet_quadrature(id, lf, n, true, false, r, st);
      else
        r.n_per.push_back(0);
    } else {
      if (ext_facet)
        facet_quadrature(id, lf, n, false, false, r, st);
      else if (facet_has_unf(id, lf))
        facet_quadrature(id, lf, n, false, true, r, st);
      else
        r.n_per.push_back(0);
    }
  }
};

// Storage a Func may point into: Bezier coefficients and the composite description.
struct FuncStore
{
  std::vector<real> coefs[7];  // [0] plain Bezier function, [1..6] Bezier leaves of a composite
  Composite comp;
};

// one non-composite function from params; BEZIER: params = order[dim] followed by the coefficients
static Func make_leaf(int dim, int kind, int negative, const double *params, int nparams, std::vector<real> &store)
{
  Func f;
  std::memset(&f, 0, sizeof(f));
  f.kind = kind;
  f.dim = dim;
  f.negative = negative;
  if (kind == K_BEZIER) {
    for (int d = 0; d < dim; ++d) f.order[d] = (int)params[d];
    store.assign(params + dim, params + nparams);
    f.coefs = store.data();
  } else {
    for (int i = 0; i < nparams && i < 16; ++i) f.p[i] = params[i];
  }
  return f;
}

// params -> Func.  COMPOSITE: the postfix program qugar_b200/cpp.py emits (3, n_leaf, n_aff, n_instr, leaves (kind,
// negative, np, p[np]), affine maps (12 each), instructions (op, arg): 0 LEAF l, 1 XPUSH a, 2 XPOP a, 3 NEG, 4 ADD, 5 SUB),
// rebuilt here into the expression TREE the reference evaluates (XPUSH a ... XPOP a brackets a transformed node).
static Func make_func(int dim, int kind, int negative, const double *params, int nparams, FuncStore &store)
{
  if (kind != K_COMPOSITE) return make_leaf(dim, kind, negative, params, nparams, store.coefs[0]);
  Func f;
  std::memset(&f, 0, sizeof(f));
  f.kind = kind;
  f.dim = dim;
  f.negative = negative;
  Composite &c = store.comp;
  c.leaves.clear();
  c.affs.clear();
  c.nodes.clear();
  const int nleaf = (int)params[1], naff = (int)params[2], ninstr = (int)params[3];
  const double *q = params + 4;
  for (int l = 0; l < nleaf; ++l) {
    const int np = (int)q[2];
    c.leaves.push_back(make_leaf(dim, (int)q[0], (int)q[1], q + 3, np, store.coefs[1 + l]));
    q += 3 + np;
  }
  for (int a = 0; a < naff; ++a) {
    c.affs.push_back(std::vector<real>(q, q + 12));
    q += 12;
  }
  std::vector<int> stack;
  for (int i = 0; i < ninstr; ++i, q += 2) {
    const int op = (int)q[0], arg = (int)q[1];
    CompNode nd;
    nd.op = 0; nd.leaf = 0; nd.aff = 0; nd.child[0] = nd.child[1] = -1;
    if (op == 1) continue;  // XPUSH: the matching XPOP creates the transformed node
    if (op == 0) {
      nd.op = 0;
      nd.leaf = arg;
    } else if (op == 2) {
      nd.op = 1;
      nd.aff = arg;
      nd.child[0] = stack.back();
      stack.pop_back();
    } else if (op == 3) {
      nd.op = 2;
      nd.child[0] = stack.back();
      stack.pop_back();
    } else {
      nd.op = op == 4 ? 3 : 4;
      nd.child[1] = stack.back();
      stack.pop_back();
      nd.child[0] = stack.back();
      stack.pop_back();
    }
    c.nodes.push_back(nd);
    stack.push_back((int)c.nodes.size() - 1);
  }
  c.root = stack.back();
  f.comp = &c;
  return f;
}

struct AnyDomain
{
  int dim;
  FuncStore store;  // Bezier coefficients / composite description the domain's Func points into
  std::unique_ptr<Domain<2>> d2;
  std::unique_ptr<Domain<3>> d3;
};

static void merge(Rule &dst, const Rule &src)
{
  dst.n_per.insert(dst.n_per.end(), src.n_per.begin(), src.n_per.end());
  dst.pts.insert(dst.pts.end(), src.pts.begin(), src.pts.end());
  dst.wts.insert(dst.wts.end(), src.wts.begin(), src.wts.end());
  dst.nrm.insert(dst.nrm.end(), src.nrm.begin(), src.nrm.end());
}

// Runs fn(i, rule, stats) for i in [0,nc) over nthreads contiguous chunks and concatenates in order.
template<typename Fn> static Rule *run_chunks(int64_t nc, int nthreads, int pdim, bool normals, Stats *total, Fn fn)
{
  if (nthreads < 1) nthreads = 1;
  if ((int64_t)nthreads > nc) nthreads = nc > 0 ? (int)nc : 1;
  std::vector<Rule> parts((size_t)nthreads);
  std::vector<Stats> stats((size_t)nthreads);
  auto work = [&](int t) {
    const int64_t b = nc * t / nthreads, e = nc * (t + 1) / nthreads;
    for (int64_t i = b; i < e; ++i) fn(i, parts[(size_t)t], &stats[(size_t)t]);
  };
  if (nthreads == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(work, t);
    for (auto &t : th) t.join();
  }
  Rule *out = new Rule();
  out->pdim = pdim;
  out->has_normals = normals;
  for (auto &p : parts) merge(*out, p);
  if (total) {
    for (auto &s : stats) {
      total->n_bisections += s.n_bisections;
      total->n_leaves += s.n_leaves;
      if (s.max_level > total->max_level) total->max_level = s.max_level;
    }
  }
  return out;
}

}  // namespace orc

using namespace orc;

extern "C" {

void orc_set_knobs3(int seg_le, int rf_cap_mode, int rf_use_sign, double trig_C, int newton_maxsteps)
{
  knobs().seg_le = seg_le;
  knobs().rf_cap_mode = rf_cap_mode;
  knobs().rf_use_sign = rf_use_sign;
  knobs().trig_C = trig_C;
  knobs().newton_maxsteps = newton_maxsteps;
}

// generic knob setter for the calibration sweeps (tools/calibrate_oracle.py)
int orc_set_knob(const char *name, double v)
{
  std::string n(name);
  Knobs &k = knobs();
  if (n == "trig_bound") k.trig_bound = (int)v;
  else if (n == "rootfind_maxlevel") k.rootfind_maxlevel = (int)v;
  else if (n == "dev_order") k.dev_order = (int)v;
  else if (n == "seg_tol_factor") k.seg_tol_factor = v;
  else if (n == "newton_tol_factor") k.newton_tol_factor = v;
  else if (n == "seg_le") k.seg_le = (int)v;
  else if (n == "rf_cap_mode") k.rf_cap_mode = (int)v;
  else if (n == "rf_use_sign") k.rf_use_sign = (int)v;
  else if (n == "trig_C") k.trig_C = v;
  else if (n == "newton_maxsteps") k.newton_maxsteps = (int)v;
  else if (n == "tol_scale") k.tol_scale = (int)v;
  else if (n == "seg_tol_scale") k.seg_tol_scale = (int)v;
  else if (n == "psi_default_sign") k.psi_default_sign = (int)v;
  else return -1;
  return 0;
}

void orc_set_trace(int on) { trace_on() = on; }

void orc_set_tol_scale(int tol_scale) { knobs().tol_scale = tol_scale; knobs().seg_tol_scale = tol_scale; }
void orc_set_tol_scales(int newton, int seg) { knobs().tol_scale = newton; knobs().seg_tol_scale = seg; }

void orc_set_knobs2(double seg_tol_factor, double newton_tol_factor)
{
  knobs().seg_tol_factor = seg_tol_factor;
  knobs().newton_tol_factor = newton_tol_factor;
}

void orc_set_knobs(int trig_bound, int rootfind_maxlevel, int dev_order)
{
  knobs().trig_bound = trig_bound;
  knobs().rootfind_maxlevel = rootfind_maxlevel;
  knobs().dev_order = dev_order;
}

int orc_uses_libm()
{
#ifdef ORACLE_LIBM
  return 1;
#else
  return 0;
#endif
}

void orc_sincos(double x, double *s, double *c)
{
  *s = sin_(x);
  *c = cos_(x);
}

// breaks: concatenated per-direction arrays, n_breaks[dim] entries each
void *orc_domain_create_mt(int dim, int kind, int negative, const double *params, int nparams, const double *breaks,
  const int64_t *n_breaks, int nthreads);

void *orc_domain_create(int dim, int kind, int negative, const double *params, int nparams, const double *breaks,
  const int64_t *n_breaks)
{
  return orc_domain_create_mt(dim, kind, negative, params, nparams, breaks, n_breaks, 1);
}

void *orc_domain_create_mt(int dim, int kind, int negative, const double *params, int nparams, const double *breaks,
  const int64_t *n_breaks, int nthreads)
{
  AnyDomain *ad = new AnyDomain();
  const Func f = make_func(dim, kind, negative, params, nparams, ad->store);
  ad->dim = dim;
  const double *b = breaks;
  if (dim == 2) {
    ad->d2.reset(new Domain<2>());
    ad->d2->phi = f;
    for (int d = 0; d < 2; ++d) {
      ad->d2->breaks[d].assign(b, b + n_breaks[d]);
      ad->d2->n[d] = n_breaks[d] - 1;
      b += n_breaks[d];
    }
    ad->d2->build(nthreads);
  } else {
    ad->d3.reset(new Domain<3>());
    ad->d3->phi = f;
    for (int d = 0; d < 3; ++d) {
      ad->d3->breaks[d].assign(b, b + n_breaks[d]);
      ad->d3->n[d] = n_breaks[d] - 1;
      b += n_breaks[d];
    }
    ad->d3->build(nthreads);
  }
  return ad;
}

void orc_domain_free(void *h) { delete (AnyDomain *)h; }

int64_t orc_domain_num_cells(void *h)
{
  AnyDomain *ad = (AnyDomain *)h;
  return ad->dim == 2 ? ad->d2->num_cells() : ad->d3->num_cells();
}

int64_t orc_domain_num_facet_records(void *h)
{
  AnyDomain *ad = (AnyDomain *)h;
  return ad->dim == 2 ? (int64_t)ad->d2->facets.size() : (int64_t)ad->d3->facets.size();
}

void orc_domain_status(void *h, unsigned char *out)
{
  AnyDomain *ad = (AnyDomain *)h;
  const auto &s = ad->dim == 2 ? ad->d2->status : ad->d3->status;
  std::memcpy(out, s.data(), s.size());
}

// cells sorted ascending; statuses 2*dim per record
void orc_domain_facet_records(void *h, int64_t *cells, unsigned char *statuses)
{
  AnyDomain *ad = (AnyDomain *)h;
  size_t i = 0;
  if (ad->dim == 2) {
    for (auto &kv : ad->d2->facets) {
      cells[i] = kv.first;
      for (int k = 0; k < 4; ++k) statuses[4 * i + k] = kv.second[k];
      ++i;
    }
  } else {
    for (auto &kv : ad->d3->facets) {
      cells[i] = kv.first;
      for (int k = 0; k < 6; ++k) statuses[6 * i + k] = kv.second[k];
      ++i;
    }
  }
}

void *orc_quad_cells(void *h, const int64_t *cells, int64_t nc, int n, int nthreads, long long *stats_out)
{
  AnyDomain *ad = (AnyDomain *)h;
  Stats st;
  Rule *r;
  if (ad->dim == 2)
    r = run_chunks(nc, nthreads, 2, false, &st, [&](int64_t i, Rule &rr, Stats *s) { ad->d2->cell_quadrature(cells[i], n, rr, s); });
  else
    r = run_chunks(nc, nthreads, 3, false, &st, [&](int64_t i, Rule &rr, Stats *s) { ad->d3->cell_quadrature(cells[i], n, rr, s); });
  if (stats_out) {
    stats_out[0] = st.n_bisections;
    stats_out[1] = st.n_leaves;
    stats_out[2] = st.max_level;
  }
  return r;
}

// Timing entry for the CPU baseline: runs the three reference calls over `cells` on nthreads threads, each
// thread keeping its own containers (no concatenation); returns total points of the interior rule.
long long orc_time_rules(void *h, const int64_t *cells, int64_t nc, int n, int nthreads, int do_interior, int do_bdry,
  long long *n_bdry_pts)
{
  AnyDomain *ad = (AnyDomain *)h;
  if (nthreads < 1) nthreads = 1;
  std::vector<long long> ni((size_t)nthreads, 0), nb((size_t)nthreads, 0);
  auto work = [&](int t) {
    const int64_t b = nc * t / nthreads, e = nc * (t + 1) / nthreads;
    Rule ri, rb;
    Stats st;
    // the reference reserves 2 n^dim points per cell up front (cut_quadrature.cpp:114-122, 147-155)
    size_t np_cell = 2, ns_cell = 2;
    for (int d = 0; d < ad->dim; ++d) np_cell *= (size_t)n;
    for (int d = 0; d + 1 < ad->dim; ++d) ns_cell *= (size_t)n;
    if (do_interior) {
      ri.pts.reserve((size_t)(e - b) * np_cell * ad->dim);
      ri.wts.reserve((size_t)(e - b) * np_cell);
    }
    if (do_bdry) {
      rb.pts.reserve((size_t)(e - b) * ns_cell * ad->dim);
      rb.nrm.reserve((size_t)(e - b) * ns_cell * ad->dim);
      rb.wts.reserve((size_t)(e - b) * ns_cell);
    }
    for (int64_t i = b; i < e; ++i) {
      if (do_interior) {
        if (ad->dim == 2)
          ad->d2->cell_quadrature(cells[i], n, ri, &st);
        else
          ad->d3->cell_quadrature(cells[i], n, ri, &st);
      }
      if (do_bdry) {
        if (ad->dim == 2)
          ad->d2->cell_unf_bound_quadrature(cells[i], n, false, false, rb, &st);
        else
          ad->d3->cell_unf_bound_quadrature(cells[i], n, false, false, rb, &st);
      }
    }
    ni[(size_t)t] = (long long)ri.wts.size();
    nb[(size_t)t] = (long long)rb.wts.size();
  };
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t) th.emplace_back(work, t);
  for (auto &t : th) t.join();
  long long a = 0, b = 0;
  for (int t = 0; t < nthreads; ++t) {
    a += ni[(size_t)t];
    b += nb[(size_t)t];
  }
  if (n_bdry_pts) *n_bdry_pts = b;
  return a;
}

void *orc_quad_unf_bdry(void *h, const int64_t *cells, int64_t nc, int n, int include_facet, int exclude_ext,
  int nthreads)
{
  AnyDomain *ad = (AnyDomain *)h;
  if (ad->dim == 2)
    return run_chunks(nc, nthreads, 2, true, nullptr, [&](int64_t i, Rule &rr, Stats *s) {
      ad->d2->cell_unf_bound_quadrature(cells[i], n, include_facet != 0, exclude_ext != 0, rr, s);
    });
  return run_chunks(nc, nthreads, 3, true, nullptr, [&](int64_t i, Rule &rr, Stats *s) {
    ad->d3->cell_unf_bound_quadrature(cells[i], n, include_facet != 0, exclude_ext != 0, rr, s);
  });
}

void *orc_quad_facets(void *h, const int64_t *cells, const int32_t *facets, int64_t nc, int n, int exterior,
  int nthreads)
{
  AnyDomain *ad = (AnyDomain *)h;
  if (ad->dim == 2)
    return run_chunks(nc, nthreads, 1, false, nullptr, [&](int64_t i, Rule &rr, Stats *s) {
      ad->d2->facets_quadrature_entry(cells[i], facets[i], n, exterior != 0, rr, s);
    });
  return run_chunks(nc, nthreads, 2, false, nullptr, [&](int64_t i, Rule &rr, Stats *s) {
    ad->d3->facets_quadrature_entry(cells[i], facets[i], n, exterior != 0, rr, s);
  });
}

void orc_rule_sizes(void *h, int64_t *n_entities, int64_t *n_pts, int *pdim, int *has_normals)
{
  Rule *r = (Rule *)h;
  *n_entities = (int64_t)r->n_per.size();
  *n_pts = (int64_t)r->wts.size();
  *pdim = r->pdim;
  *has_normals = r->has_normals ? 1 : 0;
}

void orc_rule_copy(void *h, int32_t *n_per, double *pts, double *wts, double *nrm)
{
  Rule *r = (Rule *)h;
  if (n_per) std::memcpy(n_per, r->n_per.data(), r->n_per.size() * sizeof(int));
  if (pts) std::memcpy(pts, r->pts.data(), r->pts.size() * sizeof(double));
  if (wts) std::memcpy(wts, r->wts.data(), r->wts.size() * sizeof(double));
  if (nrm && r->has_normals) std::memcpy(nrm, r->nrm.data(), r->nrm.size() * sizeof(double));
}

void orc_rule_free(void *h) { delete (Rule *)h; }

// phi and grad at points (for function-level tests)
void orc_func_eval(int dim, int kind, int negative, const double *params, int nparams, const double *pts, int64_t n,
  double *val, double *grd)
{
  FuncStore store;
  const Func f = make_func(dim, kind, negative, params, nparams, store);
  for (int64_t i = 0; i < n; ++i) {
    if (dim == 2) {
      val[i] = eval<2, real>(f, pts + 2 * i);
      if (grd) grad<2, real>(f, pts + 2 * i, grd + 2 * i);
    } else {
      val[i] = eval<3, real>(f, pts + 3 * i);
      if (grd) grad<3, real>(f, pts + 3 * i, grd + 3 * i);
    }
  }
}

// Interval evaluation of phi on a box: returns alpha, maxDeviation (for kernel-level parity tests)
void orc_func_eval_interval(int dim, int kind, int negative, const double *params, int nparams, const double *lo,
  const double *hi, double *alpha, double *maxdev)
{
  FuncStore store;
  const Func f = make_func(dim, kind, negative, params, nparams, store);
  if (dim == 2) {
    Interval<2> x[2];
    for (int d = 0; d < 2; ++d) {
      x[d] = Interval<2>::variable(0.5 * (lo[d] + hi[d]), d);
      Interval<2>::delta(d) = 0.5 * (hi[d] - lo[d]);
    }
    Interval<2> r = eval<2, Interval<2>>(f, x);
    *alpha = r.alpha;
    *maxdev = r.maxDeviation();
  } else {
    Interval<3> x[3];
    for (int d = 0; d < 3; ++d) {
      x[d] = Interval<3>::variable(0.5 * (lo[d] + hi[d]), d);
      Interval<3>::delta(d) = 0.5 * (hi[d] - lo[d]);
    }
    Interval<3> r = eval<3, Interval<3>>(f, x);
    *alpha = r.alpha;
    *maxdev = r.maxDeviation();
  }
}

// BezierTP::rescale_domain (bezier_tp.cpp:403-415) = algoim::bernstein::deCasteljau(coefficients, a, b): the Bernstein
// coefficients of the polynomial restricted (or extended) to the box [lo,hi], one direction after the other; per direction
// the order of the two one-sided subdivisions is chosen for stability, as Algoim does (restated from the published
// algorithm, not from source).  coefs: direction 0 fastest (bezier_tp.cpp:53-61).  BezierTP::sign (:417-431): +1 / -1 if
// all coefficients are positive / negative, else 0.
static void bz_left(double *c, int P, long long stride, double tau)
{
  for (int i = 1; i < P; ++i)
    for (int j = P - 1; j >= i; --j) c[j * stride] = c[j * stride] * tau + c[(j - 1) * stride] * (1.0 - tau);
}
static void bz_right(double *c, int P, long long stride, double tau)
{
  for (int i = 1; i < P; ++i)
    for (int j = 0; j < P - i; ++j) c[j * stride] = c[j * stride] * (1.0 - tau) + c[(j + 1) * stride] * tau;
}
int orc_bezier_rescale(int dim, const int *order, const double *coefs, const double *lo, const double *hi, double *out)
{
  long long n = 1;
  for (int d = 0; d < dim; ++d) n *= order[d];
  for (long long i = 0; i < n; ++i) out[i] = coefs[i];
  long long stride = 1;
  for (int d = 0; d < dim; ++d) {
    const int P = order[d];
    const double a = lo[d], b = hi[d];
    // every line of direction d
    for (long long base = 0; base < n; ++base) {
      if ((base / stride) % P != 0) continue;
      double *c = out + base;
      if (std::fabs(b) >= std::fabs(a - 1.0)) {
        bz_left(c, P, stride, b);
        bz_right(c, P, stride, a / b);
      } else {
        bz_right(c, P, stride, a);
        bz_left(c, P, stride, (b - a) / (1.0 - a));
      }
    }
    stride *= P;
  }
  int pos = 1, neg = 1;
  for (long long i = 0; i < n; ++i) {
    pos &= out[i] > 0.0;
    neg &= out[i] < 0.0;
  }
  return pos ? 1 : (neg ? -1 : 0);
}

// ---- reparameterization (oracle_reparam.hpp) -------------------------------------------------------------
struct AnyMesh
{
  int dim;
  std::unique_ptr<ReparamMesh<2>> m2;
  std::unique_ptr<ReparamMesh<3>> m3;
};

// impl_reparam_general.cpp:885-935: one box.  levelset: 0 volume, 1 zero level set; facet >= 0: that face of the box.
// merge_tol < 0: no merge (the reference's tests merge with 2e4 eps).  stable: stable sorts in make_points_unique.
void *orc_reparam_general(int dim, int kind, int negative, const double *params, int nparams, const double *lo,
  const double *hi, int order, int levelset, int facet, double merge_tol, int stable)
{
  FuncStore store;
  const Func f = make_func(dim, kind, negative, params, nparams, store);
  AnyMesh *am = new AnyMesh();
  am->dim = dim;
  const int pd = (levelset || facet >= 0) ? dim - 1 : dim;
  if (dim == 2) {
    am->m2.reset(new ReparamMesh<2>(pd, order));
    am->m2->merge_stable = stable;
    Box<2> b;
    for (int d = 0; d < 2; ++d) { b.lo[d] = lo[d]; b.hi[d] = hi[d]; }
    reparam_general<2>(f, b, *am->m2, levelset != 0, facet);
    if (merge_tol >= 0.0) am->m2->merge_coincident_points(merge_tol);
  } else {
    am->m3.reset(new ReparamMesh<3>(pd, order));
    am->m3->merge_stable = stable;
    Box<3> b;
    for (int d = 0; d < 3; ++d) { b.lo[d] = lo[d]; b.hi[d] = hi[d]; }
    reparam_general<3>(f, b, *am->m3, levelset != 0, facet);
    if (merge_tol >= 0.0) am->m3->merge_coincident_points(merge_tol);
  }
  return am;
}

// impl_reparam.cpp:46-113: the cut cells among `cells` (sorted), then the full cells (volume only, with wirebasket),
// then the optional merge with 1e4 eps.
extern "C++" {
template<int N> static void reparam_domain_impl(Domain<N> &dom, const int64_t *cells, int64_t nc, ReparamMesh<N> &mesh,
  bool levelset, bool merge)
{
  std::vector<int64_t> sorted(cells, cells + nc);
  std::sort(sorted.begin(), sorted.end());
  for (int64_t c : sorted)
    if (dom.status[(size_t)c] == CELL_CUT) reparam_general<N>(dom.phi, dom.cell_box(c), mesh, levelset, -1);
  if (!levelset)
    for (int64_t c : sorted)
      if (dom.status[(size_t)c] == CELL_FULL) mesh.add_full_cell(dom.cell_box(c), true);
  if (merge) mesh.merge_coincident_points(1.0e4 * EPS);
}
}  // extern "C++"

void *orc_reparam_domain(void *h, const int64_t *cells, int64_t nc, int order, int levelset, int merge, int stable)
{
  AnyDomain *ad = (AnyDomain *)h;
  AnyMesh *am = new AnyMesh();
  am->dim = ad->dim;
  const int pd = levelset ? ad->dim - 1 : ad->dim;
  if (ad->dim == 2) {
    am->m2.reset(new ReparamMesh<2>(pd, order));
    am->m2->merge_stable = stable;
    reparam_domain_impl<2>(*ad->d2, cells, nc, *am->m2, levelset != 0, merge != 0);
  } else {
    am->m3.reset(new ReparamMesh<3>(pd, order));
    am->m3->merge_stable = stable;
    reparam_domain_impl<3>(*ad->d3, cells, nc, *am->m3, levelset != 0, merge != 0);
  }
  return am;
}

void orc_reparam_sizes(void *h, int64_t *npts, int64_t *nconn, int64_t *nwires, int *pd, int *order)
{
  AnyMesh *am = (AnyMesh *)h;
  if (am->dim == 2) {
    *npts = (int64_t)am->m2->points.size(); *nconn = (int64_t)am->m2->conn.size(); *nwires = (int64_t)am->m2->wires.size();
    *pd = am->m2->pd; *order = am->m2->order;
  } else {
    *npts = (int64_t)am->m3->points.size(); *nconn = (int64_t)am->m3->conn.size(); *nwires = (int64_t)am->m3->wires.size();
    *pd = am->m3->pd; *order = am->m3->order;
  }
}

void orc_reparam_copy(void *h, double *pts, int64_t *conn, int64_t *wires)
{
  AnyMesh *am = (AnyMesh *)h;
  if (am->dim == 2) {
    for (size_t i = 0; i < am->m2->points.size(); ++i) for (int d = 0; d < 2; ++d) pts[2 * i + d] = am->m2->points[i][d];
    for (size_t i = 0; i < am->m2->conn.size(); ++i) conn[i] = (int64_t)am->m2->conn[i];
    for (size_t i = 0; i < am->m2->wires.size(); ++i) wires[i] = (int64_t)am->m2->wires[i];
  } else {
    for (size_t i = 0; i < am->m3->points.size(); ++i) for (int d = 0; d < 3; ++d) pts[3 * i + d] = am->m3->points[i][d];
    for (size_t i = 0; i < am->m3->conn.size(); ++i) conn[i] = (int64_t)am->m3->conn[i];
    for (size_t i = 0; i < am->m3->wires.size(); ++i) wires[i] = (int64_t)am->m3->wires[i];
  }
}

void orc_reparam_free(void *h) { delete (AnyMesh *)h; }

}  // extern "C"
