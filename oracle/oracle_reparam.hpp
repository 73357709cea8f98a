// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle_math.hpp header).
//
// CPU restatement of the reference's reparameterization of implicit domains by a GENERAL level set:
//   /root/reference/cpp/qugar/impl_reparam_general.cpp:70-875   ImplicitGeneralReparam (recursion, reference
//                                                               intervals, "similar" intervals, point insertion)
//   /root/reference/cpp/qugar/impl_utils.cpp:89-206             RootsIntervals (add_root, adjust_roots)
//   /root/reference/cpp/qugar/reparam_mesh.cpp:349-735,759-785  ReparamMesh (cells, edges, wirebasket, merge)
//   /root/reference/cpp/qugar/impl_reparam_mesh.cpp:47-345      orientation, wirebasket predicates
//   /root/reference/cpp/qugar/point.cpp:34-134                  make_points_unique
//   /root/reference/cpp/qugar/lagrange_tp_utils.cpp:38-150      Lagrange basis at the cell midpoint
//   /root/reference/cpp/qugar/impl_reparam.cpp:46-113           create_reparameterization (batch driver)
// Recursive and map-driven like the reference (the device code is a flat per-leaf state machine).  Pinned to the golden
// sizes and centroids of every case of cpp/test/test_impl_general_reparam_{0,1}.cpp (and to their connectivity averages /
// moments where those do not depend on the tie order of an unstable std::sort); the Fischer-Koch case of _2.cpp is 12 points
// off (tests/test_oracle_reparam.py).
// From memory (not in tree): algoim::bernstein::modifiedChebyshevNode; rootFind / Interval as in oracle_quad.hpp.
#pragma once
#include "oracle_quad.hpp"
#include <array>
#include <cmath>
#include <functional>
#include <map>

namespace orc {

constexpr int CHEBYSHEV_ORDER = 7;  // reparam_mesh.hpp:53

// algoim::bernstein::modifiedChebyshevNode(i, n): Chebyshev roots rescaled so that the end nodes are 0 and 1.
inline real modified_chebyshev_node(int i, int n)
{
  if (n == 1) return 0.5;
  const real tn = PI / (2 * n);
  return 0.5 - 0.5 * std::cos((2 * i + 1) * tn) / std::cos(tn);
}

// Tolerance (tolerance.cpp:64-82): absolute comparisons.
inline bool tol_equal(real tol, real a, real b) { return std::fabs(a - b) <= tol; }
inline bool tol_zero(real tol, real a) { return std::fabs(a) <= tol; }

// ---- ReparamMesh<dim, range> -----------------------------------------------------------------------------
template<int N> struct ReparamMesh
{
  int pd;     // parametric dimension of the cells
  int order;  // points per direction
  std::vector<std::array<real, N>> points;
  std::vector<size_t> conn, wires;
  int merge_stable = 0;  // 1: std::stable_sort in make_points_unique (what the device implements)

  ReparamMesh(int pd_, int order_) : pd(pd_), order(order_) {}
  size_t npc() const
  {
    size_t n = 1;
    for (int i = 0; i < pd; ++i) n *= (size_t)order;
    return n;
  }
  size_t num_cells() const { return conn.size() / npc(); }
  bool use_chebyshev() const { return order >= CHEBYSHEV_ORDER; }
  real node01(int i) const { return use_chebyshev() ? modified_chebyshev_node(i, order) : real(i) / real(order - 1); }

  int allocate_cells(size_t n)
  {
    const size_t nc = num_cells();
    conn.resize(conn.size() + n * npc());
    return (int)nc;
  }
  void insert_cell_point(const real *p, int cell, int pt_id)
  {
    conn[(size_t)cell * npc() + (size_t)pt_id] = points.size();
    std::array<real, N> a;
    for (int d = 0; d < N; ++d) a[d] = p[d];
    points.push_back(a);
  }
  static size_t flat(const int *t, int n, int order)
  {
    size_t f = 0, s = 1;
    for (int d = 0; d < n; ++d) {
      f += (size_t)t[d] * s;
      s *= (size_t)order;
    }
    return f;
  }

  // reparam_mesh.cpp:516-529
  void sort_edge_points(size_t *e) const
  {
    bool reverse = e[order - 1] < e[0];
    if (e[0] == e[order - 1] && order > 2) reverse = e[order - 2] < e[1];
    if (reverse) std::reverse(e, e + order);
  }
  // reparam_mesh.cpp:535-571 (+ impl_utils.cpp:59-87 for the edge numbering)
  std::vector<size_t> get_edge_points(int cell, int edge) const
  {
    int cdirs[2], sides[2];
    if (pd == 2) {
      cdirs[0] = edge / 2;
      sides[0] = edge % 2;
    } else {
      if (edge < 4) {
        cdirs[0] = 0;
        cdirs[1] = 1;
      } else if (edge < 8) {
        cdirs[0] = 0;
        cdirs[1] = 2;
      } else {
        cdirs[0] = 1;
        cdirs[1] = 2;
      }
      sides[0] = edge % 2;
      sides[1] = (edge / 2) % 2;
    }
    int lo[3] = { 0, 0, 0 }, hi[3] = { order, order, order };
    for (int i = 0; i < pd - 1; ++i) {
      if (sides[i] == 0)
        hi[cdirs[i]] = 1;
      else
        lo[cdirs[i]] = order - 1;
    }
    std::vector<size_t> ids;
    const size_t off = (size_t)cell * npc();
    int t[3];
    for (t[2] = (pd > 2 ? lo[2] : 0); t[2] < (pd > 2 ? hi[2] : 1); ++t[2])
      for (t[1] = lo[1]; t[1] < hi[1]; ++t[1])
        for (t[0] = lo[0]; t[0] < hi[0]; ++t[0]) ids.push_back(conn[off + flat(t, pd, order)]);
    sort_edge_points(ids.data());
    return ids;
  }
  bool coincident(real tol, const std::array<real, N> &a, const std::array<real, N> &b) const
  {
    for (int d = 0; d < N; ++d)
      if (!tol_equal(tol, a[d], b[d])) return false;
    return true;
  }
  bool subentity_degenerate(const std::vector<size_t> &ids, real tol) const
  {
    for (size_t i : ids)
      if (!coincident(tol, points[ids.front()], points[i])) return false;
    return true;
  }
  // reparam_mesh.cpp:575-605
  bool edge_in_subdomain(int sub_dim, const std::vector<size_t> &ids, const Box<N> &box, real tol) const
  {
    auto bounds = [&](size_t id, int *b) {
      for (int d = 0; d < N; ++d) {
        b[d] = -1;
        if (tol_equal(tol, points[id][d], box.lo[d]))
          b[d] = 0;
        else if (tol_equal(tol, points[id][d], box.hi[d]))
          b[d] = 1;
      }
    };
    int b0[N];
    bounds(ids.front(), b0);
    for (size_t id : ids) {
      int b[N];
      bounds(id, b);
      int nb = 0;
      for (int d = 0; d < N; ++d)
        if (b0[d] >= 0 && b0[d] == b[d]) ++nb;
      if (nb < sub_dim) return false;
    }
    return true;
  }
  // reparam_mesh.cpp:352-390 (dim == range)
  void add_full_cell(const Box<N> &box, bool wirebasket)
  {
    int t[3] = { 0, 0, 0 };
    for (t[2] = 0; t[2] < (N > 2 ? order : 1); ++t[2])
      for (t[1] = 0; t[1] < order; ++t[1])
        for (t[0] = 0; t[0] < order; ++t[0]) {
          std::array<real, N> p;
          for (int d = 0; d < N; ++d) p[d] = box.lo[d] + (box.hi[d] - box.lo[d]) * node01(t[d]);
          conn.push_back(points.size());
          points.push_back(p);
        }
    const int cell = (int)num_cells() - 1;
    if (wirebasket) {
      const int ne = N == 2 ? 4 : 12;
      for (int e = 0; e < ne; ++e) {
        const auto ids = get_edge_points(cell, e);
        wires.insert(wires.end(), ids.begin(), ids.end());
      }
    }
  }
  // reparam_mesh.cpp:759-785
  void permute_cell_directions(size_t cell)
  {
    const size_t off = cell * npc();
    if (pd == 1) {
      for (int i = 0; i < order / 2; ++i) std::swap(conn[off + (size_t)i], conn[off + (size_t)(order - 1 - i)]);
      return;
    }
    int t[3] = { 0, 0, 0 };
    for (t[2] = 0; t[2] < (pd > 2 ? order : 1); ++t[2])
      for (t[1] = 0; t[1] < order; ++t[1])
        for (t[0] = 0; t[0] < order; ++t[0]) {
          if (t[0] <= t[1]) continue;
          int u[3] = { t[1], t[0], t[2] };
          std::swap(conn[off + flat(t, pd, order)], conn[off + flat(u, pd, order)]);
        }
  }

  // lagrange_tp_utils.cpp:38-92 with chebyshev == false, i.e. (sic) the modified Chebyshev nodes, at x = 1/2
  void lagrange_mid_1d(std::vector<real> &val, std::vector<real> &der) const
  {
    const real x = 0.5;
    auto c = [&](int i) { return modified_chebyshev_node(i, order); };
    val.clear();
    der.clear();
    for (int i = 0; i < order; ++i) {
      real prod = 1.0;
      for (int j = 0; j < order; ++j)
        if (i != j) prod *= (x - c(j)) / (c(i) - c(j));
      val.push_back(prod);
    }
    for (int i = 0; i < order; ++i) {
      real sum = 0.0;
      for (int j = 0; j < order; ++j) {
        if (i == j) continue;
        real term = 1.0 / (c(i) - c(j));
        for (int k = 0; k < order; ++k)
          if (k != i && k != j) term *= (x - c(k)) / (c(i) - c(k));
        sum += term;
      }
      der.push_back(sum);
    }
  }
  // tensor-product basis / derivative values at the midpoint, flat index dir 0 fastest (lagrange_tp_utils.cpp:96-150)
  void lagrange_mid(std::vector<real> &basis, std::vector<std::vector<real>> &ders) const
  {
    std::vector<real> v, d;
    lagrange_mid_1d(v, d);
    const size_t n = npc();
    basis.assign(n, 1.0);
    ders.assign((size_t)pd, std::vector<real>(n, 1.0));
    for (size_t f = 0; f < n; ++f) {
      int t[3];
      size_t r = f;
      for (int i = 0; i < pd; ++i) {
        t[i] = (int)(r % (size_t)order);
        r /= (size_t)order;
      }
      for (int i = 0; i < pd; ++i) basis[f] *= v[(size_t)t[i]];
      for (int dd = 0; dd < pd; ++dd)
        for (int i = 0; i < pd; ++i) ders[(size_t)dd][f] *= (i == dd) ? d[(size_t)t[i]] : v[(size_t)t[i]];
    }
  }
  void cell_jacobian(size_t cell, const std::vector<real> &basis, const std::vector<std::vector<real>> &ders,
    real jac[3][N], real *eval_pt) const
  {
    for (int a = 0; a < 3; ++a)
      for (int d = 0; d < N; ++d) jac[a][d] = 0.0;
    for (int d = 0; d < N; ++d) eval_pt[d] = 0.0;
    const size_t off = cell * npc();
    for (size_t i = 0; i < npc(); ++i) {
      const auto &p = points[conn[off + i]];
      for (int a = 0; a < pd; ++a)
        for (int d = 0; d < N; ++d) jac[a][d] += p[d] * ders[(size_t)a][i];
      for (int d = 0; d < N; ++d) eval_pt[d] += p[d] * basis[i];
    }
  }
  // impl_reparam_mesh.cpp:183-214 (dim == range)
  void orient_cells_positively(const std::vector<int> &ids)
  {
    std::vector<real> basis;
    std::vector<std::vector<real>> ders;
    lagrange_mid(basis, ders);
    for (int c : ids) {
      real J[3][N], ep[N];
      cell_jacobian((size_t)c, basis, ders, J, ep);
      real det;
      if (N == 3)
        det = (J[0][0] * (J[1][1] * J[2][2 % N] - J[1][2 % N] * J[2][1])) + (J[0][1] * (J[1][2 % N] * J[2][0] - J[1][0] * J[2][2 % N]))
              + (J[0][2 % N] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]));
      else
        det = (J[0][0] * J[1][1]) - (J[0][1] * J[1][0]);
      if (det < 0) permute_cell_directions((size_t)c);
    }
  }
  // impl_reparam_mesh.cpp:229-278 (range == dim + 1)
  void orient_levelset_cells_positively(const std::vector<int> &ids, const std::function<void(const real *, real *)> &outer_normal)
  {
    std::vector<real> basis;
    std::vector<std::vector<real>> ders;
    lagrange_mid(basis, ders);
    for (int c : ids) {
      real J[3][N], ep[N];
      cell_jacobian((size_t)c, basis, ders, J, ep);
      real nrm[N] = {};
      if (pd == 2) {
        nrm[0] = (J[0][1] * J[1][2 % N]) - (J[0][2 % N] * J[1][1]);
        nrm[1] = (J[0][2 % N] * J[1][0]) - (J[0][0] * J[1][2 % N]);
        nrm[2 % N] = (J[0][0] * J[1][1]) - (J[0][1] * J[1][0]);
      } else {
        nrm[0] = J[0][1];
        nrm[1] = -J[0][0];
      }
      real s = 0.0;
      for (int d = 0; d < N; ++d) s += nrm[d] * nrm[d];
      s = std::sqrt(s);
      for (int d = 0; d < N; ++d) nrm[d] /= s;
      real tn[N];
      outer_normal(ep, tn);
      real dot = 0.0;
      for (int d = 0; d < N; ++d) dot += nrm[d] * tn[d];
      if (dot < 0) permute_cell_directions((size_t)c);
    }
  }
  // impl_reparam_mesh.cpp:100-178 with a single implicit function
  void generate_wirebasket(const Func &phi, const std::vector<int> &ids, const Box<N> &box, real tol)
  {
    if (pd < 2) return;
    const int ne = pd == 2 ? 4 : 12;
    auto on_ls = [&](size_t id) { return tol_zero(tol, eval<N, real>(phi, points[id].data())); };
    for (int c : ids)
      for (int e = 0; e < ne; ++e) {
        const auto ep = get_edge_points(c, e);
        if (subentity_degenerate(ep, tol)) continue;
        bool add = edge_in_subdomain(N - 1, ep, box, tol);
        if (!add) {
          if (edge_in_subdomain(1, ep, box, tol)) {
            add = true;
            for (size_t id : ep) add = add && on_ls(id);
          } else {
            // check_internal: needs range-1 functions vanishing at the point; one function is given
            add = true;
            for (size_t id : ep) add = add && (1 >= N - 1 && on_ls(id));
          }
        }
        if (add) wires.insert(wires.end(), ep.begin(), ep.end());
      }
  }

  // point.cpp:34-134
  void find_coincident(int sub_dim, real tol, size_t *idx, size_t n, std::vector<size_t> &map) const
  {
    if (n < 2) return;
    auto cmp = [&](size_t a, size_t b) { return points[a][sub_dim] < points[b][sub_dim]; };
    if (merge_stable)
      std::stable_sort(idx, idx + n, cmp);
    else
      std::sort(idx, idx + n, cmp);
    for (size_t i = 0; i < n - 1;) {
      const size_t i0 = i++;
      for (; i < n; ++i)
        if (!tol_equal(tol, points[idx[i - 1]][sub_dim], points[idx[i]][sub_dim])) break;
      const size_t nco = i - i0;
      if (nco >= 2) {
        if (sub_dim == 0) {
          for (size_t j = i0 + 1; j < i; ++j) map[idx[j]] = idx[i0];
        } else {
          find_coincident(sub_dim - 1, tol, idx + i0, nco, map);
        }
      }
    }
  }
  // reparam_mesh.cpp:607-626, 659-735
  void merge_coincident_points(real tol)
  {
    if (points.size() < 2) return;
    const size_t n = points.size();
    std::vector<size_t> otu(n), idx(n);
    for (size_t i = 0; i < n; ++i) otu[i] = idx[i] = i;
    find_coincident(N - 1, tol, idx.data(), n, otu);
    std::vector<size_t> otn(n);
    std::vector<std::array<real, N>> np;
    for (size_t i = 0; i < n; ++i)
      if (otu[i] == i) {
        otn[i] = np.size();
        np.push_back(points[i]);
      }
    std::vector<size_t> fin(n);
    for (size_t i = 0; i < n; ++i) fin[i] = otn[otu[i]];
    points.swap(np);
    for (auto &c : conn) c = fin[c];
    for (auto &c : wires) c = fin[c];
    purge_duplicate_wirebasket_edges();
  }
  void sort_wirebasket_edges()
  {
    const size_t o = (size_t)order, ne = wires.size() / o;
    for (size_t w = 0; w < ne; ++w) sort_edge_points(wires.data() + w * o);
    std::vector<size_t> ind(ne);
    for (size_t i = 0; i < ne; ++i) ind[i] = i;
    const size_t *e = wires.data();
    std::sort(ind.begin(), ind.end(), [&](size_t l, size_t r) {
      for (size_t d = 0; d < o; ++d)
        if (e[l * o + d] != e[r * o + d]) return e[l * o + d] < e[r * o + d];
      return false;
    });
    const std::vector<size_t> old = wires;
    for (size_t i = 0; i < ne; ++i)
      for (size_t d = 0; d < o; ++d) wires[i * o + d] = old[ind[i] * o + d];
  }
  // reparam_mesh.cpp:698-735, iterator for iterator (edges that occur more than once are dropped altogether)
  void purge_duplicate_wirebasket_edges()
  {
    sort_wirebasket_edges();
    std::vector<size_t> nw;
    const ptrdiff_t o = order;
    auto it0 = wires.cbegin();
    const auto end = wires.cend();
    bool repetition = false;
    for (auto it = it0; it != end;) {
      const auto e0 = it;
      repetition = false;
      for (it += o; it < end; it += o) {
        if (std::equal(e0, e0 + o, it)) {
          if (!repetition) {
            repetition = true;
            nw.insert(nw.end(), it0, it);
          }
        } else {
          if (repetition) it0 = it;
          break;
        }
      }
    }
    if (!repetition) nw.insert(nw.end(), it0, end);
    wires = nw;
  }
};

// ---- RootsIntervals (impl_utils.cpp:89-206) -----------------------------------------------------------------
template<int N> struct RootsIntervals
{
  real point[N];
  std::vector<real> roots;
  std::vector<int> func_ids;
  std::vector<char> active;
  void clear()
  {
    roots.clear();
    func_ids.clear();
    active.clear();
    for (int d = 0; d < N; ++d) point[d] = 0.0;
  }
  void add_root(real r, int f)
  {
    roots.push_back(r);
    func_ids.push_back(f);
  }
  int num_roots() const { return (int)roots.size(); }
  void adjust_roots(real tol, real x0, real x1)
  {
    if (roots.empty()) return;
    std::vector<std::pair<real, int>> ri;
    for (size_t i = 0; i < roots.size(); ++i) ri.emplace_back(roots[i], func_ids[i]);
    std::sort(ri.begin(), ri.end(), [](const auto &l, const auto &r) { return l.first < r.first; });
    std::vector<real> ro;
    std::vector<int> fo;
    for (auto &p : ri) {
      ro.push_back(p.first);
      fo.push_back(p.second);
    }
    roots.clear();
    func_ids.clear();
    roots.push_back(x0);
    func_ids.push_back(-1);
    bool has_x0 = false, has_x1 = false;
    for (int i = 0; i < (int)ro.size(); ++i) {
      real root = ro[(size_t)i];
      const int f = fo[(size_t)i];
      if (tol_equal(tol, root, x0)) {
        if (f == -1) {
          has_x0 = true;
          continue;
        }
        root = x0;
      } else if (tol_equal(tol, root, x1)) {
        if (f == -1) {
          has_x1 = true;
          continue;
        }
        root = x1;
      } else if (0 < i && tol_equal(tol, root, ro[(size_t)i - 1])) {
        root = ro[(size_t)i - 1];
      }
      roots.push_back(root);
      func_ids.push_back(f);
    }
    if (!has_x0) {
      roots.erase(roots.begin());
      func_ids.erase(func_ids.begin());
    }
    if (has_x1) {
      roots.push_back(x1);
      func_ids.push_back(-1);
    }
  }
};

template<int N> struct TIdx
{
  int v[N];
  TIdx()
  {
    for (int d = 0; d < N; ++d) v[d] = 0;
  }
  bool operator<(const TIdx &o) const
  {
    for (int d = 0; d < N; ++d)
      if (v[d] != o.v[d]) return v[d] < o.v[d];
    return false;
  }
};

// ---- ImplicitGeneralReparam<dim, range, R, S, T> (impl_reparam_general.cpp:70-875) --------------------------
template<int N> struct GeneralReparam
{
  static const int MAXPSI = 1 << (N - 1);
  static const int MAX_SUBDIV = 16;

  const Func &phi;
  GeneralReparam *up;       // R when !T: the higher-dimension instance
  ReparamMesh<N> *mesh;     // R when T
  int M;                    // `dim`
  bool S, T;
  bool free_[N];
  PsiCode<N> psi[MAXPSI];
  int psiCount;
  Box<N> xrange;
  int order;
  int k = -1;
  std::map<TIdx<N>, RootsIntervals<N>> ref_intervals;
  std::map<TIdx<N>, int> cells_map;

  real node01(int i) const
  {
    return order >= CHEBYSHEV_ORDER ? modified_chebyshev_node(i, order) : real(i) / real(order - 1);
  }
  real interp(int i, int d) const { return xrange.lo[d] + (xrange.extent(d) * node01(i)); }

  bool prune(Interval<N> *xint)
  {
    for (int i = 0; i < psiCount;) {
      for (int d = 0; d < N; ++d)
        if (!free_[d]) xint[d].alpha = xrange.side(psi[i].side[d], d);
      const Interval<N> res = eval<N, Interval<N>>(phi, xint);
      if (res.uniformSign()) {
        if ((res.alpha >= 0.0 && psi[i].sign >= 0) || (res.alpha <= 0.0 && psi[i].sign <= 0)) {
          --psiCount;
          std::swap(psi[i], psi[psiCount]);
        } else {
          return false;
        }
      } else {
        ++i;
      }
    }
    return true;
  }
  void clear_cells_map()
  {
    if (T)
      cells_map.clear();
    else
      up->clear_cells_map();
  }
  void clear_ref_intervals()
  {
    ref_intervals.clear();
    if (!T) up->clear_ref_intervals();
  }
  // Tolerance().update(value * extent): the LARGER of 10 eps and 10 eps x extent (tolerance.cpp:47-55)
  real tolerance() const { return std::max(10.0 * EPS, 10.0 * EPS * xrange.extent(k)); }
  void set_psi_bounds(int i, real *x) const
  {
    for (int d = 0; d < N; ++d)
      if (!free_[d]) x[d] = xrange.side(psi[i].side[d], d);
  }
  std::vector<real> compute_roots(const real *point, int i) const
  {
    real x[N];
    for (int d = 0; d < N; ++d) x[d] = point[d];
    set_psi_bounds(i, x);
    std::vector<real> roots;
    root_find<N>(phi, x, k, xrange.lo[k], xrange.hi[k], roots, M > 1);
    return roots;
  }
  bool check_active_interval(real *x, real x0, real x1) const
  {
    if (tol_equal(tolerance(), x1, x0)) return false;
    bool okay = true;
    x[k] = 0.5 * (x0 + x1);
    for (int i = 0; i < psiCount && okay; ++i) {
      set_psi_bounds(i, x);
      okay &= eval<N, real>(phi, x) > 0.0 ? (psi[i].sign >= 0) : (psi[i].sign <= 0);
    }
    return okay;
  }
  void compute_all_intervals(const real *point, RootsIntervals<N> &iv) const
  {
    iv.clear();
    real x[N];
    for (int d = 0; d < N; ++d) x[d] = iv.point[d] = point[d];
    const real x0 = xrange.lo[k], x1 = xrange.hi[k];
    iv.add_root(x0, -1);
    iv.add_root(x1, -1);
    for (int i = 0; i < psiCount; ++i)
      for (real r : compute_roots(x, i)) iv.add_root(r, i);
    iv.adjust_roots(tolerance(), x0, x1);
    if (!S) {
      const int n = iv.num_roots() - 1;
      iv.active.resize((size_t)n);
      for (int i = 0; i < n; ++i) iv.active[(size_t)i] = check_active_interval(x, iv.roots[(size_t)i], iv.roots[(size_t)i + 1]);
    }
  }
  void compute_ref_intervals(const real *point_, TIdx<N> cell_tid)
  {
    if (M == 1 && !T) up->clear_ref_intervals();
    real point[N];
    for (int d = 0; d < N; ++d) point[d] = point_[d];
    RootsIntervals<N> iv;
    compute_all_intervals(point, iv);
    ref_intervals.emplace(cell_tid, iv);
    for (int i = 0; i < (int)iv.active.size(); ++i) {
      if (!iv.active[(size_t)i]) continue;
      cell_tid.v[k] = i;
      point[k] = 0.5 * (iv.roots[(size_t)i] + iv.roots[(size_t)i + 1]);
      if (!T) up->compute_ref_intervals(point, cell_tid);
    }
  }
  void insert_point(const real *point, const TIdx<N> &cell_tid, const int *pt_tid, int npt)
  {
    int cell;
    auto it = cells_map.find(cell_tid);
    if (it != cells_map.end()) {
      cell = it->second;
    } else {
      cell = mesh->allocate_cells(1);
      cells_map.emplace(cell_tid, cell);
    }
    mesh->insert_cell_point(point, cell, (int)ReparamMesh<N>::flat(pt_tid, npt, order));
  }
  void tensor_product_full()
  {
    if (M == N) {
      mesh->add_full_cell(xrange, false);
      return;
    }
    real point[N];
    for (int d = 0; d < N; ++d) point[d] = 0.0;
    int fix = -1;
    for (int d = 0; d < N; ++d)
      if (!free_[d]) {
        fix = d;
        point[d] = xrange.side(psi[0].side[d], d);
        break;
      }
    const TIdx<N> cell_tid;
    int t[3] = { 0, 0, 0 };
    for (t[2] = 0; t[2] < (M > 2 ? order : 1); ++t[2])
      for (t[1] = 0; t[1] < (M > 1 ? order : 1); ++t[1])
        for (t[0] = 0; t[0] < order; ++t[0]) {
          for (int d = 0, l = 0; d < N; ++d)
            if (d != fix) point[d] = interp(t[l++], d);
          insert_point(point, cell_tid, t, M);
        }
  }
  void tensor_product_base()
  {
    real x[N];
    for (int d = 0; d < N; ++d) x[d] = free_[d] ? xrange.midpoint(d) : 0.0;
    up->clear_ref_intervals();
    up->compute_ref_intervals(x, TIdx<N>());
    int t[3] = { 0, 0, 0 };
    for (t[2] = 0; t[2] < (M > 2 ? order : 1); ++t[2])
      for (t[1] = 0; t[1] < (M > 1 ? order : 1); ++t[1])
        for (t[0] = 0; t[0] < order; ++t[0]) {
          real xx[N];
          TIdx<N> pt;
          for (int d = 0, l = 0; d < N; ++d) {
            xx[d] = 0.0;
            if (free_[d]) {
              xx[d] = interp(t[l], d);
              pt.v[d] = t[l];
              ++l;
            }
          }
          up->process(xx, TIdx<N>(), pt);
        }
  }
  // (sic) the root at index psi_id is pushed once per root of that psi (impl_reparam_general.cpp:399-411)
  std::vector<real> ref_roots_of(const RootsIntervals<N> &ref, int i) const
  {
    std::vector<real> r;
    for (int j = 0; j < ref.num_roots(); ++j)
      if (ref.func_ids[(size_t)j] == i) r.push_back(ref.roots[(size_t)i]);
    return r;
  }
  std::vector<real> compute_similar_roots(const RootsIntervals<N> &ref, int i, const real *point_) const
  {
    auto ref_roots = ref_roots_of(ref, i);
    if (ref_roots.empty()) return ref_roots;
    real point[N];
    for (int d = 0; d < N; ++d) point[d] = point_[d];
    auto roots = compute_roots(point, i);
    const real tol = tolerance();
    const real x0 = xrange.lo[k], x1 = xrange.hi[k];
    roots.erase(std::remove_if(roots.begin(), roots.end(), [&](real r) { return tol_equal(tol, r, x0) || tol_equal(tol, r, x1); }),
      roots.end());
    const size_t nref = ref_roots.size();
    const size_t nr = roots.size();
    if (nr == nref) return roots;
    if (nr > nref) return ref_roots;
    set_psi_bounds(i, point);
    real p0[N], p1[N];
    for (int d = 0; d < N; ++d) p0[d] = p1[d] = point[d];
    p0[k] = x0;
    p1[k] = x1;
    const real v0 = eval<N, real>(phi, p0), v1 = eval<N, real>(phi, p1);
    bool r0 = false, r1 = false;
    for (int t = 0; t < 4; ++t) {
      const real nt = tol * std::pow(10.0, t);
      r0 = tol_zero(nt, v0);
      r1 = tol_zero(nt, v1);
      if (r0 || r1) break;
    }
    if (r0 ^ r1) {
      roots.insert(roots.end(), nref - nr, r0 ? x0 : x1);
    } else if (r0 && r1) {
      if (nr + 2 == nref) {
        roots.push_back(x0);
        roots.push_back(x1);
      } else if (nr + 1 == nref) {
        real xx[N];
        for (int d = 0; d < N; ++d) xx[d] = ref.point[d];
        set_psi_bounds(i, xx);
        xx[k] = 0.5 * (x0 + ref_roots.front());
        const bool sign = eval<N, real>(phi, xx) > 0;
        std::sort(roots.begin(), roots.end());
        const real xmid = roots.empty() ? x1 : roots.front();
        point[k] = 0.5 * (x0 + xmid);
        const bool new_sign = eval<N, real>(phi, point) > 0;
        roots.push_back(sign == new_sign ? x1 : x0);
      }
    }
    if (roots.size() == nref) return roots;
    return ref_roots;
  }
  void compute_similar_intervals(const RootsIntervals<N> &ref, const real *point, RootsIntervals<N> &iv) const
  {
    iv.clear();
    for (int d = 0; d < N; ++d) iv.point[d] = point[d];
    if (ref.num_roots() == 2) {
      if (!S) iv = ref;
      return;
    }
    for (int i = 0; i < psiCount; ++i)
      for (real r : compute_similar_roots(ref, i, point)) iv.add_root(r, i);
    const real x0 = xrange.lo[k], x1 = xrange.hi[k];
    if (!S) {
      iv.add_root(x0, -1);
      iv.add_root(x1, -1);
      iv.active = ref.active;
    }
    iv.adjust_roots(tolerance(), x0, x1);
  }
  void process(const real *point_, TIdx<N> cell_tid, TIdx<N> pt_tid)
  {
    real point[N];
    for (int d = 0; d < N; ++d) point[d] = point_[d];
    if (M == 1) compute_ref_intervals(point, cell_tid);
    RootsIntervals<N> iv;
    const RootsIntervals<N> &ref = ref_intervals.at(cell_tid);
    if (M == 1)
      iv = ref;
    else
      compute_similar_intervals(ref, point, iv);
    if (S) {
      for (real r : iv.roots) {
        point[k] = r;
        cell_tid.v[k] = 0;
        int t[N];
        int n = 0;
        for (int d = 0; d < N; ++d)
          if (d != k) t[n++] = pt_tid.v[d];
        insert_point(point, cell_tid, t, N - 1);
      }
      return;
    }
    for (int i = 0; i < (int)iv.active.size(); ++i) {
      if (!iv.active[(size_t)i]) continue;
      const real x0 = iv.roots[(size_t)i], x1 = iv.roots[(size_t)i + 1];
      cell_tid.v[k] = i;
      for (int j = 0; j < order; ++j) {
        point[k] = x0 + (x1 - x0) * node01(j);
        pt_tid.v[k] = j;
        if (T && M < N) {
          int fix = -1;
          for (int d = 0; d < N; ++d)
            if (!free_[d]) fix = d;
          set_psi_bounds(0, point);
          int t[N];
          int n = 0;
          for (int d = 0; d < N; ++d)
            if (d != fix) t[n++] = pt_tid.v[d];
          insert_point(point, cell_tid, t, N - 1);
        } else if (T) {
          insert_point(point, cell_tid, pt_tid.v, N);
        } else {
          up->process(point, cell_tid, pt_tid);
        }
      }
    }
  }

  GeneralReparam(const Func &phi_, GeneralReparam *up_, ReparamMesh<N> *mesh_, int M_, bool S_, bool T_, const bool *free,
    const PsiCode<N> *psi_, int psiCount_, const Box<N> &xrange_, int order_, int level = 0)
    : phi(phi_), up(up_), mesh(mesh_), M(M_), S(S_), T(T_), psiCount(psiCount_), xrange(xrange_), order(order_)
  {
    for (int d = 0; d < N; ++d) free_[d] = free[d];
    for (int i = 0; i < MAXPSI; ++i) psi[i] = psi_[i];
    if (M == 1) {
      for (int d = 0; d < N; ++d)
        if (free_[d]) k = d;
      real x0[N];
      for (int d = 0; d < N; ++d) x0[d] = 0.0;
      process(x0, TIdx<N>(), TIdx<N>());
      return;
    }
    Interval<N> xint[N];
    for (int d = 0; d < N; ++d) {
      if (free_[d]) {
        xint[d] = Interval<N>::variable(xrange.midpoint(d), d);
        Interval<N>::delta(d) = xrange.extent(d) * 0.5;
      } else {
        xint[d] = Interval<N>(0.0);
        Interval<N>::delta(d) = 0.0;
      }
    }
    if (!prune(xint)) return;
    if (psiCount == 0) {
      if (!S) {
        if (T)
          tensor_product_full();
        else
          tensor_product_base();
      }
      return;
    }
    real max_quan = 0.0;
    for (int d = 0; d < N; ++d)
      if (!free_[d]) xint[d].alpha = xrange.side(psi[0].side[d], d);
    Interval<N> g[N];
    grad<N, Interval<N>>(phi, xint, g);
    for (int d = 0; d < N; ++d) {
      if (free_[d] && std::fabs(g[d].alpha) > 1.001 * g[d].maxDeviation()) {
        const real quan = std::fabs(g[d].alpha) * xrange.extent(d);
        if (quan > max_quan) {
          max_quan = quan;
          k = d;
        }
      }
    }
    PsiCode<N> newPsi[MAXPSI];
    for (int i = 0; i < MAXPSI; ++i) {  // unused slots stay as a default-constructed PsiCode would be
      for (int d = 0; d < N; ++d) newPsi[i].side[d] = 0;
      newPsi[i].sign = knobs().psi_default_sign;
    }
    int nNew = 0;
    for (int i = 0; i < psiCount && nNew + 2 <= MAXPSI; ++i) {
      for (int d = 0; d < N; ++d)
        if (!free_[d]) xint[d].alpha = xrange.side(psi[i].side[d], d);
      bool direction_okay = k != -1;
      if (direction_okay) {
        Interval<N> gi[N];
        grad<N, Interval<N>>(phi, xint, gi);
        direction_okay = gi[k].uniformSign();
      }
      if (!direction_okay) {
        if (level < MAX_SUBDIV) {
          real max_ext = 0.0;
          int ind = -1;
          for (int d = 0; d < N; ++d)
            if (free_[d] && xrange.extent(d) > max_ext) {
              max_ext = xrange.extent(d);
              ind = d;
            }
          const real xmid = xrange.midpoint(ind);
          Box<N> b0 = xrange, b1 = xrange;
          b0.hi[ind] = xmid;
          b1.lo[ind] = xmid;
          { GeneralReparam sub(phi, up, mesh, M, S, T, free_, psi, psiCount, b0, order, level + 1); }
          clear_cells_map();
          { GeneralReparam sub(phi, up, mesh, M, S, T, free_, psi, psiCount, b1, order, level + 1); }
          return;
        }
        real xp[N];
        for (int d = 0; d < N; ++d) xp[d] = xrange.midpoint(d);
        bool okay = true;
        for (int j = 0; j < MAXPSI && okay; ++j) {
          for (int d = 0; d < N; ++d)
            if (!free_[d]) xp[d] = xrange.side(psi[j].side[d], d);
          okay &= eval<N, real>(phi, xp) >= 0.0 ? (psi[j].sign >= 0) : (psi[j].sign <= 0);
        }
        if (okay && !S) {
          if (T)
            tensor_product_full();
          else
            tensor_product_base();
        }
        return;
      }
      int bottom, top;
      determine_signs(S, g[k].alpha > 0.0, psi[i].sign, bottom, top);
      newPsi[nNew] = psi[i];
      newPsi[nNew].side[k] = 0;
      newPsi[nNew].sign = bottom;
      ++nNew;
      newPsi[nNew] = psi[i];
      newPsi[nNew].side[k] = 1;
      newPsi[nNew].sign = top;
      ++nNew;
    }
    bool nfree[N];
    for (int d = 0; d < N; ++d) nfree[d] = free_[d];
    nfree[k] = false;
    GeneralReparam sub(phi, this, nullptr, M - 1, false, false, nfree, newPsi, nNew, xrange, order);
  }
};

// impl_reparam_general.cpp:823-850 (reparameterize) and :852-883 (reparameterize_facet); facet < 0: cell
template<int N> void reparam_general(const Func &phi, const Box<N> &box, ReparamMesh<N> &mesh, bool levelset, int facet)
{
  PsiCode<N> psi[1 << (N - 1)];
  for (int i = 0; i < (1 << (N - 1)); ++i) {
    for (int d = 0; d < N; ++d) psi[i].side[d] = 0;
    psi[i].sign = knobs().psi_default_sign;
  }
  psi[0].sign = -1;
  bool free[N];
  for (int d = 0; d < N; ++d) free[d] = true;
  int M = N;
  if (facet >= 0) {
    psi[0].side[facet / 2] = (unsigned char)(facet % 2);
    free[facet / 2] = false;
    M = N - 1;
  }
  const int n0 = (int)mesh.num_cells();
  { GeneralReparam<N> impl(phi, nullptr, &mesh, M, levelset, true, free, psi, 1, box, mesh.order); }
  const int n1 = (int)mesh.num_cells();
  std::vector<int> ids;
  for (int c = n0; c < n1; ++c) ids.push_back(c);
  if (facet >= 0) {
    const int dir = facet / 2, side = facet % 2;
    mesh.orient_levelset_cells_positively(ids, [&](const real *, real *n) {
      for (int d = 0; d < N; ++d) n[d] = 0.0;
      n[dir] = side == 0 ? -1.0 : 1.0;
    });
  } else if (levelset) {
    mesh.orient_levelset_cells_positively(ids, [&](const real *x, real *n) {
      grad<N, real>(phi, x, n);
      real s = 0.0;
      for (int d = 0; d < N; ++d) s += n[d] * n[d];
      s = std::sqrt(s);
      for (int d = 0; d < N; ++d) n[d] /= s;
    });
  } else {
    mesh.orient_cells_positively(ids);
  }
  mesh.generate_wirebasket(phi, ids, box, 1.0e4 * EPS);
}

}  // namespace orc
