/* qugar_b200 -- C ABI of the B200-native cut-cell quadrature engine.
 *
 * Drop-in boundary for QUGaR's implicit-domain hot path.  QUGaR itself has no C ABI: its Python module
 * `qugar.cpp` is a set of nanobind wrappers over C++ classes.  Each entry point below names the
 * reference interface it replaces (paths relative to /root/reference); INTEGRATION.md shows the
 * nanobind-side stub that forwards to it.
 *
 * Conventions: every pointer is HOST memory; every function returns 0 on success and a negative
 * qg_status otherwise (never throws, never aborts); qg_last_error() gives the message of the last
 * failure on the calling thread.  Objects are immutable after creation and may be shared between
 * threads; calls that launch GPU work serialise on the domain's stream.  There is no CPU fallback:
 * without a usable CUDA device every compute call fails with QG_ERR_NO_DEVICE.
 */
#ifndef QUGAR_B200_H
#define QUGAR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qg_grid qg_grid;
typedef struct qg_func qg_func;
typedef struct qg_domain qg_domain;
typedef struct qg_rule qg_rule;

typedef enum {
  QG_OK = 0,
  QG_ERR_INVALID = -1,   /* bad argument (the reference would assert / UB) */
  QG_ERR_NO_DEVICE = -2, /* CUDA device or driver missing */
  QG_ERR_CUDA = -3,      /* CUDA runtime failure */
  QG_ERR_CAPACITY = -4,  /* a fixed on-device capacity was exceeded (roots per line, recursion stack) */
  QG_ERR_UNSUPPORTED = -5
} qg_status;

/* Level-set kinds.  params layout:
 *   QG_SPHERE (disk in 2-D): radius, center[dim]        primitive_funcs_lib.cpp:59-74, impl_functions.cpp:123-195 (use_bzr=False)
 *   QG_PLANE  (line in 2-D): origin[dim], normal[dim]   primitive_funcs_lib.cpp:842-857
 *   QG_BEZIER (BezierTP<dim,1>, bezier_tp.hpp:41-285): order[dim] (integers stored as doubles, 1..8) followed by the
 *             prod(order) Bernstein coefficients on [0,1]^dim, direction 0 fastest.  Evaluated by de Casteljau on
 *             double and Interval (bezier_tp.cpp:165-255) and integrated with the GENERAL algorithm; the reference
 *             routes BezierTP to algoim::ImplicitPolyQuadrature instead.
 *   QG_CONSTANT: value;  QG_CYLINDER (3-D): radius, origin[3], unit axis[3];  QG_ELLIPSOID (ellipse in 2-D): semi_axes[dim],
 *             origin[dim], orthonormal basis[dim][dim] (row d = axis d);  QG_ANNULUS (2-D): inner radius, outer radius,
 *             center[2];  QG_TORUS (3-D): major radius, minor radius, center[3], unit axis[3]
 *             (primitive_funcs_lib.cpp:772-780, 211-225, 327-350, 484-498, 705-729; use_bzr=False variants)
 *   QG_DIM_LINEAR: the 2^dim vertex values of a bi-/tri-linear function on [0,1]^dim, direction 0 fastest (impl_funcs_lib.cpp:96-152)
 *   QG_SQUARE: no parameters; max_d |x_d| - 1, the box [-1,1]^dim (Square<dim>, impl_funcs_lib.cpp:44-98); placed elsewhere by
 *             the affine map of a QG_COMPOSITE leaf, as Square(transf) does (impl_funcs_lib_macros.hpp:172-199)
 *   QG_COMPOSITE: TransformedFunction / AddFunctions / SubtractFunctions (impl_funcs_lib.cpp:168-290) over one or two
 *             non-composite leaves: op (0 single, 1 add, 2 subtract), nleaf, then per leaf: kind, negative, has_aff,
 *             aff[12] (the INVERSE affine map the reference stores: dim*dim matrix row major, then translation; unused
 *             slots 0), np, and the leaf's own np parameters
 *   TPMS, 3-D: periods m,n,q; 2-D: periods m,n then z   tpms_lib.cpp, impl_functions.cpp:836-900 */
typedef enum {
  QG_SPHERE = 0,
  QG_PLANE = 1,
  QG_BEZIER = 2,
  QG_CONSTANT = 3,
  QG_CYLINDER = 4,
  QG_ELLIPSOID = 5,
  QG_ANNULUS = 6,
  QG_TORUS = 7,
  QG_COMPOSITE = 8,
  QG_DIM_LINEAR = 9,
  QG_SCHOEN = 10,
  QG_SCHOEN_IWP = 11,
  QG_SCHOEN_FRD = 12,
  QG_FISCHER_KOCH_S = 13,
  QG_SCHWARZ_DIAMOND = 14,
  QG_SCHWARZ_PRIMITIVE = 15,
  QG_SQUARE = 16
} qg_func_kind;

/* ImmersedCellStatus (cpp/include/qugar/unfitted_domain_binary_part.hpp:34-39) */
typedef enum { QG_CELL_CUT = 0, QG_CELL_FULL = 1, QG_CELL_EMPTY = 2 } qg_cell_status;

/* ImmersedFacetStatus (cpp/include/qugar/unfitted_domain.hpp:36-48), same order */
typedef enum {
  QG_FACET_CUT = 0,
  QG_FACET_FULL,
  QG_FACET_EMPTY,
  QG_FACET_CUT_UNF_BDRY,
  QG_FACET_CUT_EXT_BDRY,
  QG_FACET_CUT_UNF_BDRY_EXT_BDRY,
  QG_FACET_FULL_UNF_BDRY,
  QG_FACET_FULL_EXT_BDRY,
  QG_FACET_UNF_BDRY,
  QG_FACET_EXT_BDRY,
  QG_FACET_UNF_BDRY_EXT_BDRY
} qg_facet_status;

/* Facet queries of UnfittedDomain (cpp/qugar/unfitted_domain.cpp:119-246) */
typedef enum {
  QG_FACETS_EMPTY = 0,
  QG_FACETS_FULL = 1,
  QG_FACETS_CUT = 2,
  QG_FACETS_FULL_UNF_BDRY = 3,
  QG_FACETS_UNF_BDRY = 4
} qg_facet_query;

const char *qg_last_error(void);
/* number of usable CUDA devices (0 when none); never fails */
int qg_device_count(void);
const char *qg_version(void);

/* CartGridTP<dim>(breaks)  -- cpp/qugar/cart_grid_tp.cpp:79-111; python create_cart_grid, common.cpp:160-187.
 * breaks[d] has n_breaks[d] strictly increasing values; they are used as given. */
int qg_grid_create(int dim, const double *const *breaks, const int64_t *n_breaks, qg_grid **out);
void qg_grid_free(qg_grid *);
int qg_grid_num_cells_dir(const qg_grid *, int64_t n_cells[3]);

/* ImplicitFunc<dim> factories -- python/nanobind_wrappers/impl_functions.cpp:123-195, 836-900;
 * create_negative, impl_functions.cpp:718-735 (negative != 0 wraps the function in Negative<dim>). */
int qg_func_create(int kind, int dim, const double *params, int n_params, int negative, qg_func **out);
void qg_func_free(qg_func *);
/* ImplicitFunc::eval / grad on n points (impl_functions.cpp:49-99); grad may be NULL. Runs on the GPU. */
int qg_func_eval(const qg_func *, const double *points, int64_t n, double *values, double *grad);

/* UnfittedImplDomain<dim>(phi, grid) -- cpp/qugar/impl_unfitted_domain.cpp:415-513;
 * python create_unfitted_impl_domain, unf_domain.cpp:264-305.  device: CUDA ordinal. */
int qg_domain_create(const qg_func *, const qg_grid *, int device, qg_domain **out);
/* Same, classifying only the cells whose index in the LAST direction lies in [k_begin, k_end) (one rank's
 * slab when a grid is sharded over GPUs; k_end < 0 means "to the end").  The kd-tree of the whole grid is
 * walked, as in the reference's target-cells path (impl_unfitted_domain.cpp:447-450), so cells get exactly
 * the status they have in the full domain; cells outside the slab stay unclassified (status 255). */
int qg_domain_create_slab(const qg_func *, const qg_grid *, int device, int64_t k_begin, int64_t k_end, qg_domain **out);
void qg_domain_free(qg_domain *);
/* counts[0..2] = full, empty, cut cells; counts[3] = cells that keep facet records */
int qg_domain_counts(const qg_domain *, int64_t counts[4]);
/* get_{full,empty,cut}_cells (unfitted_domain.cpp:82-117): sorted ids; ids must hold the matching count */
int qg_domain_cells(const qg_domain *, int status, int64_t *ids);
/* status of every cell, num_total_cells bytes (qg_cell_status) */
int qg_domain_cell_status(const qg_domain *, uint8_t *status);
/* facet records (facets_status_, unfitted_domain.hpp:150): cells sorted, 2*dim statuses per cell */
int qg_domain_facet_records(const qg_domain *, int64_t *cells, uint8_t *statuses);
/* get_*_facets without targets (unfitted_domain.cpp:119-246): first call with cells == NULL to get *n,
 * then with buffers of that size.  Output sorted by (cell, local facet). */
int qg_domain_facets(const qg_domain *, int query, int64_t *cells, int32_t *facets, int64_t *n);

/* create_quadrature<dim> -- cpp/qugar/cut_quadrature.cpp:107-138, impl_quadrature.cpp:493-537;
 * python create_quadrature, cut_quad.cpp:179-189.  One entry per input cell, in input order. */
int qg_quad_cells(const qg_domain *, const int64_t *cells, int64_t n_cells, int n_pts_dir, qg_rule **out);
/* create_unfitted_bound_quadrature<dim> -- cut_quadrature.cpp:140-175, impl_quadrature.cpp:540-592;
 * python create_unfitted_bound_quadrature, cut_quad.cpp:196-213. */
int qg_quad_unf_bdry(const qg_domain *, const int64_t *cells, int64_t n_cells, int n_pts_dir,
  int include_facet_unf_bdry, int exclude_ext_bdry, qg_rule **out);
/* create_facets_quadrature_{interior,exterior}_integral<dim> -- cut_quadrature.cpp:178-262,
 * impl_quadrature.cpp:595-649; python cut_quad.cpp:222-258. */
int qg_quad_facets(const qg_domain *, const int64_t *cells, const int32_t *facets, int64_t n, int n_pts_dir,
  int exterior, qg_rule **out);

/* CutCellsQuad / CutUnfBoundsQuad / CutIsoBoundsQuad views (cut_quadrature.hpp:36-83): pinned host arrays
 * owned by the rule (valid until qg_rule_free).  point_dim = dim, or dim-1 for facet rules.
 * normals is NULL unless the rule is an unfitted-boundary rule. */
int qg_rule_view(const qg_rule *, int64_t *n_entities, int64_t *n_points, int *point_dim,
  const int32_t **n_pts_per_entity, const double **points, const double **weights, const double **normals);
void qg_rule_free(qg_rule *);

/* Device-resident variants used by the benchmark and by callers that keep the rule on the GPU:
 * the same computation as qg_quad_cells / qg_quad_unf_bdry on cells already uploaded with
 * qg_domain_upload_cells; nothing is copied back except the totals.  d_points/d_weights/d_normals
 * receive DEVICE pointers owned by the rule.  kernel_ms (may be NULL) receives per-pass device times. */
typedef struct qg_dcells qg_dcells;
int qg_domain_upload_cells(const qg_domain *, const int64_t *cells, int64_t n_cells, qg_dcells **out);
void qg_dcells_free(qg_dcells *);
/* the domain's cut cells (ascending ids) as a device-resident list */
int qg_domain_cut_cells_device(const qg_domain *, qg_dcells **out);
int qg_dcells_count(const qg_dcells *, int64_t *n);
int qg_quad_cells_device(const qg_domain *, const qg_dcells *, int n_pts_dir, qg_rule **out);
int qg_quad_unf_bdry_device(const qg_domain *, const qg_dcells *, int n_pts_dir, qg_rule **out);
int qg_rule_device_view(const qg_rule *, int64_t *n_entities, int64_t *n_points, const int32_t **d_n_pts_per_entity,
  const double **d_points, const double **d_weights, const double **d_normals);
/* per-pass device times of the call that produced the rule: ctl, k0, k1, k2, k3, scans, other (ms) */
int qg_rule_pass_ms(const qg_rule *, double ms[8]);
/* number of CUDA kernels this library has launched in the calling process so far (bench.py: gpu_launches) */
long long qg_launch_count(void);
/* bytes the download of this rule moved from device to host (0 before the download).  3-D interior rules of the host API
 * travel in a transport encoding (per group of n_pts_dir nodes: the two fixed reference coordinates once; per node the
 * coordinate along the line and the weight) and are laid out by host threads: 0.6 of the plain size. */
long long qg_rule_d2h_bytes(const qg_rule *);
/* device-to-device copy on `device` (lets a consumer, e.g. a torch tensor, take a copy of a device-resident rule array).
 * Ordering contract: the copy is queued on the engine's stream of that device and the call returns when it has completed,
 * so it sees every rule produced by earlier calls and is finished before any later call (or qg_rule_free) can reuse the
 * source.  Work the CALLER has in flight on dst / src on another stream is not ordered: synchronise that stream first. */
int qg_memcpy_device(void *dst, const void *src, int64_t bytes, int device);
/* work units of the call that produced the rule: chain items, stage-1 calls, stage-2 lines, nodes before purge */
int qg_rule_pass_counts(const qg_rule *, long long counts[4]);
/* Custom-coefficient packing on the device: the consumer of the rules in the reference's FEniCSx layer
 * (python/qugar/dolfinx/custom_coefficients.py).  Replaces, for cell and unfitted-boundary integrals with real
 * coefficients, _allocate_new_coeffs (:278-312), _compute_n_vals_per_custom_entity (:314-368), _compute_offsets (:370-449),
 * _compute_new_vals (:730-775) and _copy_vals (:451-512): the result is the array `new_coeffs`
 *     rows 0 .. n_rows-1 : the old coefficients | ptrdiff_t offset in the trailing column(s): -1 standard entity,
 *                          0 empty entity, else position of the entity's block relative to the first slot of its row
 *     appended rows      : per custom entity, per quadrature: n_pts | points | weights | normals (unfitted boundary)
 * elem_size 8 (float64) or 4 (float32, two trailing columns per offset).  rules[q]: device-resident rules (qg_quad_*_device)
 * with one entity per custom id, in the order of the integrand's quadratures.  old_coeffs, custom_ids, empty_ids: host.
 * to_host != 0 also copies the packed array to pinned host memory -- the only bulk transfer of the whole path. */
typedef struct qg_coeffs qg_coeffs;
int qg_pack_custom_coeffs(int device, int elem_size, const void *old_coeffs, int64_t n_rows, int64_t n_old_cols,
                          const int64_t *custom_ids, int64_t n_custom, const int64_t *empty_ids, int64_t n_empty,
                          const qg_rule *const *rules, int n_rules, int to_host, qg_coeffs **out);
int qg_coeffs_view(const qg_coeffs *, int64_t *n_rows, int64_t *n_cols, int *elem_size, const void **host, const void **dev,
                   int64_t *d2h_bytes);
void qg_coeffs_free(qg_coeffs *);

/* BezierTP<dim,1>::rescale_domain + sign for a batch of boxes (cpp/qugar/bezier_tp.cpp:403-431; the per-cell copy the
 * reference makes at impl_quadrature.cpp:165-168 and impl_unfitted_domain.cpp:351-360).  boxes: n_boxes x (lo[dim], hi[dim]),
 * host; coefs_out: n_boxes x n_coefs (direction 0 fastest, like the function's own coefficients), sign_out: +1 / -1 / 0 =
 * algoim::bernstein::uniformSign of the restricted coefficients.  The function must be a QG_BEZIER with <= 216 coefficients. */
int qg_bezier_rescale(const qg_func *bezier, const double *boxes, int64_t n_boxes, int device, double *coefs_out, int8_t *sign_out);

/* Memory pools.  Work buffers come from a per-device caching allocator and host copies from a pinned pool; neither
 * shrinks on its own.  These give the cached (currently unused) blocks back to the driver, e.g. before handing the GPU
 * to another library in the same process.  bytes_released may be NULL.  (No reference counterpart: QUGaR's containers
 * are std::vector, cut_quadrature.hpp:36-83.) */
int qg_cache_release(int device, int64_t *bytes_released);
int qg_pinned_release(int64_t *bytes_released);

/* copy a device-resident rule to its pinned host arrays (done implicitly by the host variants) */
int qg_rule_download(qg_rule *);

/* CUDA-event timer on the engine's stream of a device (used by bench.py to time device-resident steps) */
typedef struct qg_timer qg_timer;
int qg_timer_create(int device, qg_timer **out);
int qg_timer_start(qg_timer *);
int qg_timer_stop(qg_timer *, double *ms);
void qg_timer_free(qg_timer *);

/* Gauss-Legendre rule on [0,1]^dim, direction 0 fastest (Quadrature<dim>::create_Gauss_01,
 * cpp/qugar/quadrature.cpp:118-133,199-277; python create_Gauss_quad_01, quad.cpp:57-60).
 * points: n_pts_dir^dim x dim, weights: n_pts_dir^dim.  Host-only table lookup. */
int qg_gauss_01(int dim, int n_pts_dir, double *points, double *weights);

/* ---- reparameterization ----------------------------------------------------------------------------
 * Lagrange cells of order^param_dim points over the cut (and, for the volume, full) cells of `cells`:
 * create_reparameterization<dim, levelset>(unf_domain, cells, n_pts_dir, merge_points) (cpp/qugar/impl_reparam.cpp:46-113,
 * python create_reparameterization / create_reparameterization_levelset, python/nanobind_wrappers/reparam.cpp:97-155) for
 * general (non-Bezier) level sets: reparam_general / reparam_general_facet (cpp/qugar/impl_reparam_general.cpp:823-935),
 * ImplReparamMesh orientation and wirebasket (impl_reparam_mesh.cpp:100-330), ReparamMesh::merge_coincident_points
 * (reparam_mesh.cpp:607-735).
 *   mode   0: volume {phi < 0} (param_dim = dim); 1: zero level set (param_dim = dim - 1); 2 + f: face f of each cell
 *   flags  bit 0: merge coincident points with tolerance merge_tol (the reference uses 1e4 eps; its merge picks the
 *          representative of a coincident group with an unstable std::sort, here the sorts are stable);
 *          bit 1: reparameterize every listed cell whatever its status and add no full cells (= reparam_general on a box)
 * The mesh lives in HBM; qg_mesh_view downloads it once (points n_points x dim, connectivity n_cells x order^param_dim,
 * wirebasket n_wire_edges x order; ids are 0-based point indices, direction 0 fastest inside a cell). */
typedef struct qg_mesh qg_mesh;
int qg_reparam(const qg_domain *, const int64_t *cells, int64_t n_cells, int order, int mode, int flags, double merge_tol,
               qg_mesh **out);
int qg_mesh_view(qg_mesh *, int *dim, int *param_dim, int *order, int64_t *n_points, int64_t *n_cells, int64_t *n_wire_edges,
                 const double **points, const int64_t **connectivity, const int64_t **wires_connectivity);
int qg_mesh_device_view(const qg_mesh *, int64_t *n_points, int64_t *n_cells, int64_t *n_wire_edges, const double **d_points,
                        const int64_t **d_connectivity, const int64_t **d_wires_connectivity);
void qg_mesh_free(qg_mesh *);

#ifdef __cplusplus
}
#endif
#endif
