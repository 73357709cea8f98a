// qugar_b200 -- one translation unit per specialised level-set kind: nvcc -DQG_KIND=<kind or -1> -DQG_KIND_TAG=<name>.
// Exports qg::launch_pipe_<tag>, the launcher of the CTL / K0 / K1 / K2 / K3 kernels instantiated for that kind
// (-1 keeps the run-time switch and serves every kind without a unit of its own).
#include "qg_kernels_kind.cuh"

#if !defined(QG_KIND) || !defined(QG_KIND_TAG)
#error "compile with -DQG_KIND=<int> -DQG_KIND_TAG=<identifier>"
#endif
#define QG_CAT2(a, b) a##b
#define QG_CAT(a, b) QG_CAT2(a, b)

namespace qg {
void QG_CAT(launch_pipe_, QG_KIND_TAG)(int which, int dim, unsigned grid, cudaStream_t s, const PipeArgs &a, const long long *shift)
{
  launch_pipe_kind<QG_KIND>(which, dim, grid, s, a, shift);
}
void QG_CAT(launch_rp_, QG_KIND_TAG)(int emit, int dim, unsigned grid, cudaStream_t s, const RpLeafArgs &a)
{
  launch_rp_kind<QG_KIND>(emit, dim, grid, s, a);
}
}  // namespace qg
