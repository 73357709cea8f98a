// qugar_b200 -- launch configuration shared by the engine (which sizes the grids) and the per-kind kernel units
#pragma once

namespace qg {

constexpr int kBlock = 128;       // threads per block of CTL / K0 / K1 / K2 and of the small bookkeeping kernels
constexpr int kExpParents = 64;   // K1 / K2: parents (chain items / stage-1 calls) owned by one block
constexpr int kK3Lines = 256;     // generic K3: stage-2 lines owned by one block
constexpr int kK3Threads = 256;

// the kernels of one pipeline pass, launched through launch_pipe (qg_engine.cu) -> launch_pipe_kind<KIND> (qg_kind.cu)
enum PipeKernel { PK_CTL_COUNT = 0, PK_CTL_EMIT, PK_K0, PK_K1, PK_K2, PK_K3 };

}  // namespace qg
