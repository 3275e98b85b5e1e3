// tile_sweep_large.cu -- instantiates the tile-sweep kernel (tile_sweep.cuh) for rq_program_large, linear tile layout.
#include "tile_sweep.cuh"
extern "C" int rq_sweep_configure_large_lin(void) { return configure<rq_program_large, false>(); }
extern "C" int rq_launch_sweep_large_lin(rq_cplx* state, const rq_program_large* prog, void* stream) { return launch<rq_program_large, false>(state, prog, stream); }
