// tile_sweep_small.cu -- instantiates the tile-sweep kernel (tile_sweep.cuh) for rq_program_small, linear tile layout.
#include "tile_sweep.cuh"
extern "C" int rq_sweep_configure_small_lin(void) { return configure<rq_program_small, false>(); }
extern "C" int rq_launch_sweep_small_lin(rq_cplx* state, const rq_program_small* prog, void* stream) { return launch<rq_program_small, false>(state, prog, stream); }
