// tile_sweep_small_swz.cu -- instantiates the tile-sweep kernel (tile_sweep.cuh) for rq_program_small, XOR-swizzled tile layout.
#include "tile_sweep.cuh"
extern "C" int rq_sweep_configure_small_swz(void) { return configure<rq_program_small, true>(); }
extern "C" int rq_launch_sweep_small_swz(rq_cplx* state, const rq_program_small* prog, void* stream) { return launch<rq_program_small, true>(state, prog, stream); }
