// tile_sweep_large_swz.cu -- instantiates the tile-sweep kernel (tile_sweep.cuh) for rq_program_large, XOR-swizzled tile layout.
#include "tile_sweep.cuh"
extern "C" int rq_sweep_configure_large_swz(void) { return configure<rq_program_large, true>(); }
extern "C" int rq_launch_sweep_large_swz(rq_cplx* state, const rq_program_large* prog, void* stream) { return launch<rq_program_large, true>(state, prog, stream); }
