// tile_sweep.cu -- dispatch of a sweep program to the instantiation that matches its size and tile layout.
// The kernel itself is in tile_sweep.cuh; it is compiled in four translation units so that the build parallelises.
#include "sv_internal.h"

extern "C" {
int rq_sweep_configure_small_lin(void);
int rq_sweep_configure_small_swz(void);
int rq_sweep_configure_large_lin(void);
int rq_sweep_configure_large_swz(void);
int rq_launch_sweep_small_lin(rq_cplx*, const rq_program_small*, void*);
int rq_launch_sweep_small_swz(rq_cplx*, const rq_program_small*, void*);
int rq_launch_sweep_large_lin(rq_cplx*, const rq_program_large*, void*);
int rq_launch_sweep_large_swz(rq_cplx*, const rq_program_large*, void*);

int rq_sweep_configure(void) {
    int e = rq_sweep_configure_small_lin();
    if (!e) e = rq_sweep_configure_small_swz();
    if (!e) e = rq_sweep_configure_large_lin();
    if (!e) e = rq_sweep_configure_large_swz();
    return e;
}
int rq_launch_sweep_small(rq_cplx* state, const rq_program_small* prog, void* stream) {
    return prog->hdr.swz ? rq_launch_sweep_small_swz(state, prog, stream) : rq_launch_sweep_small_lin(state, prog, stream);
}
int rq_launch_sweep_large(rq_cplx* state, const rq_program_large* prog, void* stream) {
    return prog->hdr.swz ? rq_launch_sweep_large_swz(state, prog, stream) : rq_launch_sweep_large_lin(state, prog, stream);
}
}
