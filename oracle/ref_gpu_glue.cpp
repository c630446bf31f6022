// TEST INFRASTRUCTURE (oracle) — symbols ref_harness.cpp expects from the solver side when the
// UNMODIFIED reference cuda.cu is linked instead of ref_spmat_cpu.cpp.  The real solver has no
// iteration-cap knob (cuda.cu:438 hard-codes 1000 iterations; with tolerance 0 and max_restarts 0 it
// runs exactly that many) and reports nothing, so force_iters is ignored and fixed_iters says 1000.
int cup2d_ref_last_iters = -1;
double cup2d_ref_last_err = -1;
int cup2d_ref_force_iters = -1;
int cup2d_ref_fixed_iters = 1000;
