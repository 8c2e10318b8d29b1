// ORACLE — TEST INFRASTRUCTURE ONLY.
// Thin extern "C" wrapper around the VERBATIM reference L-BFGS
// (/root/reference/src/plan_manage/include/se3gcopter/lbfgs.hpp — libc only, compiles stand-alone).
// Built by oracle/Makefile into oracle/_ref/libref_lbfgs.so from the sources where they lie; the
// reference source is never copied into this repository.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "lbfgs.hpp"  // -I /root/reference/src/plan_manage/include/se3gcopter

extern "C" {

// Drive an arbitrary lbfgs_evaluate_t evaluator with the reference's own lbfgs_optimize.
int ref_lbfgs_optimize(int n, double* x, double* fx, lbfgs::lbfgs_evaluate_t eval, void* instance, int mem_size,
                       int past, double delta, double g_epsilon, double min_step, int max_iterations) {
    lbfgs::lbfgs_parameter_t p;
    lbfgs::lbfgs_load_default_parameters(&p);
    p.mem_size = mem_size;
    p.past = past;
    p.delta = delta;
    p.g_epsilon = g_epsilon;
    p.min_step = min_step;
    p.max_iterations = max_iterations;
    return lbfgs::lbfgs_optimize(n, x, fx, eval, nullptr, nullptr, instance, &p);
}

}  // extern "C"
