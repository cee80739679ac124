/* TEST INFRASTRUCTURE (oracle / CPU-baseline build only) -- Cilk runtime API shim over OpenMP. */
#ifndef DPPR_ORACLE_CILK_API_SHIM_OMP_H
#define DPPR_ORACLE_CILK_API_SHIM_OMP_H
#include <omp.h>
#include <cstdlib>
#include <cstring>
static inline int __cilkrts_get_nworkers() { return omp_get_max_threads(); }
static inline int __cilkrts_set_param(const char *name, const char *value) {
    if (std::strcmp(name, "nworkers") == 0) {
        int n = std::atoi(value);
        if (n > 0) omp_set_num_threads(n);
    }
    return 0;
}
#endif
