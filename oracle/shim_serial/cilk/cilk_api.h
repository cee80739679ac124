/* TEST INFRASTRUCTURE (oracle build only) -- Cilk runtime API shim, serial. */
#ifndef DPPR_ORACLE_CILK_API_SHIM_SERIAL_H
#define DPPR_ORACLE_CILK_API_SHIM_SERIAL_H
static inline int __cilkrts_get_nworkers() { return 1; }
static inline int __cilkrts_set_param(const char *, const char *) { return 0; }
#endif
