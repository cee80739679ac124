/* TEST INFRASTRUCTURE (oracle / CPU-baseline build only) -- Cilk Plus shim backed
 * by OpenMP.  `cilk_for` becomes an OpenMP work-shared loop; nested cilk_for
 * (cpu/PPRCPUMTCilkRev.h:239) becomes a nested (inactive) parallel region, i.e.
 * runs serially inside the outer worker -- NOT Cilk work stealing.  Every number
 * produced with this build is labelled "OpenMP-backed cilk_for". */
#ifndef DPPR_ORACLE_CILK_SHIM_OMP_H
#define DPPR_ORACLE_CILK_SHIM_OMP_H
#include <omp.h>
#define DPPR_PRAGMA(x) _Pragma(#x)
#define cilk_for DPPR_PRAGMA(omp parallel for schedule(dynamic, 64)) for
#define cilk_spawn
#define cilk_sync
#endif
