/* TEST INFRASTRUCTURE (oracle build only) -- Cilk Plus shim, serial elision.
 * The reference (cpu/CilkUtil.h:4-7) includes <cilk/cilk.h> and uses cilk_for;
 * Cilk Plus was removed from GCC 8+.  Cilk's defined serial elision is
 * `cilk_for` == `for`, which makes the reference deterministic. */
#ifndef DPPR_ORACLE_CILK_SHIM_SERIAL_H
#define DPPR_ORACLE_CILK_SHIM_SERIAL_H
#define cilk_for for
#define cilk_spawn
#define cilk_sync
#endif
