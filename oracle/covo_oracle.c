/* covo_oracle.c -- C restatement of the two heavy loops of the CoVO-MPC hot path.
 *
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY.  Never linked into or called by the product
 * (covo_mpc_b200); used by oracle/oracle_c.py for `bench.py`'s cpu_baseline / `--impl reference`
 * legs and cross-checked against oracle/oracle_np.py in tests/test_oracle_c.py.
 * Pinned the way oracle_np.py is: tests/test_oracle_c.py holds it to the NumPy oracle, which tests/test_reference_golden.py holds to
 * outputs of the reference's own source executed under a NumPy shim (tests/golden/reference_*.npz); the JAX PRNG / XLA rounding
 * themselves stay unpinned.
 *
 * It follows the REFERENCE's algorithm, not the product's:
 *   oracle_rollout_costs  <-> vmap over N of the lax.scan over H of step_env with reward freeze
 *                             (quadjax/controllers/covo.py:227-263, envs/quadrotor.py:215-263,
 *                              dynamics/free.py:74-155, dynamics/utils.py:266-294, quadrotor.py:479-503)
 *   oracle_hessian_fof    <-> jacfwd(jacfwd(get_cumulated_cost)) (controllers/covo.py:134-185):
 *                             forward-over-forward, one second-order tangent lane (i <= j) at a time
 *                             through the whole unrolled H-step rollout -- all (4H)(4H+1)/2 lanes.
 * OpenMP over samples / tangent lanes (the reference's XLA:CPU runs on all host cores too).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* float32 instance: oracle_rollout_costs, oracle_hessian_fof (the reference's arithmetic) */
#define REAL float
#define N_(x) x
#define K(x) x##f
#define M_(fn) fn##f
#include "covo_oracle_impl.h"
#undef REAL
#undef N_
#undef K
#undef M_
/* float64 instance: oracle_rollout_costs_f64, oracle_hessian_fof_f64 (same inputs, float64 arithmetic and outputs) */
#define REAL double
#define N_(x) x##_f64
#define K(x) x
#define M_(fn) fn
#include "covo_oracle_impl.h"
#undef REAL
#undef N_
#undef K
#undef M_

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
