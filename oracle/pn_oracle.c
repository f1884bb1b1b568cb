/*
 * pn_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU oracle for the PointNeighbors.jl hot path (GridNeighborhoodSearch + FullGridCellList ->
 * initialize!/update! -> foreach_point_neighbor, PrecomputedNeighborhoodSearch, PeriodicBox).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (pointneighbors.jl_b200/) never does.
 *
 * The reference is pure Julia and there is no `julia` binary in this image, so
 * there is no oracle/_ref build: the restatement is pinned against the golden vectors the
 * reference's own tests hold (tests/golden/, tests/test_oracle_golden.py).
 *
 * Build:  make -C oracle      (gcc -O2 -fopenmp -ffp-contract=off, never -ffast-math)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define REAL float
#define SUF _f32
#define REAL_IS_FLOAT 1
#include "pn_oracle_impl.h"
#undef REAL
#undef SUF
#undef REAL_IS_FLOAT

#define REAL double
#define SUF _f64
#define REAL_IS_FLOAT 0
#include "pn_oracle_impl.h"
#undef REAL
#undef SUF
#undef REAL_IS_FLOAT

#include "pn_oracle_mixed.h"

#ifdef _OPENMP
#include <omp.h>
int pno_max_threads(void) { return omp_get_max_threads(); }
void pno_set_threads(int n) { omp_set_num_threads(n); }
#else
int pno_max_threads(void) { return 1; }
void pno_set_threads(int n) { (void)n; }
#endif
