/* oracle/vi_oracle.c -- TEST INFRASTRUCTURE.  Placeholder translation unit for the value-iteration
 * oracle (reference pomdpy/solvers/value_iteration.py:24-181); filled in with the VI row. */
int orc_vi_placeholder(void) { return 0; }
