// TEST INFRASTRUCTURE -- host build of the product's correctly rounded pow / exp (fractulus_b200/csrc/frx_math.cuh: every
// operation is an explicitly rounded IEEE op, so the host build gives the bits of the sm_100a build) for the
// -DORC_CR_POW variant of oracle/frx_oracle.c.  Compile with -ffp-contract=off (oracle/Makefile).
#include "../fractulus_b200/csrc/frx_math.cuh"
extern "C" double orc_cr_pow(double x, double y) { return frx_pow_cr(x, y); }
extern "C" double orc_cr_exp(double x) { return frx_exp_cr(x); }
