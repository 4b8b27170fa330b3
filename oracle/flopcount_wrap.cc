// Counting build of the CPU oracle (test tooling): the same C source compiled as C++ with `double` replaced by a
// wrapper whose operators count every add / mul / div / sqrt the algorithm executes.  tools/count_flops.py builds it
// into oracle/_count/ and measures the algorithmic FLOP per env-step that bench.py's roofline numerator uses.
#include "flopcount.h"
extern "C" {
long long o_flops[8];
#include "mjc_oracle.c"
}
