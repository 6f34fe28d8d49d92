/* Minimal htslib stand-in for building the oracle (test infrastructure only). */
#include "shim_all.h"
