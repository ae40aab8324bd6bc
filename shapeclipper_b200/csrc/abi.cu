// abi.cu — ABI bookkeeping for libsc_b200.so (see include/sc_b200.h).
#include "sc_b200.h"

extern "C" int sc_abi_version(void) { return SC_B200_ABI_VERSION; }
