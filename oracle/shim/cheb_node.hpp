// PVFMM stand-in (test infrastructure only): forwards to the single shim header.
#include "pvfmm_shim.hpp"
