// Overlay for the drop-in demonstration (oracle/Makefile, target _ref/run_tests_b200): placed FIRST on the include path
// so that the reference's own, unmodified sources (src/app, src/methods, src/matrix_ops, src/planning, src/timing)
// resolve <gpu-api.h> to the alias table of this back end instead of src/gpu-api/gpu-api.h (the cuBLAS one).
// Nothing else of the reference is shadowed.
#pragma once
#include "../../rtatblas_b200/include/rtat/gpu-api.h"
