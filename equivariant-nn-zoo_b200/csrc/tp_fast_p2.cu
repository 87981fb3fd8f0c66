// Part 2 of the GENERATED tensor-product convolution kernels (tp_generated.cuh spreads its structures over
// N_PARTS translation units so that they compile in parallel; part 0 and the table live in tp_fast.cu).
#include "tp_fast.h"

#define E3B_TP_PART 2
#include "tp_generated.cuh"
