// Interface between the ABI translation unit and the generated-kernel translation unit.
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"

struct GenEntry {
  int n_in; int in_l[E3B_MAX_BLOCKS]; int n_sh; int sh_l[E3B_MAX_BLOCKS]; int n_paths;
  int path_in[E3B_MAX_PATHS], path_sh[E3B_MAX_PATHS], path_lout[E3B_MAX_PATHS], path_slot[E3B_MAX_PATHS];
  int path_ybase[E3B_MAX_PATHS], path_ykstride[E3B_MAX_PATHS];  // output layout the kernel was generated for
  int n_groups;
  void (*fwd)(const TpArgs<float>&, int64_t grid, cudaStream_t);
  void (*bwd)(const TpArgs<float>&, int64_t grid, cudaStream_t);
  int paired_bwd_parts;   // d/dY partial-sum rows per edge the paired (multiplicity 64) backward kernel writes
  bool paired_bwd_ok;     // the generator's default choice for that kernel
};

// d/dY partial-sum rows per edge ([E, n_part, sh_dim]) the backward launcher of `g` writes at multiplicity `mul`
int e3b_gen_bwd_parts(const GenEntry* g, int mul);

// the generated kernel whose structure equals `d` (ignoring mul and parities), or nullptr
const GenEntry* e3b_find_generated(const e3b_tp_desc* d, const int32_t* y_base, const int32_t* y_kstride);
