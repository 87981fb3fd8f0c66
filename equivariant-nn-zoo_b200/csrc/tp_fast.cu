// Translation unit of the GENERATED tensor-product convolution kernels (tp_generated.cuh).
#include "tp_fast.h"

#include <cstdlib>

// E3B_TP_PIPELINED=0 selects the direct-load kernels (A/B comparison and bisection)
bool e3b_tp_pipelined_enabled() {
  static const int on = [] { const char* v = getenv("E3B_TP_PIPELINED"); return (v && v[0] == '0') ? 0 : 1; }();
  return on != 0;
}

#include "tp_generated.cuh"

const GenEntry* e3b_find_generated(const e3b_tp_desc* d, const int32_t* y_base, const int32_t* y_kstride) {
  for (int e = 0; e < kNumGenEntries; ++e) {
    const GenEntry& g = kGenEntries[e];
    if (g.n_in != d->n_in || g.n_sh != d->n_sh || g.n_paths != d->n_paths) continue;
    bool ok = true;
    for (int b = 0; b < g.n_in && ok; ++b) ok = g.in_l[b] == d->in_l[b];
    for (int s = 0; s < g.n_sh && ok; ++s) ok = g.sh_l[s] == d->sh_l[s];
    for (int q = 0; q < g.n_paths && ok; ++q)
      ok = g.path_in[q] == d->path_in[q] && g.path_sh[q] == d->path_sh[q] && g.path_lout[q] == d->path_lout[q] &&
           g.path_slot[q] == d->path_slot[q] && g.path_ybase[q] == y_base[q] && g.path_ykstride[q] == y_kstride[q];
    if (ok) return &g;
  }
  return nullptr;
}
