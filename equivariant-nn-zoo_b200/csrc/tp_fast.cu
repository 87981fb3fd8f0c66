// Translation unit of the GENERATED tensor-product convolution kernels (tp_generated.cuh).
#include "tp_fast.h"

#include <cstdlib>

// E3B_TP_PIPELINED=0 selects the direct-load kernels (A/B comparison and bisection)
bool e3b_tp_pipelined_enabled() {
  static const int on = [] { const char* v = getenv("E3B_TP_PIPELINED"); return (v && v[0] == '0') ? 0 : 1; }();
  return on != 0;
}

// E3B_TP_PAIRED=0 keeps one channel per thread in the pipelined backward kernels (A/B comparison)
// (E3B_TP_PAIRED=1: backward only, 2: forward only, default 3: both)
static int tp_paired_mask() {
  static const int m = [] { const char* v = getenv("E3B_TP_PAIRED"); return (v && v[0] >= '0' && v[0] <= '3') ? v[0] - '0' : 3; }();
  return m;
}
// E3B_TP_DECOUPLED=0: the one-channel-per-thread backward keeps its CTA barrier per edge and the staged TMA reduce-add (A/B)
bool e3b_tp_decoupled_enabled() {
  static const int on = [] { const char* v = getenv("E3B_TP_DECOUPLED"); return (v && v[0] == '0') ? 0 : 1; }();
  return on != 0;
}
static bool tp_paired_force() {
  static const int on = [] { const char* v = getenv("E3B_TP_PAIRED_FORCE"); return (v && v[0] == '1') ? 1 : 0; }();
  return on != 0;
}
// `preferred`: the generator's per-structure choice (measured, see gen_tp.py); E3B_TP_PAIRED_FORCE=1 overrides it
bool e3b_tp_paired_enabled(bool preferred) { return (tp_paired_mask() & 1) != 0 && (preferred || tp_paired_force()); }
bool e3b_tp_paired_fwd_enabled(bool preferred) { return (tp_paired_mask() & 2) != 0 && (preferred || tp_paired_force()); }

// Ring depth of the paired kernels: 3 like the one-channel kernels.  Measured at W2: 3 / 4 / 5 stages make no difference at equal
// CTAs per SM (the kernels are bound by arithmetic and instruction fetch, not by bytes in flight), and a deeper ring only
// costs occupancy.  E3B_TP_STAGES_B / E3B_TP_STAGES_F override (experiments).
int e3b_tp_stages(int bwd, size_t per_stage, size_t fixed) {
  static const int env_b = [] { const char* v = getenv("E3B_TP_STAGES_B"); return v ? atoi(v) : 0; }();
  static const int env_f = [] { const char* v = getenv("E3B_TP_STAGES_F"); return v ? atoi(v) : 0; }();
  long n = (bwd ? env_b : env_f);
  if (n <= 0) n = 3;
  const long n_max = (long)((227 * 1024 - 64 - fixed) / per_stage);
  if (n > n_max) n = n_max;
  if (n < 2) n = 2;
  return (int)n;
}

#define E3B_TP_PART 0   // this translation unit: part 0 of the generated kernels + the table (parts 1.. in tp_fast_p*.cu)
#include "tp_generated.cuh"

int e3b_gen_bwd_parts(const GenEntry* g, int mul) {
  if (mul == 64 && e3b_tp_pipelined_enabled() && e3b_tp_paired_enabled(g->paired_bwd_ok)) return g->paired_bwd_parts;
  return g->n_groups * ((mul + 31) / 32);
}

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
