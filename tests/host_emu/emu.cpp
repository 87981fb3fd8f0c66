// CPU emulation of the GENERATED tensor-product contraction code (tests only; never part of
// the product library).  Compiles csrc/tp_generated.cuh with g++ so that the generator's
// algebra (CG constants, offsets, layouts, gradients) is checked against the oracle without a GPU.
#define E3B_HOST_EMU 1
#include "../../equivariant-nn-zoo_b200/csrc/common.cuh"
#include "../../equivariant-nn-zoo_b200/csrc/tp_generated.cuh"

extern "C" int emu_groups(int sid) { return kEmuGroups[sid]; }

extern "C" int emu_tp_f64(int sid, int bwd, int64_t n_nodes, int mul, int64_t x_dim, int64_t sh_dim, int64_t w_dim,
                          int64_t y_dim, const double* x, const double* sh, const double* w, const double* gy,
                          const int64_t* in_ptr, const int32_t* in_nbr, const int32_t* in_eid, double* y,
                          double* gx_edge, double* gsh, double* gw) {
  TpArgs<double> a;
  a.x = x; a.sh = sh; a.w = w; a.gy = gy; a.y = y; a.gx_edge = gx_edge; a.gsh = gsh; a.gw = gw;
  a.in_ptr = in_ptr; a.in_nbr = in_nbr; a.in_eid = in_eid;
  a.n_nodes = n_nodes; a.x_dim = x_dim; a.sh_dim = sh_dim; a.w_dim = w_dim; a.y_dim = y_dim;
  a.mul = mul; a.n_chunks = (mul + 31) / 32; a.n_part = a.n_chunks * kEmuGroups[sid];
  return emu_tp<double>(sid, bwd, a);
}
