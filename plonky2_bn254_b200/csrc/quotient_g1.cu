// K8 instantiation for kind 0 (g1); one translation unit per kind so they compile in parallel.
#include "quotient_impl.cuh"

namespace quot {

int num_constraints(int kind, int nch) {
  int base = kind == 0 ? LY<0>::BASE_CONSTRAINTS : kind == 1 ? LY<1>::BASE_CONSTRAINTS : LY<2>::BASE_CONSTRAINTS;
  int nh = kind == 0 ? LY<0>::NH : kind == 1 ? LY<1>::NH : LY<2>::NH;
  return base + (nh + 2) * nch + 4 * nch;
}

void run_g1(const Params& p, pbStream s) {
  pb_launch_lb<64, 16>("quotient g1 pass 0", QuotientK<0, 0>{p}, p.count, s);
  pb_launch("quotient g1 pass 1", QuotientK<0, 1>{p}, p.count, s, 64);
  pb_launch_lb<64, 16>("quotient g1 pass 2", QuotientK<0, 2>{p}, p.count, s);  // 64 registers: no scratch arrays in this pass
  pb_launch_lb<64, 16>("quotient g1 pass 3", QuotientK<0, 3>{p}, p.count, s);
}

}  // namespace quot
