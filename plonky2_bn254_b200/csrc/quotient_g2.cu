// K8 instantiation for kind 1 (g2); one translation unit per kind so they compile in parallel.
#include "quotient_impl.cuh"

namespace quot {

void run_g2(const Params& p, pbStream s) {
  pb_launch_lb<64, 16>("quotient g2 pass 0", QuotientK<1, 0>{p}, p.count, s);
  pb_launch("quotient g2 pass 1", QuotientK<1, 1>{p}, p.count, s, 64);
  pb_launch_lb<64, 16>("quotient g2 pass 2", QuotientK<1, 2>{p}, p.count, s);  // 64 registers: no scratch arrays in this pass
  pb_launch_lb<64, 16>("quotient g2 pass 3", QuotientK<1, 3>{p}, p.count, s);
}

}  // namespace quot
