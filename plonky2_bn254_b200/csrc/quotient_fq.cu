// K8 instantiation for kind 2 (fq); one translation unit per kind so they compile in parallel.
#include "quotient_impl.cuh"

namespace quot {

void run_fq(const Params& p, pbStream s) {
  pb_launch_lb<64, 16>("quotient fq pass 0", QuotientK<2, 0>{p}, p.count, s);
  pb_launch_lb<64, 16>("quotient fq pass 2", QuotientK<2, 2>{p}, p.count, s);  // 64 registers: no scratch arrays in this pass
  pb_launch_lb<64, 16>("quotient fq pass 3", QuotientK<2, 3>{p}, p.count, s);
}

}  // namespace quot
