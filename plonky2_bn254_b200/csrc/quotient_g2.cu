// K8 instantiation for kind 1 (g2); one translation unit per kind so they compile in parallel.
#include "quotient_impl.cuh"

namespace quot {

void run_g2(const Params& p, pbStream s) { pb_launch("quotient g2", QuotientK<1>{p}, p.size, s, 64); }

}  // namespace quot
