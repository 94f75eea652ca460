// K8 instantiation for kind 2 (fq); one translation unit per kind so they compile in parallel.
#include "quotient_impl.cuh"

namespace quot {

void run_fq(const Params& p, pbStream s) { pb_launch("quotient fq", QuotientK<2>{p}, p.size, s, 64); }

}  // namespace quot
