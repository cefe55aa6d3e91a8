// mut_dsge.cu -- instantiates the mutation / evaluation kernels of these likelihood functors (see mutate_kernel.cuh)
#include "mutate_kernel.cuh"

namespace smc {
void register_dsge(std::vector<KernelEntry>& t)
{
    t.push_back(make_entry<ASLik>());   /* An-Schorfheide DSGE, config C4 */
}
}  // namespace smc
