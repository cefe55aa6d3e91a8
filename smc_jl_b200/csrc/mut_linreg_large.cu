// mut_linreg_large.cu -- instantiates the mutation / evaluation kernels of these likelihood functors (see mutate_kernel.cuh)
#include "mutate_kernel.cuh"

namespace smc {
void register_linreg_large(std::vector<KernelEntry>& t)
{
    t.push_back(LINREG(24));
    t.push_back(LINREG(32));
}
}  // namespace smc
