// mut_linreg_20.cu -- instantiates the mutation / evaluation kernels of these likelihood functors (see mutate_kernel.cuh)
#include "mutate_kernel.cuh"

namespace smc {
void register_linreg_20(std::vector<KernelEntry>& t)
{
    t.push_back(LINREG(20));
}
}  // namespace smc
