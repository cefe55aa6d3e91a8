// mut_linreg_mid.cu -- instantiates the mutation / evaluation kernels of these likelihood functors (see mutate_kernel.cuh)
#include "mutate_kernel.cuh"

namespace smc {
void register_linreg_mid(std::vector<KernelEntry>& t)
{
    t.push_back(LINREG(8));
    t.push_back(LINREG(10));
    t.push_back(LINREG(12));
    t.push_back(LINREG(16));
}
}  // namespace smc
