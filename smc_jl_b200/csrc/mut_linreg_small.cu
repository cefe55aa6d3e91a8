// mut_linreg_small.cu -- instantiates the mutation / evaluation kernels of these likelihood functors (see mutate_kernel.cuh)
#include "mutate_kernel.cuh"

namespace smc {
void register_linreg_small(std::vector<KernelEntry>& t)
{
    t.push_back(LINREG(1));
    t.push_back(LINREG(2));
    t.push_back(LINREG(3));
    t.push_back(LINREG(4));
    t.push_back(LINREG(5));
    t.push_back(LINREG(6));
}
}  // namespace smc
