// regression sizes between the fully specialised ones (general kernel only): K = 30, 31
#include "mutate_kernel.cuh"

namespace smc {
void register_linreg_fill_f(std::vector<KernelEntry>& t)
{
    t.push_back(LINREG_LITE(30));
    t.push_back(LINREG_LITE(31));
}
}  // namespace smc
